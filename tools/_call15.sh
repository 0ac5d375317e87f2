mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_c15.log 2>&1; tail -5 gpurun_out/gputest_c15.log
timeout 120 python tools/timeline.py eval > gpurun_out/tl_eval_c15.txt 2>&1; grep -A12 "totals" gpurun_out/tl_eval_c15.txt; grep "^evaluate" gpurun_out/tl_eval_c15.txt
LTG_TOPK_PREFILTER=0 timeout 120 python tools/timeline.py eval > gpurun_out/tl_eval_c15_nopre.txt 2>&1; grep -A6 "totals" gpurun_out/tl_eval_c15_nopre.txt
