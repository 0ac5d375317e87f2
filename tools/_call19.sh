mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -q ) > gpurun_out/gputest_c19.log 2>&1; tail -6 gpurun_out/gputest_c19.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c19_bench.json 2>gpurun_out/c19_bench.err; tail -c 300 gpurun_out/c19_bench.err; python tools/_show.py gpurun_out/c19_bench.json
timeout 120 python tools/timeline.py eval > gpurun_out/tl_eval_c19.txt 2>&1; grep "^evaluate" gpurun_out/tl_eval_c19.txt
