mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -q ) > gpurun_out/gputest_c16.log 2>&1; tail -8 gpurun_out/gputest_c16.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c16_$n.json 2>gpurun_out/c16_$n.err; tail -c 300 gpurun_out/c16_$n.err; python tools/_show.py gpurun_out/c16_$n.json; }
run s0 LTG_SPLIT_D=0
run s1 LTG_SPLIT_D=1
run s2 LTG_SPLIT_D=2
run s3 LTG_SPLIT_D=3
LTG_SPLIT_D=3 timeout 120 python tools/timeline.py step > gpurun_out/tl_step_c16_s3.txt 2>&1
LTG_SPLIT_D=1 timeout 120 python tools/timeline.py step > gpurun_out/tl_step_c16_s1.txt 2>&1
