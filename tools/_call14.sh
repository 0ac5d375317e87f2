mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_c14.log 2>&1; tail -5 gpurun_out/gputest_c14.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c14_$n.json 2>gpurun_out/c14_$n.err; tail -c 300 gpurun_out/c14_$n.err; python tools/_show.py gpurun_out/c14_$n.json; }
run new A=1
run old LTG_ADAM_ONE_ROUND=0 LTG_REUSE_GATHER=0 LTG_SMALL_ADAM_EARLY=0 LTG_TOPK_PREFILTER=0
run sp32 LTG_D_SP=32
run sp24 LTG_D_SP=24 LTG_D_SP3=16
timeout 120 python tools/timeline.py step > gpurun_out/tl_step_c14.txt 2>&1
