mkdir -p gpurun_out
for v in "early:" "late:_late"; do
  n=${v%%:*}; sfx=${v##*:}
  LTG_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --no-dp-check > gpurun_out/bench_v8_$n.json 2>gpurun_out/bench_v8_$n.err; tail -c 200 gpurun_out/bench_v8_$n.err
  python tools/_show.py gpurun_out/bench_v8_$n.json
done
LTG_LIB_SUFFIX=_late timeout 300 python tools/timeline.py step > gpurun_out/tl_step_v8_late.txt 2>&1
timeout 300 python tools/timeline.py step > gpurun_out/tl_step_v8_early.txt 2>&1
LTG_LIB_SUFFIX=_late timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
