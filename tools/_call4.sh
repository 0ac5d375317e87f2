mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
python bench.py --no-cpu-baseline > gpurun_out/bench_v4.json 2>gpurun_out/bench_v4.err; tail -c 300 gpurun_out/bench_v4.err
python tools/_show.py gpurun_out/bench_v4.json
python tools/timeline.py step > gpurun_out/tl_step_v4.txt 2>&1
for c in netflix askubuntu msd; do
  timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_${c}_n1.json 2>gpurun_out/bench_${c}_n1.err; tail -c 200 gpurun_out/bench_${c}_n1.err
  python tools/_show.py gpurun_out/bench_${c}_n1.json
done
