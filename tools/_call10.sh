mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_full_n2.log 2>&1; tail -6 gpurun_out/gputest_full_n2.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v10.json 2>gpurun_out/bench_v10.err; tail -c 300 gpurun_out/bench_v10.err
python tools/_show.py gpurun_out/bench_v10.json
