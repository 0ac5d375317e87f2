mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v7_pdl1.json 2>gpurun_out/bench_v7.err; tail -c 300 gpurun_out/bench_v7.err
python tools/_show.py gpurun_out/bench_v7_pdl1.json
LTG_PDL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v7_pdl0.json 2>gpurun_out/bench_v7b.err; tail -c 300 gpurun_out/bench_v7b.err
python tools/_show.py gpurun_out/bench_v7_pdl0.json
timeout 300 python tools/timeline.py step > gpurun_out/tl_step_v7.txt 2>&1
