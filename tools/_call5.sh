mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q ) > gpurun_out/gputest_k.log 2>&1; tail -4 gpurun_out/gputest_k.log
( timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_kernels_gpu.py ) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v5.json 2>gpurun_out/bench_v5.err; tail -c 300 gpurun_out/bench_v5.err
python tools/_show.py gpurun_out/bench_v5.json
timeout 300 python tools/timeline.py step > gpurun_out/tl_step_v5.txt 2>&1
