mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -q ) > gpurun_out/gputest_c20.log 2>&1; tail -25 gpurun_out/gputest_c20.log | cut -c1-220
