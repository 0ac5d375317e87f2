mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_train_cli_gpu.py -m gpu -q -k "resume" ) > gpurun_out/gputest_c21.log 2>&1; tail -12 gpurun_out/gputest_c21.log | cut -c1-200
