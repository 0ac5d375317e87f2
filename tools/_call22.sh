mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "d_step or g_step or three_phases or evaluate" ) > gpurun_out/gputest_c22.log 2>&1; tail -4 gpurun_out/gputest_c22.log | cut -c1-200
