mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v5.json 2>gpurun_out/bench_v5.err; tail -c 300 gpurun_out/bench_v5.err
python tools/_show.py gpurun_out/bench_v5.json
( timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_vocab_parallel_gpu.py -m gpu -x -q ) > gpurun_out/gputest_n2.log 2>&1; tail -5 gpurun_out/gputest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_ml20m_n2.json 2>gpurun_out/bench_ml20m_n2.err; tail -c 300 gpurun_out/bench_ml20m_n2.err
python tools/_show.py gpurun_out/bench_ml20m_n2.json
