mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_vocab_parallel_gpu.py tests/test_train_cli_gpu.py -m gpu -x -q ) > gpurun_out/gputest_c18_n2.log 2>&1; tail -5 gpurun_out/gputest_c18_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 > gpurun_out/c18_bench_n2.out 2>gpurun_out/c18_bench_n2.err; tail -c 400 gpurun_out/c18_bench_n2.err | tail -3
grep "^{" gpurun_out/c18_bench_n2.out > gpurun_out/c18_bench_n2.json; python tools/_show.py gpurun_out/c18_bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/c18_bench_n2.json')); print(d.get('dp_check'))"
