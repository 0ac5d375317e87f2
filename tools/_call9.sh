mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --config x1m > gpurun_out/bench_x1m_n8.out 2>gpurun_out/bench_x1m_n8.err; tail -c 400 gpurun_out/bench_x1m_n8.err | tail -3
timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 > gpurun_out/bench_ml20m_n8.out 2>gpurun_out/bench_ml20m_n8.err; tail -c 400 gpurun_out/bench_ml20m_n8.err | tail -3
timeout 400 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 > gpurun_out/bench_ml20m_n4.out 2>gpurun_out/bench_ml20m_n4.err; tail -c 400 gpurun_out/bench_ml20m_n4.err | tail -3
for f in x1m_n8 ml20m_n8 ml20m_n4; do grep "^{" gpurun_out/bench_$f.out > gpurun_out/bench_$f.json; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$f.json')); print('$f', round(d['value']), d['ms_per_step'], d.get('e2e',{}).get('value'), (d.get('dp_check') or d.get('vp_check') or {}).get('ok'))"; done
