python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/cur.json 2>gpurun_out/cur.err; tail -c 300 gpurun_out/cur.err
LTG_FUSED_DZ12=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/cur_b.json 2>/dev/null
python tools/timeline.py step > gpurun_out/tl_step.txt 2>&1; python tools/timeline.py d > gpurun_out/tl_d.txt 2>&1
