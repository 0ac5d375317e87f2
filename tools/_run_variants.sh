python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in base "LTG_EARLY_ADAM=0" "LTG_GRAPH_PRIORITY=0" "LTG_EARLY_ADAM=0 LTG_GRAPH_PRIORITY=0"; do
  n=$(echo $v | tr ' =' '__')
  if [ "$v" = base ]; then env python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v_$n.json 2>gpurun_out/v_$n.err;
  else env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v_$n.json 2>gpurun_out/v_$n.err; fi
done
python tools/timeline.py step > gpurun_out/tl_step.txt 2>&1; python tools/timeline.py g > gpurun_out/tl_g.txt 2>&1
