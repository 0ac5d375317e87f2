for v in "LTG_ADAM_AFTER_MID=1" "LTG_ADAM_AFTER_MID=0" "LTG_ADAM_AFTER_MID=1 LTG_FUSED_DZ12=0" "LTG_ADAM_AFTER_MID=1 LTG_EARLY_ADAM=0"; do
  n=$(echo $v | tr ' =' '__')
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$n.json 2>gpurun_out/sw_$n.err
  python tools/_show.py gpurun_out/sw_$n.json; tail -c 300 gpurun_out/sw_$n.err
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/timeline.py step > gpurun_out/tl_step.txt 2>&1
