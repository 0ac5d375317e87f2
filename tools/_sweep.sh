for v in "LTG_DEC_CHUNKS=1" "LTG_DEC_CHUNKS=2" "LTG_DEC_CHUNKS=4" "LTG_DEC_CHUNKS=6" "LTG_DEC_CHUNKS=8" "LTG_DEC_CHUNKS=12"; do
  n=$(echo $v | tr ' =' '__')
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$n.json 2>gpurun_out/sw_$n.err
  python tools/_show.py gpurun_out/sw_$n.json; tail -c 300 gpurun_out/sw_$n.err
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
