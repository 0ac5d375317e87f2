mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest.log 2>&1; tail -15 gpurun_out/gputest.log
python bench.py --no-cpu-baseline > gpurun_out/bench_v3.json 2>gpurun_out/bench_v3.err; tail -c 300 gpurun_out/bench_v3.err
python tools/_show.py gpurun_out/bench_v3.json
python tools/timeline.py step > gpurun_out/tl_step_v3.txt 2>&1
REP=gpurun_out/r2_gemms_v3
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --kernel-name 'regex:gemm_tcgen05|disc_fused|enc_gather' -f -o $REP python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
ls -la $REP.ncu-rep; du -sh gpurun_out
