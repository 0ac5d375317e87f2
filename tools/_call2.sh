mkdir -p gpurun_out
REP=/tmp/r2_step_v2
LTG_NCU_EVAL=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --kernel-name 'regex:disc_fused|gemm_tcgen05|enc_gather|sample_pairs|dec_row_bwd|topk|vae_mid_fwd_a|tanh_bwd4|disc_gather' \
  -f -o $REP python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log; ls -la $REP.ncu-rep
python tools/ncu_table.py $REP.ncu-rep gpurun_out/r2_ncu_step_table_v2.txt gpurun_out/r2_ncu_step_table_v2.json > /dev/null 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_v2.csv 2>/dev/null
for k in disc_fused EpiLogitsStats enc_gather sample_pairs dec_row_bwd topk; do
  python tools/ncu_src.py $REP.ncu-rep $k 0 45 > gpurun_out/src_$k.txt 2>&1
done
# dz12 GEMM = first gemm<256,0,0,2,EpiStore>; wgrad = gemm<128,1,1,2,EpiStore> (first launch of that instantiation is the decoder wgrad)
python tools/ncu_src.py $REP.ncu-rep 'gemm_tcgen05_kernel<256, false, false, 2' 0 45 > gpurun_out/src_dz12.txt 2>&1
python tools/ncu_src.py $REP.ncu-rep 'gemm_tcgen05_kernel<128, true, true, 2' 0 45 > gpurun_out/src_wgrad.txt 2>&1
python tools/ncu_src.py $REP.ncu-rep 'gemm_tcgen05_kernel<128, false, true, 4' 0 45 > gpurun_out/src_dgrad.txt 2>&1
du -sh gpurun_out
