# one ncu --set full capture per kernel (first matching launch is the warm-up call of scripts/prof_rowkernels.py)
for k in attn_fwd_small_kernel layernorm_bwd_kernel colsum_partial_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_$k python scripts/prof_rowkernels.py > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
python scripts/prof_rowkernels.py
