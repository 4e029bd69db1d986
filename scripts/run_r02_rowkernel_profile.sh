python scripts/prof_rowkernels.py
ncu --set full --clock-control none --import-source on -k regex:'layernorm_bwd|attn_fwd_kernel|colsum_partial|layernorm_fwd' -c 8 -f -o gpurun_out/r02_rowkernels python scripts/prof_rowkernels.py > gpurun_out/ncu_rowk.log 2>&1
tail -3 gpurun_out/ncu_rowk.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_pretrain_large.csv python bench.py --workload pretrain_large --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
wc -l gpurun_out/r02_launches_pretrain_large.csv
