timeout 600 python -m pytest tests/test_gpu_model.py -q -x -k "semseg_adapter_interpolate or encoder_full_batch" 2>&1 | tail -3
python scripts/prof_attn_bwd_shapes.py
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 2 -c 1 -f -o gpurun_out/r02_attn_bwd_n99 python scripts/prof_attn_bwd_shapes.py > gpurun_out/ncu_ab1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 10 -c 1 -f -o gpurun_out/r02_attn_bwd_n257 python scripts/prof_attn_bwd_shapes.py > gpurun_out/ncu_ab2.log 2>&1
tail -2 gpurun_out/ncu_ab1.log gpurun_out/ncu_ab2.log
