# A/B inside one box: LayerNorm folded into the GEMM epilogues (deterministic per-part statistics) vs standalone kernels
timeout 600 python -m pytest tests/test_gpu_lnfold.py -q -x 2>&1 | tail -3
MB_LN_FOLD=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_seg.py -q -x 2>&1 | tail -3
for v in 0 1 0 1; do
  MB_LN_FOLD=$v python bench.py --workload encoder_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_fold_${v}.json
  python - gpurun_out/ab_fold_${v}.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d['kernels']
print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d['e2e']['value'], {n:k[n]['ms'] for n in ('gemm','layernorm_fwd','attn_fwd') if n in k}, d['clocks']['sm_mhz'])
P
done
MB_LN_FOLD=1 python bench.py --workload encoder_base --batch 1 --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base b1 fold', d['value'], d['ms_per_step'])"
MB_LN_FOLD=0 python bench.py --workload encoder_base --batch 1 --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base b1 nofold', d['value'], d['ms_per_step'])"
