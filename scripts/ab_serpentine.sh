# A/B inside one box: consumers walk their rows opposite to their producer (L2 reuse) vs everything upwards
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1 0 1; do
  for wl in encoder_large pretrain_large; do
    MB_SERPENTINE=$v python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_serp_${wl}_${v}.json
    python - gpurun_out/ab_serp_${wl}_${v}.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d['kernels']
print(sys.argv[1], d.get('value'), d.get('ms_per_step'), {n:k[n]['ms'] for n in ('gemm','layernorm_fwd','layernorm_bwd','attn_fwd') if n in k}, d['clocks']['sm_mhz'])
P
  done
done
