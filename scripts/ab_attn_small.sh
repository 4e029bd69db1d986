timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attn64" 2>&1 | tail -15
for sm in 0 1; do MB_ATTN_SMALL=$sm python scripts/prof_rowkernels.py 2>&1 | grep attn; done
for sm in 0 1 0 1; do
  MB_ATTN_SMALL=$sm python bench.py --workload pretrain_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_small_${sm}.json
  python - gpurun_out/ab_small_${sm}.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d['kernels']['attn_fwd'])
P
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
