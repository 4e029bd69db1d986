timeout 600 python -m pytest tests/test_gpu_visible.py -q -x 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for v in 0 1 0 1; do
  MB_EMBED_VISIBLE=$v python bench.py --workload pretrain_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_vis_${v}.json
  python - gpurun_out/ab_vis_${v}.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d['kernels']['gemm'], d.get('gpu_launches'))
P
done
