timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "rowops or training" 2>&1 | tail -2
for v in 0 1 0 1; do
  MB_COLSUM_REVERSE=$v python bench.py --workload pretrain_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('colsum_reverse=$v', d['value'], d['ms_per_step'], k['colsum'])"
done
