# A/B of programmatic dependent launch inside ONE box (results differ +-3-5 % between boxes).
# MB_PDL bit mask: 1 = tensor kernels, 2 = row kernels.
for pdl in 0 1 2 3 0; do
  for wl in pretrain_large encoder_large; do
    MB_PDL=$pdl python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_pdl_${wl}_${pdl}.json
    python - gpurun_out/ab_pdl_${wl}_${pdl}.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
done
MB_PDL=3 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
