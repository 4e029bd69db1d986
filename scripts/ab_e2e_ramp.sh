for r in 18 8 12 24 32 18; do
  python bench.py --workload encoder_large --steps 10 --warmup 3 --no-cpu-baseline --e2e-ramp $r 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ramp $r', d['value'], d['e2e']['value'], round(d['e2e']['value']/d['value'],4), d['e2e'].get('e2e_uint8_inputs'))"
done
