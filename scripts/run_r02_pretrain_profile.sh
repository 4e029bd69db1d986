timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --workload pretrain_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_pretrain_large_v5.json
python - gpurun_out/r02_bench_pretrain_large_v5.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('model_tflops'), d.get('gpu_launches'))
print({k:(v['ms'],v['launches']) for k,v in d['kernels'].items()})
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_pretrain_large_v2.csv python bench.py --workload pretrain_large --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
wc -l gpurun_out/r02_launches_pretrain_large_v2.csv
