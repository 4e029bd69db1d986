set -x
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default_1gpu_v3.json 2> gpurun_out/bench_v3.err; tail -c 600 gpurun_out/r02_bench_default_1gpu_v3.json | head -c 300
python bench.py --workload encoder_base --batch 1 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_encoder_base_b1.json 2>/dev/null
python bench.py --workload encoder_base --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_encoder_base.json 2>/dev/null
python bench.py --workload pretrain_base --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pretrain_base.json 2>/dev/null
python bench.py --workload cls_large --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cls_large_v2.json 2>/dev/null
python bench.py --workload cls_large --linear-probe --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cls_large_linear_probe.json 2>/dev/null
python bench.py --workload cls_large --label-smoothing 0.1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cls_large_smooth.json 2>/dev/null
for f in gpurun_out/r02_bench_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'))
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r02_launches_default.csv
