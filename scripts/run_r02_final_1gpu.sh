set -x
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default_1gpu_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/r02_bench_default_1gpu_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; head -c 400 gpurun_out/r02_bench_reference_arm.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_default_1gpu_final.json').read().strip().splitlines()[-1])
print(len(open('gpurun_out/r02_bench_default_1gpu_final.json').read().strip().splitlines()), 'line(s)')
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'])
p=d['pretrain']; print(p['value'], p['ms_per_step'], p['e2e']['value'], p['model_tflops'])
P
