timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --workload pretrain_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_pretrain_large_v6.json
python - gpurun_out/r02_bench_pretrain_large_v6.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('model_tflops'), d.get('gpu_launches'))
print({k:(v['ms'],v['launches']) for k,v in d['kernels'].items() if k in ('gemm','dec_assemble_bwd','masked_ce_fwd','masked_ce_bwd','masked_mse_fwd')})
P
python bench.py --workload pretrain_base --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_pretrain_base_v2.json
python bench.py --workload cls_large --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_cls_large_v3.json
for f in gpurun_out/r02_bench_pretrain_base_v2.json gpurun_out/r02_bench_cls_large_v3.json; do python - $f <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'))
P
done
