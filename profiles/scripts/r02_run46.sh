set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest46.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest46.log
tail -4 gpurun_out/r2_pytest46.log
python bench.py > gpurun_out/r2_bench46.log 2> gpurun_out/r2_bench46.err; echo "bench rc=$?" >> gpurun_out/r2_bench46.err; tail -c 300 gpurun_out/r2_bench46.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench46.log') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype','clocks')})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'approx', d['approx_bf16']['value'], d['approx_bf16']['roofline']['frac'])
for k in ('eval_c2','eval_c3'):
    print(k, json.dumps(d.get(k))[:1200])
PY
