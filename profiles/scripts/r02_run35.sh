set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke35.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke35.log; tail -3 gpurun_out/r2_smoke35.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest35.log
tail -5 gpurun_out/r2_pytest35.log
python bench.py > gpurun_out/r2_bench35.log 2> gpurun_out/r2_bench35.err; echo "bench rc=$?" >> gpurun_out/r2_bench35.err; tail -c 300 gpurun_out/r2_bench35.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench35.log') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype','clocks')})
print('roofline', d['roofline'])
print('e2e', d['e2e']['value'], 'approx', d['approx_bf16']['value'], d['approx_bf16']['roofline'])
for k in ('train','eval_c2','eval_c3','train_large_table','train_large_table_sortseg_fused','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:700])
PY
