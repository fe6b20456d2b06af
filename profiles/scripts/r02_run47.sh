timeout 400 python -m pytest tests/test_gpu_rank_tc.py -x -q > gpurun_out/r2_pytest47.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest47.log; tail -2 gpurun_out/r2_pytest47.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench47.log 2> gpurun_out/r2_bench47.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench47.log') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'approx', d['approx_bf16']['value'], d['approx_bf16']['roofline']['frac'])
print('c3', {m:(d['eval_c3'][m]['value'], d['eval_c3'][m]['k_rank_tc_ms_one_side']) for m in ('exact','approx_bf16')}, 'c2', d['eval_c2']['value'])
PY
