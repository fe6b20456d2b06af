set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke52.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke52.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest52.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest52.log
tail -4 gpurun_out/r2_pytest52.log
python bench.py > gpurun_out/r2_bench52.log 2> gpurun_out/r2_bench52.err; echo "bench rc=$?" >> gpurun_out/r2_bench52.err; tail -c 200 gpurun_out/r2_bench52.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench52.log') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['pipe_frac_issued_mmas'], 'e2e', d['e2e']['value'], 'approx', d['approx_bf16']['value'], d['approx_bf16']['roofline']['frac'])
print('c3', {m:(d['eval_c3'][m]['value'], d['eval_c3'][m]['ms_per_pass'], d['eval_c3'][m]['k_rank_tc_ms_one_side']) for m in ('exact','approx_bf16')}, 'c2', d['eval_c2']['value'], d['eval_c2']['roofline']['frac'])
print('train', d['train']['value'], 'large', d['train_large_table']['value'], d['train_large_table_sortseg_fused']['value'])
PY
