timeout 300 python -m pytest tests/test_gpu_rank_tc.py -x -q -k "bf16 or stress or approx" > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest26.log
tail -3 gpurun_out/r2_pytest26.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench26.log 2> gpurun_out/r2_bench26.err; echo "bench rc=$?" >> gpurun_out/r2_bench26.err; tail -c 300 gpurun_out/r2_bench26.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench26.log') if x.startswith('{')]
d=json.loads(l[-1])
print('exact', d['value'], d['ms_per_step'], 'approx', json.dumps(d['approx_bf16']))
print('c3 approx', json.dumps(d['eval_c3']['approx_bf16']))
PY
