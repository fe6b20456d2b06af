set -x
timeout 600 python -m pytest tests/test_gpu_rank_tc.py tests/test_gpu_fullsize.py tests/test_gpu_canaries.py tests/test_gpu_siblings.py tests/test_gpu_full_goldens.py -x -q > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest25.log
tail -15 gpurun_out/r2_pytest25.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench25.log 2> gpurun_out/r2_bench25.err; echo "bench rc=$?" >> gpurun_out/r2_bench25.err; tail -c 300 gpurun_out/r2_bench25.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench25.log') if x.startswith('{')]
d=json.loads(l[-1])
print('exact', d['value'], d['ms_per_step'], 'approx', json.dumps(d['approx_bf16']))
print('c3', json.dumps(d['eval_c3'])[:900])
PY
