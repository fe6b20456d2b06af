set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_rank_tc.py -q -m gpu -k grouped > gpurun_out/r2_grouped55.log 2>&1; echo "grouped rc=$?"; tail -15 gpurun_out/r2_grouped55.log
timeout 60 python bench.py --skip-secondary --no-cpu-baseline > gpurun_out/r2_bench55.log 2> gpurun_out/r2_bench55.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench55.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench55.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['kernel_ms'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])
PY
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest55.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest55.log
