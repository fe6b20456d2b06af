set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke53.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke53.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest53.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest53.log
timeout 200 python bench.py > gpurun_out/r2_bench53.log 2> gpurun_out/r2_bench53.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r2_bench53.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench53.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype','clocks')})
print('roofline',d['roofline']['frac'],'e2e',d['e2e']['value'],'approx',d['approx_bf16']['value'],d['approx_bf16']['roofline']['frac'])
PY
