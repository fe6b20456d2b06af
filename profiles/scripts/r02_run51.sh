for p in 0 1; do
KGEB_TC_2CTA=$p python bench.py --no-cpu-baseline --skip-secondary > gpurun_out/r2_bench51_$p.log 2> gpurun_out/r2_bench51_$p.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2_bench51_$p.log') if x.startswith('{')]
d=json.loads(l[-1])
print('pair=$p', {k:d[k] for k in ('value','ms_per_step')}, d['clocks']['sm_mhz'], 'kernel_ms', d['kernel_ms'], 'roofline', d['roofline']['frac'], d['roofline']['pipe_frac_issued_mmas'], 'e2e', d['e2e']['value'], 'approx', d['approx_bf16']['value'], d['approx_bf16']['k_rank_tc_ms'], d['approx_bf16']['roofline']['frac'])
PY
done
