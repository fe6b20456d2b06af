cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r01e.log 2> gpurun_out/bench_r01e.err; echo rc=$?; tail -c 300 gpurun_out/bench_r01e.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r01e.log') if l.startswith('{')][-1])
print('value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d['kernel_ms'], d['clocks'])
print(json.dumps(d['other_shapes'], indent=1))
print(d['eval']['value'], {k:v['queries_per_s'] for k,v in d['eval_tensor'].items() if isinstance(v,dict) and 'queries_per_s' in v})
PY
