cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_final.log 2> gpurun_out/bench_n2_final.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2_final.log') if l.startswith('{')][-1])
print('n_gpus', d['n_gpus'], 'value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), 'eval %.4g' % d['eval']['value'], {k: '%.4g' % v['queries_per_s'] for k,v in d['eval_tensor'].items() if isinstance(v,dict) and 'queries_per_s' in v}, d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
