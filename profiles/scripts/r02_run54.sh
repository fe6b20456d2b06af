timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/r2_pytest54.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest54.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench54_n2.log 2> gpurun_out/r2_bench54_n2.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/r2_bench54_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench54_n2.log') if l.startswith('{')][-1])
print('n_gpus', d['n_gpus'], 'value %.4g ms %.4g e2e %.4g approx %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['approx_bf16']['value']), d['clocks'])
for k in ('eval_c2','eval_c3','train_sharded_c5'):
    print(k, json.dumps(d.get(k))[:500])
PY
