timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench57_n8.log 2> gpurun_out/r2_bench57_n8.err; echo "bench n8 rc=$?"
tail -c 300 gpurun_out/r2_bench57_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench57_n8.log') if l.startswith('{')][-1])
print('n_gpus', d['n_gpus'], 'value %.4g ms %.4g e2e %.4g approx %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['approx_bf16']['value']), d['clocks'])
print('c2 %.4g' % d['eval_c2']['value'], 'c3 exact %.4g approx %.4g' % (d['eval_c3']['exact']['value'], d['eval_c3']['approx_bf16']['value']), d['eval_c3']['exact']['ms_per_pass'])
for k in ('train_sharded_c5','train_sharded_c5_local_negatives','train_sharded_c5_alltoall'):
    v=d.get(k); print(k, None if v is None else (v['value'], v['ms_per_step']))
PY
