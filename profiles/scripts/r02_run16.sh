set -x
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench16_n8.log 2> gpurun_out/r2_bench16_n8.err; echo "bench rc=$?" >> gpurun_out/r2_bench16_n8.err
tail -c 1200 gpurun_out/r2_bench16_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench16_n4.log 2> gpurun_out/r2_bench16_n4.err; echo "bench rc=$?" >> gpurun_out/r2_bench16_n4.err
python - <<'PY'
import json
for n in (8, 4):
    l=[x for x in open(f'gpurun_out/r2_bench16_n{n}.log') if x.startswith('{')]
    if l:
        d=json.loads(l[-1])
        print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'approx', d['approx_bf16']['value'], 'e2e', d['e2e']['value'])
        for k in ('eval_c2','eval_c3','train_sharded_c5','train_sharded_c5_local_negatives','train_sharded_c5_alltoall'):
            v=d.get(k); print('  ',k, json.dumps(v)[:420] if v else v)
PY
