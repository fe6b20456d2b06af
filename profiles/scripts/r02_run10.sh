set -x
timeout 420 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log
tail -30 gpurun_out/r2_pytest10.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench10_n2.log 2> gpurun_out/r2_bench10_n2.err; echo "bench rc=$?" >> gpurun_out/r2_bench10_n2.err
tail -c 1500 gpurun_out/r2_bench10_n2.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench10_n2.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    for k in ('value','ms_per_step','eval_c2','eval_c3','train_sharded_c5','train_sharded_c5_remote_red','train_sharded_c5_alltoall'):
        print(k, json.dumps(d.get(k))[:700])
PY
