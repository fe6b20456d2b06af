set -x
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n profiles/exp_peer_train.py >> gpurun_out/r2_peer_train.jsonl 2>> gpurun_out/r2_peer_train.err
done
timeout 120 python profiles/exp_peer_train.py >> gpurun_out/r2_peer_train.jsonl 2>> gpurun_out/r2_peer_train.err
cat gpurun_out/r2_peer_train.jsonl; tail -5 gpurun_out/r2_peer_train.err
