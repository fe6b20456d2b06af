cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -c 400 gpurun_out/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 profiles/c5_bench.py --steps 5 --warmup 2 --queries 32768 > gpurun_out/c5_n8.log 2> gpurun_out/c5_n8.err; echo "c5 n8 rc=$?"
tail -c 300 gpurun_out/c5_n8.err
grep "^{" gpurun_out/c5_n8.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 profiles/c5_bench.py --steps 2 --warmup 1 --queries 32768 --eval-mode exact > gpurun_out/c5_n8_exact.log 2> gpurun_out/c5_n8_exact.err; echo "c5 exact n8 rc=$?"
grep "^{" gpurun_out/c5_n8_exact.log | cut -c1-900
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
