set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench7_n2.log 2> gpurun_out/r2_bench7_n2.err; echo "bench rc=$?" >> gpurun_out/r2_bench7_n2.err
tail -25 gpurun_out/r2_pytest7.log; tail -c 2500 gpurun_out/r2_bench7_n2.err; tail -c 7000 gpurun_out/r2_bench7_n2.log
