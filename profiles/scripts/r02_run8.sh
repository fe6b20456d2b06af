set -x
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
tail -30 gpurun_out/r2_pytest8.log
