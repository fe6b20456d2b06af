timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_hake.py tests/test_gpu_canaries.py tests/test_gpu_api.py tests/test_gpu_rank_tc.py -x -q > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest28.log
tail -25 gpurun_out/r2_pytest28.log
