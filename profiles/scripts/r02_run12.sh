set -x
timeout 900 python -m pytest tests/test_gpu_rank_tc.py tests/test_gpu_full_goldens.py tests/test_gpu_api.py tests/test_gpu_canaries.py tests/test_gpu_siblings.py -x -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log
tail -40 gpurun_out/r2_pytest12.log
