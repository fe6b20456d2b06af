timeout 120 python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile3_tma.json 2> gpurun_out/r2_rank_tile3.err; echo "rc=$?"
cat gpurun_out/r2_rank_tile3_tma.json; tail -5 gpurun_out/r2_rank_tile3.err
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py tests/test_gpu_full_goldens.py tests/test_gpu_api.py -x -q -k "rank or validate or golden or tile" > gpurun_out/r2_pytest39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest39.log
tail -6 gpurun_out/r2_pytest39.log
