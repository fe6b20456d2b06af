python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile_packed_v2.json 2> gpurun_out/r2_rank_tile.err
cat gpurun_out/r2_rank_tile_packed_v2.json; tail -5 gpurun_out/r2_rank_tile.err
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py tests/test_gpu_full_goldens.py tests/test_gpu_api.py -x -q -k "rank or validate or golden or tile" > gpurun_out/r2_pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest33.log
tail -6 gpurun_out/r2_pytest33.log
