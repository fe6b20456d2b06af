python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile_packed.json 2> gpurun_out/r2_rank_tile.err
KGEB_RANK_TILE=scalar python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile_scalar.json 2>> gpurun_out/r2_rank_tile.err
cat gpurun_out/r2_rank_tile_packed.json gpurun_out/r2_rank_tile_scalar.json; tail -5 gpurun_out/r2_rank_tile.err
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py tests/test_gpu_api.py tests/test_gpu_full_goldens.py -x -q -k "rank or validate or golden or tile or 1n" > gpurun_out/r2_pytest30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest30.log
tail -6 gpurun_out/r2_pytest30.log
