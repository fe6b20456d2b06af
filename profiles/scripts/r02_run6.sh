set -x
python -m pytest tests/test_gpu_rank_tc.py tests/test_gpu_siblings.py tests/test_gpu_fullsize.py tests/test_gpu_canaries.py tests/test_gpu_full_goldens.py tests/test_gpu_api.py -x -q -k "not FB15k" > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench6.log 2> gpurun_out/r2_bench6.err; echo "bench rc=$?" >> gpurun_out/r2_bench6.err
tail -15 gpurun_out/r2_pytest6.log; tail -c 1500 gpurun_out/r2_bench6.err; tail -c 6000 gpurun_out/r2_bench6.log
