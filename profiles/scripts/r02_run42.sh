# CTA-pair variant: correctness (approximate + exact tests in a child environment) under a hard timeout, then C3/C5 timing
KGEB_TC_2CTA=1 timeout 300 python -m pytest tests/test_gpu_rank_tc.py -x -q -k "not ts_variant" > gpurun_out/r2_pytest42.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest42.log
tail -5 gpurun_out/r2_pytest42.log
