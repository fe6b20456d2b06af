set -x
timeout 600 python -m pytest tests/test_gpu_literals.py -x -q > gpurun_out/r2_pytest18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest18.log
tail -30 gpurun_out/r2_pytest18.log
KGEB_GEMM=ffma timeout 300 python profiles/exp_harness_prof.py > gpurun_out/r2_harness_ffma.log 2>&1; tail -5 gpurun_out/r2_harness_ffma.log
timeout 300 python profiles/exp_harness_prof.py > gpurun_out/r2_harness_tc.log 2>&1; tail -5 gpurun_out/r2_harness_tc.log
