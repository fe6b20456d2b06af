timeout 600 python -m pytest tests/test_gpu_literals.py -x -q > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log
tail -4 gpurun_out/r2_pytest21.log
timeout 120 python profiles/exp_gemm_tc.py > gpurun_out/r2_gemm3.jsonl 2> gpurun_out/r2_gemm3.err
cat gpurun_out/r2_gemm3.jsonl; tail -3 gpurun_out/r2_gemm3.err
