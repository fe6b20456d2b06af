KGEB_GEMM=ffma timeout 120 python profiles/exp_gemm_tc.py > gpurun_out/r2_gemm.jsonl 2> gpurun_out/r2_gemm.err
timeout 120 python profiles/exp_gemm_tc.py >> gpurun_out/r2_gemm.jsonl 2>> gpurun_out/r2_gemm.err
cat gpurun_out/r2_gemm.jsonl; tail -3 gpurun_out/r2_gemm.err
