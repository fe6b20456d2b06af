timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 30 -c 3 -o gpurun_out/r2_gemm_tc -f python profiles/exp_gemm_tc.py > gpurun_out/r2_ncu22.log 2>&1
tail -3 gpurun_out/r2_ncu22.log
