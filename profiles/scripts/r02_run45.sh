timeout 400 python -m pytest tests/test_gpu_rank_tc.py -x -q > gpurun_out/r2_pytest45.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest45.log
tail -3 gpurun_out/r2_pytest45.log
for p in 0 1; do KGEB_TC_2CTA=$p timeout 300 python profiles/exp_tc_pair.py > gpurun_out/r2_tc_pair_b$p.json 2> gpurun_out/r2_tc_pair$p.err; echo "rc=$?"; cat gpurun_out/r2_tc_pair_b$p.json; tail -3 gpurun_out/r2_tc_pair$p.err; done
