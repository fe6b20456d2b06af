KGEB_TC_2CTA=1 timeout 300 python -m pytest tests/test_gpu_rank_tc.py -x -q -k "not ts_variant" > gpurun_out/r2_pytest44.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest44.log
tail -4 gpurun_out/r2_pytest44.log
if grep -q "rc=0" gpurun_out/r2_pytest44.log; then
for p in 1 0; do KGEB_TC_2CTA=$p timeout 300 python profiles/exp_tc_pair.py > gpurun_out/r2_tc_pair$p.json 2> gpurun_out/r2_tc_pair$p.err; echo "rc=$?"; cat gpurun_out/r2_tc_pair$p.json; tail -3 gpurun_out/r2_tc_pair$p.err; done
fi
