set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
tail -8 gpurun_out/r2_pytest13.log
timeout 300 python profiles/exp_l2_tc.py > gpurun_out/r2_l2_tc.jsonl 2> gpurun_out/r2_l2_tc.err; cat gpurun_out/r2_l2_tc.jsonl; tail -3 gpurun_out/r2_l2_tc.err
