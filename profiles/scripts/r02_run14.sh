set -x
timeout 600 python -m pytest tests/test_gpu_sortseg.py -x -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest14.log
tail -25 gpurun_out/r2_pytest15.log
timeout 600 python profiles/exp_sortseg.py > gpurun_out/r2_sortseg3.jsonl 2> gpurun_out/r2_sortseg3.err; cat gpurun_out/r2_sortseg3.jsonl; tail -3 gpurun_out/r2_sortseg3.err
