set -x
python -m pytest tests/test_gpu_sortseg.py -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
python profiles/exp_sortseg.py > gpurun_out/r2_sortseg.jsonl 2> gpurun_out/r2_sortseg.err
tail -15 gpurun_out/r2_pytest3.log; cat gpurun_out/r2_sortseg.jsonl; tail -5 gpurun_out/r2_sortseg.err
