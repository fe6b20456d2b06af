set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err; echo "bench rc=$?" >> gpurun_out/r2_bench4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench4_ref.log 2>&1
tail -5 gpurun_out/r2_pytest4.log; tail -c 1500 gpurun_out/r2_bench4.err; tail -c 6000 gpurun_out/r2_bench4.log
