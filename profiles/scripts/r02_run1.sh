set -x
python -m pytest tests -m gpu -x -q -k "not WN18RR_TransE and not FB15k" > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err; echo "bench rc=$?" >> gpurun_out/r2_bench1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench1_ref.log 2>&1
tail -5 gpurun_out/r2_pytest1.log; tail -c 1500 gpurun_out/r2_bench1.err; tail -c 3000 gpurun_out/r2_bench1.log
