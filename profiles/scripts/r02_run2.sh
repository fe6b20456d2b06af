set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err; echo "bench rc=$?" >> gpurun_out/r2_bench2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches2.csv python bench.py --steps 2 --warmup 3 --skip-secondary --no-cpu-baseline > gpurun_out/r2_ncu2.log 2>&1
tail -c 1500 gpurun_out/r2_bench2.err; tail -c 6000 gpurun_out/r2_bench2.log
