set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench11.log 2> gpurun_out/r2_bench11.err; echo "bench rc=$?" >> gpurun_out/r2_bench11.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches11.csv python bench.py --steps 2 --warmup 3 --skip-secondary --no-cpu-baseline > gpurun_out/r2_ncu11.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rank_tc -s 2 -c 2 -o gpurun_out/r2_rank_tc_c5 -f python bench.py --steps 1 --warmup 3 --skip-secondary --no-cpu-baseline > gpurun_out/r2_ncu11_full.log 2>&1
ls -la gpurun_out/ | tail -8
tail -c 600 gpurun_out/r2_bench11.err; tail -3 gpurun_out/r2_ncu11.log; tail -3 gpurun_out/r2_ncu11_full.log
