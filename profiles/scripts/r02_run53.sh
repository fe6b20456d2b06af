set -o pipefail
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-secondary"
$CMD > gpurun_out/r2_plain53.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_headline_final.csv $CMD > gpurun_out/r2_ncu53a.log 2>&1
echo "launches rc=$?"
$CMD > gpurun_out/r2_plain53b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rank_tc_t -s 6 -c 2 -o gpurun_out/r2_rank_tc_pair_c5 -f $CMD > gpurun_out/r2_ncu53b.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu53b.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench53_ref.log 2>&1; tail -c 400 gpurun_out/r2_bench53_ref.log
