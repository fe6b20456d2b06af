set -o pipefail
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-secondary"
$CMD > gpurun_out/r2_plain55.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rank_tc_t -s 12 -c 2 -o gpurun_out/r2_rank_tc_pair_c5_bf16 -f $CMD > gpurun_out/r2_ncu55.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu55.log
