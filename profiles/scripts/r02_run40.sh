set -o pipefail
python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rank_tile3 -s 2 -c 2 -o gpurun_out/r2_rank_tile3 -f python profiles/exp_rank_tile.py > gpurun_out/r2_ncu40.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_ncu40.log
