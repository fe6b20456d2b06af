timeout 120 python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile3.json 2> gpurun_out/r2_rank_tile3.err; echo "rc=$?"
cat gpurun_out/r2_rank_tile3.json; tail -5 gpurun_out/r2_rank_tile3.err
