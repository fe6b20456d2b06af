python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile_packed2.json 2> gpurun_out/r2_rank_tile.err
KGEB_RANK_TILE=packed4 python profiles/exp_rank_tile.py > gpurun_out/r2_rank_tile_packed4.json 2>> gpurun_out/r2_rank_tile.err
cat gpurun_out/r2_rank_tile_packed2.json gpurun_out/r2_rank_tile_packed4.json; tail -5 gpurun_out/r2_rank_tile.err
