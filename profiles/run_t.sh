cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rank_tile' -c 1 -o gpurun_out/prof_r01_ranktile2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 --skip-large > gpurun_out/ncu_ranktile2.log 2>&1; echo rc=$?
