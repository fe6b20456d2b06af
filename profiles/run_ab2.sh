cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_train_pipe|k_train_ring' -s 3 -c 1 -o gpurun_out/prof_r01_pipe python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 --skip-large > gpurun_out/ncu_pipe.log 2>&1; echo rc=$?
grep -v "^[0-9]*\. " gpurun_out/ncu_pipe.log | tail -4 | cut -c1-200
