set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01d.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_r01d.log
timeout 600 python bench.py > gpurun_out/bench_r01d.log 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r01d.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rank_tile' -c 2 -o gpurun_out/prof_r01_ranktile python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 --skip-large > gpurun_out/ncu_ranktile.log 2>&1; echo "ncu1 rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_train_ring<.*3, .*4|k_train_ringILi3ELi4' -c 2 -o gpurun_out/prof_r01_ring_c5 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 > gpurun_out/ncu_ring_c5.log 2>&1; echo "ncu2 rc=$?"
tail -5 gpurun_out/ncu_ring_c5.log
ls -la gpurun_out/*.ncu-rep
