set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'k_train_ringILi3ELi2ELi1' -s 2 -c 1 -o gpurun_out/prof_r01_ring_c5 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 > gpurun_out/ncu_ring_c5.log 2>&1; echo "ncu2 rc=$?"
grep -v "^[0-9]*\. _Z" gpurun_out/ncu_ring_c5.log | tail -5 | cut -c1-300
ls -la gpurun_out/*.ncu-rep
