set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_final.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-steps 1 --skip-large --skip-shapes > gpurun_out/ncu_launches_final.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_train_ring|k_opt_update2' -s 8 -c 2 -o gpurun_out/prof_final_train python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 --skip-large --skip-shapes > gpurun_out/ncu_final_train.log 2>&1; echo "ncu train rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rank_tc' -c 2 -o gpurun_out/prof_final_tc python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-steps 1 --skip-large --skip-shapes > gpurun_out/ncu_final_tc.log 2>&1; echo "ncu tc rc=$?"
ls -la gpurun_out/*final*
