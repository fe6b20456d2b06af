set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_seqk_pairs -s 1 -c 1 -o gpurun_out/prof_r2_seqk_pairs_c5 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-secondary > gpurun_out/ncu_r2_seqk.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_r2_seqk.log | cut -c1-300
ls -la gpurun_out/prof_r2_seqk_pairs_c5*
