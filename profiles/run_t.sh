cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python profiles/sweep_bench.py | tee gpurun_out/sweep_r01.jsonl
