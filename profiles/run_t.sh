cd $GRAFT_REPO_ROOT
python profiles/exp_cpu_overhead.py 2>&1 | head -40
