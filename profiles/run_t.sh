cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_api.py -m gpu -x -q --durations=8 2>&1 | tail -30
