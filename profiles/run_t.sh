cd $GRAFT_REPO_ROOT
python profiles/exp_l2warm.py
