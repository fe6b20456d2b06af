cd $GRAFT_REPO_ROOT
for cfg in "KGEB_TRAIN_KERNEL=ring" "KGEB_TRAIN_KERNEL=ring KGEB_RING_D=3" "KGEB_TRAIN_KERNEL=ring KGEB_RING_D=2"; do env $cfg python profiles/exp_train.py 2>&1 | grep "flush True"; done
