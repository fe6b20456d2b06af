cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_rank_tc.py -m gpu -x -q 2>&1 | tail -2
TAG=roles timeout 120 python profiles/exp_tc.py
for v in noepi noload; do TAG=roles-$v KGEB_LIB_PATH=$PWD/kge-keras_b200/lib_$v/libkgeb200.so timeout 120 python profiles/exp_tc.py; done
TAG=roles SHAPE=C5 timeout 120 python profiles/exp_tc.py
