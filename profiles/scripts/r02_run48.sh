for rep in 1 2; do
KGEB_EXP_SMALL=1 KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_oldtc.so python profiles/exp_tc_pair.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old', {k:v['k_rank_tc_ms'] for k,v in d.items() if isinstance(v,dict)})"
KGEB_EXP_SMALL=1 python profiles/exp_tc_pair.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new', {k:v['k_rank_tc_ms'] for k,v in d.items() if isinstance(v,dict)})"
done
KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_oldtc.so python profiles/exp_tc_pair.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old', {k:v['k_rank_tc_ms'] for k,v in d.items() if isinstance(v,dict)})"
python profiles/exp_tc_pair.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new', {k:v['k_rank_tc_ms'] for k,v in d.items() if isinstance(v,dict)})"
