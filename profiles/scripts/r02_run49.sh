KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_oldtc.so python profiles/exp_tc_pair.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old   ', {k:v['k_rank_tc_ms'] for k,v in d.items() if isinstance(v,dict)})"
for p in 0 1; do KGEB_TC_2CTA=$p python profiles/exp_tc_pair.py 2>/dev/null > gpurun_out/r2_tc_pair_c$p.json; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_tc_pair_c$p.json').read()); print('pair=$p', {k:v['k_rank_tc_ms'] for k,v in d.items() if isinstance(v,dict)})"; done
KGEB_TC_2CTA=1 timeout 300 python -m pytest tests/test_gpu_rank_tc.py -x -q -k "not ts_variant" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_rank_tc.py -x -q 2>&1 | tail -1
