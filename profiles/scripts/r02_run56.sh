timeout 300 python -m pytest tests/test_gpu_rank_tc.py -x -q 2>&1 | tail -1
for p in 1 0; do KGEB_TC_2CTA=$p timeout 300 python profiles/exp_tc_pair.py 2>/dev/null > gpurun_out/r2_tc_pair_e$p.json; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_tc_pair_e$p.json').read()); print('pair=$p', {k:(v['k_rank_tc_ms'], v['checksum']) for k,v in d.items() if isinstance(v,dict)})"; done
