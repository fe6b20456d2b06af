cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_rank_tc.py tests/test_gpu_fullsize.py tests/test_gpu_siblings.py -m gpu -x -q -k "tc or c3 or c5 or tensor" 2>&1 | tail -2
for md in bf16 exact; do timeout 120 python profiles/tc_bench.py --shape C3 --mode $md --iters 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pair=2 $md', d['ms'], d['algorithmic_tflops'])"; done
