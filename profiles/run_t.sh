cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_multi.py -m gpu -x -q -k "train or dropout or world1" 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --eval-steps 1 --skip-large --skip-shapes 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['kernel_ms'])"; done
