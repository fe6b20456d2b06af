cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -x -q -k "train or fit or multi_step" 2>&1 | tail -3
for cfg in "KGEB_TRAIN_KERNEL=ring" "KGEB_PIPE_DEPTH=2" "KGEB_PIPE_DEPTH=3" "KGEB_PIPE_DEPTH=4"; do
  env $cfg timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --eval-steps 1 --skip-large 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$cfg', 'value %.4g' % d['value'], 'kernel_ms', d['kernel_ms'], 'e2e %.4g' % d['e2e']['value'])"
done
