cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for rep in 1 2; do for cfg in "KGEB_OPT_FUSED=0" "KGEB_OPT_FUSED=1"; do
env $cfg timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --eval-steps 1 --skip-large 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$cfg', 'value %.4g' % d['value'], 'kernel_ms', d['kernel_ms'], 'e2e %.4g' % d['e2e']['value'], 'launches', d['gpu_launches'])"
done; done
