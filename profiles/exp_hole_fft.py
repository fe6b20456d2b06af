"""C4 shape (HolE d=256, K=64): score_fwd / score_bwd on the negative list, direct O(d^2) kernels against the Fourier-domain mode."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, B, K = 123182, 37, int(os.environ.get("W", 256)), 2048, 64
rng = np.random.default_rng(0)
eng = Engine(abi.HOLE, N, R, W, W, 12.0)
pos = torch.from_numpy(np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)).cuda()
neg = eng.negatives(pos, K, 1, 0)
d = torch.ones(len(neg), device="cuda")
eng._ensure_train_buffers(abi.SGD)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mode in (False, True):
    eng.set_hole_mode(mode)
    print("fft" if mode else "direct", "score_fwd %.3f ms  score_bwd %.3f ms  (%d triples)" % (timed(lambda: eng.score(neg)), timed(lambda: eng.score_bwd(neg, d)), len(neg)))
