"""C4 shape (HolE d=256, 2048 x 64 negatives, pointwize_logistic): the fused Fourier-domain step, for timing and ncu."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, B, K = 123182, 37, 256, 2048, 64
rng = np.random.default_rng(0)
eng = Engine(abi.HOLE, N, R, W, W, 12.0)
eng.set_hole_mode(True)
pos = torch.from_numpy(np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)).cuda()
kw = dict(K=K, loss_kind=abi.POINTWISE_LOGISTIC, seed=1, opt_kind=abi.LAZY_ADAM)
for i in range(3):
    eng.train_step(pos, None, offset=i, **kw)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); eng.forward_backward(pos, None, offset=9, **kw); e1.record(); eng.apply_optimizer(abi.LAZY_ADAM); e2.record()
torch.cuda.synchronize()
print("fused k_hole_fft_train %.3f ms, optimizer %.3f ms (%d triples)" % (e0.elapsed_time(e1), e1.elapsed_time(e2), B * (1 + K)))
