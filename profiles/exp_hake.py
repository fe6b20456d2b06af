"""HAKE (as written in the reference) on the WN18RR shape: d=100 (entity rows 200, relation rows 300), 16 384 x (1+10) triples."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, d, B, K = 40943, 11, 100, 16384, 10
rng = np.random.default_rng(0)
eng = Engine(abi.HAKE, N, R, 2 * d, 3 * d, 12.0, param=(float(((12.0 + 2.0) / (2 * d) / 2) / np.pi), 0.5, 1.0))   # embedding_range / pi, phase_weight, mod_weight
pos = torch.from_numpy(np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)).cuda()
kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, seed=1, opt_kind=abi.LAZY_ADAM)
for i in range(3):
    eng.train_step(pos, None, offset=i, **kw)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(10):
    eng.train_step(pos, None, offset=10 + i, **kw)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print("HAKE step %.3f ms -> %.3g triples/s" % (ms, B * (1 + K) / ms * 1e3))
