"""Experiment: C2 ring kernel with a cold L2 (flushed) vs a warm L2 (the same launch repeated right after an untimed one)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, B, K = 40943, 11, 200, 16384, 10
rng = np.random.default_rng(0)
eng = Engine(abi.TRANSE_L1, N, R, W, W, 12.0)
nb = 40
pos = torch.from_numpy(np.stack([rng.integers(0, N, (nb, B)), rng.integers(0, R, (nb, B)), rng.integers(0, N, (nb, B))], 2).astype(np.int32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, seed=1, opt_kind=abi.LAZY_ADAM)
for mode in ("cold", "warm", "warm-table-only"):
    ts = []
    for i in range(nb):
        flush.zero_()
        if mode == "warm":
            eng.forward_backward(pos[(i + 1) % nb], None, offset=1000 + i, **kw)      # untimed: leaves E and gE in L2
        elif mode == "warm-table-only":
            _ = eng.E.sum()                                                           # streams the table through L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward_backward(pos[i], None, offset=i, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        eng.gE.zero_(); eng.gR.zero_(); eng.tE.zero_(); eng.tR.zero_()
    print(mode, "median ms %.4f" % float(np.median(ts[5:])))
