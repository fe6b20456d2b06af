"""Experiment: time of the fused train kernel (C2 shape) with and without the gradient reductions (hinge inactive)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, B, K = 40943, 11, 200, 16384, 10
rng = np.random.default_rng(0)
eng = Engine(abi.TRANSE_L1, N, R, W, W, 12.0)
nb = 30
pos = torch.from_numpy(np.stack([rng.integers(0, N, (nb, B)), rng.integers(0, R, (nb, B)), rng.integers(0, N, (nb, B))], 2).astype(np.int32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
import itertools
for margin, do_flush in itertools.product((1.0, -1e9), (True, False)):
    ts = []
    for i in range(nb):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward_backward(pos[i], None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=margin, seed=1, offset=i, opt_kind=abi.LAZY_ADAM)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        eng.gE.zero_(); eng.gR.zero_()
    print(os.environ.get("KGEB_TRAIN_KERNEL", "pipe"), os.environ.get("KGEB_PIPE_DEPTH", ""), "margin", margin, "flush", do_flush, "median ms %.4f" % float(np.median(ts[5:])))
