"""Breakdown of the harness-flavoured step (tests.py:72-87: m = 20 480 oversampled positives and negatives, all-pairs hinge)."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, m = 40943, 11, 200, 20480
rng = np.random.default_rng(0)
eng = Engine(abi.TRANSE_L1, N, R, W, W, 12.0)
eng._ensure_train_buffers(abi.LAZY_ADAM)
pos = torch.from_numpy(np.stack([rng.integers(0, N, m), rng.integers(0, R, m), rng.integers(0, N, m)], 1).astype(np.int32)).cuda()
neg = torch.from_numpy(np.stack([rng.integers(0, N, m), rng.integers(0, R, m), rng.integers(0, N, m)], 1).astype(np.int32)).cuda()
def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t, f = eng.score(pos), eng.score(neg)
print("score_fwd x2        %.1f us" % timed(lambda: (eng.score(pos), eng.score(neg))))
for lk, name in ((abi.PAIRWISE_HINGE, "hinge"), (abi.PAIRWISE_LOGISTIC, "logistic")):
    print("loss allpairs %-8s %.1f us" % (name, timed(lambda: eng.loss(t, f, 1, lk, abi.ALLPAIRS, 1.0, 1.0))))
_, dt, df = eng.loss(t, f, 1, abi.PAIRWISE_HINGE, abi.ALLPAIRS, 1.0, 1.0)
print("score_bwd x2        %.1f us" % timed(lambda: (eng.score_bwd(pos, dt), eng.score_bwd(neg, df))))
print("opt_update2         %.1f us" % timed(lambda: eng.apply_optimizer(abi.LAZY_ADAM)))
print("whole step          %.1f us" % timed(lambda: eng.train_step(pos, neg, opt_kind=abi.LAZY_ADAM, K=1, loss_kind=abi.PAIRWISE_HINGE, pair_mode=abi.ALLPAIRS, margin=1.0)))
