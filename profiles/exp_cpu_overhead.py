"""CPU cost of one Engine.train_step call (launch path only: the GPU is never waited for inside the loop)."""
import os, sys, time, cProfile, pstats, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, B, K = 40943, 11, 200, 2048, 10
rng = np.random.default_rng(0)
eng = Engine(abi.TRANSE_L1, N, R, W, W, 12.0)
pos = torch.from_numpy(np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)).cuda()
kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, seed=1, opt_kind=abi.LAZY_ADAM, lr=1e-3)
for i in range(20): eng.train_step(pos, None, offset=i, **kw)
torch.cuda.synchronize()
n = 300   # 600 launches: below the driver's launch-queue depth, so the CPU never waits for the GPU
t0 = time.perf_counter()
for i in range(n): eng.train_step(pos, None, offset=i, **kw)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("CPU time per train_step call: %.1f us (GPU step ~35 us at this batch)" % ((t1 - t0) / n * 1e6))
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(n): eng.train_step(pos, None, offset=i, **kw)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
