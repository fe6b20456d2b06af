"""Three C5 steps per gradient mode (for an ncu capture: profiles/scripts/r02_run27.sh)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
N, R, W, B, K = 1_000_000, 1000, 512, 16384, 64
for mode in ("red", "sortseg_fused"):
    rng = np.random.default_rng(0)
    lim = float(np.sqrt(6.0 / (N + W)))
    eng = Engine(abi.COMPLEX, N, R, W, W, E=torch.empty(N, W, device=dev).uniform_(-lim, lim), R=rng.uniform(-0.1, 0.1, (R, W)).astype(np.float32))
    eng.set_grad_mode(mode)
    b = torch.from_numpy(np.stack([rng.integers(0, N, (3, B)), rng.integers(0, R, (3, B)), rng.integers(0, N, (3, B))], axis=2).astype(np.int32)).to(dev)
    for i in range(3):
        eng.train_step(b[i], None, offset=i, K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, seed=5, opt_kind=abi.LAZY_ADAM, lr=1e-3)
    torch.cuda.synchronize()
    del eng
    torch.cuda.empty_cache()
