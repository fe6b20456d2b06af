"""C2 train step with the reference's default Dropout(0.2) on every gathered row vs dp = 0 (fused ring kernel, both flavours)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
for kind, N, R, W, B, K, name in ((abi.TRANSE_L1, 40943, 11, 200, 16384, 10, "C2 TransE-L1 d=200"), (abi.COMPLEX, 14541, 237, 400, 16384, 10, "C3 ComplEx d=200")):
    rng = np.random.default_rng(0)
    eng = Engine(kind, N, R, W, W, 12.0)
    nb = 30
    pos = torch.from_numpy(np.stack([rng.integers(0, N, (nb, B)), rng.integers(0, R, (nb, B)), rng.integers(0, N, (nb, B))], 2).astype(np.int32)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for dp in (0.0, 0.2):
        ts = []
        for i in range(nb):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.forward_backward(pos[i], None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, seed=1, offset=i, opt_kind=abi.LAZY_ADAM, dropout=dp)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            eng.apply_optimizer(abi.LAZY_ADAM, lr=1e-3)
        print(name, "dropout", dp, "k_train_ring median ms %.4f" % float(np.median(ts[5:])))
