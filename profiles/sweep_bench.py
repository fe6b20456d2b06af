"""SURVEY 8(d) sweeps on the C2 shape (TransE-L1 d=200, N=40 943, R=11, K=10, pairwize_hinge, LazyAdam): batch of positives
B in {2 048 (the harness' batch, tests.py:119), 16 384, 131 072} x entity distribution of the positives in {uniform, Zipf(1.0)}
(Zipf concentrates the positives' rows -- and their gradient reductions -- on a few hot entities; the negatives stay uniform,
as generate_negative draws them).  One JSON line per point; CUDA events, L2 flushed before every step."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine

N, R, W, K = 40943, 11, 200, 10
rng = np.random.default_rng(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
zipf_p = 1.0 / np.arange(1, N + 1)
zipf_p /= zipf_p.sum()
perm = rng.permutation(N)   # hot entities are scattered over the table, not the first rows
for dist in ("uniform", "zipf"):
    for B in (2048, 16384, 131072):
        eng = Engine(abi.TRANSE_L1, N, R, W, W, 12.0)
        steps, warm = 20, 3
        def ents(n):
            return rng.integers(0, N, n) if dist == "uniform" else perm[rng.choice(N, n, p=zipf_p)]
        nb = steps + warm
        pos = torch.from_numpy(np.stack([ents(nb * B).reshape(nb, B), rng.integers(0, R, (nb, B)), ents(nb * B).reshape(nb, B)], 2).astype(np.int32)).cuda()
        kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, seed=3, opt_kind=abi.LAZY_ADAM)
        for i in range(warm):
            eng.train_step(pos[i], None, offset=i, lr=1e-3, **kw)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            eng.forward_backward(pos[warm + i], None, offset=warm + i, **kw)
            ev[i][1].record()
            eng.apply_optimizer(abi.LAZY_ADAM, lr=1e-3)
            ev[i][2].record()
        torch.cuda.synchronize()
        fb = float(np.median([e[0].elapsed_time(e[1]) for e in ev])); op = float(np.median([e[1].elapsed_time(e[2]) for e in ev]))
        print(json.dumps({"shape": "C2", "positives": dist, "B": B, "K": K, "k_train_ring_ms": round(fb, 4), "k_opt_update2_ms": round(op, 4),
                          "triples_per_s": B * (1 + K) / ((fb + op) * 1e-3), "hottest_entity_share": float(zipf_p[0]) if dist == "zipf" else 1.0 / N}))
        del eng, pos
