"""k_rank_tile (scalar, round 1) against k_rank_tile2 (packed FADD2, flat work list) on the C2 shape: TransE-L1 d=200, N=40943, 3134 test
triples, both sides.  Run once per variant (the choice is read once per process): KGEB_RANK_TILE=scalar python profiles/exp_rank_tile.py"""
import os, sys, json, hashlib, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
rng = np.random.default_rng(0)
out = {"variant": os.environ.get("KGEB_RANK_TILE", "packed")}
for name, kind, N, R, d, T in (("C2 TransE-L1 d=200 WN18RR shape", abi.TRANSE_L1, 40943, 11, 200, 3134),
                               ("TransE-L1 d=100 N=14541 (FB15k-237 shape)", abi.TRANSE_L1, 14541, 237, 100, 20466),
                               ("ragged TransE-L1 d=52 N=1001", abi.TRANSE_L1, 1001, 7, 52, 333)):
    lim = np.sqrt(6.0 / (N + d))
    E = rng.uniform(-lim, lim, (N, d)).astype(np.float32); Rm = rng.uniform(-lim, lim, (R, d)).astype(np.float32)
    eng = Engine(kind, N, R, d, d, 12.0, E=E, R=Rm)
    tri = torch.from_numpy(np.stack([rng.integers(0, N, T), rng.integers(0, R, T), rng.integers(0, N, T)], 1).astype(np.int32)).cuda()
    res = {}
    for side, sname in ((abi.SIDE_TAIL, "tail"), (abi.SIDE_HEAD, "head")):
        c = eng.rank_counts(tri, side)
        torch.cuda.synchronize()
        eng.set_kernel_timing(True)
        for _ in range(10): c = eng.rank_counts(tri, side)
        torch.cuda.synchronize()
        ms, n = eng.kernel_ms_sum(abi.TIMED_RANK_TILE)
        eng.set_kernel_timing(False)
        fadd = (2 if side == abi.SIDE_TAIL else 3) * T * N * d          # lane-FADDs of the L1 distance
        res[sname] = {"kernel_ms": round(ms / n, 4), "counts_sha": hashlib.sha1(c.cpu().numpy().tobytes()).hexdigest()[:12],
                      "gfadd_per_s": round(fadd / (ms / n * 1e-3) / 1e9, 1)}
    out[name] = res
print(json.dumps(out))
