"""Pass time of the captured RankPlan with the two sides serial (KGEB_RANK_FORK=0) or as parallel graph branches (=1), full table and a
1/8 row shard (what one rank of 8 runs), C2 (TransE-L1, CUDA cores) and C3 (ComplEx, tcgen05 exact / bf16)."""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
rng = np.random.default_rng(0)
out = {"fork": os.environ.get("KGEB_RANK_FORK", "auto")}
def csr(Q, N, lam):
    n = rng.poisson(lam, Q); offs = np.zeros(Q + 1, np.int64); offs[1:] = np.cumsum(n)
    ids = np.concatenate([np.sort(rng.choice(N, k, replace=False)) for k in n]).astype(np.int32) if offs[-1] else np.zeros(0, np.int32)
    return offs, ids
for name, kind, N, R, W, T, modes in (("C2", abi.TRANSE_L1, 40943, 11, 200, 3134, ("cuda",)), ("C3", abi.COMPLEX, 14541, 237, 400, 20466, ("exact", "bf16"))):
    lim = np.sqrt(6.0 / (N + W))
    eng = Engine(kind, N, R, W, W, 12.0, E=rng.uniform(-lim, lim, (N, W)).astype(np.float32), R=rng.uniform(-lim, lim, (R, W)).astype(np.float32))
    tri = torch.from_numpy(np.stack([rng.integers(0, N, T), rng.integers(0, R, T), rng.integers(0, N, T)], 1).astype(np.int32)).cuda()
    ct, ch = csr(T, N, 5.4), csr(T, N, 23.5)
    for mode in modes:
        for shard, rr in (("full", (0, N)), ("1/8", (0, (N + 7) // 8))):
            plan = eng.rank_plan(tri, ct, ch, row_range=rr, mode=mode).capture()
            for _ in range(3): plan.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): c = plan.run()
            e1.record(); torch.cuda.synchronize()
            out[f"{name} {mode} {shard}"] = {"ms_per_pass": round(e0.elapsed_time(e1) / 20, 4), "forked": plan.fork_sides, "checksum": int(c.sum().item())}
            del plan
print(json.dumps(out))
