"""Wall time of validate() through the public API on bundled-dataset shapes (host CSR build + upload + ranking + read-back)."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
import KGEkeras
from KGEkeras import utils as ku
rng = np.random.default_rng(0)
for name, cls, kw, N, R, F, T in (("WN18RR-shape TransE-L1 d=200", KGEkeras.TransE, dict(norm=1, gamma=12), 40943, 11, 89869, 3134),
                                  ("FB15k-237-shape ComplEx d=200", KGEkeras.ComplEx, {}, 14541, 237, 289650, 20466)):
    m = cls(e_dim=200, r_dim=200, num_entities=N, num_relations=R, dp=0.0, **kw)
    filt = np.stack([rng.integers(0, N, F), rng.integers(0, R, F), rng.integers(0, N, F)], 1)
    test = np.stack([rng.integers(0, N, T), rng.integers(0, R, T), rng.integers(0, N, T)], 1)
    key = lambda a: (a[:, 0] * R + a[:, 1]) * N + a[:, 2]
    filt = filt[~np.isin(key(filt), key(test))]
    ku.validate(m, test, N, 32, filt); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): res = ku.validate(m, test, N, 32, filt)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("%-32s validate() %.2f ms for %d queries -> %.3g queries/s (mrr %.4f)" % (name, dt * 1e3, 2 * T, 2 * T / dt, res["mrr"]))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); ku.validate(m, test, N, 32, filt); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
