"""Throughput of the reference-facing API: model.fit() on a WN18RR-shaped synthetic graph (86 835 train triples), batch 2 048,
10 negatives, pairwize_hinge, through (a) a plain (n,3) array (in-kernel negatives), (b) KGEkeras.DataGenerator with the
device-side batch pipeline (the harness' all-pairs loss), (c) the same generator on the host (reference-style pipeline)."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
import KGEkeras
N, R, n = 40943, 11, 86835
rng = np.random.default_rng(0)
train = np.stack([rng.integers(0, N, n), rng.integers(0, R, n), rng.integers(0, N, n)], 1).astype(np.int32)
def model(pair_mode):
    m = KGEkeras.TransE(e_dim=200, r_dim=200, num_entities=N, num_relations=R, norm=1, gamma=12, loss_function="pairwize_hinge",
                        negative_samples=10, batch_size=2048, dp=0.0)
    m.pair_mode = pair_mode
    m.compile(optimizer="adam")
    return m
def timed(label, m, x, epochs=3, **kw):
    m.fit(x, epochs=1, verbose=0, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); m.fit(x, epochs=epochs, verbose=0, **kw); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / epochs
    print("%-44s %.3f s/epoch  %.3g positives/s  %.3g triples/s" % (label, dt, n / dt, n * 11 / dt))
m = model("paired"); timed("fit(array), in-kernel negatives, paired", m, train, batch_size=2048)
m = model("allpairs"); timed("fit(DataGenerator(engine)), all-pairs hinge", m, KGEkeras.DataGenerator(train, ns=10, batch_size=2048, engine=m._engine_()))
m = model("allpairs"); timed("fit(DataGenerator host), all-pairs hinge", m, KGEkeras.DataGenerator(train, ns=10, batch_size=2048), epochs=1)
import cProfile, pstats
m = model("paired"); m.fit(train, epochs=1, verbose=0, batch_size=2048); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); m.fit(train, epochs=3, verbose=0, batch_size=2048); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)
