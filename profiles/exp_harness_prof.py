"""Where does a harness-configuration batch (literals + tanh Dense + all-pairs hinge) spend its time?  cProfile over fit()
plus per-stage device timings."""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "kge-keras_b200"))
import numpy as np, torch
import KGEkeras
rng = np.random.default_rng(41)
N, R, n = 14541, 237, 2048 * 24
train = np.stack([rng.integers(0, N, n), rng.integers(0, R, n), rng.integers(0, N, n)], axis=1).astype(np.int32)
for lit in (True, False):
    m = KGEkeras.DistMult(e_dim=100, r_dim=100, num_entities=N, num_relations=R, literals=rng.random((N, 5)) if lit else None,
                          literal_activation="tanh" if lit else None, loss_function="pairwize_hinge", margin=1, negative_samples=10, batch_size=2048)
    m.pair_mode = "allpairs"
    m.compile(optimizer="adam")
    gen = KGEkeras.DataGenerator(train, ns=10, batch_size=2048, engine=m._engine_())
    m.fit(gen, epochs=1, verbose=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); m.fit(gen, epochs=2, verbose=0); torch.cuda.synchronize()
    print("literals", lit, "ms/batch", (time.perf_counter() - t0) / (2 * len(gen)) * 1e3)
    # stages
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bs = [gen[i] for i in range(len(gen))]
    torch.cuda.synchronize(); print("  batch gen ms", (time.perf_counter() - t0) / len(gen) * 1e3)
    t0 = time.perf_counter()
    for (x, _) in bs:
        m.train_on_batch(x)
    torch.cuda.synchronize(); print("  train_on_batch ms", (time.perf_counter() - t0) / len(gen) * 1e3)
    pr = cProfile.Profile(); pr.enable(); m.fit(gen, epochs=1, verbose=0); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
