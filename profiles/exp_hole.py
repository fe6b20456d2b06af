"""C4 shape (HolE d=256, N=123 182, R=37, K=64, pointwise logistic): time of the unfused train step and of 1-N eval."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
N, R, W, B, K = 123182, 37, 256, int(os.environ.get("B", 2048)), 64
rng = np.random.default_rng(0)
eng = Engine(abi.HOLE, N, R, W, W, 12.0)
pos = torch.from_numpy(np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)).cuda()
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
ms = timed(lambda: eng.train_step(pos, None, opt_kind=abi.LAZY_ADAM, K=K, loss_kind=abi.POINTWISE_LOGISTIC, seed=1, offset=0))
print("HolE C4 train step B=%d K=%d: %.2f ms -> %.3g triples/s" % (B, K, ms, B * (1 + K) / ms * 1e3))
neg = eng.negatives(pos, K, 1, 0)
ms_f = timed(lambda: eng.score(neg))
print("  score_fwd on %d triples: %.2f ms" % (len(neg), ms_f))
d = torch.ones(len(neg), device="cuda")
ms_b = timed(lambda: eng.score_bwd(neg, d))
print("  score_bwd on %d triples: %.2f ms" % (len(neg), ms_b))
Q = 5000
tri = pos[:Q] if B >= Q else pos
for mode in (abi.TC_BF16X3_EXACT, abi.TC_BF16):
    packed = eng.pack_entities_tc(mode)
    ms_e = timed(lambda: eng.rank_counts_tc(tri, 0, mode=mode, packed=packed), 3)
    print("  1-N eval tail %d queries mode %d: %.2f ms -> %.3g q/s" % (len(tri), mode, ms_e, len(tri) / ms_e * 1e3))
eng.set_kernel_timing(True)
for mode in (abi.TC_BF16X3_EXACT, abi.TC_BF16):
    packed = eng.pack_entities_tc(mode)
    st = {}
    eng.rank_counts_tc(tri, 0, mode=mode, packed=packed, stats=st)
    torch.cuda.synchronize()
    print("mode", mode, st, "k_rank_tc ms %.3f refine ms %.3f" % (eng.last_kernel_ms(abi.TIMED_RANK_TC), eng.last_kernel_ms(abi.TIMED_RANK_REFINE)))
from KGEkeras._engine import _ptr, _stream, as_idx
import ctypes as C
t = as_idx(tri, eng.device); Q = t.shape[0]
q = torch.zeros(Q, W, device="cuda"); ts = torch.zeros(Q, device="cuda")
ms_q = timed(lambda: eng._check(eng.lib.kgeb_query_prep(eng.ctx, _ptr(eng.E), 0, eng.N, _ptr(eng.Rel), _ptr(t), Q, 0, _ptr(q), _stream())))
ms_t = timed(lambda: eng._check(eng.lib.kgeb_target_scores(eng.ctx, _ptr(eng.E), 0, eng.N, _ptr(q), _ptr(t), Q, 0, _ptr(ts), _stream())))
print("query_prep ms %.3f target_scores ms %.3f" % (ms_q, ms_t))
import cProfile, pstats
eng.set_kernel_timing(False)
packed = eng.pack_entities_tc(abi.TC_BF16X3_EXACT)
def run():
    for _ in range(5):
        eng.rank_counts_tc(tri, 0, mode=abi.TC_BF16X3_EXACT, packed=packed)
    torch.cuda.synchronize()
run()
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
