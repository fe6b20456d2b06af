import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200"), os.path.join(ROOT, "tests")]
from oracle import kge_oracle as ko
from test_gpu_fullsize import engine, triples
N, R, W, B, K = 40943, 11, 200, int(os.environ.get("B", 16384)), 10
eng, E, Rm = engine(ko.TRANSE_L1, N, R, W, seed=1)
rng = np.random.default_rng(2)
pos = triples(rng, B, N, R)
res = eng.forward_backward(pos, None, K=K, loss_kind=ko.PAIRWISE_HINGE, margin=1.0, seed=7, offset=3, want_scores=True, want_neg=True, opt_kind=ko.OPT_LAZY_ADAM)
neg = res["neg"].cpu().numpy()
g_f = eng.gR.clone().cpu().numpy().astype(np.float64); gE_f = eng.gE.clone().cpu().numpy().astype(np.float64)
eng.gE.zero_(); eng.gR.zero_(); eng.tE.zero_(); eng.tR.zero_()
eng.forward_backward(pos, neg, K=K, loss_kind=ko.PAIRWISE_HINGE, margin=1.0, force_unfused=True, opt_kind=ko.OPT_LAZY_ADAM)
g_u = eng.gR.clone().cpu().numpy().astype(np.float64); gE_u = eng.gE.clone().cpu().numpy().astype(np.float64)
eng.gE.zero_(); eng.gR.zero_(); eng.tE.zero_(); eng.tR.zero_()
eng.forward_backward(pos, neg, K=K, loss_kind=ko.PAIRWISE_HINGE, margin=1.0, opt_kind=ko.OPT_LAZY_ADAM)   # fused, index-driven (MODE 0)
g_i = eng.gR.clone().cpu().numpy().astype(np.float64)
# fp64 oracle
t = ko.score_triples(ko.TRANSE_L1, E, Rm, pos, 12.0, np.float64); f = ko.score_triples(ko.TRANSE_L1, E, Rm, neg, 12.0, np.float64)
L, dt, df = ko.loss_and_coefs(ko.PAIRWISE_HINGE, ko.PAIRED, t, f, K, 1.0)
idx = np.concatenate([pos, neg]); w = np.concatenate([dt, df])
s, p, o = ko.gather(E.astype(np.float64), Rm.astype(np.float64), idx)
r = s + p - o
gp = -np.sign(r) * w[:, None]
gR = np.zeros((R, W)); np.add.at(gR, idx[:, 1], gp)
for name, g in (("fused-philox", g_f), ("unfused", g_u), ("fused-index", g_i)):
    d = np.abs(g - gR)
    print(name, "max|diff| %.3e" % d.max(), "rows with diff>1e-7:", np.nonzero(d.max(1) > 1e-7)[0], "max|g| %.3e" % np.abs(gR).max())
print("gE fused vs unfused max diff %.3e" % np.abs(gE_f - gE_u).max())
print("row0 sample oracle", gR[0, :6], "\n fused", g_f[0, :6], "\n unfused", g_u[0, :6])
