# diagnostic: where does the fused kernel's TransE-L2 gradient differ from the fp64 truth by > 1e-6 of the summed term magnitudes?
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "kge-keras_b200"))
from oracle import kge_oracle as ko
from KGEkeras._engine import Engine
N, R, d, B, K = 120, 6, 16, 64, 5
kind = ko.TRANSE_L2
rng = np.random.default_rng(21)
E, Rm = ko.glorot_uniform(rng, N, d) * 4.0, ko.glorot_uniform(rng, R, d) * 4.0
rng = np.random.default_rng(31)
pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], axis=1).astype(np.int32)
neg = ko.philox_negatives(pos, N, K, 7, 0, ko.CORRUPT_EITHER).astype(np.int32)
neg[::3, 1] = (neg[::3, 1] + 1) % R
lk = ko.LOSS_NAMES["pairwize_hinge"]
sE = [np.zeros_like(E), np.zeros_like(E)]; sR = [np.zeros_like(Rm), np.zeros_like(Rm)]
w = ko.train_step(kind, lk, ko.PAIRED, E, Rm, pos, neg, K, 1.0, 2.0, ko.OPT_ADAM_DENSE, sE, sR, 0.05, 1, np.float64, eps=1e-3)
idx_all = np.concatenate([pos, neg]); ds = np.concatenate([np.ravel(w["dpos"]), np.ravel(w["dneg"])])
aE, aR = ko.dense_grads_abs(kind, E, Rm, idx_all, ds, 2.0)
for opt in (ko.OPT_LAZY_ADAM, ko.OPT_ADAM_DENSE):
    for rep in range(3):
        eng = Engine(kind, N, R, d, d, 2.0, E=E, R=Rm)
        res = eng.forward_backward(pos, neg, K=K, loss_kind=lk, pair_mode=ko.PAIRED, margin=1.0, want_scores=True, opt_kind=opt)
        gE = eng.gE.cpu().numpy()[:, :d].astype(np.float64)
        err = np.abs(gE - w["gE"]) / (aE + 1e-30)
        order = np.argsort(err.ravel())[::-1][:5]
        print("opt", opt, "rep", rep, "max", err.max())
        for f in order:
            i, k = np.unravel_index(f, err.shape)
            m = (idx_all[:, 0] == i) | (idx_all[:, 2] == i)
            print("   row", i, "k", k, "ratio %.3g got %.9g want %.9g abs %.4g refs %d active %d" % (err[i, k], gE[i, k], w["gE"][i, k], aE[i, k], m.sum(), (ds[m] != 0).sum()))
        ps = res["pos_scores"].cpu().numpy(); ns = res["neg_scores"].cpu().numpy() if "neg_scores" in res else None
        print("   pos score err", np.abs(ps - w["pos_scores"]).max())
        if ns is not None and "neg_scores" in w:
            hv = 1.0 + np.ravel(w["neg_scores"]).reshape(B, K) - w["pos_scores"][:, None]
            print("   smallest |hinge argument|", np.abs(hv).min())
