"""GPU parity against golden vectors made by the REFERENCE'S OWN SOURCE (tests/golden/fp_ref.npz).

The fixture holds what /root/reference/KGEkeras/models.py (`EmbeddingModel.call` -> `func`, lines 134-162, 171-237, 383-489)
and the loss functions of utils.py:288-319 return when they are executed unmodified over the torch-backed `tensorflow`
stand-in of oracle/ref_tf_exec.py, in fp32 and in fp64, plus torch autograd through that op sequence
(tests/golden/make_golden_fp.py).  /root/reference does not exist on the GPU box: only the committed file is read here.
The CUDA path is called through the C ABI (KGEkeras._engine -> ctypes -> libkgeb200.so).

Bars, against the fp64 run of the reference source:
  * scores: 1e-5 * sum|terms| (oracle.kge_oracle.score_abs_terms); HAKE (one Frobenius norm over the call) 2e-5 relative;
  * gradients of sum_i c_i score_i: 1e-5 of the summed magnitudes of an entry's terms for the four hot models
    (dense_grads_abs), 1e-4 relative + 1e-5 of the largest entry for the sibling kinds;
  * losses: 2e-6 on the value, 2e-5 relative on d/dtrue, d/dfalse;
  * composed steps (model + loss as the harness wires them, tests.py:74-83): loss 1e-5, gradients 1e-4 relative + 1e-5 of
    the largest entry.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import fp_cases as fc  # noqa: E402

pytestmark = pytest.mark.gpu

FX = np.load(fc.FIXTURE)
KIND = {n: getattr(ko, n) for n in ("DISTMULT", "TRANSE_L1", "TRANSE_L2", "COMPLEX", "HOLE", "ROTATE", "PROTATE", "MODE_L1", "MODE_L2", "HAKE")}
HOT = ("DISTMULT", "TRANSE_L1", "TRANSE_L2", "COMPLEX", "HOLE")


def fx(prefix, key):
    return FX[f"{prefix}/{key}"]


def engine_for(prefix, kind):
    from KGEkeras._engine import Engine
    E, R = fx(prefix, "E"), fx(prefix, "R")
    gamma = float(fx(prefix, "gamma")[0]) if f"{prefix}/gamma" in FX else 12.0
    param = 1.0
    if f"{prefix}/param" in FX:
        p = FX[f"{prefix}/param"]
        param = tuple(float(x) for x in p) if kind == "HAKE" else float(p[0])
    eng = Engine(KIND[kind], E.shape[0], R.shape[0], E.shape[1], R.shape[1], gamma, E=E, R=R, param=param)
    return eng, E, R, gamma


@pytest.mark.parametrize("case", fc.SCORE_CASES, ids=[c[0] for c in fc.SCORE_CASES])
def test_scores_match_the_reference_source(case):
    name, _, _, kind, *_ = case
    pre = "score/" + name
    eng, E, R, gamma = engine_for(pre, kind)
    idx, want = fx(pre, "idx"), fx(pre, "score64")
    modes = [False, True] if (kind == "HOLE" and E.shape[1] >= 32) else [False]   # direct O(d^2) and Fourier-domain HolE kernels
    for fft in modes:
        if kind == "HOLE":
            eng.set_hole_mode(fft)
        got = eng.score(idx).cpu().numpy().astype(np.float64)
        if kind == "HAKE":
            np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * np.abs(want).max())
            continue
        bound = 1e-5 * ko.score_abs_terms(KIND[kind], E, R, idx, gamma)
        assert np.all(np.abs(got - want) <= bound), (fft, float(np.max(np.abs(got - want) / bound)))


@pytest.mark.parametrize("case", [c for c in fc.SCORE_CASES if c[3] != "HAKE"], ids=[c[0] for c in fc.SCORE_CASES if c[3] != "HAKE"])
def test_gradients_match_the_reference_autodiff(case):
    name, _, _, kind, *_ = case
    pre = "score/" + name
    eng, E, R, gamma = engine_for(pre, kind)
    idx, c = fx(pre, "idx"), fx(pre, "c")
    eng.score_bwd(idx, torch.from_numpy(c))
    got_gE = eng.gE.cpu().numpy()[:, :E.shape[1]].astype(np.float64)
    got_gR = eng.gR.cpu().numpy()[:, :R.shape[1]].astype(np.float64)
    wE, wR = fx(pre, "gE64"), fx(pre, "gR64")
    if kind in HOT:
        aE, aR = ko.dense_grads_abs(KIND[kind], E, R, idx, c, gamma)
        assert np.all(np.abs(got_gE - wE) <= 1e-5 * aE + 1e-30), float(np.max(np.abs(got_gE - wE) / (aE + 1e-30)))
        assert np.all(np.abs(got_gR - wR) <= 1e-5 * aR + 1e-30), float(np.max(np.abs(got_gR - wR) / (aR + 1e-30)))
    else:
        gs = max(np.abs(wE).max(), np.abs(wR).max())
        np.testing.assert_allclose(got_gE, wE, rtol=1e-4, atol=1e-5 * gs)
        np.testing.assert_allclose(got_gR, wR, rtol=1e-4, atol=1e-5 * gs)


LOSS_CASES = [(l, m, mg) for l in fc.LOSS_NAMES for m in (("PAIRED", "ALLPAIRS") if l.startswith("pairwize") else ("PAIRED",))
              for mg in fc.LOSS_MARGINS]


@pytest.mark.parametrize("lname,mode,margin", LOSS_CASES)
def test_losses_match_the_reference_source(lname, mode, margin):
    from KGEkeras._engine import Engine
    pre = f"loss/{lname}/{mode}/m{margin}"
    t, f = fx(pre, "t"), fx(pre, "f")
    eng = Engine(ko.DISTMULT, 10, 3, 4, 4)
    loss, dt, df = eng.loss(torch.from_numpy(t), torch.from_numpy(f), fc.LOSS_K, ko.LOSS_NAMES[lname], getattr(ko, mode), margin, 1.0)
    # Keras' clip bound 1 - 1e-7 is itself rounded in fp32 (0.99999988), in TF as in the kernel: BCE is held to the reference's fp32 run
    sfx = "32" if lname == "pointwize_cross_entropy" else "64"
    L, wdt, wdf = float(fx(pre, "loss" + sfx)[0]), fx(pre, "dt" + sfx).astype(np.float64), fx(pre, "df" + sfx).astype(np.float64)
    assert abs(float(loss.item()) - L) <= 2e-6 * max(1.0, abs(L))
    np.testing.assert_allclose(dt.cpu().numpy(), wdt, rtol=2e-5, atol=2e-7 * np.max(np.abs(wdt)))
    np.testing.assert_allclose(df.cpu().numpy(), wdf, rtol=2e-5, atol=2e-7 * np.max(np.abs(wdf)))


@pytest.mark.parametrize("case", fc.STEP_CASES, ids=[c[0] for c in fc.STEP_CASES])
def test_steps_match_the_reference_source(case):
    name, base, lname, mode, extra = case
    pre = "step/" + name
    kind = next(c for c in fc.SCORE_CASES if c[0] == base)[3]
    eng, E, R, _ = engine_for(pre, kind)
    pos, neg, K, margin = fx(pre, "pos"), fx(pre, "neg"), int(fx(pre, "K")[0]), float(fx(pre, "margin")[0])
    lk, pm = ko.LOSS_NAMES[lname], getattr(ko, mode)
    if mode == "ALLPAIRS":
        K = 1                  # the harness call: loss(score, fscore, margin, 1) on two equally long lists
    grads = {}
    if "literals" in extra:
        from KGEkeras._literal import LiteralHead
        head = LiteralHead(eng, fx(pre, "L"), extra["literal_activation"], rng=np.random.default_rng(0))
        head.set_weights(fx(pre, "W"), fx(pre, "b"))
        res = head.forward_backward(pos, neg, K=K, loss_kind=lk, pair_mode=pm, margin=margin, opt_kind=ko.OPT_SGD, want_scores=True)
        grads.update(gW=head.gW[:, :E.shape[1]], gb=head.gb[0, :E.shape[1]])
    else:
        res = eng.forward_backward(pos, neg, K=K, loss_kind=lk, pair_mode=pm, margin=margin, opt_kind=ko.OPT_SGD, want_scores=True)
    want_loss = float(fx(pre, "loss64")[0])
    if "regularization" in extra:
        eng.l3_reg(extra["regularization"], res["loss"])          # models.py:10-11,80-92: Keras adds the regulariser to the loss
        want_loss = float(fx(pre, "total64")[0])
    assert abs(float(res["loss"].item()) - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    for key in ("pos_scores", "neg_scores"):
        w = fx(pre, key + "64")
        np.testing.assert_allclose(res[key].cpu().numpy(), w, rtol=1e-4, atol=1e-4 * np.abs(w).max(), err_msg=key)
    grads.update(gE=eng.gE[:, :E.shape[1]], gR=eng.gR[:, :R.shape[1]])
    gscale = max(np.abs(fx(pre, k + "64")).max() for k in grads)
    for k, got in grads.items():
        np.testing.assert_allclose(got.cpu().numpy(), fx(pre, k + "64"), rtol=2e-4, atol=2e-5 * gscale, err_msg=k)
