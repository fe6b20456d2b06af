"""Dropout (models.py:122,144-150) in the CUDA path against its host restatement (oracle.kge_oracle.dropout_keep).

The masks are counter-based Philox draws, so -- like the negative sampler -- the kernel and the oracle must agree bit for
bit on WHICH elements are dropped; scores, loss and gradients then match to the usual fp32 bar.  The keep rate itself is
checked statistically against 1 - dp."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko
from test_gpu_kernels import make_engine, rand_triples, width

pytestmark = pytest.mark.gpu

KINDS = [ko.DISTMULT, ko.TRANSE_L1, ko.TRANSE_L2, ko.COMPLEX, ko.HOLE]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("d", [8, 100, 200])
def test_score_fwd_with_dropout_matches_oracle_masks(kind, d):
    if kind == ko.HOLE and d > 100:
        pytest.skip("HolE O(d^2) oracle is slow")
    p, seed, off = 0.2, 77, 5
    N, R, m = 300, 7, 257
    eng, E, Rm = make_engine(kind, N, R, d, seed=d, scale=3.0)
    W = width(kind, d)
    idx = rand_triples(np.random.default_rng(1), m, N, R)
    for stream in (0, 1):
        got = eng.score(idx, dropout=(p, seed, off, stream)).cpu().numpy().astype(np.float64)
        s, pr, o = ko.gather(E.astype(np.float64), Rm.astype(np.float64), idx)
        ms = ko.dropout_mult(kind, W, m, 0, stream, p, seed, off)
        mp = ko.dropout_mult(kind, W, m, 1, stream, p, seed, off)
        mo = ko.dropout_mult(kind, W, m, 2, stream, p, seed, off)
        want = ko.score(kind, s * ms, pr * mp, o * mo, 12.0, hole_fft=False)
        scale = np.abs(want).max() + 1.0
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * scale)
    # inference call afterwards is untouched by the armed-then-disarmed state
    np.testing.assert_allclose(eng.score(idx).cpu().numpy(), ko.score_triples(kind, E, Rm, idx, 12.0, np.float64, hole_fft=False),
                               rtol=0, atol=2e-5 * scale)


def test_dropout_keep_rate_and_independence():
    for p in (0.1, 0.2, 0.5):
        k0 = ko.dropout_keep(ko.TRANSE_L1, 200, 4000, 0, 0, p, 5, 1)
        k1 = ko.dropout_keep(ko.TRANSE_L1, 200, 4000, 0, 1, p, 5, 1)
        k2 = ko.dropout_keep(ko.TRANSE_L1, 200, 4000, 2, 0, p, 5, 1)
        for k in (k0, k1, k2):
            assert abs(k.mean() - (1 - p)) < 3e-3
        assert abs((k0 == k1).mean() - ((1 - p) ** 2 + p ** 2)) < 5e-3      # independent masks per call
        assert abs((k0 == k2).mean() - ((1 - p) ** 2 + p ** 2)) < 5e-3      # ... and per slot


@pytest.mark.parametrize("kind", [ko.DISTMULT, ko.TRANSE_L1, ko.COMPLEX, ko.HOLE])
@pytest.mark.parametrize("explicit_neg", [True, False])
def test_train_step_with_dropout(kind, explicit_neg):
    """Fused kernel (ring, DP flavour) or the unfused HolE sequence: loss and dense gradients against the oracle."""
    p, seed, off = 0.25, 4242, 3
    N, R, d, B, K = 500, 5, 24, 64, 4
    eng, E, Rm = make_engine(kind, N, R, d, seed=21, scale=3.0)
    rng = np.random.default_rng(2)
    pos = rand_triples(rng, B, N, R)
    pos[:4, 2] = pos[:4, 0]
    neg = ko.philox_negatives(pos, N, K, seed, off, ko.CORRUPT_EITHER).astype(np.int32)
    res = eng.forward_backward(pos, neg if explicit_neg else None, K=K, loss_kind=ko.PAIRWISE_LOGISTIC, margin=1.0,
                               corrupt_mode=ko.CORRUPT_EITHER, seed=seed, offset=off, want_scores=True, opt_kind=ko.OPT_SGD, dropout=p)
    want = ko.train_step_dropout(kind, ko.PAIRWISE_LOGISTIC, E, Rm, pos, neg, K, 1.0, 12.0, p, seed, off)
    np.testing.assert_allclose(res["pos_scores"].cpu().numpy(), want["pos_scores"], rtol=0, atol=3e-5 * (np.abs(want["pos_scores"]).max() + 1))
    np.testing.assert_allclose(res["neg_scores"].cpu().numpy(), want["neg_scores"], rtol=0, atol=3e-5 * (np.abs(want["neg_scores"]).max() + 1))
    assert abs(float(res["loss"].item()) - want["loss"]) <= 1e-5 * max(1.0, abs(want["loss"]))
    W = width(kind, d)
    gE = eng.gE.cpu().numpy()[:, :W].astype(np.float64)
    gR = eng.gR.cpu().numpy()[:, :W].astype(np.float64)
    np.testing.assert_allclose(gE, want["gE"], rtol=2e-4, atol=2e-6 * np.abs(want["gE"]).max())
    np.testing.assert_allclose(gR, want["gR"], rtol=2e-4, atol=2e-6 * np.abs(want["gR"]).max())


def test_model_fit_applies_dropout():
    """The drop-in model: dp > 0 changes the training loss (masks are drawn), predict() stays deterministic."""
    from KGEkeras import models as km
    rng = np.random.default_rng(0)
    N, R, d = 200, 4, 16
    E, Rm = ko.glorot_uniform(rng, N, d), ko.glorot_uniform(rng, R, d)
    tri = rand_triples(rng, 512, N, R)
    losses = {}
    for dp in (0.0, 0.3):
        m = km.DistMult(e_dim=d, r_dim=d, num_entities=N, num_relations=R, dp=dp, init_entities=E, init_relations=Rm,
                        loss_function="pairwize_hinge", negative_samples=4)
        m.compile(optimizer="sgd")
        losses[dp] = float(m.train_on_batch(tri).item())
        a, b = m.predict(tri), m.predict(tri)
        np.testing.assert_array_equal(a, b)
        if dp:
            c1, c2 = m(tri, training=True).cpu().numpy(), m(tri, training=True).cpu().numpy()
            assert not np.array_equal(c1, c2) and not np.array_equal(c1, a)
    assert losses[0.0] != losses[0.3]
