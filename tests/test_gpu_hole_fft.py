"""HolE through the Fourier domain (kgeb_set_hole_mode(KGEB_HOLE_FFT); the reference's own formulation, models.py:233-234)
against the fp64 direct circular correlation of the oracle and against the library's direct kernels: scores within 1e-5 of the
summed magnitudes, gradients to fp32 FFT rounding, runs of equal relations (the Fourier-domain accumulation of the relation
gradient), dropout, widths that are not powers of two (direct kernels keep serving them)."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


def make(N, R, d, seed=0, scale=3.0):
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(seed)
    E, Rm = ko.glorot_uniform(rng, N, d) * scale, ko.glorot_uniform(rng, R, d) * scale
    eng = Engine(ko.HOLE, N, R, d, d, 12.0, E=E, R=Rm)
    eng.set_hole_mode(True)
    return eng, E, Rm


@pytest.mark.parametrize("d", [32, 64, 128, 256, 512])
@pytest.mark.parametrize("m", [1, 5, 1000])
def test_scores(d, m):
    N, R = 300, 7
    eng, E, Rm = make(N, R, d, seed=d)
    rng = np.random.default_rng(m)
    idx = np.stack([rng.integers(0, N, m), rng.integers(0, R, m), rng.integers(0, N, m)], 1).astype(np.int32)
    idx[: min(m, 3), 2] = idx[: min(m, 3), 0]
    got = eng.score(idx).cpu().numpy().astype(np.float64)
    if d <= 256:
        want = ko.score_triples(ko.HOLE, E, Rm, idx, 12.0, np.float64, hole_fft=False)
    else:
        want = ko.score_triples(ko.HOLE, E, Rm, idx, 12.0, np.float64, hole_fft=True)
    s, p, o = ko.gather(np.abs(E.astype(np.float64)), np.abs(Rm.astype(np.float64)), idx)
    bound = 1e-5 * np.sum(p * np.real(np.fft.ifft(np.conj(np.fft.fft(s)) * np.fft.fft(o))), axis=-1)   # summed magnitudes
    assert np.all(np.abs(got - want) <= bound), float(np.max(np.abs(got - want) / bound))
    eng.set_hole_mode(False)
    direct = eng.score(idx).cpu().numpy().astype(np.float64)
    assert np.all(np.abs(got - direct) <= 2 * bound)


@pytest.mark.parametrize("d", [32, 256])
@pytest.mark.parametrize("runs", [1, 16])
def test_gradients(d, runs):
    N, R, M = 200, 5, 960
    eng, E, Rm = make(N, R, d, seed=11)
    rng = np.random.default_rng(5)
    idx = np.stack([rng.integers(0, N, M), np.repeat(rng.integers(0, R, M // runs), runs), rng.integers(0, N, M)], 1).astype(np.int32)
    idx[:4, 2] = idx[:4, 0]
    dscore = rng.normal(size=M).astype(np.float32)
    dscore[10:20] = 0.0
    eng.score_bwd(idx, torch.from_numpy(dscore))
    gE, gR, tE, tR = ko.dense_grads(ko.HOLE, E, Rm, idx, dscore, 12.0)
    gs = max(np.max(np.abs(gE)), np.max(np.abs(gR)))
    np.testing.assert_allclose(eng.gE.cpu().numpy()[:, :d], gE, rtol=1e-4, atol=2e-5 * gs)
    np.testing.assert_allclose(eng.gR.cpu().numpy()[:, :d], gR, rtol=1e-4, atol=2e-5 * gs)
    np.testing.assert_array_equal(eng.tE.cpu().numpy().astype(bool), tE)
    np.testing.assert_array_equal(eng.tR.cpu().numpy().astype(bool), tR)


@pytest.mark.parametrize("p", [0.0, 0.3])
def test_step_with_dropout_matches_direct_kernels(p):
    """A whole unfused step (scores, loss, gradient buffers) in FFT mode against the direct kernels, same dropout masks."""
    N, R, d, B, K = 150, 6, 64, 64, 8
    eng, E, Rm = make(N, R, d, seed=2)
    rng = np.random.default_rng(1)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    neg = ko.philox_negatives(pos, N, K, 3, 1, ko.CORRUPT_BOTH).astype(np.int32)
    out = {}
    for mode in (True, False):
        eng.set_hole_mode(mode)
        if eng.gE is not None:
            eng.gE.zero_(); eng.gR.zero_(); eng.tE.zero_(); eng.tR.zero_()
        res = eng.forward_backward(pos, neg, K=K, loss_kind=ko.POINTWISE_LOGISTIC, seed=9, offset=4, opt_kind=ko.OPT_SGD, dropout=p, want_scores=True)
        out[mode] = (float(res["loss"].item()), res["pos_scores"].cpu().numpy(), res["neg_scores"].cpu().numpy(), eng.gE.cpu().numpy().copy(),
                     eng.gR.cpu().numpy().copy(), eng.tE.cpu().numpy().copy())
    a, b = out[True], out[False]
    assert abs(a[0] - b[0]) <= 1e-5 * max(1.0, abs(b[0]))
    sscale = max(np.abs(b[1]).max(), np.abs(b[2]).max())
    np.testing.assert_allclose(a[1], b[1], rtol=1e-4, atol=1e-5 * sscale)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-4, atol=1e-5 * sscale)
    gs = max(np.abs(b[3]).max(), np.abs(b[4]).max())
    np.testing.assert_allclose(a[3], b[3], rtol=1e-4, atol=2e-5 * gs)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-4, atol=2e-5 * gs)
    np.testing.assert_array_equal(a[5], b[5])


def test_other_widths_keep_the_direct_kernels():
    eng, E, Rm = make(50, 3, 100, seed=4)       # 100 is not a power of two
    rng = np.random.default_rng(0)
    idx = np.stack([rng.integers(0, 50, 200), rng.integers(0, 3, 200), rng.integers(0, 50, 200)], 1).astype(np.int32)
    got = eng.score(idx).cpu().numpy()
    eng.set_hole_mode(False)
    np.testing.assert_array_equal(got, eng.score(idx).cpu().numpy())


@pytest.mark.parametrize("loss", [ko.POINTWISE_LOGISTIC, ko.POINTWISE_HINGE, ko.POINTWISE_SQUARE])
@pytest.mark.parametrize("p", [0.0, 0.25])
@pytest.mark.parametrize("given_neg", [True, False])
def test_fused_step_matches_unfused_sequence(loss, p, given_neg):
    """kgeb_train_fwd_bwd in the Fourier-domain mode (one kernel: transform, score, loss term, gradients) against
    score_fwd -> loss_fwd_bwd -> score_bwd of the same mode, with listed negatives and with the in-kernel Philox draw."""
    N, R, d, B, K = 180, 5, 128, 96, 6
    eng, E, Rm = make(N, R, d, seed=6)
    rng = np.random.default_rng(2)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    neg = ko.philox_negatives(pos, N, K, 3, 7, ko.CORRUPT_BOTH).astype(np.int32)
    out = {}
    for unfused in (False, True):
        if eng.gE is not None:
            eng.gE.zero_(); eng.gR.zero_(); eng.tE.zero_(); eng.tR.zero_()
        l0 = eng.launches
        res = eng.forward_backward(pos, neg if given_neg else None, K=K, loss_kind=loss, margin=1.0, seed=3, offset=7, opt_kind=ko.OPT_SGD, dropout=p,
                                   want_scores=True, want_neg=True, force_unfused=unfused)
        out[unfused] = (float(res["loss"].item()), res["pos_scores"].cpu().numpy(), res["neg_scores"].cpu().numpy(), eng.gE.cpu().numpy().copy(),
                        eng.gR.cpu().numpy().copy(), eng.tE.cpu().numpy().copy(), eng.tR.cpu().numpy().copy(), res["neg"].cpu().numpy(), eng.launches - l0)
    f, u = out[False], out[True]
    assert f[8] == 1 and u[8] > 1                                  # one launch against the unfused sequence
    np.testing.assert_array_equal(f[7], neg)                       # the in-kernel draw is the oracle's Philox stream
    assert abs(f[0] - u[0]) <= 2e-6 * max(1.0, abs(u[0]))
    np.testing.assert_allclose(f[1], u[1], rtol=1e-6, atol=1e-6 * np.abs(u[1]).max())
    np.testing.assert_allclose(f[2], u[2], rtol=1e-6, atol=1e-6 * np.abs(u[2]).max())
    gs = max(np.abs(u[3]).max(), np.abs(u[4]).max())
    np.testing.assert_allclose(f[3], u[3], rtol=1e-4, atol=1e-5 * gs)
    np.testing.assert_allclose(f[4], u[4], rtol=1e-4, atol=1e-5 * gs)
    np.testing.assert_array_equal(f[5], u[5])
    np.testing.assert_array_equal(f[6], u[6])
    # and against the fp64 oracle (direct correlation)
    if p == 0.0:
        t = ko.score_triples(ko.HOLE, E, Rm, pos, 12.0, np.float64, hole_fft=False)
        fs = ko.score_triples(ko.HOLE, E, Rm, neg, 12.0, np.float64, hole_fft=False)
        L, dt_, df_ = ko.loss_and_coefs(loss, ko.PAIRED, t, fs, K, 1.0)
        assert abs(f[0] - L) <= 2e-5 * max(1.0, abs(L))
        gE1, gR1, _, _ = ko.dense_grads(ko.HOLE, E, Rm, pos, dt_, 12.0)
        gE2, gR2, _, _ = ko.dense_grads(ko.HOLE, E, Rm, neg, df_, 12.0)
        np.testing.assert_allclose(f[3][:, :d], gE1 + gE2, rtol=2e-4, atol=2e-5 * gs)
        np.testing.assert_allclose(f[4][:, :d], gR1 + gR2, rtol=2e-4, atol=2e-5 * gs)
