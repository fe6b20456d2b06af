"""GPU parity of HAKE (models.py:362-400) as the reference wrote it: the modulus term is one Frobenius norm over the whole
scoring call, so scores and gradients of a call are coupled.  Oracle: oracle.kge_oracle.hake_scores / hake_row_grads /
hake_step (fp64), themselves checked against a torch restatement written like the reference (tests/test_cpu_oracle_and_host.py)."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


def make(N, R, d, seed=0, gamma=12.0, eps=2.0, pw=0.5, mw=1.0):
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(seed)
    E = (rng.normal(size=(N, 2 * d)) * 0.05).astype(np.float32)
    Rm = (rng.normal(size=(R, 3 * d)) * 0.6).astype(np.float32)
    Rm[:, 2 * d:] += 0.5                                   # biases on both sides of the clip at 1 and of -|mod_p|
    hp = ko.hake_params(gamma, eps, 2 * d, pw, mw)
    return Engine(ko.HAKE, N, R, 2 * d, 3 * d, gamma, E=E, R=Rm, param=hp), E, Rm, hp


@pytest.mark.parametrize("d,m", [(4, 1), (6, 37), (50, 300), (100, 2048), (33, 129)])
def test_hake_scores(d, m):
    N, R = 64, 5
    eng, E, Rm, hp = make(N, R, d, seed=d)
    rng = np.random.default_rng(m)
    idx = np.stack([rng.integers(0, N, m), rng.integers(0, R, m), rng.integers(0, N, m)], 1).astype(np.int32)
    got = eng.score(idx).cpu().numpy()
    s, p, o = ko.gather(E.astype(np.float64), Rm.astype(np.float64), idx)
    want = ko.hake_scores(s, p, o, 12.0, *hp)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * np.abs(want).max())
    # the coupling itself: scoring a prefix of the call changes every score by the same modulus-norm difference
    half = eng.score(idx[: max(1, m // 2)]).cpu().numpy()
    w2 = ko.hake_scores(s[: max(1, m // 2)], p[: max(1, m // 2)], o[: max(1, m // 2)], 12.0, *hp)
    np.testing.assert_allclose(half, w2, rtol=2e-5, atol=2e-5 * np.abs(w2).max())


@pytest.mark.parametrize("loss", [ko.PAIRWISE_HINGE, ko.PAIRWISE_LOGISTIC, ko.POINTWISE_LOGISTIC])
@pytest.mark.parametrize("p", [0.0, 0.3])
@pytest.mark.parametrize("d", [6, 40])
def test_hake_step_against_oracle(loss, p, d):
    N, R, B, K = 50, 4, 32, 3
    eng, E, Rm, hp = make(N, R, d, seed=3 + d)
    rng = np.random.default_rng(9)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    neg = ko.philox_negatives(pos, N, K, 5, 1, ko.CORRUPT_BOTH).astype(np.int32)
    res = eng.forward_backward(pos, neg, K=K, loss_kind=loss, margin=1.0, seed=11, offset=2, opt_kind=ko.OPT_SGD, dropout=p, want_scores=True)
    want = ko.hake_step(loss, ko.PAIRED, E, Rm, pos, neg, K, 1.0, 12.0, hp, p=p, seed=11, offset=2)
    assert abs(float(res["loss"].item()) - want["loss"]) <= 3e-5 * max(1.0, abs(want["loss"]))
    np.testing.assert_allclose(res["pos_scores"].cpu().numpy(), want["pos_scores"], rtol=3e-5, atol=3e-5 * np.abs(want["pos_scores"]).max())
    np.testing.assert_allclose(res["neg_scores"].cpu().numpy(), want["neg_scores"], rtol=3e-5, atol=3e-5 * np.abs(want["neg_scores"]).max())
    for name, got, w in (("gE", eng.gE[:, :2 * d], want["gE"]), ("gR", eng.gR[:, :3 * d], want["gR"])):
        np.testing.assert_allclose(got.cpu().numpy(), w, rtol=3e-4, atol=3e-5 * np.abs(w).max(), err_msg=name)
    touched = np.zeros(N, bool)
    touched[np.concatenate([pos[:, [0, 2]].reshape(-1), neg[:, [0, 2]].reshape(-1)])] = True
    np.testing.assert_array_equal(eng.tE.cpu().numpy().astype(bool), touched)


def test_hake_model_trains_and_refuses_ranking():
    import KGEkeras
    from KGEkeras._abi import KgebError
    rng = np.random.default_rng(0)
    N, R, n = 200, 6, 4096
    train = np.stack([rng.integers(0, N, n), rng.integers(0, R, n), rng.integers(0, N, n)], 1).astype(np.int32)
    m = KGEkeras.HAKE(e_dim=16, r_dim=16, num_entities=N, num_relations=R, loss_function="pairwize_logistic", negative_samples=4, batch_size=512)
    m.compile(optimizer="adam")
    assert m.e_dim == 32 and m.r_dim == 48
    hist = m.fit(train, epochs=6, verbose=0)
    losses = hist.history["loss"]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    sc = m.predict(train[:64])
    assert np.asarray(sc).reshape(-1).shape == (64,)
    with pytest.raises(KgebError, match="HAKE has no 1-N ranking"):
        m._engine_().rank_counts(train[:8], 0)
