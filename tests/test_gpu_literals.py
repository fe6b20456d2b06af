"""GPU parity of the literal branch (models.py:114-120,152-158; the configuration the reference harness trains, tests.py:113,
125-126: DistMult + 5 literals + tanh Dense + all-pairs hinge): gathers with dropout, the Dense GEMMs, activation forward and
backward, gradient scatter and the optimizer step against oracle.kge_oracle.literal_step (fp64); inference table; model API."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu
ACTS = {"tanh": ko.ACT_TANH, "relu": ko.ACT_RELU, "sigmoid": ko.ACT_SIGMOID, "linear": ko.ACT_LINEAR}


def make(kind, N, R, d, wl, aname, seed=0):
    from KGEkeras._engine import Engine
    from KGEkeras._literal import LiteralHead
    rng = np.random.default_rng(seed)
    W = 2 * d if kind == ko.COMPLEX else d
    E, Rm = ko.glorot_uniform(rng, N, W) * 4, ko.glorot_uniform(rng, R, W) * 4
    L = rng.random((N, wl)).astype(np.float32)
    eng = Engine(kind, N, R, W, W, 12.0, E=E, R=Rm)
    head = LiteralHead(eng, L, aname, rng=rng)
    Wd, b = head.get_weights()
    b = (rng.normal(size=b.shape) * 0.1).astype(np.float32)
    head.set_weights(Wd, b)
    return eng, head, E, Rm, L, Wd, b


def test_sgemm_all_transposes():
    from KGEkeras._engine import _ptr, _stream
    eng, head, *_ = make(ko.DISTMULT, 10, 3, 8, 2, "tanh")
    rng = np.random.default_rng(1)
    for M, N, K in ((1, 1, 1), (65, 33, 17), (300, 100, 105), (7, 130, 64), (257, 100, 100), (40, 260, 31)):
        A, B = rng.normal(size=(M, K)).astype(np.float32), rng.normal(size=(K, N)).astype(np.float32)
        want = A.astype(np.float64) @ B.astype(np.float64)
        mag = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)       # sum_k |a_k b_k|: the parity bar is 1e-5 of it;
        for ta in (0, 1):                                                      # the tf32 x 3 split stays 5 times inside
            for tb in (0, 1):
                a = torch.from_numpy(np.ascontiguousarray(A.T if ta else A)).cuda()
                b = torch.from_numpy(np.ascontiguousarray(B.T if tb else B)).cuda()
                c = torch.full((M, N + 3), 7.0, device="cuda")
                head._gemm(ta, tb, M, N, K, a, a.shape[1], b, b.shape[1], c, N + 3, 0)
                assert np.all(np.abs(c[:, :N].cpu().numpy() - want) <= 2e-6 * mag + 1e-30), (M, N, K, ta, tb)
                assert float(c[:, N:].min()) == 7.0          # padding columns untouched
                head._gemm(ta, tb, M, N, K, a, a.shape[1], b, b.shape[1], c, N + 3, 1)   # accumulate
                assert np.all(np.abs(c[:, :N].cpu().numpy() - 2 * want) <= 4e-6 * mag + 1e-30), (M, N, K, ta, tb)


@pytest.mark.parametrize("M,N,K", [(105, 100, 40960), (5, 3, 1031), (64, 64, 100000), (130, 7, 5000)])
def test_sgemm_split_k(M, N, K):
    """The weight gradient X^T gY has a short M, N and the batch as K: the K axis is sliced across CTAs and the slices are added
    in slice order, so two runs agree bit for bit (no atomics) and the value matches fp64 to fp32 accumulation error."""
    eng, head, *_ = make(ko.DISTMULT, 10, 3, 8, 2, "tanh")
    rng = np.random.default_rng(M + K)
    A, B = rng.normal(size=(K, M)).astype(np.float32), rng.normal(size=(K, N)).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.full((M, N + 1), 3.0, device="cuda")
    head._gemm(1, 0, M, N, K, a, M, b, N, c, N + 1, 0)
    first = c.clone()
    tol = 4e-7 * np.sqrt(K) * 8
    np.testing.assert_allclose(c[:, :N].cpu().numpy(), want, rtol=1e-5, atol=tol * np.abs(want).max())
    assert float(c[:, N:].min()) == 3.0
    c.fill_(3.0)
    head._gemm(1, 0, M, N, K, a, M, b, N, c, N + 1, 0)
    assert torch.equal(c, first)
    head._gemm(1, 0, M, N, K, a, M, b, N, c, N + 1, 1)
    np.testing.assert_allclose(c[:, :N].cpu().numpy(), 2 * want, rtol=1e-5, atol=2 * tol * np.abs(want).max())


@pytest.mark.parametrize("kind", [ko.DISTMULT, ko.TRANSE_L1, ko.COMPLEX])
@pytest.mark.parametrize("aname", list(ACTS))
@pytest.mark.parametrize("p", [0.0, 0.25])
def test_literal_step_against_oracle(kind, aname, p):
    N, R, d, wl, B, K = 300, 5, 20, 5, 64, 3
    eng, head, E, Rm, L, Wd, b = make(kind, N, R, d, wl, aname, seed=len(aname) + kind)
    rng = np.random.default_rng(7)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    neg = ko.philox_negatives(pos, N, K, 5, 1, ko.CORRUPT_BOTH).astype(np.int32)
    res = head.forward_backward(pos, neg, K=K, loss_kind=ko.PAIRWISE_LOGISTIC, margin=0.0, seed=21, offset=4, opt_kind=ko.OPT_SGD, dropout=p,
                                want_scores=True)
    want = ko.literal_step(kind, ko.PAIRWISE_LOGISTIC, ko.PAIRED, E, Rm, L, Wd, b, ACTS[aname], pos, neg, K, 0.0, 12.0, p=p, seed=21, offset=4)
    assert abs(float(res["loss"].item()) - want["loss"]) <= 2e-5 * max(1.0, abs(want["loss"]))
    np.testing.assert_allclose(res["pos_scores"].cpu().numpy(), want["pos_scores"], rtol=1e-4, atol=1e-4 * np.abs(want["pos_scores"]).max())
    W_ = E.shape[1]
    # one absolute scale for all four gradients: entries that cancel analytically (TransE: d/ds' = -d/do', so the bias gradient of
    # a linear Dense is exactly zero) carry fp32 cancellation noise proportional to the step's gradient magnitude, not their own
    gscale = max(np.abs(want[k]).max() for k in ("gE", "gR", "gW", "gb"))
    for name, got, w in (("gE", eng.gE[:, :W_], want["gE"]), ("gR", eng.gR[:, :W_], want["gR"]),
                         ("gW", head.gW[:, :W_], want["gW"]), ("gb", head.gb[0, :W_], want["gb"])):
        np.testing.assert_allclose(got.cpu().numpy(), w, rtol=2e-4, atol=2e-5 * gscale, err_msg=name)
    touched = np.zeros(N, bool)
    touched[np.concatenate([pos[:, [0, 2]].reshape(-1), neg[:, [0, 2]].reshape(-1)])] = True
    np.testing.assert_array_equal(eng.tE.cpu().numpy().astype(bool), touched)
    # one SGD step moves E, R, W, b by -lr * gradient and hands every gradient buffer back clean
    head.apply_optimizer(ko.OPT_SGD, lr=0.1)
    np.testing.assert_allclose(eng.get_tables()[0], E - 0.1 * want["gE"], rtol=1e-4, atol=1e-5)
    Wn, bn = head.get_weights()
    np.testing.assert_allclose(Wn, Wd - 0.1 * want["gW"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bn, b - 0.1 * want["gb"], rtol=1e-4, atol=1e-5)
    assert float(head.gW.abs().max()) == 0.0 and float(head.gb.abs().max()) == 0.0 and float(eng.gE.abs().max()) == 0.0


def test_literal_inference_table_and_ranking():
    eng, head, E, Rm, L, Wd, b = make(ko.DISTMULT, 200, 4, 16, 5, "tanh", seed=3)
    Ep = ko.literal_table(E, L, Wd, b, ko.ACT_TANH)
    eff = head.effective_engine()
    np.testing.assert_allclose(eff.get_tables()[0], Ep, rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(2)
    tri = np.stack([rng.integers(0, 200, 50), rng.integers(0, 4, 50), rng.integers(0, 200, 50)], 1).astype(np.int32)
    got = head.score(tri).cpu().numpy()
    want = ko.score_triples(ko.DISTMULT, Ep, Rm, tri, 12.0, np.float64)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # training-mode scores draw masks: they equal the oracle's with the same (p, seed, offset)
    sc = head.score(tri, dropout=(0.5, 3, 9)).cpu().numpy()
    m = len(tri)
    mX = ko.dropout_mult(ko.DISTMULT, 16 + 5, 2 * m, 0, 0, 0.5, 3, 9)
    mP = ko.dropout_mult(ko.DISTMULT, 16, m, 0, 1, 0.5, 3, 9)
    refs = np.concatenate([tri[:, 0], tri[:, 2]])
    Y = np.tanh((np.concatenate([E[refs], L[refs]], 1) * mX) @ Wd.astype(np.float64) + b)
    np.testing.assert_allclose(sc, (Y[:m] * (Rm[tri[:, 1]] * mP) * Y[m:]).sum(-1), rtol=1e-4, atol=1e-5)


def test_harness_configuration_trains_and_validates():
    """tests.py:113-134 on bundled kinship: DistMult, 5 random literals, tanh, all-pairs hinge, Adam, the harness' generator."""
    import random
    import os
    import KGEkeras
    from KGEkeras import utils as ku
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "datasets", "kinship.npz"))
    N, R, train, valid, test = int(d["num_entities"]), int(d["num_relations"]), d["train"], d["valid"], d["test"][:200]
    np.random.seed(0); random.seed(0)
    literals = np.random.rand(N, 5)
    m = KGEkeras.DistMult(e_dim=32, r_dim=32, num_entities=N, num_relations=R, literals=literals, literal_activation="tanh",
                          loss_function="pairwize_hinge", margin=1, negative_samples=4, dp=0.2)
    m.pair_mode = "allpairs"
    m.compile(optimizer=type("Adam", (), {"learning_rate": 0.02})())
    filt = np.concatenate([train, valid])
    before = ku.validate(m, test, N, 32, filt)
    hist = m.fit(KGEkeras.DataGenerator(train, ns=4, batch_size=512, engine=m._engine_()), epochs=8, verbose=0)
    after = ku.validate(m, test, N, 32, filt)
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    assert after["mrr"] > before["mrr"] * 1.5, (before["mrr"], after["mrr"])
    assert len(m.get_weights()) == 4 and m.get_weights()[2].shape == (32 + 5, 32)
    assert m.predict(test).shape == (200,)
