"""GPU tests of the sort-segmented entity backward (kgeb_set_grad_sortseg, csrc/sortseg_kernels.cu): the north star's K2.

Bars: the gradient buffers equal the fp64 oracle within the same 1e-5 * max|g| bar as the vector-reduction path, the two
paths agree with each other, the touched flags are identical, a full step (LazyAdam) lands on the oracle's tables -- and,
what the reduction path cannot offer, two runs of the same step give BIT-IDENTICAL entity gradients (stable sort +
one in-order writer per row)."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko
from test_gpu_kernels import make_engine, rand_triples

pytestmark = pytest.mark.gpu

FUSED = [ko.DISTMULT, ko.TRANSE_L1, ko.TRANSE_L2, ko.COMPLEX]


def _abi():
    from KGEkeras import _abi as abi
    return abi


@pytest.mark.parametrize("kind", FUSED)
@pytest.mark.parametrize("d,N,B,K", [(8, 50, 37, 3), (100, 300, 200, 10), (200, 1000, 256, 10), (128, 64, 300, 40)])
@pytest.mark.parametrize("loss", ["pairwize_hinge", "pointwize_logistic"])
def test_sortseg_gradients_equal_oracle_and_reduction_path(kind, d, N, B, K, loss):
    abi = _abi()
    lk = abi.LOSS_KINDS[loss]
    R = 5
    rng = np.random.default_rng(d + B)
    pos = rand_triples(rng, B, N, R)
    pos[:3, 2] = pos[:3, 0]                                   # self loops: both slots of a triple land in one segment
    neg = np.stack([rng.integers(0, N, B * K), np.repeat(pos[:, 1], K), rng.integers(0, N, B * K)], 1).astype(np.int32)
    neg[:K, 0] = pos[0, 0]                                    # negatives sharing the positive's subject
    out = {}
    for mode in ("red", "sortseg"):
        eng, E, Rm = make_engine(kind, N, R, d, seed=7, scale=3.0)
        eng.set_grad_mode(mode)
        res = eng.forward_backward(pos, neg, K=K, loss_kind=lk, margin=1.0, opt_kind=abi.LAZY_ADAM)
        out[mode] = (float(res["loss"].item()), eng.gE.cpu().numpy().astype(np.float64), eng.gR.cpu().numpy().astype(np.float64),
                     eng.tE.cpu().numpy().copy(), eng.tR.cpu().numpy().copy())
    want = ko.train_step(kind, ko.LOSS_NAMES[loss], ko.PAIRED, E, Rm, pos, neg, K, 1.0, 12.0, ko.OPT_LAZY_ADAM,
                         [np.zeros_like(E), np.zeros_like(E)], [np.zeros_like(Rm), np.zeros_like(Rm)], 1e-3, 1)
    gs = max(np.max(np.abs(want["gE"])), np.max(np.abs(want["gR"])), 1e-30)
    W = E.shape[1]
    for mode in ("red", "sortseg"):
        L, gE, gR, tE, tR = out[mode]
        assert abs(L - want["loss"]) <= 1e-5 * max(1.0, abs(want["loss"])), mode
        np.testing.assert_allclose(gE[:, :W], want["gE"], rtol=1e-4, atol=1e-5 * gs, err_msg=mode)
        np.testing.assert_allclose(gR[:, :W], want["gR"], rtol=1e-4, atol=1e-5 * gs, err_msg=mode)
    np.testing.assert_array_equal(out["red"][3], out["sortseg"][3])
    np.testing.assert_array_equal(out["red"][4], out["sortseg"][4])
    np.testing.assert_allclose(out["sortseg"][1], out["red"][1], rtol=1e-4, atol=1e-5 * gs)


@pytest.mark.parametrize("kind", FUSED)
@pytest.mark.parametrize("corrupt", ["both", "either"])
def test_sortseg_philox_step_is_reproducible_and_lands_on_the_oracle(kind, corrupt):
    abi = _abi()
    cm = {"both": abi.CORRUPT_BOTH, "either": abi.CORRUPT_EITHER}[corrupt]
    N, R, d, B, K = 400, 4, 64, 500, 20      # 20 000 references onto 400 rows: long segments, heavy collisions
    rng = np.random.default_rng(3)
    pos = rand_triples(rng, B, N, R)
    grads = []
    for rep in range(2):
        eng, E, Rm = make_engine(kind, N, R, d, seed=5, scale=3.0)
        eng.set_grad_mode("sortseg")
        res = eng.forward_backward(pos, None, K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, corrupt_mode=cm, seed=11, offset=2,
                                   want_neg=True, opt_kind=abi.LAZY_ADAM)
        grads.append(eng.gE.cpu().numpy().copy())
        neg = res["neg"].cpu().numpy()
    np.testing.assert_array_equal(grads[0], grads[1])          # bit-identical entity gradients, run to run
    np.testing.assert_array_equal(neg, ko.philox_negatives(pos, N, K, 11, 2, {"both": ko.CORRUPT_BOTH, "either": ko.CORRUPT_EITHER}[corrupt]).astype(np.int32))
    want = ko.train_step(kind, ko.PAIRWISE_LOGISTIC, ko.PAIRED, E, Rm, pos, neg, K, 1.0, 12.0, ko.OPT_LAZY_ADAM,
                         [np.zeros_like(E), np.zeros_like(E)], [np.zeros_like(Rm), np.zeros_like(Rm)], 1e-3, 1, eps=1e-3)
    gs = np.max(np.abs(want["gE"]))
    np.testing.assert_allclose(grads[0][:, :E.shape[1]].astype(np.float64), want["gE"], rtol=1e-4, atol=1e-5 * gs)
    eng.apply_optimizer(abi.LAZY_ADAM, lr=1e-3, eps=1e-3)
    np.testing.assert_allclose(eng.get_tables()[0], want["E"], rtol=1e-4, atol=2e-5)
    # the next step reuses the armed workspace
    eng.train_step(pos, None, K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, corrupt_mode=cm, seed=11, offset=3, opt_kind=abi.LAZY_ADAM)
    assert int(eng.tE.sum().item()) == 0 and float(eng.gE.abs().max().item()) == 0.0


def test_sortseg_refuses_dropout_and_disarms():
    abi = _abi()
    eng, E, Rm = make_engine(ko.DISTMULT, 100, 3, 16, seed=1)
    eng.set_grad_mode("sortseg")
    pos = rand_triples(np.random.default_rng(0), 40, 100, 3)
    with pytest.raises(abi.KgebError):
        eng.forward_backward(pos, None, K=4, dropout=0.2, seed=1, offset=0, opt_kind=abi.LAZY_ADAM)
    eng.set_grad_mode("red")
    eng.forward_backward(pos, None, K=4, dropout=0.2, seed=1, offset=0, opt_kind=abi.LAZY_ADAM)
    with pytest.raises(ValueError):
        eng.set_grad_mode("atomic")


# ---- the optimizer fused at segment end (kgeb_train_step_sortseg): no gradient buffer, two table copies ----------------------------
@pytest.mark.parametrize("kind_name,d", [("TRANSE_L1", 40), ("COMPLEX", 32), ("DISTMULT", 100), ("TRANSE_L2", 64)])
@pytest.mark.parametrize("opt_name", ["LAZY_ADAM", "ADAGRAD", "SGD"])
def test_fused_sortseg_step_equals_reduction_path(kind_name, d, opt_name):
    """Several steps with changing batches (rows touched by one step and not by the next exercise the copy-over of stale rows):
    tables, moments and loss equal to the vector-reduction path + kgeb_opt_update2 within fp32 summation-order noise."""
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    kind, opt = getattr(abi, kind_name), getattr(abi, opt_name)
    N, R, B, K = 3000, 6, 256, 5
    rng = np.random.default_rng(d)
    E0, R0 = ko.glorot_uniform(rng, N, d) * 3, ko.glorot_uniform(rng, R, d) * 3
    a, b = Engine(kind, N, R, d, d, 12.0, E=E0, R=R0), Engine(kind, N, R, d, d, 12.0, E=E0, R=R0)
    b.set_grad_mode("sortseg_fused")
    for step in range(5):
        lo, hi = (0, N) if step % 2 == 0 else (step * 300, step * 300 + 600)          # odd steps touch a small id window only
        pos = np.stack([rng.integers(lo, hi, B), rng.integers(0, R, B), rng.integers(lo, hi, B)], 1).astype(np.int32)
        neg = np.stack([rng.integers(lo, hi, B * K), np.repeat(pos[:, 1], K), rng.integers(lo, hi, B * K)], 1).astype(np.int32)
        kw = dict(K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, opt_kind=opt, lr=0.01)
        la = a.train_step(pos, neg, **kw)["loss"].item()
        lb = b.train_step(pos, neg, **kw)["loss"].item()
        assert abs(la - lb) <= 1e-6 * max(1.0, abs(la))
        np.testing.assert_allclose(b.get_tables()[0], a.get_tables()[0], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(b.get_tables()[1], a.get_tables()[1], rtol=2e-5, atol=2e-6)
    assert b.gE is None                                                               # no entity gradient buffer was ever made
    # both copies agree on every row the last step did not touch
    touched_last = b.tE2.bool().cpu().numpy()
    np.testing.assert_array_equal(b.E.cpu().numpy()[~touched_last], b.E2.cpu().numpy()[~touched_last])
    if opt != abi.SGD:
        np.testing.assert_allclose(b.opt_state[1].cpu().numpy(), a.opt_state[1].cpu().numpy(), rtol=2e-5, atol=1e-7)


def test_fused_sortseg_is_reproducible_and_serves_philox_negatives():
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    N, R, d, B, K = 2000, 4, 64, 512, 8
    rng = np.random.default_rng(1)
    E0, R0 = ko.glorot_uniform(rng, N, d) * 3, ko.glorot_uniform(rng, R, d) * 3
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    outs, first = [], []
    for _ in range(2):
        e = Engine(abi.COMPLEX, N, R, d, d, 12.0, E=E0, R=R0)
        e.set_grad_mode("sortseg_fused")
        for step in range(3):
            e.train_step(pos, None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01, seed=9, offset=step)
            if step == 0:
                first.append(e.get_tables()[0])
        outs.append(e.get_tables()[0])
    # entity rows of ONE step are bit-identical run to run (every sum runs in sorted order); later steps inherit the low-bit noise of
    # the relation table, whose gradient still arrives by vector reductions
    np.testing.assert_array_equal(first[0], first[1])
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-5, atol=1e-6)
    ref = Engine(abi.COMPLEX, N, R, d, d, 12.0, E=E0, R=R0)
    for step in range(3):
        ref.train_step(pos, None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01, seed=9, offset=step)
    np.testing.assert_allclose(outs[0], ref.get_tables()[0], rtol=2e-5, atol=2e-6)


def test_fused_sortseg_falls_back_for_dense_adam():
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    N, R, d, B, K = 500, 3, 16, 64, 2
    rng = np.random.default_rng(2)
    E0, R0 = ko.glorot_uniform(rng, N, d), ko.glorot_uniform(rng, R, d)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    a, b = Engine(abi.DISTMULT, N, R, d, d, 12.0, E=E0, R=R0), Engine(abi.DISTMULT, N, R, d, d, 12.0, E=E0, R=R0)
    b.set_grad_mode("sortseg_fused")
    for step in range(2):
        kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, opt_kind=abi.ADAM_DENSE, lr=0.01, seed=1, offset=step)
        a.train_step(pos, None, **kw)
        b.train_step(pos, None, **kw)
    np.testing.assert_allclose(b.get_tables()[0], a.get_tables()[0], rtol=2e-5, atol=2e-6)
    assert torch.equal(b.E, b.E2)
