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
