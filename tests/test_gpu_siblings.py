"""GPU parity of the sibling score functions: RotatE models.py:421-459, pRotatE 461-489, ModE 403-418,
through the C ABI (score_fwd / score_bwd / rank_1n / rank_1n_tc) against the CPU oracle.

Bars: fp32 scores within 1e-5 * (sum of |terms|) of the fp64 restatement; gradients within 1e-4 relative (+1e-5 of the
largest entry); ModE and RotatE rank counts bit-exact against the oracle's sequential-k fp32 emulation (RotatE's cos/sin
query vector is taken from the device so the comparison is about the ranking arithmetic, the vector itself is checked to
1e-6); pRotatE ranks (a transcendental per term: CUDA sinf and NumPy sin may differ in the last bit) equal the fp64 ranks
except on pairs whose fp64 scores lie within 1e-4 of the target's.
"""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

SIB = [ko.ROTATE, ko.PROTATE, ko.MODE_L1, ko.MODE_L2]
MODULUS = 0.4375


def widths(kind, d):
    return (2 * d, d) if kind == ko.ROTATE else (d, d)


def make_engine(kind, N, R, d, seed=0, gamma=12.0, scale=3.0, dyadic=False):
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(seed)
    We, Wr = widths(kind, d)
    if dyadic:
        E, Rm = ko.dyadic_table(rng, N, We), ko.dyadic_table(rng, R, Wr)
    else:
        E, Rm = ko.glorot_uniform(rng, N, We) * scale, ko.glorot_uniform(rng, R, Wr) * scale
    return Engine(kind, N, R, We, Wr, gamma, E=E, R=Rm, param=MODULUS if kind == ko.PROTATE else 1.0), E, Rm


def rand_triples(rng, n, N, R):
    return np.stack([rng.integers(0, N, n), rng.integers(0, R, n), rng.integers(0, N, n)], axis=1).astype(np.int32)


def par(kind):
    return MODULUS if kind == ko.PROTATE else 1.0


@pytest.mark.parametrize("kind", SIB)
@pytest.mark.parametrize("d", [8, 100, 200])
@pytest.mark.parametrize("gamma", [12.0, 0.0])
def test_sibling_score_fwd(kind, d, gamma):
    eng, E, Rm = make_engine(kind, 500, 7, d, seed=3 + d, gamma=gamma, scale=12.0)   # scale 12: phases beyond +-pi occur
    rng = np.random.default_rng(2)
    idx = rand_triples(rng, 1000, 500, 7)
    idx[:5, 2] = idx[:5, 0]
    got = eng.score(idx).cpu().numpy().astype(np.float64)
    want = ko.score_triples(kind, E, Rm, idx, gamma, np.float64, param=par(kind))
    bound = 1e-5 * ko.score_abs_terms(kind, E, Rm, idx, gamma)
    assert np.all(np.abs(got - want) <= bound), float(np.max(np.abs(got - want) / bound))


@pytest.mark.parametrize("kind", [ko.MODE_L1, ko.MODE_L2])
def test_mode_score_dyadic_exact(kind):
    eng, E, Rm = make_engine(kind, 300, 11, 100, seed=5, dyadic=True)
    idx = rand_triples(np.random.default_rng(1), 777, 300, 11)
    got = eng.score(idx).cpu().numpy()
    want = ko.score_triples(kind, E, Rm, idx, 12.0, np.float32)
    if kind == ko.MODE_L1:
        np.testing.assert_array_equal(got, want)          # every product and sum is exact
    else:
        np.testing.assert_allclose(got, want, rtol=2e-7)  # one correctly rounded sqrt either side


@pytest.mark.parametrize("kind", SIB)
@pytest.mark.parametrize("d", [8, 100])
@pytest.mark.parametrize("gamma", [12.0, 0.0])
def test_sibling_score_bwd(kind, d, gamma):
    N, R, M = 200, 5, 900
    eng, E, Rm = make_engine(kind, N, R, d, seed=11, gamma=gamma)
    rng = np.random.default_rng(5)
    idx = rand_triples(rng, M, N, R)
    idx[:4, 2] = idx[:4, 0]
    dscore = rng.normal(size=M).astype(np.float32)
    dscore[10:20] = 0.0
    eng.score_bwd(idx, torch.from_numpy(dscore))
    gE, gR, tE, tR = ko.dense_grads(kind, E, Rm, idx, dscore, gamma, param=par(kind))
    got_gE = eng.gE.cpu().numpy()[:, :E.shape[1]].astype(np.float64)
    got_gR = eng.gR.cpu().numpy()[:, :Rm.shape[1]].astype(np.float64)
    gs = max(np.max(np.abs(gE)), np.max(np.abs(gR)))
    np.testing.assert_allclose(got_gE, gE, rtol=1e-4, atol=1e-5 * gs)
    np.testing.assert_allclose(got_gR, gR, rtol=1e-4, atol=1e-5 * gs)
    np.testing.assert_array_equal(eng.tE.cpu().numpy().astype(bool), tE)
    np.testing.assert_array_equal(eng.tR.cpu().numpy().astype(bool), tR)
    # padding columns of the device tables never receive gradient
    assert float(eng.gE[:, E.shape[1]:].abs().sum()) == 0.0


@pytest.mark.parametrize("kind", SIB)
def test_sibling_dropout_masks_match_oracle(kind):
    d, N, R, M, p, seed, off = 100, 50, 4, 64, 0.25, 77, 5
    eng, E, Rm = make_engine(kind, N, R, d, seed=4)
    idx = rand_triples(np.random.default_rng(6), M, N, R)
    got = eng.score(idx, dropout=(p, seed, off, 1)).cpu().numpy().astype(np.float64)
    We, Wr = widths(kind, d)
    ms = ko.dropout_mult(kind, We, M, 0, 1, p, seed, off)
    mp = ko.dropout_mult(kind, Wr, M, 1, 1, p, seed, off)
    mo = ko.dropout_mult(kind, We, M, 2, 1, p, seed, off)
    s, pp, o = ko.gather(E.astype(np.float64), Rm.astype(np.float64), idx)
    want = ko.score(kind, s * ms, pp * mp, o * mo, 12.0, param=par(kind))
    scale = np.abs(12.0) + np.sum(np.abs(s) + np.abs(o), axis=-1) * 2 + np.sum(np.abs(pp), axis=-1) * 2
    assert np.all(np.abs(got - want) <= 1e-5 * scale)


def _rank_case(kind, N, R, d, Q, seed, gamma=12.0):
    eng, E, Rm = make_engine(kind, N, R, d, seed=seed, gamma=gamma)
    rng = np.random.default_rng(seed + 1)
    tri = rand_triples(rng, Q, N, R)
    known = rand_triples(rng, 4 * Q, N, R)
    known[:Q, :2] = tri[:, :2]
    known[Q:2 * Q, 1:] = tri[:, 1:]
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in known])
    return eng, E, Rm, tri, fh, ft


def _device_qvec(eng, tri, side):
    from KGEkeras._engine import _ptr, _stream, as_idx
    t = as_idx(tri, eng.device)
    qw = eng.lib.kgeb_query_width(eng.ctx, side)
    q = torch.zeros(t.shape[0], qw, dtype=torch.float32, device=eng.device)
    eng._check(eng.lib.kgeb_query_prep(eng.ctx, _ptr(eng.E), 0, eng.N, _ptr(eng.Rel), _ptr(t), t.shape[0], side, _ptr(q), _stream()))
    return q.cpu().numpy()


@pytest.mark.parametrize("kind", [ko.MODE_L1, ko.MODE_L2, ko.ROTATE])
@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("d,N,Q", [(8, 130, 70), (100, 1000, 257)])
@pytest.mark.parametrize("gamma", [12.0, 0.0])
def test_sibling_rank_counts_bit_exact(kind, side, d, N, Q, gamma):
    eng, E, Rm, tri, fh, ft = _rank_case(kind, N, 5, d, Q, seed=17 + d, gamma=gamma)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    sd = 0 if side == "tail" else 1
    counts = eng.rank_counts(tri, sd, offs, ids).cpu().numpy()
    qv = ko.query_vectors(kind, E, Rm, tri, side)
    if kind == ko.ROTATE:
        dq = _device_qvec(eng, tri, sd)
        want_q = qv if gamma > 0 else -qv
        np.testing.assert_allclose(dq, want_q, rtol=0, atol=1e-6)
        sc = ko.seqk_scores(ko.DISTMULT, dq, E, side, gamma)     # the dot product the kernel ranks by, device query vector
    else:
        sc = ko.seqk_scores(kind, qv, E, side, gamma)
    want = ko.ranks_from_scores(sc, tri, side, filt) - 1
    np.testing.assert_array_equal(counts, want.astype(np.int32))
    # and the ranks agree with the model's own fp64 scores except at near-ties
    sc64 = ko.all_entity_scores(kind, E, Rm, tri, side, gamma, np.float64)
    _assert_ranks_match_up_to_near_ties(counts, sc64, tri, side, filt)


def _assert_ranks_match_up_to_near_ties(counts, sc64, tri, side, filt, tol=1e-4):
    r64 = ko.ranks_from_scores(sc64, tri, side, filt)
    for qi in np.nonzero(r64 - 1 != counts)[0]:
        tgt = tri[qi, 2] if side == "tail" else tri[qi, 0]
        gap = np.sort(np.abs(sc64[qi] - sc64[qi, tgt]))[1:1 + abs(int(r64[qi] - 1 - counts[qi]))]
        assert np.all(gap <= tol * max(1.0, np.max(np.abs(sc64[qi])))), (qi, gap)


@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("d,N,Q", [(8, 130, 70), (100, 1000, 257)])
@pytest.mark.parametrize("gamma", [12.0, 0.0])
def test_protate_rank_counts(side, d, N, Q, gamma):
    kind = ko.PROTATE
    eng, E, Rm, tri, fh, ft = _rank_case(kind, N, 5, d, Q, seed=23 + d, gamma=gamma)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    counts = eng.rank_counts(tri, 0 if side == "tail" else 1, offs, ids).cpu().numpy()
    sc64 = ko.all_entity_scores(kind, E, Rm, tri, side, gamma, np.float64, param=MODULUS)
    _assert_ranks_match_up_to_near_ties(counts, sc64, tri, side, filt)
    # the fp32 emulation agrees on all but a handful of queries (sinf vs numpy.sin last-bit differences)
    sc = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, tri, side), E, side, gamma, param=MODULUS)
    want = ko.ranks_from_scores(sc, tri, side, filt) - 1
    assert np.mean(counts == want) > 0.95


@pytest.mark.parametrize("side", ["tail", "head"])
def test_rotate_tensor_core_counts_equal_cuda_core(side):
    from KGEkeras import _abi as abi
    eng, E, Rm, tri, fh, ft = _rank_case(ko.ROTATE, 2500, 5, 100, 300, seed=41)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    sd = 0 if side == "tail" else 1
    a = eng.rank_counts(tri, sd, offs, ids).cpu().numpy()
    b = eng.rank_counts_tc(tri, sd, offs, ids, mode=abi.TC_BF16X3_EXACT).cpu().numpy()
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("cls,kw", [("RotatE", {}), ("pRotatE", {}), ("ModE", dict(norm=1)), ("ModE", dict(norm=2))])
def test_sibling_models_fit_and_validate(cls, kw):
    """The reference constructors (width multipliers kept) train through the unfused kernel sequence and rank."""
    import KGEkeras
    from KGEkeras import utils as ku
    rng = np.random.default_rng(0)
    N, R = 60, 4
    train = rand_triples(rng, 600, N, R)
    test = train[:40]
    # widths as the reference multiplies them: RotatE e*2 / r*1, pRotatE e*2 / r*2, ModE e*2 / r*3 (-> 2*e_dim == 3*r_dim)
    e_dim, r_dim = (12, 8) if cls == "ModE" else (8, 8)
    m = getattr(KGEkeras, cls)(e_dim=e_dim, r_dim=r_dim, num_entities=N, num_relations=R, negative_samples=4, batch_size=128,
                               loss_function="pairwize_hinge", dp=0.0, **kw)
    m.compile(optimizer="adam")
    m._opt.lr = 0.05
    h = m.fit(train, epochs=12, verbose=0)
    assert h.history["loss"][-1] < h.history["loss"][0]
    res = ku.validate(m, [tuple(int(x) for x in t) for t in test], N, 32, filtering_triples=[])
    E, Rm = m.get_weights()
    kind = m._kind_()
    sc_t = ko.all_entity_scores(kind, E, Rm, test, "tail", m._gamma_(), np.float64, param=m._param_())
    r_t = ko.ranks_from_scores(sc_t, test, "tail", {})
    assert abs(res["tail_mrr"] - float(np.mean(1.0 / r_t))) < 0.02
    # training-mode call applies dropout; inference call does not
    m.dp.rate = 0.5
    a, b = m.call(test, training=True).cpu().numpy(), m.call(test).cpu().numpy()
    assert not np.allclose(a, b)
