"""GPU parity tests: the C-ABI kernels (through KGEkeras._engine, i.e. ctypes -> libkgeb200.so) against the CPU oracle.

Bars (stated here, as the task demands):
  * integer / index work (Philox negatives, filtered rank counts): bit-exact;
  * scores on dyadic tables (every product and sum exact in fp32): bit-exact, any summation order;
  * fp32 scores / losses / gradients on real-valued tables: |gpu - fp64 truth| <= 1e-5 * (sum of |terms|), i.e. 1e-5
    relative to the magnitude the rounding error scales with (a plain relative bound is meaningless for sums that cancel);
  * rank kernel scores on real-valued tables: bit-exact against the sequential-k fp32 emulation in the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

KINDS = [ko.DISTMULT, ko.TRANSE_L1, ko.TRANSE_L2, ko.COMPLEX, ko.HOLE]
KNAME = {ko.DISTMULT: "DistMult", ko.TRANSE_L1: "TransE_L1", ko.TRANSE_L2: "TransE_L2", ko.COMPLEX: "ComplEx", ko.HOLE: "HolE"}
RTOL = 1e-5


def width(kind, d):
    return 2 * d if kind == ko.COMPLEX else d


def make_engine(kind, N, R, d, seed=0, dyadic=False, gamma=12.0, scale=1.0):
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(seed)
    W = width(kind, d)
    if dyadic:
        E, Rm = ko.dyadic_table(rng, N, W), ko.dyadic_table(rng, R, W)
    else:
        E, Rm = ko.glorot_uniform(rng, N, W) * scale, ko.glorot_uniform(rng, R, W) * scale
    return Engine(kind, N, R, W, W, gamma, E=E, R=Rm), E, Rm


def rand_triples(rng, n, N, R):
    return np.stack([rng.integers(0, N, n), rng.integers(0, R, n), rng.integers(0, N, n)], axis=1).astype(np.int32)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("d", [4, 100, 200, 256, 320])
def test_score_fwd_dyadic_exact(kind, d):
    if kind == ko.HOLE and d > 256:
        pytest.skip("HolE O(d^2) oracle is slow")
    eng, E, Rm = make_engine(kind, 300, 11, d, seed=d, dyadic=True)
    rng = np.random.default_rng(1)
    idx = rand_triples(rng, 777, 300, 11)
    idx[:5, 2] = idx[:5, 0]  # self loops
    got = eng.score(idx).cpu().numpy()
    # fp32 op-by-op oracle: on dyadic tables every intermediate is exact, so any summation order gives these bits
    want = ko.score_triples(kind, E, Rm, idx, 12.0, np.float32, hole_fft=False)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("d", [8, 100, 200])
def test_score_fwd_real(kind, d):
    eng, E, Rm = make_engine(kind, 500, 7, d, seed=3 + d, scale=3.0)
    rng = np.random.default_rng(2)
    idx = rand_triples(rng, 1000, 500, 7)
    got = eng.score(idx).cpu().numpy().astype(np.float64)
    want = ko.score_triples(kind, E, Rm, idx, 12.0, np.float64, hole_fft=False)
    bound = RTOL * ko.score_abs_terms(kind, E, Rm, idx, 12.0)
    assert np.all(np.abs(got - want) <= bound), float(np.max(np.abs(got - want) / bound))
    # and the fp32 op-by-op restatement of the reference sits inside the same band
    ref32 = ko.score_triples(kind, E, Rm, idx, 12.0, np.float32).astype(np.float64)
    assert np.all(np.abs(ref32 - want) <= bound)


def test_score_empty_and_errors():
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    eng, _, _ = make_engine(ko.DISTMULT, 10, 3, 8)
    assert eng.score(np.zeros((0, 3), np.int32)).numel() == 0
    with pytest.raises(ValueError):
        eng.score(np.zeros((4, 2), np.int32))
    with pytest.raises(abi.KgebError):  # ComplEx needs dim % 4 == 0: reported, never a silent slow path
        Engine(ko.COMPLEX, 10, 3, 12, 12).score(np.zeros((1, 3), np.int32))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("d", [8, 100])
def test_score_bwd(kind, d):
    N, R, M = 200, 5, 900
    eng, E, Rm = make_engine(kind, N, R, d, seed=11, scale=3.0)
    rng = np.random.default_rng(5)
    idx = rand_triples(rng, M, N, R)
    idx[:4, 2] = idx[:4, 0]
    dscore = rng.normal(size=M).astype(np.float32)
    dscore[10:20] = 0.0
    eng.score_bwd(idx, torch.from_numpy(dscore))
    gE, gR, tE, tR = ko.dense_grads(kind, E, Rm, idx, dscore, 12.0)
    got_gE = eng.gE.cpu().numpy()[:, :E.shape[1]].astype(np.float64)
    got_gR = eng.gR.cpu().numpy()[:, :Rm.shape[1]].astype(np.float64)
    # the bar of the scores applied to every gradient entry: 1e-5 of the summed magnitudes of its terms (one term per row reference)
    aE, aR = ko.dense_grads_abs(kind, E, Rm, idx, dscore, 12.0)
    assert np.all(np.abs(got_gE - gE) <= 1e-5 * aE + 1e-30), float(np.max(np.abs(got_gE - gE) / (aE + 1e-30)))
    assert np.all(np.abs(got_gR - gR) <= 1e-5 * aR + 1e-30), float(np.max(np.abs(got_gR - gR) / (aR + 1e-30)))
    np.testing.assert_array_equal(eng.tE.cpu().numpy().astype(bool), tE)
    np.testing.assert_array_equal(eng.tR.cpu().numpy().astype(bool), tR)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [ko.CORRUPT_BOTH, ko.CORRUPT_HEAD, ko.CORRUPT_TAIL, ko.CORRUPT_EITHER])
def test_philox_negatives_bit_exact(mode):
    N = 40943
    eng, _, _ = make_engine(ko.DISTMULT, N, 11, 4)
    rng = np.random.default_rng(9)
    pos = rand_triples(rng, 1000, N, 11)
    for seed, offset, K in [(0, 0, 1), (12345678901234567, 3, 10), (2 ** 63 + 5, 2 ** 40 + 1, 7)]:
        got = eng.negatives(pos, K, seed, offset, mode).cpu().numpy()
        want = ko.philox_negatives(pos, N, K, seed, offset, mode)
        np.testing.assert_array_equal(got, want.astype(np.int32))


@pytest.mark.parametrize("lname", list(ko.LOSS_NAMES))
@pytest.mark.parametrize("pm", [ko.PAIRED, ko.ALLPAIRS])
def test_loss_fwd_bwd(lname, pm):
    lk = ko.LOSS_NAMES[lname]
    eng, _, _ = make_engine(ko.DISTMULT, 10, 3, 4)
    rng = np.random.default_rng(lk * 2 + pm)
    B, K = 300, 7
    t = rng.normal(size=B).astype(np.float32) * 2
    f = rng.normal(size=B * K).astype(np.float32) * 2
    if lk == ko.POINTWISE_BCE:   # half of the scores inside the clip interval (0, 1), the rest clipped (zero gradient)
        t[::2], f[::2] = rng.uniform(0.01, 0.99, len(t[::2])), rng.uniform(0.01, 0.99, len(f[::2]))
    loss, dt, df = eng.loss(torch.from_numpy(t), torch.from_numpy(f), K, lk, pm, 1.0, 1.0)
    # the clip bound 1 - 1e-7 is itself rounded in fp32 (0.99999988), in TF as here: the BCE restatement runs in fp32
    odt = np.float32 if lk == ko.POINTWISE_BCE else np.float64
    L, wdt, wdf = ko.loss_and_coefs(lk, pm, t.astype(odt), f.astype(odt), K, 1.0)
    assert abs(float(loss.item()) - L) <= 2e-6 * max(1.0, abs(L))
    np.testing.assert_allclose(dt.cpu().numpy(), wdt, rtol=2e-5, atol=2e-7 * np.max(np.abs(wdt)))
    np.testing.assert_allclose(df.cpu().numpy(), wdf, rtol=2e-5, atol=2e-7 * np.max(np.abs(wdf)))
    # the reference function itself (restated) gives the same value
    ref = ko.loss_function_lookup(lname)
    if pm == ko.PAIRED:
        Lref = ref(t.reshape(-1, 1) if ko.PAIRWISE_HINGE <= lk <= ko.PAIRWISE_SQUARE else t, f, 1.0, K)
    else:
        Lref = ref(t, f, 1.0, 1)
    assert abs(float(Lref) - L) <= (1e-2 if lk == ko.POINTWISE_BCE else 1e-5) * max(1.0, abs(L))   # BCE: fp64 vs fp32 clip bound


# ------------------------------------------------------------------------------------------------------------------
def _state(kind_opt, E, Rm):
    if kind_opt == ko.OPT_SGD:
        return [], []
    if kind_opt == ko.OPT_ADAGRAD:
        return [np.full_like(E, 0.1)], [np.full_like(Rm, 0.1)]
    return [np.zeros_like(E), np.zeros_like(E)], [np.zeros_like(Rm), np.zeros_like(Rm)]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("lname,pm", [("pairwize_hinge", ko.PAIRED), ("pairwize_logistic", ko.PAIRED),
                                      ("pointwize_logistic", ko.PAIRED), ("pointwize_hinge", ko.PAIRED),
                                      ("pairwize_hinge", ko.ALLPAIRS), ("pairwize_square_loss", ko.PAIRED),
                                      ("pointwize_square_loss", ko.PAIRED), ("pointwize_cross_entropy", ko.PAIRED)])
@pytest.mark.parametrize("opt", [ko.OPT_LAZY_ADAM, ko.OPT_ADAM_DENSE, ko.OPT_ADAGRAD, ko.OPT_SGD])
def test_train_step_index_driven(kind, lname, pm, opt):
    if opt in (ko.OPT_ADAGRAD, ko.OPT_SGD) and lname not in ("pairwize_hinge",):
        pytest.skip("optimizer variants are exercised on one loss")
    lk = ko.LOSS_NAMES[lname]
    N, R, d, B, K = 120, 6, 16, 64, 5
    eng, E, Rm = make_engine(kind, N, R, d, seed=21, scale=4.0, gamma=2.0)
    rng = np.random.default_rng(31)
    pos = rand_triples(rng, B, N, R)
    neg = ko.philox_negatives(pos, N, K, 7, 0, ko.CORRUPT_EITHER).astype(np.int32)
    neg[::3, 1] = (neg[::3, 1] + 1) % R  # index-driven negatives may carry any relation
    lr = 0.05
    sE, sR = _state(opt, E, Rm)
    # eps = 1e-3 keeps Adam's m/(sqrt(v)+eps) well conditioned for gradients that cancel to ~0 (the fp32 atomics'
    # summation order is free); the default eps = 1e-7 is covered with exact inputs in test_opt_update_exact_inputs
    # (the BCE clip bound 1 - 1e-7 only exists rounded in fp32, so its restatement runs in fp32 like the reference would)
    want = ko.train_step(kind, lk, pm, E, Rm, pos, neg, K, 1.0, 2.0, opt, sE, sR, lr, 1, np.float32 if lk == ko.POINTWISE_BCE else np.float64, eps=1e-3)
    res = eng.forward_backward(pos, neg, K=K, loss_kind=lk, pair_mode=pm, margin=1.0, want_scores=True, opt_kind=opt)
    W = E.shape[1]
    gE = eng.gE.cpu().numpy()[:, :W].astype(np.float64)
    gR = eng.gR.cpu().numpy()[:, :W].astype(np.float64)
    L = float(res["loss"].item())
    assert abs(L - want["loss"]) <= 1e-5 * max(1.0, abs(want["loss"])), (L, want["loss"])
    tol_s = RTOL * ko.score_abs_terms(kind, E, Rm, pos, 2.0)
    assert np.all(np.abs(res["pos_scores"].cpu().numpy() - want["pos_scores"]) <= tol_s)
    # gradients: 1e-5 of the summed magnitudes of each entry's terms (dL/dscore from the oracle's fp64 pass)
    idx_all = np.concatenate([pos, neg])
    aE, aR = ko.dense_grads_abs(kind, E, Rm, idx_all, np.concatenate([np.ravel(want["dpos"]), np.ravel(want["dneg"])]), 2.0)
    # (BCE: `want` is itself an fp32 pass -- the clip bound only exists rounded -- and dL/dscore = 1/p, 1/(1-p) amplifies the last-bit
    # differences of two fp32 passes near the clip, so two fp32 results are compared at 1e-4 instead of one against the truth at 1e-5)
    gtol = 1e-4 if lk == ko.POINTWISE_BCE else 1e-5
    assert np.all(np.abs(gE - want["gE"]) <= gtol * aE + 1e-30), float(np.max(np.abs(gE - want["gE"]) / (aE + 1e-30)))
    assert np.all(np.abs(gR - want["gR"]) <= gtol * aR + 1e-30), float(np.max(np.abs(gR - want["gR"]) / (aR + 1e-30)))
    eng.apply_optimizer(opt, lr=lr, eps=1e-3)
    E1, R1 = eng.get_tables()
    np.testing.assert_allclose(E1, want["E"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(R1, want["R"], rtol=1e-4, atol=2e-5)
    # gradient buffers and flags are handed back clean
    assert float(eng.gE.abs().max()) == 0.0 and float(eng.gR.abs().max()) == 0.0
    assert int(eng.tE.sum()) == 0 and int(eng.tR.sum()) == 0


@pytest.mark.parametrize("kind", [ko.DISTMULT, ko.TRANSE_L1, ko.COMPLEX])
@pytest.mark.parametrize("mode", [ko.CORRUPT_BOTH, ko.CORRUPT_EITHER])
def test_train_step_philox_equals_index_driven(kind, mode):
    """In-kernel Philox negatives == the host restatement, and the fused pass on them == the unfused sequence."""
    N, R, d, B, K = 500, 4, 100, 256, 10
    eng, E, Rm = make_engine(kind, N, R, d, seed=5, scale=3.0)
    rng = np.random.default_rng(6)
    pos = rand_triples(rng, B, N, R)
    lk = ko.PAIRWISE_HINGE
    res = eng.forward_backward(pos, None, K=K, loss_kind=lk, margin=1.0, corrupt_mode=mode, seed=99, offset=4,
                               want_scores=True, want_neg=True)
    neg_want = ko.philox_negatives(pos, N, K, 99, 4, mode).astype(np.int32)
    np.testing.assert_array_equal(res["neg"].cpu().numpy(), neg_want)
    g1 = eng.gE.clone(), eng.gR.clone()
    L1 = float(res["loss"].item())
    eng.gE.zero_(); eng.gR.zero_(); eng.tE.zero_(); eng.tR.zero_()
    res2 = eng.forward_backward(pos, neg_want, K=K, loss_kind=lk, margin=1.0, force_unfused=True)
    assert abs(L1 - float(res2["loss"].item())) <= 1e-6 * max(1.0, abs(L1))
    gs = float(g1[0].abs().max())
    assert float((g1[0] - eng.gE).abs().max()) <= 1e-5 * gs
    assert float((g1[1] - eng.gR).abs().max()) <= 1e-5 * max(gs, float(g1[1].abs().max()))


def test_multi_step_training_reduces_loss_and_tracks_oracle():
    """20 LazyAdam steps on a fixed batch: loss decreases and the table stays within fp32 drift of the fp64 oracle."""
    kind, N, R, d, B, K = ko.DISTMULT, 80, 4, 16, 128, 4
    eng, E, Rm = make_engine(kind, N, R, d, seed=8, scale=2.0)
    rng = np.random.default_rng(8)
    pos = rand_triples(rng, B, N, R)
    neg = ko.philox_negatives(pos, N, K, 1, 0, ko.CORRUPT_BOTH).astype(np.int32)
    E64, R64 = E.astype(np.float64), Rm.astype(np.float64)
    sE, sR = _state(ko.OPT_LAZY_ADAM, E64, R64)
    losses = []
    for step in range(1, 21):
        res = eng.train_step(pos, neg, opt_kind=ko.OPT_LAZY_ADAM, lr=0.01, eps=1e-4, K=K, loss_kind=ko.PAIRWISE_LOGISTIC, margin=0.0)
        losses.append(float(res["loss"].item()))
        w = ko.train_step(kind, ko.PAIRWISE_LOGISTIC, ko.PAIRED, E64, R64, pos, neg, K, 0.0, 12.0, ko.OPT_LAZY_ADAM,
                          sE, sR, 0.01, step, np.float64, eps=1e-4)
        assert abs(losses[-1] - w["loss"]) <= 1e-4 * abs(w["loss"])
        E64, R64, sE, sR = w["E"], w["R"], w["state_E"], w["state_R"]
    assert losses[-1] < losses[0]
    E1, R1 = eng.get_tables()
    np.testing.assert_allclose(E1, E64, rtol=1e-3, atol=1e-4)


# ------------------------------------------------------------------------------------------------------------------
def _rank_case(kind, N, R, d, Q, seed, dyadic):
    eng, E, Rm = make_engine(kind, N, R, d, seed=seed, dyadic=dyadic, scale=3.0)
    rng = np.random.default_rng(seed + 1)
    tri = rand_triples(rng, Q, N, R)
    known = rand_triples(rng, 4 * Q, N, R)
    known[:Q, :2] = tri[:, :2]   # some known-true triples share (s,p) / (p,o) with the queries
    known[Q:2 * Q, 1:] = tri[:, 1:]
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in known])
    return eng, E, Rm, tri, fh, ft


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("d,N,Q", [(8, 130, 70), (100, 1000, 257), (36, 64, 64)])
def test_rank_counts_bit_exact_real_valued(kind, side, d, N, Q):
    eng, E, Rm, tri, fh, ft = _rank_case(kind, N, 5, d, Q, seed=17 + d, dyadic=False)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    sd = 0 if side == "tail" else 1
    counts = eng.rank_counts(tri, sd, offs, ids).cpu().numpy()
    qv = ko.query_vectors(kind, E, Rm, tri, side)
    sc = ko.seqk_scores(kind, qv, E, side, 12.0)
    want = ko.ranks_from_scores(sc, tri, side, filt) - 1
    np.testing.assert_array_equal(counts, want.astype(np.int32))
    # near-tie-aware agreement with the fp64 truth: ranks may differ only where fp64 scores are within the fp32 band
    sc64 = ko.all_entity_scores(kind, E, Rm, tri, side, 12.0, np.float64)
    r64 = ko.ranks_from_scores(sc64, tri, side, filt)
    diff = np.nonzero(r64 - 1 != counts)[0]
    for qi in diff:
        tgt = tri[qi, 2] if side == "tail" else tri[qi, 0]
        gap = np.sort(np.abs(sc64[qi] - sc64[qi, tgt]))[1:1 + abs(int(r64[qi] - 1 - counts[qi]))]
        assert np.all(gap <= 1e-4 * max(1.0, np.max(np.abs(sc64[qi])))), (qi, gap)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("side", ["tail", "head"])
def test_rank_counts_dyadic_exact_with_ties(kind, side):
    eng, E, Rm, tri, fh, ft = _rank_case(kind, 700, 5, 12, 200, seed=3, dyadic=True)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    counts = eng.rank_counts(tri, 0 if side == "tail" else 1, offs, ids).cpu().numpy()
    sc = ko.all_entity_scores(kind, E, Rm, tri, side, 12.0, np.float32)   # exact on dyadic tables
    want = ko.ranks_from_scores(sc, tri, side, filt) - 1
    np.testing.assert_array_equal(counts, want.astype(np.int32))


@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("gamma", [12.0, 0.0])
def test_rank_tile_flat_work_list(side, gamma):
    """The packed TransE-L1 tile kernels cut ONE list of (query tile, entity tile) items into equal ranges over the CTAs (148 for the
    TMA-fed kernel, <= 296 for the register-staged one): with 11 x 101 items every CTA owns several, ranges start and end inside a query
    tile and cross into the next one (counts flushed, target scores and the resident query tile reloaded, the stage ring running across
    the boundary, odd and even numbers of items per run for the two consumer groups); ragged last tiles on both axes (TMA zero fill) and
    a ragged last k-chunk (52 = 32 + 20).  Integer counts equal the oracle's sequential-k fp32 ranks."""
    kind, N, R, d, Q = ko.TRANSE_L1, 6400 + 17, 5, 52, 640 + 5
    eng, E, Rm = make_engine(kind, N, R, d, seed=77, gamma=gamma)
    rng = np.random.default_rng(78)
    tri = rand_triples(rng, Q, N, R)
    sd = 0 if side == "tail" else 1
    counts = eng.rank_counts(tri, sd).cpu().numpy()
    qv = ko.query_vectors(kind, E, Rm, tri, side)
    sc = ko.seqk_scores(kind, qv, E, side, gamma)
    tgt = tri[:, 2] if side == "tail" else tri[:, 0]
    want = (sc > sc[np.arange(Q), tgt][:, None]).sum(axis=1)
    np.testing.assert_array_equal(counts, want.astype(np.int32))
    # any row partition adds up to the same integers (each shard is its own flat list)
    acc = np.zeros_like(counts)
    for r0, r1 in ((0, 1000), (1000, 1063), (1063, N)):
        acc += eng.rank_counts(tri, sd, row_range=(r0, r1)).cpu().numpy()
    np.testing.assert_array_equal(acc, counts)


@pytest.mark.parametrize("side", ["tail", "head"])
def test_rank_tile_wide_rows(side):
    """d = 320: the tail side's resident query tile still fits the TMA-fed kernel (k_rank_tile3), the head side's two-part tile does not
    and runs on k_rank_tile2 -- both give the oracle's integers."""
    kind, N, R, d, Q = ko.TRANSE_L1, 1000 + 3, 4, 320, 130
    eng, E, Rm = make_engine(kind, N, R, d, seed=79)
    tri = rand_triples(np.random.default_rng(80), Q, N, R)
    counts = eng.rank_counts(tri, 0 if side == "tail" else 1).cpu().numpy()
    sc = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, tri, side), E, side, 12.0)
    tgt = tri[:, 2] if side == "tail" else tri[:, 0]
    np.testing.assert_array_equal(counts, (sc > sc[np.arange(Q), tgt][:, None]).sum(axis=1).astype(np.int32))


@pytest.mark.parametrize("variant", ["scalar", "packed2"])
def test_tile_kernel_variants_in_child_process(variant):
    """KGEB_RANK_TILE (read once per process) selects the round-1 scalar tile kernel ("scalar") or the register-staged packed kernel
    ("packed2") instead of the TMA-fed one for A/B runs: the TransE-L1 rank tests pass on each of them too."""
    import os
    import subprocess
    import sys
    if os.environ.get("KGEB_RANK_TILE"):
        pytest.skip("already inside the child")
    env = dict(os.environ, KGEB_RANK_TILE=variant)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k",
                        "rank_tile_flat_work_list or rank_tile_wide or rank_shards_sum or (rank_counts and not HOLE)"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


@pytest.mark.parametrize("opt", [ko.OPT_SGD, ko.OPT_ADAGRAD, ko.OPT_LAZY_ADAM, ko.OPT_ADAM_DENSE])
def test_opt_update_exact_inputs(opt):
    """Optimizer kernel alone, fed the same fp32 gradient as the oracle, Keras default eps = 1e-7, 3 steps."""
    N, W = 333, 24
    eng, E, Rm = make_engine(ko.DISTMULT, N, 4, W, seed=2)
    eng._ensure_train_buffers(opt)
    rng = np.random.default_rng(12)
    P = E.astype(np.float64)
    st = [x.astype(np.float64) for x in _state(opt, E, Rm)[0]]
    for step in range(1, 4):
        touched = rng.random(N) < 0.3
        G = (rng.normal(size=(N, W)) * touched[:, None]).astype(np.float32)
        eng.gE.copy_(torch.from_numpy(G))
        eng.tE.copy_(torch.from_numpy(touched.astype(np.uint8)))
        eng.apply_optimizer(opt, lr=0.01)
        P, st = ko.opt_update(opt, P, G.astype(np.float64), touched, st, 0.01, step)
        np.testing.assert_allclose(eng.E.cpu().numpy(), P, rtol=2e-5, atol=1e-6)
        assert float(eng.gE.abs().max()) == 0.0 and int(eng.tE.sum()) == 0


def test_rank_shards_sum_to_whole():
    """Integer partial counts over any row partition add up to the single-shard result (the multi-GPU contract)."""
    kind = ko.TRANSE_L1
    eng, E, Rm, tri, fh, ft = _rank_case(kind, 1000, 5, 20, 150, seed=4, dyadic=False)
    offs, ids = ko.filter_csr(tri, "tail", ft)
    whole = eng.rank_counts(tri, 0, offs, ids).cpu().numpy()
    for G in (2, 3, 8):
        per = -(-1000 // G)
        acc = np.zeros_like(whole)
        for g in range(G):
            r0, r1 = g * per, min(1000, (g + 1) * per)
            acc += eng.rank_counts(tri, 0, offs, ids, row_range=(r0, r1)).cpu().numpy()
        np.testing.assert_array_equal(acc, whole)


def test_l3_regularization_step():
    """l3_reg (models.py:10-11): loss += w*sum|E|^3, dense gradient 3 w E|E| on every entity row, all rows updated."""
    kind, N, R, d, B, K, w = ko.DISTMULT, 90, 4, 16, 32, 3, 0.01
    eng, E, Rm = make_engine(kind, N, R, d, seed=13, scale=3.0)
    rng = np.random.default_rng(14)
    pos = rand_triples(rng, B, N, R)
    neg = ko.philox_negatives(pos, N, K, 3, 0, ko.CORRUPT_BOTH).astype(np.int32)
    res = eng.forward_backward(pos, neg, K=K, loss_kind=ko.PAIRWISE_HINGE, margin=1.0, opt_kind=ko.OPT_SGD)
    eng.l3_reg(w, res["loss"])
    want = ko.train_step(kind, ko.PAIRWISE_HINGE, ko.PAIRED, E, Rm, pos, neg, K, 1.0, 12.0, ko.OPT_SGD, [], [], 0.1, 1)
    rl, rg = ko.l3_reg(E.astype(np.float64), w)
    assert abs(float(res["loss"].item()) - (want["loss"] + rl)) <= 1e-5 * (want["loss"] + rl)
    g = eng.gE.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(g, want["gE"] + rg, rtol=1e-4, atol=1e-6 * np.abs(want["gE"] + rg).max())
    assert int(eng.tE.sum()) == N
    eng.apply_optimizer(ko.OPT_SGD, lr=0.1)
    np.testing.assert_allclose(eng.get_tables()[0], E - 0.1 * (want["gE"] + rg), rtol=1e-4, atol=1e-6)


def test_device_batch_pipeline_bit_exact():
    """Device-side DataGenerator.__get_data (tests.py:55-61): Philox negatives + Philox oversampling against the host restatement."""
    eng, E, Rm = make_engine(ko.DISTMULT, 500, 7, 8, seed=1)
    rng = np.random.default_rng(3)
    pos = rand_triples(rng, 100, 500, 7)
    K, seed, step = 10, 99, 12
    dpos, dneg = eng.device_batch(pos, K, seed, step)
    neg = ko.philox_negatives(pos, 500, K, seed, step, ko.CORRUPT_BOTH).astype(np.int32)
    m = len(neg)
    np.testing.assert_array_equal(dpos.cpu().numpy(), pos[ko.philox_oversample_indices(len(pos), m, seed, step, 0)])
    np.testing.assert_array_equal(dneg.cpu().numpy(), neg[ko.philox_oversample_indices(m, m, seed, step, 1)])
    idx = ko.philox_oversample_indices(100, 200000, 5, 0, 0)
    assert idx.min() == 0 and idx.max() == 99 and abs(np.bincount(idx, minlength=100).std() / 2000 - np.sqrt(0.99 / 2000)) < 0.01


@pytest.mark.parametrize("B,Bf", [(20480, 20480), (1000, 3333), (1, 5), (1025, 1024), (7, 1)])
def test_allpairs_hinge_sorted_path_equals_dense_walk(B, Bf, monkeypatch):
    """The harness' (m,m) hinge (tests.py:81): the chunk-sort + binary-search kernels give the same coefficient vectors as the
    m x m walk -- integer counts times one scale, where the walk adds per-slice partial products in fp32 -- and the counts
    are exactly the fp64 restatement's; ties and scores sitting exactly on the margin included."""
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(B + Bf)
    t = (rng.normal(size=B) * 2).astype(np.float32)
    f = (rng.normal(size=Bf) * 2).astype(np.float32)
    if B > 4 and Bf > 4:
        f[:3] = t[:3] - 1.0          # margin + f - t == 0 exactly: not counted (strict >)
        t[3:5] = t[0]                # ties
    monkeypatch.setenv("KGEB_ALLPAIRS", "dense")
    dense = Engine(ko.DISTMULT, 4, 2, 4, 4)
    monkeypatch.delenv("KGEB_ALLPAIRS")
    fast = Engine(ko.DISTMULT, 4, 2, 4, 4)
    tt, ff = torch.from_numpy(t), torch.from_numpy(f)
    L0, dt0, df0 = dense.loss(tt, ff, 1, ko.PAIRWISE_HINGE, ko.ALLPAIRS, 1.0, 0.7)
    L1, dt1, df1 = fast.loss(tt, ff, 1, ko.PAIRWISE_HINGE, ko.ALLPAIRS, 1.0, 0.7)
    torch.testing.assert_close(dt0, dt1, rtol=2e-6, atol=0)
    torch.testing.assert_close(df0, df1, rtol=2e-6, atol=0)
    Lw, wdt, wdf = ko.loss_and_coefs(ko.PAIRWISE_HINGE, ko.ALLPAIRS, t.astype(np.float64), f.astype(np.float64), 1, 1.0)
    cnt = float(B) * float(Bf)
    # (the restatement decides margin + f - t > 0 in fp64; on fp32 inputs that differs from the fp32 predicate only at exact
    # cancellations, which the constructed ties cover: they are excluded by both)
    cf = np.rint(df1.cpu().numpy().astype(np.float64) * cnt / 0.7)
    ct = np.rint(dt1.cpu().numpy().astype(np.float64) * cnt / 0.7)
    if B * Bf <= 4_000_000:   # the fp32 predicate itself, all pairs: the integer counts are exact
        pred = (np.float32(1.0) + (f[:, None] - t[None, :])) > 0
        np.testing.assert_array_equal(cf, pred.sum(1))
        np.testing.assert_array_equal(ct, -pred.sum(0))
    else:                     # against fp64: a pair within one fp32 rounding of the margin may fall on either side
        assert np.max(np.abs(cf - np.rint(wdf * cnt))) <= 1 and np.mean(cf != np.rint(wdf * cnt)) < 1e-3
        assert np.max(np.abs(ct - np.rint(wdt * cnt))) <= 1 and np.mean(ct != np.rint(wdt * cnt)) < 1e-3
    assert abs(float(L1.item()) - 0.7 * Lw) <= 2e-6 * max(1.0, abs(Lw))
    assert abs(float(L0.item()) - float(L1.item())) <= 2e-6 * max(1.0, abs(Lw))
