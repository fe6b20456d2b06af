"""GPU parity tests of the tcgen05 1-N ranking path (kgeb_rank_1n_tc) for the dot-product scores.

Bars:
  * KGEB_TC_BF16X3_EXACT: integer filtered-rank counts IDENTICAL to the fp32 CUDA-core kernel (kgeb_rank_1n), which is
    itself bit-exact against the oracle's sequential-k fp32 scores (tests/test_gpu_kernels.py) -- checked against both;
  * the error bound the exactness rests on is checked directly: the share of undecided pairs must stay small (the
    tensor-core scores are far inside the proven band) and dyadic tables (every product and sum exact) must need no
    re-scoring beyond genuine ties;
  * KGEB_TC_BF16 (approximate): ranks may move only where the fp64 score gap to the target is below the bf16 band
    2^-7 * ||q|| * ||e|| (stated tolerance).
"""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko
from test_gpu_kernels import _rank_case, make_engine, rand_triples

pytestmark = pytest.mark.gpu

DOT_KINDS = [ko.DISTMULT, ko.COMPLEX, ko.HOLE]


def _abi():
    from KGEkeras import _abi as abi
    return abi


@pytest.mark.parametrize("kind", DOT_KINDS)
@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("d,N,Q", [(8, 130, 70), (100, 1000, 257), (36, 64, 64), (200, 2500, 300)])
def test_tc_exact_counts_equal_cuda_core_and_oracle(kind, side, d, N, Q):
    abi = _abi()
    if kind == ko.HOLE and d > 100:
        pytest.skip("HolE O(d^2) oracle is slow")
    eng, E, Rm, tri, fh, ft = _rank_case(kind, N, 5, d, Q, seed=31 + d, dyadic=False)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    sd = 0 if side == "tail" else 1
    want = eng.rank_counts(tri, sd, offs, ids).cpu().numpy()
    stats = {}
    got = eng.rank_counts_tc(tri, sd, offs, ids, mode=abi.TC_BF16X3_EXACT, stats=stats).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    if N * Q * d <= 1000 * 257 * 100:
        qv = ko.query_vectors(kind, E, Rm, tri, side)
        sc = ko.seqk_scores(kind, qv, E, side, 12.0)
        np.testing.assert_array_equal(got, (ko.ranks_from_scores(sc, tri, side, filt) - 1).astype(np.int32))
    # the bound is loose by design, but the undecided set must stay a small share of the Q x N pairs
    assert stats["undecided_pairs"] <= max(4 * Q, 0.05 * Q * N), stats


@pytest.mark.parametrize("kind", DOT_KINDS)
@pytest.mark.parametrize("side", ["tail", "head"])
def test_tc_exact_dyadic_with_ties(kind, side):
    abi = _abi()
    eng, E, Rm, tri, fh, ft = _rank_case(kind, 700, 5, 12, 200, seed=3, dyadic=True)
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    sd = 0 if side == "tail" else 1
    got = eng.rank_counts_tc(tri, sd, offs, ids, mode=abi.TC_BF16X3_EXACT).cpu().numpy()
    sc = ko.all_entity_scores(kind, E, Rm, tri, side, 12.0, np.float32)   # exact on dyadic tables
    want = ko.ranks_from_scores(sc, tri, side, filt) - 1
    np.testing.assert_array_equal(got, want.astype(np.int32))


def test_tc_exact_no_filter_and_shards_sum_to_whole():
    abi = _abi()
    kind = ko.COMPLEX
    eng, E, Rm, tri, fh, ft = _rank_case(kind, 1500, 5, 40, 300, seed=9, dyadic=False)
    whole = eng.rank_counts(tri, 0).cpu().numpy()
    np.testing.assert_array_equal(eng.rank_counts_tc(tri, 0, mode=abi.TC_BF16X3_EXACT).cpu().numpy(), whole)
    offs, ids = ko.filter_csr(tri, "tail", ft)
    whole_f = eng.rank_counts(tri, 0, offs, ids).cpu().numpy()
    for G in (2, 3):
        per = -(-1500 // G)
        acc = torch.zeros(len(tri), dtype=torch.int32, device="cuda")
        for g in range(G):
            r0, r1 = g * per, min(1500, (g + 1) * per)
            eng.rank_counts_tc(tri, 0, offs, ids, row_range=(r0, r1), counts=acc, mode=abi.TC_BF16X3_EXACT)
        np.testing.assert_array_equal(acc.cpu().numpy(), whole_f)


def test_tc_worklist_overflow_is_detected_and_retried():
    abi = _abi()
    # identical entity rows: every candidate ties the target -> every pair is undecided -> a tiny worklist must overflow
    N, R, d, Q = 600, 3, 16, 90
    rng = np.random.default_rng(5)
    E = np.repeat(ko.glorot_uniform(rng, 1, d), N, axis=0)
    Rm = ko.glorot_uniform(rng, R, d)
    from KGEkeras._engine import Engine
    eng = Engine(abi.DISTMULT, N, R, d, d, 12.0, E=E, R=Rm)
    tri = rand_triples(rng, Q, N, R)
    stats = {}
    got = eng.rank_counts_tc(tri, 0, mode=abi.TC_BF16X3_EXACT, worklist_cap=128, stats=stats).cpu().numpy()
    assert stats["worklist_cap"] > 128 and stats["undecided_pairs"] >= Q * (N - 1)
    np.testing.assert_array_equal(got, np.zeros(Q, np.int32))   # all ties: nothing is strictly greater
    np.testing.assert_array_equal(got, eng.rank_counts(tri, 0).cpu().numpy())


@pytest.mark.parametrize("kind", [ko.DISTMULT, ko.COMPLEX])
def test_tc_bf16_approximate_within_stated_band(kind):
    abi = _abi()
    N, Q, d = 2000, 256, 64
    eng, E, Rm, tri, fh, ft = _rank_case(kind, N, 5, d, Q, seed=77, dyadic=False)
    offs, ids = ko.filter_csr(tri, "tail", ft)
    got = eng.rank_counts_tc(tri, 0, offs, ids, mode=abi.TC_BF16).cpu().numpy()
    sc64 = ko.all_entity_scores(kind, E, Rm, tri, "tail", 12.0, np.float64)
    r64 = ko.ranks_from_scores(sc64, tri, "tail", ft) - 1
    qv = ko.query_vectors(kind, E.astype(np.float64), Rm.astype(np.float64), tri, "tail")
    band = 2.0 ** -7 * np.linalg.norm(qv, axis=1) * np.linalg.norm(E, axis=1).max()
    for qi in range(Q):
        if got[qi] != r64[qi]:
            tgt = tri[qi, 2]
            near = int(np.sum(np.abs(sc64[qi] - sc64[qi, tgt]) <= band[qi])) - 1
            assert abs(int(got[qi]) - int(r64[qi])) <= near, (qi, got[qi], r64[qi], near)


def test_tc_rejects_transe_l1():
    abi = _abi()
    eng, E, Rm = make_engine(ko.TRANSE_L1, 100, 3, 8)
    with pytest.raises(abi.KgebError):
        eng.pack_entities_tc(abi.TC_BF16)


def _near_tie_table(rng, N, base_row, step_bits=21, spread=600):
    """N rows that are tiny multiples of one pattern: row_j = base * (1 + j * 2^-step_bits) (+ an ulp-sized wobble), so the
    candidates' scores sit from far inside to a few bands outside the target's band and every decision class of the
    tensor-core filter (counted / dropped / undecided) occurs."""
    W = base_row.shape[0]
    steps = (rng.integers(-spread, spread, size=(N, 1)) * np.float32(2.0 ** -step_bits)).astype(np.float32)
    wobble = (rng.integers(0, 4, size=(N, W)) * np.float32(2.0 ** -23)).astype(np.float32)
    return (base_row[None, :] * (np.float32(1.0) + steps) * (np.float32(1.0) + wobble)).astype(np.float32)


def _check_exact_against_cuda_core(E, Rm, kind=None, Q=128, seed=0, want_both=True):
    abi = _abi()
    from KGEkeras._engine import Engine
    kind = ko.DISTMULT if kind is None else kind
    N, W = E.shape
    eng = Engine(kind, N, Rm.shape[0], W, W, 12.0, E=E, R=Rm)
    tri = rand_triples(np.random.default_rng(seed), Q, N, Rm.shape[0])
    stats = {}
    for sd in (0, 1):
        want = eng.rank_counts(tri, sd).cpu().numpy()
        got = eng.rank_counts_tc(tri, sd, mode=abi.TC_F16X3_EXACT, stats=stats).cpu().numpy()
        np.testing.assert_array_equal(got, want)
    return stats, tri, eng


@pytest.mark.parametrize("width", [64, 200])
@pytest.mark.parametrize("pbits", [8, 11])
def test_tc_exact_adversarial_split_residuals(width, pbits):
    """Rows built so that the hi/lo split leaves its LARGEST residuals and every dropped product has the same sign: all entries
    positive and equal to 1 + 2^-p - 2^-2p, which rounds DOWN to 1.0 in a p-bit format (p = 11: the fp16 planes of the exact
    mode; p = 8: the bf16 planes of the first build), so lo ~ 2^-p and the residual ~ 2^-2p.  The dropped terms
    ql.el + rq.e + q.re then add up coherently -- the case the representation term of the band has to cover -- and, the rows
    being parallel and positive, sum|q_k e_k| = ||q|| ||e|| and every tensor-core truncation has the same sign."""
    N, R = 1024, 2
    rng = np.random.default_rng(width + pbits)
    base = np.full(width, 1.0 + 2.0 ** -pbits - 2.0 ** (-2 * pbits), np.float32)
    E = _near_tie_table(rng, N, base)
    Rm = np.ones((R, width), np.float32)
    stats, tri, eng = _check_exact_against_cuda_core(E, Rm, seed=width)
    sc = ko.seqk_scores(ko.DISTMULT, ko.query_vectors(ko.DISTMULT, E, Rm, tri, "tail"), E, "tail", 12.0)
    got = eng.rank_counts_tc(tri, 0, mode=_abi().TC_F16X3_EXACT).cpu().numpy()
    np.testing.assert_array_equal(got, (ko.ranks_from_scores(sc, tri, "tail", {}) - 1).astype(np.int32))
    assert 0 < stats["undecided_pairs"] < 128 * N     # some pairs are decided by the tensor cores, some are not


@pytest.mark.parametrize("shape", ["front_loaded", "back_loaded", "wide_range", "tiny_rows", "alternating"])
def test_tc_exact_band_covers_unfriendly_magnitude_profiles(shape):
    """The band's weighted-norm term assumes nothing about WHERE along k the mass of a row sits, and its absolute terms cover
    elements that fall below the fp16 planes' normal range after the table-wide power-of-two scale.  Profiles: all the mass in
    the first K=16 block (the accumulator is large from the first instruction on: the worst case of the weighted norms), all in
    the last block, magnitudes spread over 2^-30..1 inside every row (low planes go subnormal), rows whose norm is 2^-20 of
    the table's largest (their relative precision after the common scale is poor), alternating signs (cancellation)."""
    N, R, W = 768, 2, 208
    rng = np.random.default_rng(len(shape))
    k = np.arange(W)
    if shape == "front_loaded":
        base = np.where(k < 16, 1.0, 2.0 ** -12) * rng.uniform(0.5, 1.0, W)
    elif shape == "back_loaded":
        base = np.where(k >= W - 16, 1.0, 2.0 ** -12) * rng.uniform(0.5, 1.0, W)
    elif shape == "wide_range":
        base = 2.0 ** -rng.integers(0, 31, W).astype(np.float64) * rng.uniform(0.5, 1.0, W)
    elif shape == "tiny_rows":
        base = rng.uniform(0.5, 1.0, W)
    else:
        base = rng.uniform(0.5, 1.0, W) * np.where(k % 2 == 0, 1.0, -1.0)
    E = _near_tie_table(rng, N, base.astype(np.float32))
    if shape == "tiny_rows":
        E[: N // 2] *= np.float32(2.0 ** -20)          # half of the table is 2^-20 of the other half; near-ties inside both halves
    Rm = np.ones((R, W), np.float32)
    if shape == "alternating":
        Rm[1] = -1.0
    stats, _, _ = _check_exact_against_cuda_core(E, Rm, seed=7)
    assert stats["undecided_pairs"] > 0


@pytest.mark.parametrize("kind", [ko.COMPLEX, ko.HOLE])
def test_tc_exact_near_ties_complex_and_hole(kind):
    N, R, W = 512, 3, 64
    rng = np.random.default_rng(11)
    E = _near_tie_table(rng, N, rng.uniform(-1, 1, W).astype(np.float32))
    Rm = rng.uniform(-1, 1, (R, W)).astype(np.float32)
    _check_exact_against_cuda_core(E, Rm, kind=kind, Q=96, seed=5)


# ---- TransE with the L2 norm on the tensor cores: ||q - e||^2 = ||q||^2 - 2 (q.e - ||e||^2 / 2) ------------------------------------
@pytest.mark.parametrize("gamma", [12.0, 0.0, -1.0])
@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("d,N,Q", [(8, 130, 70), (100, 1000, 257), (36, 64, 64), (200, 2500, 300), (28, 700, 129)])
def test_tc_transe_l2_counts_equal_cuda_core_and_oracle(gamma, side, d, N, Q):
    """Integer filtered-rank counts of the tensor-core path IDENTICAL to the fp32 CUDA-core kernel and to the oracle's sequential
    scores -- gamma > 0 (score = gamma - distance), gamma <= 0 (score = +distance, models.py:189-190: the ranking is NOT flipped),
    widths whose three extra slots fall into the last real block (28, 36, 100, 200) and into a block of their own (8 -> 8..10)."""
    abi = _abi()
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(int(31 + d + 7 * gamma))
    E, Rm = ko.glorot_uniform(rng, N, d) * 3, ko.glorot_uniform(rng, 5, d) * 3
    eng = Engine(abi.TRANSE_L2, N, 5, d, d, gamma, E=E, R=Rm)
    tri = rand_triples(rng, Q, N, 5)
    known = rand_triples(rng, 4 * Q, N, 5)
    known[:Q, :2] = tri[:, :2]
    known[Q:2 * Q, 1:] = tri[:, 1:]
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in known])
    filt = ft if side == "tail" else fh
    offs, ids = ko.filter_csr(tri, side, filt)
    sd = 0 if side == "tail" else 1
    want = eng.rank_counts(tri, sd, offs, ids).cpu().numpy()
    stats = {}
    got = eng.rank_counts_tc(tri, sd, offs, ids, mode=abi.TC_F16X3_EXACT, stats=stats).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    if N * Q * d <= 1000 * 257 * 100:
        sc = ko.seqk_scores(ko.TRANSE_L2, ko.query_vectors(ko.TRANSE_L2, E, Rm, tri, side), E, side, gamma)
        np.testing.assert_array_equal(got, (ko.ranks_from_scores(sc, tri, side, filt) - 1).astype(np.int32))
    assert stats["undecided_pairs"] <= max(4 * Q, 0.05 * Q * N), stats


@pytest.mark.parametrize("gamma", [12.0, -1.0])
@pytest.mark.parametrize("profile", ["near_ties", "tiny_rows", "far_queries", "coincident"])
def test_tc_transe_l2_band_on_unfriendly_tables(gamma, profile):
    """near_ties: candidates that are ulp-sized steps of one row (distances from far inside to a few bands outside the target's);
    tiny_rows: half of the table 2^-12 of the other half; far_queries: relation rows 2^6 times the entities (||q|| >> ||e||);
    coincident: s + p == o exactly for many queries (target distance 0, score == gamma)."""
    abi = _abi()
    from KGEkeras._engine import Engine
    N, R, W, Q = 768, 3, 64, 128
    rng = np.random.default_rng(len(profile) + int(gamma))
    base = rng.uniform(-1, 1, W).astype(np.float32)
    E = _near_tie_table(rng, N, base, step_bits=18, spread=4000)
    Rm = (rng.uniform(-1, 1, (R, W)) * 2.0 ** -10).astype(np.float32)
    if profile == "tiny_rows":
        E[: N // 2] *= np.float32(2.0 ** -12)
    elif profile == "far_queries":
        Rm = (rng.uniform(-1, 1, (R, W)) * 64).astype(np.float32)
    elif profile == "coincident":
        Rm[0] = 0.0                                      # relation 0: s + p == s, and many rows of E are bit-identical pairs
        E[1::2] = E[0::2]
    eng = Engine(abi.TRANSE_L2, N, R, W, W, gamma, E=E, R=Rm)
    tri = rand_triples(rng, Q, N, R)
    if profile == "coincident":
        tri[: Q // 2, 1] = 0
        tri[: Q // 2, 2] = tri[: Q // 2, 0] ^ 1          # the target is the bit-identical twin of the subject: distance exactly 0
    stats = {}
    for sd in (0, 1):
        want = eng.rank_counts(tri, sd).cpu().numpy()
        got = eng.rank_counts_tc(tri, sd, mode=abi.TC_F16X3_EXACT, stats=stats).cpu().numpy()
        np.testing.assert_array_equal(got, want)


def test_tc_transe_l2_shards_sum_to_whole_and_validate_uses_it():
    abi = _abi()
    import KGEkeras
    from KGEkeras import utils as ku
    N, R, d = 1500, 4, 40
    rng = np.random.default_rng(2)
    E, Rm = ko.glorot_uniform(rng, N, d) * 3, ko.glorot_uniform(rng, R, d) * 3
    m = KGEkeras.TransE(e_dim=d, r_dim=d, num_entities=N, num_relations=R, norm=2, gamma=12, init_entities=E, init_relations=Rm, dp=0.0)
    eng = m._engine_()
    assert eng.tc_supported()
    tri = rand_triples(rng, 300, N, R)
    whole = eng.rank_counts(tri, 0).cpu().numpy()
    for G in (2, 3):
        per = -(-N // G)
        acc = torch.zeros(len(tri), dtype=torch.int32, device="cuda")
        for g in range(G):
            eng.rank_counts_tc(tri, 0, row_range=(g * per, min(N, (g + 1) * per)), counts=acc, mode=abi.TC_F16X3_EXACT)
        np.testing.assert_array_equal(acc.cpu().numpy(), whole)
    filt = rand_triples(rng, 2000, N, R)
    key = lambda a: (a[:, 0].astype(np.int64) * R + a[:, 1]) * N + a[:, 2]
    filt = filt[~np.isin(key(filt), key(tri))]
    res = ku.validate(m, tri, N, 32, filt)                      # default rank mode: exact, on the tensor cores for this kind now
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in filt])
    ranks = []
    for side, f in (("tail", ft), ("head", fh)):
        sc = ko.seqk_scores(ko.TRANSE_L2, ko.query_vectors(ko.TRANSE_L2, E, Rm, tri, side), E, side, 12.0)
        ranks.append(ko.ranks_from_scores(sc, tri, side, f))
    want = ko.metrics_from_ranks(ranks[0], ranks[1])
    for k in ("mrr", "hits@10", "mr"):
        assert abs(res[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, res[k], want[k])


# ---- compute-sanitizer is closed on the pool: a race in the producer / MMA / epilogue hand-offs would show up as counts that change
# from run to run or differ from the CUDA-core kernel on SOME shape; this sweeps the shapes that move every pipeline parameter ----------
def test_tc_stress_shapes_and_repeatability():
    """Widths that give 1..13 pipeline stages per tile in both modes (odd stage counts exercise the paired hand-off's tail), query and
    entity counts around the tile edges (127/128/129 queries, 255/256/257/513 entities), shards that start inside a tile, queries >> SMs
    (every CTA walks several items: both TMEM accumulator buffers and all ring stages wrap many times).  Every case: exact mode ==
    CUDA-core kernel, and 8 repeats of the same pass return the same integers (approximate mode included)."""
    abi = _abi()
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(123)
    cases = [(16, 255, 127), (32, 256, 128), (48, 257, 129), (64, 513, 64), (104, 1000, 300), (208, 777, 257), (400, 520, 140),
             (16, 40000, 2500), (136, 9000, 1100)]
    for W, N, Q in cases:
        kind = abi.COMPLEX if W % 8 == 0 and rng.integers(2) else abi.DISTMULT
        E, Rm = ko.glorot_uniform(rng, N, W) * 3, ko.glorot_uniform(rng, 4, W) * 3
        eng = Engine(kind, N, 4, W, W, 12.0, E=E, R=Rm)
        tri = rand_triples(rng, Q, N, 4)
        r0 = int(rng.integers(0, min(N // 3, 300)))
        want = torch.zeros(Q, dtype=torch.int32, device="cuda")
        for lo, hi in ((0, r0), (r0, N)):
            if hi > lo:
                eng.rank_counts(tri, 1, row_range=(lo, hi), counts=want)
        base = {}
        for rep in range(8):
            for mode in (abi.TC_F16X3_EXACT, abi.TC_BF16):
                got = torch.zeros(Q, dtype=torch.int32, device="cuda")
                for lo, hi in ((0, r0), (r0, N)):
                    if hi > lo:
                        eng.rank_counts_tc(tri, 1, row_range=(lo, hi), counts=got, mode=mode)
                if rep == 0:
                    base[mode] = got.clone()
                    if mode == abi.TC_F16X3_EXACT:
                        assert torch.equal(got, want), (W, N, Q, r0)
                else:
                    assert torch.equal(got, base[mode]), (W, N, Q, r0, mode, rep)
        del eng


def test_tc_ts_variant_passes_the_approximate_mode_tests():
    """k_rank_tc_ts (query operand in TMEM; off by default: profiles/r02_tc_ts.md) is selected by KGEB_TC_TS=1, read once per process:
    run the approximate-mode tests and the shape sweep in a child process with it set."""
    import os
    import subprocess
    import sys
    if os.environ.get("KGEB_TC_TS") == "1":
        pytest.skip("already inside the child")
    env = dict(os.environ, KGEB_TC_TS="1")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-k", "bf16_approximate or stress_shapes"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


def test_one_cta_per_tile_variant_in_child_process():
    """CTA pairs (k_rank_tc_t<2>, cta_group::2) are the default since round 2; KGEB_TC_2CTA=0 (read once per process) keeps one CTA per tile
    (k_rank_tc_t<1>): every test of this file passes on it too."""
    import os
    import subprocess
    import sys
    if os.environ.get("KGEB_TC_2CTA") == "0" or os.environ.get("KGEB_TC_TS") == "1":
        pytest.skip("already inside a child")
    env = dict(os.environ, KGEB_TC_2CTA="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k", "not variant"], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


# ---- the undecided pairs ordered by entity before the exact re-scoring (wl_group_kernels.cu, k_seqk_pairs<OP, PAIRS_GROUPED>) -------
@pytest.mark.parametrize("case", ["all_ties", "near_ties_complex", "transe_l2_tail", "transe_l2_head", "sharded"])
def test_tc_exact_with_the_worklist_grouped_by_entity(case, monkeypatch):
    """KGEB_PAIRS_GROUPED=1 forces the path the large shapes take by themselves: a counting sort of the worklist by entity on the
    device, then the re-scoring kernel whose warps stage every distinct entity row once.  The integer counts must not change:
    identical to the CUDA-core kernel with every candidate tied (each entity in Q pairs: long runs), with clustered near-ties
    (runs of mixed length, many warps that straddle runs), for the L2 operand pairs, and summed over shards."""
    abi = _abi()
    from KGEkeras._engine import Engine
    monkeypatch.setenv("KGEB_PAIRS_GROUPED", "1")        # read once, when the context is created
    rng = np.random.default_rng(23)
    if case == "all_ties":
        N, R, d, Q = 600, 3, 16, 90
        E = np.repeat(ko.glorot_uniform(rng, 1, d), N, axis=0)
        Rm = ko.glorot_uniform(rng, R, d)
        eng = Engine(abi.DISTMULT, N, R, d, d, 12.0, E=E, R=Rm)
        tri = rand_triples(rng, Q, N, R)
        stats = {}
        got = eng.rank_counts_tc(tri, 0, mode=abi.TC_BF16X3_EXACT, stats=stats).cpu().numpy()
        assert stats["undecided_pairs"] >= Q * (N - 1)
        np.testing.assert_array_equal(got, eng.rank_counts(tri, 0).cpu().numpy())
        return
    if case == "near_ties_complex":
        N, R, W = 512, 3, 64
        E = _near_tie_table(rng, N, rng.uniform(-1, 1, W).astype(np.float32))
        Rm = rng.uniform(-1, 1, (R, W)).astype(np.float32)
        stats, _, _ = _check_exact_against_cuda_core(E, Rm, kind=ko.COMPLEX, Q=96, seed=5)
        assert stats["undecided_pairs"] > 0
        return
    if case.startswith("transe_l2"):
        side = case.split("_")[-1]
        N, d, Q = 700, 28, 129
        base = ko.glorot_uniform(rng, 1, d) * 3
        E = (base + rng.normal(size=(N, d)) * 1e-6).astype(np.float32)     # clustered rows: distances agree to ~1e-6
        Rm = ko.glorot_uniform(rng, 5, d) * 3
        eng = Engine(abi.TRANSE_L2, N, 5, d, d, 12.0, E=E, R=Rm)
        tri = rand_triples(rng, Q, N, 5)
        sd = 0 if side == "tail" else 1
        stats = {}
        got = eng.rank_counts_tc(tri, sd, mode=abi.TC_F16X3_EXACT, stats=stats).cpu().numpy()
        assert stats["undecided_pairs"] > 0
        np.testing.assert_array_equal(got, eng.rank_counts(tri, sd).cpu().numpy())
        return
    # sharded: partial counts over row ranges add up to the whole, filter included
    eng, E, Rm, tri, fh, ft = _rank_case(ko.COMPLEX, 1500, 5, 40, 300, seed=9, dyadic=False)
    offs, ids = ko.filter_csr(tri, "tail", ft)
    whole_f = eng.rank_counts(tri, 0, offs, ids).cpu().numpy()
    for G in (1, 3):
        per = -(-1500 // G)
        acc = torch.zeros(len(tri), dtype=torch.int32, device="cuda")
        for g in range(G):
            eng.rank_counts_tc(tri, 0, offs, ids, row_range=(g * per, min(1500, (g + 1) * per)), counts=acc, mode=abi.TC_BF16X3_EXACT)
        np.testing.assert_array_equal(acc.cpu().numpy(), whole_f)
