"""Property tests (hypothesis) of the host-side logic: the filter-CSR builder against the reference's dict-of-sets
construction (utils.py:201-205), row sharding, the exchange plan of the row-sharded trainer, the Philox sampler's ranges."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import kge_oracle as ko

triples = st.lists(st.tuples(st.integers(0, 11), st.integers(0, 3), st.integers(0, 11)), min_size=0, max_size=60)


@settings(max_examples=150, deadline=None)
@given(test=st.lists(st.tuples(st.integers(0, 11), st.integers(0, 3), st.integers(0, 11)), min_size=1, max_size=25), filt=triples,
       side=st.sampled_from(["tail", "head"]))
def test_filter_csr_equals_reference_dict_of_sets(test, filt, side):
    from KGEkeras import utils as ku
    fh, ft = ko.build_filters(filt)               # the reference's filter dicts (utils.py:201-205), restated
    want = []
    crash = False
    for s, p, o in test:
        ids = sorted(ft.get((s, p), ())) if side == "tail" else sorted(fh.get((p, o), ()))
        tgt = o if side == "tail" else s
        crash |= tgt in ids                          # the reference's list.remove would raise here (utils.py:169-171)
        want.append(ids)
    if crash:
        with pytest.raises(ValueError):
            ku.build_filter_csr(test, filt, side)
        return
    offs, ids = ku.build_filter_csr(test, filt, side)
    assert offs[0] == 0 and len(offs) == len(test) + 1 and offs[-1] == len(ids)
    for q, w in enumerate(want):
        assert list(ids[offs[q]:offs[q + 1]]) == w
    o2, i2 = ko.filter_csr(np.asarray(test), side, ft if side == "tail" else fh)
    np.testing.assert_array_equal(offs, o2)
    np.testing.assert_array_equal(ids, i2)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 10_000), world=st.integers(1, 16))
def test_shard_rows_partition(n, world):
    from KGEkeras.utils import shard_rows
    prev = 0
    for r in range(world):
        r0, r1 = shard_rows(n, world, r)
        assert r0 == prev and r0 <= r1 <= n
        prev = r1
    assert prev == n


@settings(max_examples=60, deadline=None)
@given(ids=st.lists(st.integers(0, 49), min_size=1, max_size=80))
def test_exchange_plan_world1_remaps_onto_sorted_unique(ids):
    import torch
    from KGEkeras._sharded import ExchangePlan
    t = torch.tensor(ids, dtype=torch.int64)
    plan = ExchangePlan(t, 50)
    uniq = sorted(set(ids))
    assert plan.uniq.tolist() == uniq and plan.U == len(uniq)
    assert plan.req_counts == [len(uniq)] and plan.serve_counts == [len(uniq)]
    assert plan.serve_local.tolist() == uniq
    tri = torch.stack([t, torch.zeros_like(t), t.flip(0)], dim=1)
    rm = plan.remap_triples(tri)
    assert [uniq[i] for i in rm[:, 0].tolist()] == ids and [uniq[i] for i in rm[:, 2].tolist()] == ids[::-1]


@settings(max_examples=40, deadline=None)
@given(N=st.integers(1, 5000), K=st.integers(1, 9), seed=st.integers(0, 2 ** 63 - 1), offset=st.integers(0, 2 ** 40),
       mode=st.sampled_from([ko.CORRUPT_BOTH, ko.CORRUPT_HEAD, ko.CORRUPT_TAIL, ko.CORRUPT_EITHER]))
def test_philox_negatives_range_and_structure(N, K, seed, offset, mode):
    rng = np.random.default_rng(N + K)
    B = 17
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, 5, B), rng.integers(0, N, B)], 1)
    neg = ko.philox_negatives(pos, N, K, seed, offset, mode)
    assert neg.shape == (B * K, 3) and neg[:, [0, 2]].min() >= 0 and neg[:, [0, 2]].max() < N
    rep = np.repeat(pos, K, axis=0)
    np.testing.assert_array_equal(neg[:, 1], rep[:, 1])                         # the relation is kept (utils.py:117)
    if mode == ko.CORRUPT_HEAD:
        np.testing.assert_array_equal(neg[:, 2], rep[:, 2])
    if mode == ko.CORRUPT_TAIL:
        np.testing.assert_array_equal(neg[:, 0], rep[:, 0])
    if mode == ko.CORRUPT_EITHER:
        assert np.all((neg[:, 0] == rep[:, 0]) | (neg[:, 2] == rep[:, 2]))      # exactly one slot is redrawn
    np.testing.assert_array_equal(neg, ko.philox_negatives(pos, N, K, seed, offset, mode))   # counter based: reproducible


@settings(max_examples=60, deadline=None)
@given(d=st.integers(1, 9), m=st.integers(2, 12), seed=st.integers(0, 10_000), pw=st.floats(0.1, 2.0), mw=st.floats(0.1, 2.0))
def test_hake_call_coupling_properties(d, m, seed, pw, mw):
    """Properties of HAKE as the reference wrote it (models.py:397, norm without axis): within a call every score carries the
    same modulus term, so score differences are pure phase terms; a sub-call's scores are never below the full call's (its
    Frobenius norm is over fewer rows); the gradient of sum_i c_i score_i through the modulus term scales with sum_i c_i."""
    rng = np.random.default_rng(seed)
    s, o = rng.normal(size=(2, m, 2 * d))
    p = rng.normal(size=(m, 3 * d))
    hp = (float(np.pi / 0.35), pw, mw)
    full = ko.hake_scores(s, p, o, 12.0, *hp)
    phase_only = ko.hake_scores(s, p, o, 12.0, hp[0], pw, 0.0)
    np.testing.assert_allclose(full - full[0], phase_only - phase_only[0], rtol=0, atol=1e-9)      # one shared modulus term
    sub = ko.hake_scores(s[: m // 2 + 1], p[: m // 2 + 1], o[: m // 2 + 1], 12.0, *hp)
    assert np.all(sub >= full[: m // 2 + 1] - 1e-12)
    c = rng.normal(size=m)
    c -= c.mean()                                                                                    # sum c_i = 0 ...
    gs, gp, go = ko.hake_row_grads(s, p, o, c, 12.0, *hp)
    g0 = ko.hake_row_grads(s, p, o, c, 12.0, hp[0], pw, 0.0)
    for a, b in zip((gs, gp, go), g0):                                                               # ... kills the modulus gradient
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9 * (1 + np.abs(b).max()))


@settings(max_examples=40, deadline=None)
@given(logd=st.integers(1, 6), seed=st.integers(0, 1000))
def test_hole_fourier_identities_used_by_the_fft_kernels(logd, seed):
    """The identities hole_fft_kernels.cu relies on, in fp64: one complex FFT of s + i o yields both spectra; the score is
    (1/d) Re sum conj(P) conj(S) O; ds + i do = F^-1(c conj(P) O + i c P S); dp = Re F^-1(c conj(S) O)."""
    d = 2 ** logd
    rng = np.random.default_rng(seed)
    s, p, o = rng.normal(size=(3, d))
    c = float(rng.normal())
    Z = np.fft.fft(s + 1j * o)
    Zm = np.conj(Z[(-np.arange(d)) % d])
    S, O, P = (Z + Zm) / 2, -1j * (Z - Zm) / 2, np.fft.fft(p)
    np.testing.assert_allclose(S, np.fft.fft(s), atol=1e-9)
    np.testing.assert_allclose(O, np.fft.fft(o), atol=1e-9)
    score = ko.score(ko.HOLE, s[None], p[None], o[None], hole_fft=False)[0]
    assert abs(score - np.sum((np.conj(P) * np.conj(S) * O).real) / d) < 1e-9 * (1 + abs(score))
    ds, dp, do = ko.score_grads(ko.HOLE, s[None], p[None], o[None])
    g = np.fft.ifft(c * np.conj(P) * O + 1j * c * P * S)
    np.testing.assert_allclose(g.real, c * ds[0], atol=1e-9 * (1 + np.abs(ds).max()))
    np.testing.assert_allclose(g.imag, c * do[0], atol=1e-9 * (1 + np.abs(do).max()))
    np.testing.assert_allclose(np.fft.ifft(c * np.conj(S) * O).real, c * dp[0], atol=1e-9 * (1 + np.abs(dp).max()))


def _hf_swz(j):
    return j ^ (((j >> 4) & 3) << 2)


def _stockham_as_in_hole_fft_kernels(z, inverse=False):
    """The stage loops of warp_fft (kge-keras_b200/csrc/hole_fft_kernels.cu) restated index for index: radix-4 Stockham autosort stages
    with the per-stage compact twiddle tables w_m[k] = exp(-2 pi i m k / (4 p)), the p = 4 stage STORING through the bank permutation
    hf_swz and the stage after it (radix-4 p = 16, or the radix-2 stage when d = 32) LOADING through it, one radix-2 stage for odd log2 d."""
    d = len(z)
    logd = d.bit_length() - 1
    tw = np.exp(-2j * np.pi * np.arange(d) / d)
    ctw = np.zeros(d, complex)
    p = 1
    while 4 * p <= d:
        for e in range(3 * p):
            m = e // p + 1
            ctw[(p - 1) + e] = np.exp(-2j * np.pi * m * (e - (m - 1) * p) / (4 * p))
        p <<= 2
    x, y = np.array(z, complex), np.zeros(d, complex)
    p, lp = 1, 0
    while 4 * p <= d:
        t, rd, wr = d >> 2, p == 16, p == 4
        for i in range(t):
            k = i & (p - 1)
            j = ((i - k) << 2) + k
            w = [ctw[p - 1 + m * p + k] for m in range(3)]
            if inverse:
                w = [np.conj(v) for v in w]
            src = [i + m * t for m in range(4)]
            if rd:
                src = [_hf_swz(q) for q in src]
            u0, u1, u2, u3 = x[src[0]], x[src[1]] * w[0], x[src[2]] * w[1], x[src[3]] * w[2]
            v0, v1, v2, df = u0 + u2, u0 - u2, u1 + u3, u1 - u3
            v3 = 1j * df if inverse else -1j * df
            for m, val in enumerate((v0 + v2, v1 + v3, v0 - v2, v1 - v3)):
                y[_hf_swz(j + m * p) if wr else j + m * p] = val
        x, y = y, x
        p, lp = p << 2, lp + 2
    if p < d:
        t, sh, rd = d >> 1, logd - 1 - lp, p == 16
        for i in range(t):
            k = i & (p - 1)
            j = ((i - k) << 1) + k
            w = np.conj(tw[k << sh]) if inverse else tw[k << sh]
            a, b = (_hf_swz(i), _hf_swz(i + t)) if rd else (i, i + t)
            y[j], y[j + p] = x[a] + x[b] * w, x[a] - x[b] * w
        x, y = y, x
    return x


@pytest.mark.parametrize("d", [32, 64, 128, 256, 512, 1024])
def test_hole_fft_stage_indexing_with_the_bank_permutation(d):
    """The index maps of the conflict-free FFT stages (compact twiddle tables, hf_swz between the p = 4 and the following stage) give the
    DFT for every supported width, forward and inverse; hf_swz is a bijection of every aligned group of 64 indices that keeps the 16-element
    groups a warp reads in one wavefront intact."""
    rng = np.random.default_rng(d)
    z = rng.normal(size=d) + 1j * rng.normal(size=d)
    assert np.abs(_stockham_as_in_hole_fft_kernels(z) - np.fft.fft(z)).max() < 1e-11
    assert np.abs(_stockham_as_in_hole_fft_kernels(z, True) - np.fft.ifft(z) * d).max() < 1e-11
    idx = np.arange(d)
    img = np.array([_hf_swz(int(j)) for j in idx])
    assert sorted(img) == list(idx) and np.all(img >> 6 == idx >> 6) and np.all(img >> 4 == idx >> 4)
    # the store pattern of the p = 4 stage: one store instruction = 32 lanes writing j = 16 a + b + 4 m (a = lane >> 2, b = lane & 3);
    # through hf_swz the 32 addresses cover every 8-byte bank slot exactly twice (two wavefronts: the minimum for 256 bytes)
    for m in range(4):
        slots = [_hf_swz(16 * (lane >> 2) + (lane & 3) + 4 * m) % 16 for lane in range(32)]
        assert sorted(slots) == sorted(list(range(16)) * 2)
        plain = [(16 * (lane >> 2) + (lane & 3) + 4 * m) % 16 for lane in range(32)]
        assert len(set(plain)) == 4          # without it: four slots, eight lanes each
