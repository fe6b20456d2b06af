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
