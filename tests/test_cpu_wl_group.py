"""CPU restatement of the index arithmetic behind the entity-ordered re-scoring (kge-keras_b200/csrc/wl_group_kernels.cu and the
PAIRS_GROUPED branch of k_seqk_pairs in rank_kernels.cu): the counting sort whose slots are handed out by atomics in arbitrary
order, and the ballot / popc bookkeeping by which a warp stages each distinct entity row once and every lane finds its row.
The GPU tests (tests/test_gpu_rank_tc.py::test_tc_exact_with_the_worklist_grouped_by_entity) check the counts; this file checks
the arithmetic for every run pattern of a warp, including the 32-bit shift edge at lane 31."""
import numpy as np
import pytest

U32 = 0xFFFFFFFF


def counting_sort_by_entity(pairs, n_rows, cap, appended, rng):
    """k_wl_hist -> ExclusiveSum -> k_wl_scatter with the atomics resolved in a random order."""
    n = min(appended, cap)
    bins = np.zeros(n_rows + 1, np.uint32)
    slot = np.full(cap, U32, np.uint32)
    for x in rng.permutation(n):                       # any interleaving of the threads
        c = int(pairs[x, 1])
        if c < n_rows:
            slot[x] = bins[c]
            bins[c] += 1
    offsets = np.concatenate([[0], np.cumsum(bins[:-1], dtype=np.uint64)]).astype(np.uint32)   # exclusive sum over n_rows + 1 bins
    out = np.zeros((cap, 2), np.uint32)
    for x in range(n):
        if slot[x] != U32:
            out[offsets[pairs[x, 1]] + slot[x]] = pairs[x]
    return out, int(offsets[n_rows])


@pytest.mark.parametrize("n_rows,appended,cap", [(1, 5, 8), (7, 0, 4), (50, 400, 400), (50, 400, 300), (1000, 333, 512), (3, 64, 64)])
def test_counting_sort_is_a_permutation_ordered_by_entity(n_rows, appended, cap):
    rng = np.random.default_rng(n_rows + appended)
    stored = min(appended, cap)
    pairs = np.stack([rng.integers(0, 1 << 14, cap), rng.integers(0, n_rows, cap)], 1).astype(np.uint32)
    out, length = counting_sort_by_entity(pairs, n_rows, cap, appended, rng)
    assert length == stored                                            # what the re-scoring kernel reads as its pair count
    got = out[:length]
    assert np.all(np.diff(got[:, 1].astype(np.int64)) >= 0)            # ordered by entity
    key = lambda a: np.sort(a[:, 0].astype(np.int64) * (1 << 32) + a[:, 1])
    np.testing.assert_array_equal(key(got), key(pairs[:stored]))       # the same multiset of pairs: nothing lost, nothing doubled


def warp_slots(c):
    """PAIRS_GROUPED bookkeeping of one warp: c = the 32 lanes' entity ids (sorted runs)."""
    first = [lane == 0 or c[lane] != c[lane - 1] for lane in range(32)]          # shfl_up(c, 1) != c, lane 0 always opens a row
    firstmask = sum(1 << lane for lane in range(32) if first[lane])               # ballot
    below = lambda p: ((2 << p) & U32) - 1 & U32                                   # (2u << p) - 1u in 32-bit arithmetic (p = 31: 0 - 1)
    eslot = [bin(firstmask & below(lane)).count("1") - 1 for lane in range(32)]
    staged = {}
    for p in range(32):                                                           # the copy loop: 4 * i + sub covers p = 0..31
        if (firstmask >> p) & 1:
            staged[bin(firstmask & below(p)).count("1") - 1] = c[p]
    return firstmask, eslot, staged


@pytest.mark.parametrize("pattern", ["all_equal", "all_distinct", "runs", "tail_padding", "single_break_at_31", "random"])
def test_every_lane_reads_its_own_entity_row(pattern):
    rng = np.random.default_rng(len(pattern))
    cases = []
    if pattern == "all_equal":
        cases.append([7] * 32)
    elif pattern == "all_distinct":
        cases.append(list(range(100, 132)))
    elif pattern == "runs":
        cases.append(sorted(rng.integers(0, 6, 32).tolist()))
        cases.append([1] * 31 + [2])
        cases.append([1] + [2] * 31)
    elif pattern == "tail_padding":                 # lanes past the end of the worklist carry c = 0 after the valid, larger ids
        cases.append([40] * 10 + [41] * 5 + [0] * 17)
    elif pattern == "single_break_at_31":
        cases.append([3] * 31 + [9])
    else:
        for _ in range(200):
            cases.append(sorted(rng.integers(0, int(rng.integers(1, 40)), 32).tolist()))
    for c in cases:
        firstmask, eslot, staged = warp_slots(c)
        assert len(staged) == bin(firstmask).count("1") == len(set(zip(c, np.cumsum([i == 0 or c[i] != c[i - 1] for i in range(32)]))))
        assert sorted(staged) == list(range(len(staged)))             # slots 0 .. ndist-1, each written exactly once
        for lane in range(32):
            assert 0 <= eslot[lane] < 32
            assert staged[eslot[lane]] == c[lane], (c, lane)
