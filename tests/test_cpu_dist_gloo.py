"""world_size-2 (and 3) gloo test of the multi-GPU evaluation host logic on CPU: every rank counts only the candidates of
its row shard (KGEkeras.utils.shard_rows), the integer partial counts are summed with the same all-reduce call the GPU path
uses (KGEkeras.utils.merge_partial_counts), and the result must equal the single-shard ranks bit for bit.  The per-shard
counts come from the oracle here (no GPU in this container); on the GPU box tests/test_gpu_api.py checks the kernel shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kge_oracle as ko


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, ret):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    for p in (root, os.path.join(root, "kge-keras_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from KGEkeras import utils as ku
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)            # same data on every rank
    N, R, d, Q = 257, 4, 8, 40
    W = 2 * d if kind == ko.COMPLEX else d
    E, Rm = ko.glorot_uniform(rng, N, W), ko.glorot_uniform(rng, R, W)
    tri = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1)
    known = np.stack([tri[:, 0], tri[:, 1], rng.integers(0, N, Q)], 1)
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in known])
    sc = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, tri, "tail"), E, "tail", 12.0)
    r0, r1 = ku.shard_rows(N, world, rank)
    part = np.zeros(Q, np.int32)
    for q in range(Q):
        keep = np.zeros(N, bool)
        keep[r0:r1] = True
        keep[list(ft.get((tri[q, 0], tri[q, 1]), ()))] = False
        keep[tri[q, 2]] = False
        part[q] = np.count_nonzero(sc[q][keep] > sc[q, tri[q, 2]])
    total = ku.merge_partial_counts(torch.from_numpy(part.copy()))
    want = ko.ranks_from_scores(sc, tri, "tail", ft)
    ok = np.array_equal(total.numpy() + 1, want)
    ret[rank] = (bool(ok), int(part.sum()), (r0, r1))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rank_counts_merge_exactly(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ko.TRANSE_L1, ret), nprocs=world, join=True)
    assert len(ret) == world and all(v[0] for v in ret.values()), dict(ret)
    shards = sorted(v[2] for v in ret.values())
    assert shards[0][0] == 0 and shards[-1][1] == 257 and all(a[1] == b[0] for a, b in zip(shards, shards[1:]))


def test_shard_rows_covers_everything():
    from KGEkeras.utils import shard_rows
    for N in (1, 7, 104, 40943, 1000000):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_rows(N, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(a[0] <= a[1] for a in parts)
