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


# ---- row-sharded training: exchange plan + all-to-all on gloo, compute by the oracle ---------------------------------
def _shard_train_worker(rank, world, port, ret):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    for p in (root, os.path.join(root, "kge-keras_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from KGEkeras import _sharded as sh, utils as ku
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kind, N, R, W, B, K, lr = ko.DISTMULT, 61, 3, 8, 12, 3, 0.1
    rng = np.random.default_rng(0)                       # the same full tables on every rank
    E, Rm = ko.glorot_uniform(rng, N, W) * 3, ko.glorot_uniform(rng, R, W) * 3
    pos_all = np.stack([rng.integers(0, N, world * B), rng.integers(0, R, world * B), rng.integers(0, N, world * B)], 1)
    neg_all = ko.philox_negatives(pos_all, N, K, 5, 0, ko.CORRUPT_BOTH)
    pos, neg = pos_all[rank * B:(rank + 1) * B], neg_all[rank * B * K:(rank + 1) * B * K]
    r0, r1 = ku.shard_rows(N, world, rank)
    E_shard, g_shard = E[r0:r1].astype(np.float64), np.zeros((r1 - r0, W))
    ids = torch.from_numpy(np.concatenate([pos[:, 0], pos[:, 2], neg[:, 0], neg[:, 2]]))
    plan = sh.ExchangePlan(ids, N)
    # rows: owners -> requesters
    send = torch.from_numpy(E_shard[plan.serve_local.numpy()])
    Ec = torch.empty(plan.U, W, dtype=torch.float64)
    sh.all_to_all_rows(Ec, send, plan.req_counts, plan.serve_counts)
    ok = np.array_equal(Ec.numpy(), E.astype(np.float64)[plan.uniq.numpy()])
    # the step on the compact table (oracle), mean folded by 1/world
    w = ko.train_step(kind, ko.PAIRWISE_HINGE, ko.PAIRED, Ec.numpy(), Rm, plan.remap_triples(torch.from_numpy(pos)).numpy(),
                      plan.remap_triples(torch.from_numpy(neg)).numpy(), K, 1.0, 12.0, ko.OPT_SGD, [], [], lr, 1)
    back = torch.empty(len(plan.serve_local), W, dtype=torch.float64)
    sh.all_to_all_rows(back, torch.from_numpy(w["gE"] / world), plan.serve_counts, plan.req_counts)
    np.add.at(g_shard, plan.serve_local.numpy(), back.numpy())
    gR = torch.from_numpy(w["gR"] / world)
    dist.all_reduce(gR)
    E_new = E_shard - lr * g_shard
    # single-process truth: one step over the concatenated batch
    full = ko.train_step(kind, ko.PAIRWISE_HINGE, ko.PAIRED, E, Rm, pos_all, neg_all, K, 1.0, 12.0, ko.OPT_SGD, [], [], lr, 1)
    ok = ok and np.allclose(E_new, full["E"][r0:r1], rtol=1e-12, atol=1e-14) and np.allclose(gR.numpy(), full["gR"], rtol=1e-12, atol=1e-14)
    ok = ok and sum(plan.req_counts) == plan.U and len(plan.serve_local) == sum(plan.serve_counts)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_training_exchange_equals_single_process_step(world):
    ret = mp.Manager().dict()
    mp.spawn(_shard_train_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)
