"""Real multi-GPU check of the sharded 1-N evaluation (skipped with fewer than 2 GPUs): one process per GPU, NCCL backend,
every rank calls the public validate(); the result must equal the single-GPU ranks on every rank, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    for p in (root, os.path.join(root, "kge-keras_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from KGEkeras import models as km, utils as ku
    from oracle import kge_oracle as ko
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = np.load(os.path.join(root, "tests", "golden", "datasets", "UMLS.npz"))
    N, R = int(d["num_entities"]), int(d["num_relations"])
    rng = np.random.default_rng(5)
    E, Rm = ko.glorot_uniform(rng, N, 16), ko.glorot_uniform(rng, R, 16)
    m = km.TransE(e_dim=16, r_dim=16, num_entities=N, num_relations=R, norm=1, gamma=12, init_entities=E, init_relations=Rm, dp=0.0)
    test = d["test"][:120]
    filt = np.concatenate([d["train"], d["valid"]])
    tr, hr = ku.filtered_ranks(m, test, filt)                      # sharded over `world` ranks + NCCL all-reduce
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in filt])
    ok = True
    for side, got, f in (("tail", tr, ft), ("head", hr, fh)):
        sc = ko.seqk_scores(ko.TRANSE_L1, ko.query_vectors(ko.TRANSE_L1, E, Rm, test, side), E, side, 12.0)
        ok = ok and np.array_equal(got, ko.ranks_from_scores(sc, test, side, f))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_sharded_validate():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)


# ---- row-sharded training ------------------------------------------------------------------------------------------
def _case(world, B, K, N=5000, R=7, d=32, seed=3):
    from oracle import kge_oracle as ko
    rng = np.random.default_rng(seed)
    E, Rm = ko.glorot_uniform(rng, N, d) * 3, ko.glorot_uniform(rng, R, d) * 3
    pos = np.stack([rng.integers(0, N, world * B), rng.integers(0, R, world * B), rng.integers(0, N, world * B)], 1).astype(np.int32)
    neg = ko.philox_negatives(pos, N, K, 11, 0, ko.CORRUPT_BOTH).astype(np.int32)
    return E, Rm, pos, neg


def test_sharded_trainer_world1_equals_engine_step():
    """world = 1: plan, row gather, compact-table step, scatter-add and update must reproduce Engine.train_step."""
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    from KGEkeras._sharded import ShardedTrainer
    N, R, d, B, K = 5000, 7, 32, 256, 4
    E, Rm, pos, neg = _case(1, B, K)
    for kind in (abi.TRANSE_L1, abi.COMPLEX):
        ref = Engine(kind, N, R, d, d, 12.0, E=E, R=Rm)
        st = ShardedTrainer(kind, N, R, d, d, 12.0, E=E, R=Rm)
        for step in range(3):
            l0 = ref.train_step(pos, neg, opt_kind=abi.LAZY_ADAM, lr=0.01, K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0)["loss"].item()
            l1 = st.step(pos, neg, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01).item()
            assert abs(l0 - l1) <= 1e-6 * max(1.0, abs(l0))
        np.testing.assert_allclose(st.gather_full_table(), ref.get_tables()[0], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(st.shard.get_tables()[1], ref.get_tables()[1], rtol=1e-5, atol=1e-6)
        assert st.last["unique_rows"] == len(np.unique(np.concatenate([pos[:, [0, 2]].ravel(), neg[:, [0, 2]].ravel()])))


def _shard_worker(rank, world, port, ret):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    for p in (root, os.path.join(root, "kge-keras_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    from KGEkeras._sharded import ShardedTrainer
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    N, R, d, B, K = 5000, 7, 32, 128, 4
    E, Rm, pos, neg = _case(world, B, K)
    st = ShardedTrainer(abi.COMPLEX, N, R, d, d, 12.0, E=E, R=Rm)
    ref = Engine(abi.COMPLEX, N, R, d, d, 12.0, E=E, R=Rm)           # every rank also runs the single-GPU truth
    ok = True
    for step in range(3):
        l1 = st.step(pos[rank * B:(rank + 1) * B], neg[rank * B * K:(rank + 1) * B * K], loss_kind=abi.PAIRWISE_HINGE, margin=1.0,
                     opt_kind=abi.LAZY_ADAM, lr=0.01).item()
        l0 = ref.train_step(pos, neg, opt_kind=abi.LAZY_ADAM, lr=0.01, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0)["loss"].item()
        ok = ok and abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0))
    full = st.gather_full_table()
    ok = ok and np.allclose(full, ref.get_tables()[0], rtol=1e-4, atol=1e-5)
    ok = ok and np.allclose(st.shard.get_tables()[1], ref.get_tables()[1], rtol=1e-4, atol=1e-5)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_sharded_training_equals_single_gpu():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    ret = mp.Manager().dict()
    mp.spawn(_shard_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)
