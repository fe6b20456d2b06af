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


def _spawn(fn, args, nprocs, timeout=300):
    """mp.spawn with a deadline: a collective that never completes must fail the test, not occupy the GPUs until the box's own limit."""
    import time
    import torch.multiprocessing as mp
    ctx = mp.spawn(fn, args=args, nprocs=nprocs, join=False)
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > timeout:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"{fn.__name__}: the worker processes did not finish within {timeout} s (hung collective?)")


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
    _spawn(_worker, (world, _free_port(), ret), world)
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
    _spawn(_shard_worker, (world, _free_port(), ret), world)
    assert len(ret) == world and all(ret.values()), dict(ret)


# ---- the library's own NCCL binding (kgeb_comm_*, comm.cu): no torch.distributed anywhere in these workers --------------
def _comm_worker(rank, world, box, ret):
    import sys
    import time
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    for p in (root, os.path.join(root, "kge-keras_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ctypes as C
    from KGEkeras import _abi as abi, utils as ku
    from KGEkeras._engine import Engine, comm_unique_id
    from KGEkeras._sharded import ShardedTrainer
    from oracle import kge_oracle as ko
    torch.cuda.set_device(rank)
    if rank == 0:
        box["id"] = comm_unique_id()            # 128 bytes through any side channel (here: a Manager dict)
    while "id" not in box:
        time.sleep(0.01)
    uid = box["id"]
    ok = {}
    # (a) sharded 1-N ranking: partial counts of every rank's rows, summed by kgeb_allreduce_counts inside the captured pass
    N, R, d, Q = 3000, 5, 32, 200
    rng = np.random.default_rng(9)
    E, Rm = ko.glorot_uniform(rng, N, d) * 3, ko.glorot_uniform(rng, R, d) * 3
    tri = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int32)
    known = np.stack([rng.integers(0, N, 4 * Q), rng.integers(0, R, 4 * Q), rng.integers(0, N, 4 * Q)], 1).astype(np.int32)
    known[:Q, :2] = tri[:, :2]
    known[Q:2 * Q, 1:] = tri[:, 1:]
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in known])
    for kind, mode in ((abi.COMPLEX, "exact"), (abi.TRANSE_L1, "cuda")):
        eng = Engine(kind, N, R, d, d, 12.0, E=E, R=Rm)
        key = f"id_{kind}"
        if rank == 0:
            box[key] = comm_unique_id()
        while key not in box:
            time.sleep(0.01)
        eng.comm_init(box[key], rank, world)
        plan = eng.rank_plan(tri, ko.filter_csr(tri, "tail", ft), ko.filter_csr(tri, "head", fh), row_range=ku.shard_rows(N, world, rank), mode=mode)
        assert plan.kgeb_comm and not plan.dist_on
        tr, hr = plan.capture().ranks()
        tr2, hr2 = plan.ranks()                  # replay of the graph (all-reduce inside) gives the same ranks again
        for side, got, got2, f in (("tail", tr, tr2, ft), ("head", hr, hr2, fh)):
            sc = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, tri, side), E, side, 12.0)
            want = ko.ranks_from_scores(sc, tri, side, f)
            ok[f"rank_{kind}_{side}"] = bool(np.array_equal(got, want) and np.array_equal(got2, want))
        eng.comm_destroy()       # drops the plan's captured pass first: NCCL waits for graphs that hold its collectives
    # (b) kgeb_alltoall_rows with ragged counts (rank r sends r + g + 1 rows to rank g), kgeb_allreduce_f32 / _flags
    eng = Engine(abi.DISTMULT, 16, 2, 8, 8, 12.0)
    eng.comm_init(uid, rank, world)
    W = 12
    send_rows = [rank + g + 1 for g in range(world)]
    recv_rows = [g + rank + 1 for g in range(world)]
    send = torch.cat([torch.full((send_rows[g], W), float(100 * rank + g), device="cuda") for g in range(world)])
    recv = torch.full((sum(recv_rows), W), -1.0, device="cuda")
    snd, rcv = (C.c_int64 * world)(*send_rows), (C.c_int64 * world)(*recv_rows)
    eng._check(eng.lib.kgeb_alltoall_rows(eng.ctx, C.c_void_p(send.data_ptr()), snd, C.c_void_p(recv.data_ptr()), rcv, W,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    want = torch.cat([torch.full((recv_rows[g], W), float(100 * g + rank)) for g in range(world)])
    ok["alltoall"] = bool(torch.equal(recv.cpu(), want))
    f32 = torch.full((1000,), float(rank + 1), device="cuda")
    flg = torch.zeros(world + 3, dtype=torch.uint8, device="cuda")
    flg[rank] = 1
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    eng._check(eng.lib.kgeb_allreduce_f32(eng.ctx, C.c_void_p(f32.data_ptr()), 1000, st))
    eng._check(eng.lib.kgeb_allreduce_flags(eng.ctx, C.c_void_p(flg.data_ptr()), world + 3, st))
    ok["allreduce_f32"] = bool((f32 == world * (world + 1) / 2).all().item())
    ok["allreduce_flags"] = flg.cpu().tolist() == [1] * world + [0] * 3
    r_, w_ = C.c_int(), C.c_int()
    eng._check(eng.lib.kgeb_comm_info(eng.ctx, C.byref(r_), C.byref(w_)))
    ok["info"] = (r_.value, w_.value) == (rank, world)
    eng.comm_destroy()
    # (c) the row-sharded training step over the same binding equals the single-GPU step on the concatenated batch
    N, R, d, B, K = 5000, 7, 32, 128, 4
    Et, Rt, pos, neg = _case(world, B, K)
    if rank == 0:
        box["id_train"] = comm_unique_id()
    while "id_train" not in box:
        time.sleep(0.01)
    stn = ShardedTrainer(abi.COMPLEX, N, R, d, d, 12.0, E=Et, R=Rt, comm=(box["id_train"], rank, world))
    ref = Engine(abi.COMPLEX, N, R, d, d, 12.0, E=Et, R=Rt)
    good = True
    for step in range(3):
        l1 = stn.step(pos[rank * B:(rank + 1) * B], neg[rank * B * K:(rank + 1) * B * K], loss_kind=abi.PAIRWISE_HINGE, margin=1.0,
                      opt_kind=abi.LAZY_ADAM, lr=0.01).item()
        l0 = ref.train_step(pos, neg, opt_kind=abi.LAZY_ADAM, lr=0.01, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0)["loss"].item()
        good = good and abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0))
    good = good and np.allclose(stn.gather_full_table(), ref.get_tables()[0], rtol=1e-4, atol=1e-5)
    good = good and np.allclose(stn.shard.get_tables()[1], ref.get_tables()[1], rtol=1e-4, atol=1e-5)
    ok["sharded_train"] = bool(good) and "libkgeb200" in stn.exchange_kind
    ret[rank] = ok


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_c_abi_collectives_without_torch_distributed():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    box, ret = mgr.dict(), mgr.dict()
    _spawn(_comm_worker, (world, box, ret), world)
    assert len(ret) == world
    for r in range(world):
        assert all(ret[r].values()), (r, dict(ret[r]))


def test_c_abi_comm_refuses_calls_without_a_communicator():
    import ctypes as C
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine, comm_unique_id
    assert abi.lib().kgeb_comm_version() >= 20000
    assert len(comm_unique_id()) == abi.COMM_ID_BYTES
    eng = Engine(abi.DISTMULT, 16, 2, 8, 8, 12.0)
    t = torch.zeros(4, dtype=torch.int32, device="cuda")
    with pytest.raises(abi.KgebError, match="kgeb_comm_init first"):
        eng.allreduce_counts(t)
    # world 1 is a valid communicator: the sum of one rank's counts is itself
    eng.comm_init(comm_unique_id(), 0, 1)
    t += 3
    assert eng.allreduce_counts(t).cpu().tolist() == [3, 3, 3, 3]
    eng.comm_destroy()
