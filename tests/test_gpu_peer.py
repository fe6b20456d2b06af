"""Peer-mapped tables (kgeb_peer_*, KGEkeras/_peer.py): the row exchange of the sharded step inside the scoring kernel.

Bars: world 1 -- the step IS the single-GPU step (same kernels, same pointers' contents): loss and tables equal to Engine.train_step
within fp32 reduction-order noise (the vector reductions land in a different order); world >= 2 (skipped with one GPU) -- equal to
the single-GPU step over the concatenated batch, the same bar tests/test_gpu_multi.py holds the all-to-all trainer to."""
import os

import numpy as np
import pytest
import torch

from test_gpu_multi import _case, _spawn

pytestmark = pytest.mark.gpu


def test_peer_table_world1_is_plain_device_memory():
    from KGEkeras._peer import PeerTable
    t = PeerTable(1000, 512, 0, 1, "unit")
    assert t.rows_per_shard % 1024 == 0 and t.rows_per_shard >= 1000          # 2 MiB granules of 2 KiB rows
    assert t.local.shape == (t.rows_per_shard, 512) and t.local.data_ptr() == t.all.data_ptr()
    assert float(t.local.abs().sum().item()) == 0.0
    t.local[5, 7] = 3.0
    assert float(t.local.sum().item()) == 3.0
    t.close()


@pytest.mark.parametrize("kind_name", ["TRANSE_L1", "COMPLEX"])
def test_peer_trainer_world1_equals_engine_step(kind_name):
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    from KGEkeras._peer import PeerShardedTrainer
    kind = getattr(abi, kind_name)
    N, R, d, B, K = 5000, 7, 32, 256, 4
    E, Rm, pos, neg = _case(1, B, K)
    ref = Engine(kind, N, R, d, d, 12.0, E=E, R=Rm)
    st = PeerShardedTrainer(kind, N, R, d, d, 12.0, E=E, R=Rm, tag=kind_name, grad="pull" if kind_name == "COMPLEX" else "remote_red")
    for step in range(3):
        l0 = ref.train_step(pos, neg, opt_kind=abi.LAZY_ADAM, lr=0.01, K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0)["loss"].item()
        l1 = st.step(pos, neg, K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01).item()
        assert abs(l0 - l1) <= 1e-6 * max(1.0, abs(l0))
    np.testing.assert_allclose(st.gather_full_table(), ref.get_tables()[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.relation_table(), ref.get_tables()[1], rtol=1e-5, atol=1e-6)
    # in-kernel Philox negatives over the GLOBAL id range: same stream as the single-GPU engine
    l0 = ref.train_step(pos, None, opt_kind=abi.LAZY_ADAM, lr=0.01, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, seed=5, offset=3)["loss"].item()
    l1 = st.step(pos, None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01, seed=5, offset=3).item()
    assert abs(l0 - l1) <= 1e-6 * max(1.0, abs(l0))
    np.testing.assert_allclose(st.gather_full_table(), ref.get_tables()[0], rtol=1e-5, atol=1e-6)
    st.close()


def _peer_worker(rank, world, box, ret):
    import sys
    import time
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    for p in (root, os.path.join(root, "kge-keras_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine, comm_unique_id
    from KGEkeras._peer import PeerShardedTrainer
    torch.cuda.set_device(rank)
    os.environ["MASTER_PORT"] = str(box["port"])
    if rank == 0:
        box["id"] = comm_unique_id()
    while "id" not in box:
        time.sleep(0.01)
    N, R, d, B, K = 70000, 7, 32, 128, 4          # 128-byte rows: 16384 rows per granule -> shards of 49152 / 32768 ... rows
    E, Rm, pos, neg = _case(world, B, K, N=N)
    ok = True
    for grad in ("pull", "remote_red"):
      key = "id_" + grad
      if rank == 0:
          box[key] = comm_unique_id()
      while key not in box:
          time.sleep(0.01)
      st = PeerShardedTrainer(abi.COMPLEX, N, R, d, d, 12.0, E=E, R=Rm, comm=(box[key], rank, world), tag="w" + grad, grad=grad)
      ref = Engine(abi.COMPLEX, N, R, d, d, 12.0, E=E, R=Rm)           # every rank also runs the single-GPU truth
      for step in range(3):
        l1 = st.step(pos[rank * B:(rank + 1) * B], neg[rank * B * K:(rank + 1) * B * K], K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0,
                     opt_kind=abi.LAZY_ADAM, lr=0.01).item()
        l0 = ref.train_step(pos, neg, opt_kind=abi.LAZY_ADAM, lr=0.01, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0)["loss"].item()
        ok = ok and abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0))
      full = st.gather_full_table()
      ok = ok and np.allclose(full, ref.get_tables()[0], rtol=1e-4, atol=1e-5)
      ok = ok and np.allclose(st.relation_table(), ref.get_tables()[1], rtol=1e-4, atol=1e-5)
      ok = ok and st.per * world >= N and st.per % 16384 == 0
      st.close()
    ret[rank] = bool(ok)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_trainer_equals_single_gpu_step():
    import torch.multiprocessing as mp
    from test_gpu_multi import _free_port
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    box, ret = mgr.dict(), mgr.dict()
    box["port"] = _free_port()
    _spawn(_peer_worker, (world, box, ret), world, timeout=240)
    assert len(ret) == world and all(ret.values()), dict(ret)


def test_peer_trainer_local_negatives_world1_is_global():
    """On one GPU the rank's shard is the whole table: negatives='local' must draw the same ids as the in-kernel global stream."""
    from KGEkeras import _abi as abi
    from KGEkeras._peer import PeerShardedTrainer
    N, R, d, B, K = 5000, 7, 32, 256, 4
    E, Rm, pos, _ = _case(1, B, K)
    a = PeerShardedTrainer(abi.COMPLEX, N, R, d, d, 12.0, E=E, R=Rm, tag="la", negatives="local")
    b = PeerShardedTrainer(abi.COMPLEX, N, R, d, d, 12.0, E=E, R=Rm, tag="lb", negatives="global")
    for step in range(2):
        la = a.step(pos, None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01, seed=3, offset=step).item()
        lb = b.step(pos, None, K=K, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=0.01, seed=3, offset=step).item()
        assert abs(la - lb) <= 1e-6 * max(1.0, abs(lb))
    np.testing.assert_allclose(a.gather_full_table(), b.gather_full_table(), rtol=1e-5, atol=1e-6)
    a.close()
    b.close()
