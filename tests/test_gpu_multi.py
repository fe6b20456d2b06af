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
