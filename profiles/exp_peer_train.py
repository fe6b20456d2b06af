#!/usr/bin/env python
"""C5 row-sharded training step under torchrun: the peer-mapped trainer's two gradient modes (NVLink vector reductions / owner pull)
with the reference's global negatives.  One JSON line per mode from rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/exp_peer_train.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._engine import comm_unique_id  # noqa: E402
from KGEkeras._peer import PeerShardedTrainer  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
C5 = bench.C5
srng = np.random.default_rng(500 + rank)
steps, warm = 5, 3
sb = torch.from_numpy(np.stack([srng.integers(0, C5["N"], (steps + warm, C5["B"])), srng.integers(0, C5["R"], (steps + warm, C5["B"])),
                                srng.integers(0, C5["N"], (steps + warm, C5["B"]))], axis=2).astype(np.int32)).to(dev)
kw = dict(K=C5["K"], loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, seed=99, opt_kind=abi.LAZY_ADAM, lr=1e-3)
for grad in ("remote_red", "pull"):
    comm = None
    if world > 1:
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = (box[0], rank, world)
    st = PeerShardedTrainer(abi.COMPLEX, C5["N"], C5["R"], C5["W"], C5["W"], seed=5, comm=comm, tag="x" + grad, grad=grad)
    for i in range(warm):
        st.step(sb[i], None, offset=i, **kw)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        st.step(sb[warm + i], None, offset=warm + i, **kw)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        print(json.dumps({"world": world, "grad": grad, "ms_per_step": round(ms, 4), "triples_per_s": round(world * C5["B"] * (1 + C5["K"]) / (ms * 1e-3))}), flush=True)
    st.close()
    del st
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
