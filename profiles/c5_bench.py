"""BASELINE configs[4]: synthetic 1M-entity x 1k-relation ComplEx dim=256 (stored width 512), entity table row-sharded over the
ranks: sharded train step (rows and gradient rows exchanged by NCCL all-to-all) + 1-N filtered ranking on tcgen05 with the
integer counts all-reduced.  One process per GPU:

  python profiles/c5_bench.py                      (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/c5_bench.py

Timing: CUDA events per step, max over ranks, after warm-up; the 2 GB table does not fit the 126 MB L2.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._sharded import ShardedTrainer  # noqa: E402
from KGEkeras.utils import shard_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--R", type=int, default=1000)
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--batch", type=int, default=16384, help="positives per rank per step")
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--queries", type=int, default=8192)
    ap.add_argument("--eval-mode", default="bf16", choices=["bf16", "exact"])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, R, W, B, K = args.N, args.R, 2 * args.d, args.batch, args.K
    st = ShardedTrainer(abi.COMPLEX, N, R, W, W, 12.0, seed=0)
    rng = np.random.default_rng(100 + rank)
    nb = args.steps + args.warmup
    batches = torch.from_numpy(np.stack([rng.integers(0, N, (nb, B)), rng.integers(0, R, (nb, B)), rng.integers(0, N, (nb, B))], 2).astype(np.int32)).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kw = dict(K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, opt_kind=abi.LAZY_ADAM, lr=1e-3, seed=1)
    for i in range(args.warmup):
        st.step(batches[i], offset=i, **kw)
    barrier()
    ms = []
    for i in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.step(batches[args.warmup + i], offset=args.warmup + i, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.sum(ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    train_tps = world * args.steps * B * (1 + K) / (float(t.item()) * 1e-3)
    last = dict(st.last)

    # 1-N ranking: every rank ranks all queries against its row shard on the tensor cores, counts all-reduced
    eng = st.shard
    mode = abi.TC_BF16 if args.eval_mode == "bf16" else abi.TC_BF16X3_EXACT
    Q = args.queries
    r0, r1 = shard_rows(N, world, rank)
    erng = np.random.default_rng(7)
    tri = np.stack([erng.integers(0, N, Q), erng.integers(0, R, Q), erng.integers(0, N, Q)], 1).astype(np.int32)
    # query vectors need the source entity rows: gather them from their owners once (replicated small table of Q rows)
    src = torch.from_numpy(tri[:, 0].astype(np.int64)).to(dev)
    own = (src >= r0) & (src < r1)
    rows = torch.zeros(Q, eng.we_k, dtype=torch.float32, device=dev)
    rows[own] = eng.E[(src[own] - r0)]
    if world > 1:
        dist.all_reduce(rows)
    # a small replicated engine holding just the Q source rows + relations builds qvec/tscore through the public kernels
    from KGEkeras._engine import Engine
    qeng = Engine(abi.COMPLEX, Q, R, W, W, 12.0, E=rows.cpu().numpy(), R=eng.Rel[:, :W].cpu().numpy())
    qtri = torch.from_numpy(np.stack([np.arange(Q), tri[:, 1], np.zeros(Q)], 1).astype(np.int32)).to(dev)
    qw = W
    qvec = torch.zeros(Q, qw, dtype=torch.float32, device=dev)
    import ctypes as C
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    abi.check(qeng.lib.kgeb_query_prep(qeng.ctx, C.c_void_p(qeng.E.data_ptr()), 0, Q, C.c_void_p(qeng.Rel.data_ptr()), C.c_void_p(qtri.data_ptr()), Q, abi.SIDE_TAIL, C.c_void_p(qvec.data_ptr()), stream), qeng.ctx)
    # target scores: the owner of the target row computes them, all-reduce fills in the rest
    tri_d = torch.from_numpy(tri).to(dev)
    tscore = torch.zeros(Q, dtype=torch.float32, device=dev)
    abi.check(eng.lib.kgeb_target_scores(eng.ctx, C.c_void_p(eng.E.data_ptr()), r0, r1 - r0, C.c_void_p(qvec.data_ptr()), C.c_void_p(tri_d.data_ptr()), Q, abi.SIDE_TAIL, C.c_void_p(tscore.data_ptr()), stream), eng.ctx)
    if world > 1:
        dist.all_reduce(tscore)
    nbytes = int(eng.lib.kgeb_rank_tc_packed_bytes(eng.ctx, r1 - r0, mode))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    abi.check(eng.lib.kgeb_rank_tc_pack_entities(eng.ctx, C.c_void_p(eng.E.data_ptr()), r1 - r0, mode, C.c_void_p(packed.data_ptr()), nbytes, stream), eng.ctx)
    cap = max(1 << 20, Q * (64 + (r1 - r0) // 128))
    wsb = int(eng.lib.kgeb_rank_tc_workspace_bytes(eng.ctx, Q, mode, cap))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    eng.set_kernel_timing(True)
    ems, kms = [], []
    for i in range(2 + 5):
        counts = torch.zeros(Q, dtype=torch.int32, device=dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        abi.check(eng.lib.kgeb_rank_1n_tc(eng.ctx, C.c_void_p(eng.E.data_ptr()), r0, r1 - r0, C.c_void_p(packed.data_ptr()), C.c_void_p(qvec.data_ptr()),
                                          C.c_void_p(tscore.data_ptr()), C.c_void_p(tri_d.data_ptr()), Q, abi.SIDE_TAIL, None, None, mode,
                                          C.c_void_p(ws.data_ptr()), wsb, cap, C.c_void_p(counts.data_ptr()), stream), eng.ctx)
        if world > 1:
            dist.all_reduce(counts)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ems.append(e0.elapsed_time(e1)); kms.append(eng.last_kernel_ms(abi.TIMED_RANK_TC))
    hdr = ws[:8].view(torch.int32).cpu().tolist()
    te = torch.tensor([float(np.sum(ems))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    qps = 5 * Q / (float(te.item()) * 1e-3)
    flops = 2.0 * Q * (r1 - r0) * W
    if rank == 0:
        print(json.dumps({"workload": f"C5 ComplEx dim={args.d} N={N} R={R}, row-sharded over {world} GPU(s)", "n_gpus": world,
                          "train": {"triples_per_s": train_tps, "ms_per_step": float(t.item()) / args.steps, "positives_per_rank": B, "K": K,
                                    "scaling": "weak", **last},
                          "eval": {"mode": args.eval_mode, "queries_per_s": qps, "ms_per_pass": float(np.mean(ems)), "queries": Q, "scaling": "strong",
                                   "k_rank_tc_ms": float(np.mean(kms)), "k_rank_tc_algorithmic_tflops_per_gpu": flops / (float(np.mean(kms)) * 1e-3) / 1e12,
                                   "undecided_pairs": hdr[0] & 0xFFFFFFFF, "mean_count": float(counts.float().mean().item())}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
