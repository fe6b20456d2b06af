"""Micro-benchmark of the 1-N ranking kernels on one B200 (development aid; bench.py carries the judged numbers).

  python profiles/tc_bench.py [--shape C3|C4|C5s|C1] [--mode exact|bf16|cuda] [--iters 5] [--Q n]

Times the whole kgeb_rank_1n_tc call (query packing, tcgen05 kernel, exact re-scoring, filter correction) with CUDA
events; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._engine import Engine  # noqa: E402

SHAPES = {  # kind, N, R, d (ComplEx stores 2d), default Q
    "C1": (abi.DISTMULT, 104, 25, 100, 2148),
    "C3": (abi.COMPLEX, 14541, 237, 200, 40932),
    "C4": (abi.HOLE, 123182, 37, 256, 10000),
    "C5s": (abi.COMPLEX, 1000000, 1000, 256, 8192),
    "D3": (abi.DISTMULT, 14541, 237, 200, 40932),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="C3")
    ap.add_argument("--mode", default="exact", choices=["exact", "bf16", "cuda"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--Q", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    kind, N, R, d, Q = SHAPES[args.shape]
    Q = args.Q or Q
    W = 2 * d if kind == abi.COMPLEX else d
    rng = np.random.default_rng(0)
    lim_e, lim_r = np.sqrt(6.0 / (N + W)), np.sqrt(6.0 / (R + W))
    E = (rng.uniform(-lim_e, lim_e, (N, W)) * args.scale).astype(np.float32)
    Rm = (rng.uniform(-lim_r, lim_r, (R, W)) * args.scale).astype(np.float32)
    eng = Engine(kind, N, R, W, W, 12.0, E=E, R=Rm)
    tri = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int32)
    tri_d = torch.from_numpy(tri).cuda()
    mode = {"exact": abi.TC_BF16X3_EXACT, "bf16": abi.TC_BF16}.get(args.mode)
    stats = {}
    packed = eng.pack_entities_tc(mode) if mode is not None else None
    times = []
    for i in range(args.iters + 2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode is None:
            c = eng.rank_counts(tri_d, abi.SIDE_TAIL)
        else:
            c = eng.rank_counts_tc(tri_d, abi.SIDE_TAIL, mode=mode, packed=packed, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    flops = 2.0 * Q * N * W
    print(json.dumps({"shape": args.shape, "mode": args.mode, "Q": Q, "N": N, "W": W, "ms": ms, "queries_per_s": Q / ms * 1e3,
                      "algorithmic_tflops": flops / ms / 1e9, "executed_tflops": flops * (3 if args.mode == "exact" else 1) / ms / 1e9,
                      "mean_count": float(c.float().mean().item()), **stats}))


if __name__ == "__main__":
    main()
