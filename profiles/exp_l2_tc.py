#!/usr/bin/env python
"""TransE-L2 filtered 1-N ranking on the C2 shape (WN18RR: N = 40 943, d = 200, 3 134 test triples, both sides): the tcgen05 path
(exact mode: integer counts identical to the CUDA-core kernel) against the fp32 CUDA-core tile kernel.  One JSON line per mode."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._engine import Engine  # noqa: E402

N, R, d, nt = 40943, 11, 200, 3134
rng = np.random.default_rng(0)
eng = Engine(abi.TRANSE_L2, N, R, d, d, gamma=12.0, E=bench.glorot(rng, N, d), R=bench.glorot(rng, R, d))
erng = np.random.default_rng(7)
tri = torch.from_numpy(bench.synth_triples(erng, nt, N, R)).cuda()
ct = tuple(torch.from_numpy(a).cuda() for a in bench.synth_filter_csr(erng, nt, N, 5.4))
ch = tuple(torch.from_numpy(a).cuda() for a in bench.synth_filter_csr(erng, nt, N, 23.5))
ref = None
for mode in ("cuda", "exact"):
    plan = eng.rank_plan(tri, ct, ch, mode=mode).capture()
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        a.record()
        plan.run()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    counts = plan.run().clone()
    if ref is None:
        ref = counts
    print(json.dumps({"shape": "C2 shape, TransE-L2 d=200, N=40943, 6268 queries", "mode": mode, "tensor_cores": bool(plan.use_tc),
                      "ms_per_pass": round(ms, 4), "queries_per_s": round(2 * nt / (ms * 1e-3)), "undecided_pairs": int(sum(plan.undecided())),
                      "counts_equal_cuda_core": bool(torch.equal(counts, ref))}), flush=True)
