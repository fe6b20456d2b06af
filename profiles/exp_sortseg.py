"""Step time of the fused train step with the vector-reduction backward ('red') against the sort-segmented one ('sortseg') and the
sort-segmented one with the optimizer applied at segment end ('sortseg_fused': kgeb_train_step_sortseg, no gradient buffer, two
table copies) on the C2 shape (L2-resident table) and the C5 shape (2 GB table), with the kernels' own CUDA-event times
(in the fused mode the segment kernel's time includes the optimizer, and k_opt_update is the relation table's alone).
    python profiles/exp_sortseg.py > profiles/r02_sortseg.jsonl"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [("C2 TransE-L1 d=200 N=40943 B=16384 K=10", abi.TRANSE_L1, 40943, 11, 200, 16384, 10, abi.PAIRWISE_HINGE, abi.CORRUPT_BOTH),
          ("C2 shape, one-sided corruption", abi.TRANSE_L1, 40943, 11, 200, 16384, 10, abi.PAIRWISE_HINGE, abi.CORRUPT_EITHER),
          ("C3 ComplEx d=200 N=14541 B=16384 K=10", abi.COMPLEX, 14541, 237, 400, 16384, 10, abi.PAIRWISE_HINGE, abi.CORRUPT_BOTH),
          ("C5 ComplEx d=256 N=1e6 B=16384 K=64", abi.COMPLEX, 1_000_000, 1000, 512, 16384, 64, abi.PAIRWISE_LOGISTIC, abi.CORRUPT_BOTH)]
for name, kind, N, R, W, B, K, loss, cm in SHAPES:
    rng = np.random.default_rng(0)
    lim = float(np.sqrt(6.0 / (N + W)))
    for mode in ("red", "sortseg", "sortseg_fused"):
        eng = Engine(kind, N, R, W, W, E=torch.empty(N, W, device=dev).uniform_(-lim, lim), R=rng.uniform(-0.1, 0.1, (R, W)).astype(np.float32))
        eng.set_grad_mode(mode)
        steps, warm = 10, 3
        batches = torch.from_numpy(np.stack([rng.integers(0, N, (steps + warm, B)), rng.integers(0, R, (steps + warm, B)),
                                             rng.integers(0, N, (steps + warm, B))], axis=2).astype(np.int32)).to(dev)
        kw = dict(K=K, loss_kind=loss, margin=1.0, corrupt_mode=cm, seed=5, opt_kind=abi.LAZY_ADAM, lr=1e-3)
        for i in range(warm):
            eng.train_step(batches[i], None, offset=i, **kw)
        eng.set_kernel_timing(True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            eng.train_step(batches[warm + i], None, offset=warm + i, **kw)
            ev[i][1].record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        kt, ka, ko_ = eng.kernel_ms_sum(abi.TIMED_TRAIN), eng.kernel_ms_sum(abi.TIMED_AUX), eng.kernel_ms_sum(abi.TIMED_OPT)
        print(json.dumps({"shape": name, "grad_mode": mode, "ms_per_step": round(ms, 4), "triples_per_s": round(B * (1 + K) / (ms * 1e-3)),
                          "k_train_ring_ms": round(kt[0] / max(1, kt[1]), 4), "records+sort+segments_ms": round(ka[0] / max(1, ka[1]), 4),
                          "k_opt_update2_ms": round(ko_[0] / max(1, ko_[1]), 4)}), flush=True)
        del eng, batches
        torch.cuda.empty_cache()
