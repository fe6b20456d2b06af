#!/usr/bin/env python
"""The three GEMMs of the literal step at the harness shape (2m = 40 960 rows, kin = 105, we = 100), tensor-core path against the
CUDA-core FFMA kernel (KGEB_GEMM=ffma in the environment selects the latter).  One JSON line per GEMM."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
    sys.path.insert(0, p)
from KGEkeras import _abi as abi  # noqa: E402
from KGEkeras._engine import Engine, _ptr, _stream  # noqa: E402

eng = Engine(abi.DISTMULT, 16, 2, 8, 8, 12.0)
lib = eng.lib
M2, kin, we, ld = 40960, 105, 100, 100
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.normal(size=(M2, kin)).astype(np.float32)).cuda()
W = torch.from_numpy(rng.normal(size=(kin, ld)).astype(np.float32)).cuda()
b = torch.zeros(ld, device="cuda")
Y = torch.empty(M2, ld, device="cuda")
gZ = torch.from_numpy(rng.normal(size=(M2, ld)).astype(np.float32)).cuda()
gW = torch.zeros(kin, ld, device="cuda")
gX = torch.empty(M2, kin, device="cuda")


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


cases = {
    "fwd  Y = tanh(X W + b)   [40960x105]x[105x100]": lambda: eng._check(lib.kgeb_dense_fwd(eng.ctx, _ptr(X), kin, M2, kin, _ptr(W), ld, we, _ptr(b), 1, _ptr(Y), ld, _stream())),
    "gW += X^T gZ             [105x40960]x[40960x100]": lambda: eng._check(lib.kgeb_sgemm(eng.ctx, 1, 0, kin, we, M2, _ptr(X), kin, _ptr(gZ), ld, _ptr(gW), ld, 1, _stream())),
    "gX = gZ W^T              [40960x100]x[100x100]": lambda: eng._check(lib.kgeb_sgemm(eng.ctx, 0, 1, M2, we, we, _ptr(gZ), ld, _ptr(W), ld, _ptr(gX), kin, 0, _stream())),
}
mode = "ffma" if os.environ.get("KGEB_GEMM") == "ffma" else "tcgen05 tf32x3"
for name, fn in cases.items():
    us = timeit(fn)
    print(json.dumps({"gemm": name, "path": mode, "us": round(us, 2)}), flush=True)
