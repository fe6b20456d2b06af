"""k_rank_tc kernel time alone (CUDA events inside the library) on the C3 shape, tail side, both modes."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
shape = os.environ.get("SHAPE", "C3")
N, R, W, Q = (14541, 237, 400, 20466) if shape == "C3" else (125000, 1000, 512, 32768)
rng = np.random.default_rng(0)
eng = Engine(abi.COMPLEX, N, R, W, W, 12.0)
eng.set_kernel_timing(True)
tri = torch.from_numpy(np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int32)).cuda()
for mname, mode in (("bf16", abi.TC_BF16), ("exact", abi.TC_BF16X3_EXACT)):
    packed = eng.pack_entities_tc(mode)
    ms = []
    for i in range(8):
        eng.rank_counts_tc(tri, 0, mode=mode, packed=packed)
        torch.cuda.synchronize()
        ms.append(eng.last_kernel_ms(abi.TIMED_RANK_TC))
    k = float(np.median(ms[2:]))
    fl = 2.0 * Q * N * W
    print(os.environ.get("TAG", "base"), shape, mname, "k_rank_tc %.4f ms  algorithmic %.0f TFLOP/s  issued %.0f TFLOP/s" % (k, fl / k * 1e-9, fl * (3 if mname == "exact" else 1) / k * 1e-9))
