"""k_rank_tc_t<1> (one CTA per 128 x 256 tile) against k_rank_tc_t<2> (cta_group::2, KGEB_TC_2CTA=1: 256 x 256 per CTA pair) on the C3 and C5
shapes, both modes.  Run once per variant (the choice is read once per process).  Per-kernel device time from the library's event timers,
sides serial; counts checksums must agree between the variants."""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kge-keras_b200")]
from KGEkeras import _abi as abi
from KGEkeras._engine import Engine
out = {"pair": os.environ.get("KGEB_TC_2CTA", "0")}
shapes = [("C3", 14541, 237, 400, 20466), ("C5", 1_000_000, 1000, 512, 16384)]
if os.environ.get("KGEB_EXP_SMALL"):
    shapes = shapes[:1]
for name, N, R, W, T in shapes:
    g = torch.Generator(device="cuda").manual_seed(5)
    lim = float(np.sqrt(6.0 / (N + W)))
    E = (torch.rand(N, W, generator=g, device="cuda") * 2 - 1) * lim
    Rm = (torch.rand(R, W, generator=g, device="cuda") * 2 - 1) * lim
    eng = Engine(abi.COMPLEX, N, R, W, W, 12.0, E=E, R=Rm)
    rng = np.random.default_rng(1)
    tri = torch.from_numpy(np.stack([rng.integers(0, N, T), rng.integers(0, R, T), rng.integers(0, N, T)], 1).astype(np.int32)).cuda()
    for mode in ("bf16", "exact"):
        plan = eng.rank_plan(tri, None, None, mode=mode)
        plan.fork_sides = False
        for _ in range(2): plan.run()
        torch.cuda.synchronize()
        eng.set_kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): c = plan.run()
        e1.record(); torch.cuda.synchronize()
        ms, cnt = eng.kernel_ms_sum(abi.TIMED_RANK_TC)
        eng.set_kernel_timing(False)
        flops = 2.0 * T * N * W
        out[f"{name} {mode}"] = {"k_rank_tc_ms": round(ms / cnt, 4), "algorithmic_tflops": round(flops / (ms / cnt * 1e-3) / 1e12, 1),
                                 "pass_ms": round(e0.elapsed_time(e1) / n, 3), "undecided": int(sum(plan.undecided())), "checksum": int(c.long().sum().item())}
        del plan
    del eng, E, Rm
print(json.dumps(out))
