#!/usr/bin/env python
"""bench.py -- throughput of the B200 triple-scoring hot path on BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Headline workload (config.workload): BASELINE.json configs[4], the configuration the metric "... at 1/2/4/8 B200" is quoted
on and the one path that shards: ComplEx dim 256 (row width 512) over a synthetic 1 000 000-entity x 1 000-relation
table, filtered 1-N ranking of a FIXED set of 16 384 test triples (tail and head side: 32 768 queries) against every entity.
One "step" = one ranking pass: query prep, target scores, the tcgen05 ranking kernel with the CSR filter fused into its
epilogue, exact re-scoring of the undecided pairs -- in `exact` mode, validate()'s default, whose integer ranks are
identical to the fp32 CUDA-core kernel's and hence to the reference arithmetic -- for both sides, and at N > 1 ONE NCCL
all-reduce of the 2*Q int32 partial counts.  The entity rows are split over the N ranks, the query set is fixed:
scaling = "strong".  `value` = queries/s with inputs resident in HBM, per-step CUDA events, max over ranks.
`e2e` = the same ranking through the reference-facing call KGEkeras.utils.validate(model, test, N, bs, filtering_triples)
with HOST arrays (uploads, filter-CSR build, operand packing, both sides, read-back, the 15-key dict) in the timed region.

Secondary keys (same JSON line): `approx_bf16` (the single-plane tensor-core mode), `train` (BASELINE configs[1]: TransE-L1
dim 200 WN18RR shape, fused train step, every rank a replica), `eval_c2` / `eval_c3` (TransE CUDA-core and ComplEx tcgen05
ranking on the bundled shapes, rows sharded), `train_sharded_c5` (row-sharded table, NCCL all-to-all of rows and gradient
rows), `train_large_table` (C5 unsharded, N=1 only).  `--full` adds the remaining bundled shapes.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

C5 = dict(name="C5 ComplEx dim=256 (row width 512), synthetic N=1e6 x R=1e3: filtered 1-N ranking of 16384 test triples (tail+head), "
               "entity rows sharded over the GPUs", N=1_000_000, R=1000, W=512, triples=16384, filter_triples=200_000, B=16384, K=64)
C2 = dict(name="C2 TransE-L1 dim=200 WN18RR-shape (N=40943, R=11)", N=40943, R=11, d=200, gamma=12.0, K=10, margin=1.0, triples=3134)
C3 = dict(name="C3 ComplEx dim=200 FB15k-237-shape (N=14541, R=237)", N=14541, R=237, W=400, triples=20466)


def headline_config(args):
    """config of the headline line -- the reference arm prints the SAME keys and values."""
    return {"workload": C5["name"], "score": "ComplEx", "row_width": C5["W"], "entities": C5["N"], "relations": C5["R"],
            "queries_per_step": 2 * C5["triples"], "filter": "synthetic CSR, Poisson(5.4) tail / Poisson(23.5) head known-true ids per query",
            "rank_contract": "rank = 1 + #{c not filtered, c != target : score_c > score_target}, fp32 sequential-k scores",
            "l2": "inputs larger than L2 (packed table image 1-3 GB, fp32 table 2 GB vs 126 MB)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1590.0, source="B200_PROFILING.md fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the warm-up and the timed region run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.timed_samples, self.reasons, self.stop_flag = index, [], [], set(), False
        self.max_mhz, self.err, self.timed, self.h, self.nv = None, None, False, None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report it, never fake numbers
            self.err = repr(e)

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if ids and index < len(ids) and ids[index].strip().isdigit():
            return int(ids[index])
        return index

    def sample_once(self):
        nv, h = self.nv, self.h
        mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        (self.timed_samples if self.timed else self.samples).append(mhz)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for k, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(k)

    def run(self):
        if self.h is None:
            return
        try:
            while not self.stop_flag:
                self.sample_once()
                time.sleep(0.001)
        except Exception as e:
            self.err = repr(e)

    def summary(self):
        use = self.timed_samples or self.samples
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "error": self.err}
        return {"sm_mhz": float(np.median(use)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(use), "window": "timed region" if self.timed_samples else "warm-up (timed region shorter than one sample)"}


def ncu_traffic(key):
    """DRAM bytes per launch of a kernel from the committed ncu captures (profiles/r02_traffic.json, else r01), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                return json.load(fh)[key]["bytes"]
        except Exception:
            continue
    return None


def glorot(rng, rows, cols):
    lim = np.sqrt(6.0 / (rows + cols))
    return rng.uniform(-lim, lim, size=(rows, cols)).astype(np.float32)


def synth_filter_csr(rng, Q, N, lam):
    """Poisson(lam) known-true ids per query, sorted and unique inside a query (WN18RR tail mean 5.4, head 23.5)."""
    lens = rng.poisson(lam, Q)
    seg = np.repeat(np.arange(Q, dtype=np.int64), lens)
    key = np.unique(seg * N + rng.integers(0, N, seg.shape[0]))
    seg, ids = key // N, (key % N).astype(np.int32)
    offs = np.zeros(Q + 1, np.int64)
    np.cumsum(np.bincount(seg, minlength=Q), out=offs[1:])
    return offs, ids


def synth_triples(rng, n, N, R):
    return np.stack([rng.integers(0, N, n), rng.integers(0, R, n), rng.integers(0, N, n)], axis=1).astype(np.int32)


def c5_tables_host(seed=3):
    """The headline tables as host fp32 arrays (torch's CPU generator: 512 M uniform floats in about a second)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    limE, limR = float(np.sqrt(6.0 / (C5["N"] + C5["W"]))), float(np.sqrt(6.0 / (C5["R"] + C5["W"])))
    E = torch.empty(C5["N"], C5["W"]).uniform_(-limE, limE, generator=g)
    R = torch.empty(C5["R"], C5["W"]).uniform_(-limR, limR, generator=g)
    return E, R


def c5_queries():
    rng = np.random.default_rng(7)
    tri = synth_triples(rng, C5["triples"], C5["N"], C5["R"])
    return tri, synth_filter_csr(rng, C5["triples"], C5["N"], 5.4), synth_filter_csr(rng, C5["triples"], C5["N"], 23.5)


def rnd(x, n=6):
    return float(f"{x:.{n}g}") if isinstance(x, (float, np.floating)) else x


def run_reference(args, rank, world):
    """Reference arm: the CPU port of the reference's filtered evaluation (TensorFlow is not installable here) on the SAME
    workload, all host threads; a step is a bounded sample of the query set (the full set would take hours)."""
    if rank != 0:
        return
    from oracle import cpu_port, kge_oracle as ko
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    E, R = c5_tables_host()
    E, R = E.numpy(), R.numpy()
    tri, csr_t, csr_h = c5_queries()
    warm = max(1, min(args.warmup, 2))
    per_step = 8                       # test triples per step (16 queries): ~1-2 s of sgemm + ranking per step on 16 cores
    steps = max(1, min(args.steps, 20))
    times = []
    for i in range(warm + steps):
        lo = (i * per_step) % (C5["triples"] - per_step)
        sub_t = (csr_t[0][lo:lo + per_step + 1] - csr_t[0][lo], csr_t[1][csr_t[0][lo]:csr_t[0][lo + per_step]])
        sub_h = (csr_h[0][lo:lo + per_step + 1] - csr_h[0][lo], csr_h[1][csr_h[0][lo]:csr_h[0][lo + per_step]])
        r = cpu_port.time_eval_dot(ko.COMPLEX, E, R, tri[lo:lo + per_step], sub_t, sub_h, max_triples=per_step, chunk=per_step, budget_s=1e9)
        if i >= warm:
            times.append(r["seconds"])
        if sum(times) > 150.0:
            break
    n = len(times)
    qps = n * 2 * per_step / float(np.sum(times))
    line = {
        "impl": "reference", "metric": "filtered_eval_queries_per_s", "value": rnd(qps), "unit": "queries/s",
        "n_gpus": args.gpus, "steps": n, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": rnd(1e3 * float(np.mean(times))),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args),
        "cpu_baseline": {"value": rnd(qps), "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{n} steps of {2 * per_step} queries each (of the workload's {2 * C5['triples']}) against all 1e6 entities: NumPy/BLAS "
                                   "restatement of validate (utils.py:199-264), sgemm on all host threads; TensorFlow not installable"},
        "e2e": {"value": rnd(qps), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="positives per train step (secondary C2 train measurement)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-secondary", action="store_true", help="headline only")
    ap.add_argument("--full", action="store_true", help="also the remaining bundled shapes (C1, C3 train, C4 HolE, harness step)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import KGEkeras
    from KGEkeras import _abi as abi
    from KGEkeras import utils as ku
    from KGEkeras._engine import Engine

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warm = args.steps, max(args.warmup, 3)
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_passes(plan, eng, n_warm, n_steps, sampler=None):
        """n_steps ranking passes, each bracketed by CUDA events on the launching stream; barrier + synchronize on both sides
        of the region.  Returns (per-step ms list, launches in the region, kernel ms sums)."""
        for _ in range(n_warm):
            plan.run()
        if plan.overflowed():
            raise RuntimeError("worklist overflow during warm-up: enlarge worklist_cap")
        eng.set_kernel_timing(True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        if sampler is not None:
            sampler.timed = True
        l0 = eng.launches
        for i in range(n_steps):
            ev[i][0].record()
            plan.run()
            ev[i][1].record()
        barrier()
        if sampler is not None:
            sampler.timed = False
        launches = eng.launches - l0
        ms = [a.elapsed_time(b) for a, b in ev]
        ktc = eng.kernel_ms_sum(abi.TIMED_RANK_TC) if plan.use_tc else (0.0, 0)
        kref = eng.kernel_ms_sum(abi.TIMED_RANK_REFINE) if plan.use_tc else (0.0, 0)
        ktile = (0.0, 0) if plan.use_tc else eng.kernel_ms_sum(abi.TIMED_RANK_TILE)
        eng.set_kernel_timing(False)
        return ms, launches, dict(tc=ktc, refine=kref, tile=ktile)

    # ================= headline: C5 filtered 1-N ranking, rows sharded, fixed query set ===================================
    E_h, R_h = c5_tables_host()
    model = KGEkeras.ComplEx(e_dim=C5["W"] // 2, r_dim=C5["W"] // 2, num_entities=C5["N"], num_relations=C5["R"], dp=0.0,
                             init_entities=E_h.numpy(), init_relations=R_h.numpy())
    eng = model._engine_()
    if world > 1:
        eng.comm_init_from_torch()     # the partial counts are summed by the library's own NCCL binding (kgeb_allreduce_counts)
    tri, csr_t, csr_h = c5_queries()
    tri_d = torch.from_numpy(tri).to(dev)
    csr_t_d = tuple(torch.from_numpy(a).to(dev) for a in csr_t)
    csr_h_d = tuple(torch.from_numpy(a).to(dev) for a in csr_h)
    rows = ku.shard_rows(C5["N"], world, rank)
    Q2 = 2 * C5["triples"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    head = {}
    for mname, mode in (("exact", "exact"), ("approx_bf16", "bf16")):
        plan = eng.rank_plan(tri_d, csr_t_d, csr_h_d, row_range=rows, mode=mode)
        ms, launches, kms = timed_passes(plan, eng, warm, steps, sampler if mname == "exact" else None)
        total_ms = max_over_ranks(float(np.sum(ms)))
        und = plan.undecided()
        ktc_ms = kms["tc"][0] / max(1, kms["tc"][1])            # average duration of ONE k_rank_tc launch (one side)
        flops_launch = 2.0 * C5["triples"] * (rows[1] - rows[0]) * C5["W"]    # 2*Q*N*Kc, this rank's shard, one side
        planes = 3 if mode == "exact" else 1
        head[mname] = dict(value=steps * Q2 / (total_ms * 1e-3), ms_per_step=total_ms / steps, launches=launches,
                           k_rank_tc_ms=ktc_ms, refine_ms=kms["refine"][0] / max(1, kms["refine"][1]),
                           undecided_pairs=int(und[0] + und[1]), alg_tflops=flops_launch / (ktc_ms * 1e-3) / 1e12,
                           issued_tflops=planes * flops_launch / (ktc_ms * 1e-3) / 1e12, counts=plan.counts.clone())
        del plan
        torch.cuda.empty_cache()
    exact_counts = head["exact"].pop("counts")
    approx_counts = head["approx_bf16"].pop("counts")
    # how far the approximate mode's ranks sit from the exact ones: mean |rank_bf16 - rank_exact| / rank_exact over the queries
    rank_moved = float(((exact_counts - approx_counts).abs().double() / (exact_counts.double() + 1.0)).mean().item())

    # ---- e2e: the reference-facing call with HOST arrays (uploads, CSR build, packing, both sides, read-back) -------------
    frng = np.random.default_rng(17)
    filt_host = synth_triples(frng, C5["filter_triples"], C5["N"], C5["R"])
    key = lambda a: (a[:, 0].astype(np.int64) * C5["R"] + a[:, 1]) * C5["N"] + a[:, 2]
    filt_host = filt_host[~np.isin(key(filt_host), key(tri))]          # the reference crashes when a test triple is in its own filter
    tri_host = tri.copy()
    ku.validate(model, tri_host, C5["N"], 32, filt_host)
    barrier()
    n_e2e = 3
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        vres = ku.validate(model, tri_host, C5["N"], 32, filt_host)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
    e2e = {"value": rnd(Q2 / e2e_s), "unit": "queries/s", "h2d_bytes_per_step": int(tri_host.nbytes + filt_host.astype(np.int32).nbytes),
           "d2h_bytes_per_step": int(2 * C5["triples"] * 4 + 16), "ms_per_call": rnd(e2e_s * 1e3), "calls": n_e2e,
           "api": "KGEkeras.utils.validate(model, test_triples, N, bs, filtering_triples), host arrays, rank_mode exact; includes "
                  "kgeb_filter_csr_*, kgeb_rank_tc_pack_entities (3 GB image), both sides, D2H of the ranks, the 15-key dict",
           "mrr": rnd(float(vres["mrr"]))}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    del model, eng
    torch.cuda.empty_cache()

    sec = {}
    if not args.skip_secondary:
        for part in (sec_train_c2, sec_eval_c3, sec_train_sharded_c5, sec_train_large):
            try:       # a secondary measurement must never cost the headline line
                sec.update(part(args, rank, world, dev, barrier, max_over_ranks, timed_passes, pk))
            except Exception as ex:
                sec[part.__name__[4:]] = {"error": repr(ex)[:300]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hx, ha = head["exact"], head["approx_bf16"]
    line = {
        "metric": "filtered_eval_queries_per_s", "value": rnd(hx["value"]), "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": rnd(hx["ms_per_step"]), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16x3+f32", "data": "synthetic",
        "config": headline_config(args),
        "e2e": e2e,
        "gpu_launches": int(hx["launches"]),
        "kernel_ms": {"k_rank_tc (one side, this rank's shard)": rnd(hx["k_rank_tc_ms"]), "exact re-scoring k_seqk_pairs (one side)": rnd(hx["refine_ms"]),
                      "undecided_pairs_per_pass": hx["undecided_pairs"]},
        # the kernel runs tens of milliseconds back to back inside a long step: the SUSTAINED measured peak is its denominator
        "roofline": {"bound": "tensor", "kernel": "k_rank_tc (exact mode: 3 MMAs per k-step on the scaled hi/lo fp16 planes)",
                     "achieved": rnd(hx["alg_tflops"]), "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": rnd(hx["alg_tflops"] / pk["bf16_sustained"]),
                     "pipe_frac_issued_mmas": rnd(hx["issued_tflops"] / pk["bf16_sustained"]), "frac_of_burst_peak": rnd(hx["alg_tflops"] / pk["bf16_tflops"]),
                     "traffic": ncu_traffic("k_rank_tc_exact_c5"),
                     "peak_source": pk["source"] + ": sustained 16-bit dense rate (kernel timed inside a long step); burst = %g" % pk["bf16_tflops"],
                     "algorithmic": "2*Q*N*Kc flops per launch (SURVEY 8d), Q=16384, N=rows of this rank, Kc=512; the exact mode issues 3x that"},
        "approx_bf16": {"value": rnd(ha["value"]), "unit": "queries/s", "ms_per_step": rnd(ha["ms_per_step"]), "k_rank_tc_ms": rnd(ha["k_rank_tc_ms"]),
                        "roofline": {"bound": "tensor", "achieved": rnd(ha["alg_tflops"]), "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                                     "frac": rnd(ha["alg_tflops"] / pk["bf16_sustained"]), "frac_of_burst_peak": rnd(ha["alg_tflops"] / pk["bf16_tflops"]),
                                     "traffic": ncu_traffic("k_rank_tc_bf16_c5")},
                        "mean_relative_rank_difference_vs_exact": rnd(rank_moved)},
        "clocks": sampler.summary(),
    }
    line.update(sec)
    if not args.no_cpu_baseline and world == 1:
        from oracle import cpu_port, kge_oracle as ko
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        cb = cpu_port.time_eval_dot(ko.COMPLEX, E_h.numpy(), R_h.numpy(), tri, csr_t, csr_h, max_triples=64, chunk=8, budget_s=15.0)
        line["cpu_baseline"] = {"value": rnd(cb["queries_per_s"]), "unit": "queries/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"{cb['queries']} of the workload's {Q2} queries against all 1e6 entities, NumPy/BLAS sgemm restatement "
                                          "of validate on all host threads (TensorFlow not installable)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sec_train_c2(args, rank, world, dev, barrier, max_over_ranks, timed_passes, pk):
    """BASELINE configs[1]: the fused train step and the TransE filtered eval on the WN18RR shape."""
    import torch
    from KGEkeras import _abi as abi
    from KGEkeras import utils as ku
    from KGEkeras._engine import Engine
    out = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- C2 fused train step (replicas at N > 1: the table fits one GPU) --------------------------------------
    w = C2
    N, R, d, K, B = w["N"], w["R"], w["d"], w["K"], args.batch
    rng = np.random.default_rng(0)
    E0, R0 = glorot(rng, N, d), glorot(rng, R, d)
    eng = Engine(abi.TRANSE_L1, N, R, d, d, gamma=w["gamma"], E=E0, R=R0)
    if world > 1:
        eng.comm_init_from_torch()
    brng = np.random.default_rng(1000 + rank)
    tsteps, twarm = 50, 5
    nb = tsteps + twarm
    host_batches = torch.from_numpy(np.stack([brng.integers(0, N, (nb, B)), brng.integers(0, R, (nb, B)), brng.integers(0, N, (nb, B))], axis=2).astype(np.int32)).pin_memory()
    dev_batches = host_batches.to(dev)
    step_kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, pair_mode=abi.PAIRED, margin=w["margin"], corrupt_mode=abi.CORRUPT_BOTH, seed=1234 + rank)
    opt_kw = dict(opt_kind=abi.LAZY_ADAM, lr=1e-3)
    eng.forward_backward(dev_batches[0], None, offset=0, **step_kw, opt_kind=abi.LAZY_ADAM)
    U_E, U_R = int(eng.tE.sum().item()), int(eng.tR.sum().item())
    eng.apply_optimizer(**opt_kw)
    for i in range(twarm):
        flush.zero_()
        eng.train_step(dev_batches[i], None, offset=i, **step_kw, **opt_kw)
    eng.set_kernel_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(tsteps)]
    barrier()
    l0 = eng.launches
    for i in range(tsteps):
        flush.zero_()
        ev[i][0].record()
        eng.train_step(dev_batches[twarm + i], None, offset=twarm + i, **step_kw, **opt_kw)
        ev[i][1].record()
    barrier()
    launches = eng.launches - l0
    t_ms = max_over_ranks(float(np.sum([a.elapsed_time(b) for a, b in ev])))
    k_fb, k_op = eng.kernel_ms_sum(abi.TIMED_TRAIN), eng.kernel_ms_sum(abi.TIMED_OPT)
    eng.set_kernel_timing(False)
    fb_ms, op_ms = k_fb[0] / max(1, k_fb[1]), k_op[0] / max(1, k_op[1])
    e2e_t, n_e2e = 0.0, 30
    for i in range(3 + n_e2e):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch = host_batches[(twarm + i) % nb].to(dev, non_blocking=True)
        r = eng.train_step(batch, None, offset=10_000 + i, **step_kw, **opt_kw)
        r["loss"].item()
        if i >= 3:
            e2e_t += time.perf_counter() - t0
    e2e_t = max_over_ranks(e2e_t)
    bytes_fb = 8.0 * d * (U_E + U_R) + 12.0 * B
    bytes_step = 28.0 * d * (U_E + U_R) + 12.0 * B
    out["train"] = {"metric": "train_triples_per_s", "workload": w["name"], "scaling": "weak (replicas, no collective)",
                    "value": rnd(world * tsteps * B * (1 + K) / (t_ms * 1e-3)), "unit": "triples/s", "ms_per_step": rnd(t_ms / tsteps),
                    "batch_positives": B, "negatives_per_positive": K, "loss": "pairwize_hinge (paired)", "optimizer": "LazyAdam (touched rows)",
                    "gpu_launches_per_step": launches // tsteps, "kernel_ms": {"k_train_ring": rnd(fb_ms), "k_opt_update2": rnd(op_ms)},
                    "e2e": {"value": rnd(world * n_e2e * B * (1 + K) / e2e_t), "unit": "triples/s", "h2d_bytes_per_step": int(B * 12), "d2h_bytes_per_step": 4,
                            "api": "Engine.train_step(host int32 batch) -> loss.item()"},
                    "roofline": {"bound": "hbm", "kernel": "k_train_ring", "achieved": rnd(bytes_fb / (fb_ms * 1e-3) / 1e9), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": rnd(bytes_fb / (fb_ms * 1e-3) / 1e9 / pk["hbm_gbs"]), "traffic": ncu_traffic("k_train_ring"),
                                 "note": "L2-resident shape: bound by the L2 vector-reduction path (profiles/r01_v4_train_bound.md), not by HBM"},
                    "roofline_step": {"bound": "hbm", "achieved": rnd(bytes_step / ((fb_ms + op_ms) * 1e-3) / 1e9), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                      "frac": rnd(bytes_step / ((fb_ms + op_ms) * 1e-3) / 1e9 / pk["hbm_gbs"]), "unique_rows": [U_E, U_R]}}

    # ---- C2 TransE filtered eval (CUDA-core tile kernel), rows sharded ----------------------------------------------------
    erng = np.random.default_rng(7)
    nt = w["triples"]
    tri = torch.from_numpy(synth_triples(erng, nt, N, R)).to(dev)
    ct = tuple(torch.from_numpy(a).to(dev) for a in synth_filter_csr(erng, nt, N, 5.4))
    ch = tuple(torch.from_numpy(a).to(dev) for a in synth_filter_csr(erng, nt, N, 23.5))
    plan = eng.rank_plan(tri, ct, ch, row_range=ku.shard_rows(N, world, rank), mode="cuda").capture()
    ms, launches, kms = timed_passes(plan, eng, 3, 10)
    t_ms = max_over_ranks(float(np.sum(ms)))
    # per-kernel time of the two tile launches (tail, head) from an un-captured pass with the sides serial; the L1 distance needs 2 (tail) /
    # 3 (head) separately rounded fp32 adds per (query, entity, k): its roofline is the fp32 add pipe, SMs x 128 lanes x the maximum SM clock
    plan2 = eng.rank_plan(tri, ct, ch, row_range=ku.shard_rows(N, world, rank), mode="cuda")
    plan2.fork_sides = False
    _, _, kms2 = timed_passes(plan2, eng, 2, 5)
    tile_ms = kms2["tile"][0] / max(1, kms2["tile"][1])          # average of the tail and the head launch
    rows = ku.shard_rows(N, world, rank)
    lane_adds = 2.5 * nt * (rows[1] - rows[0]) * d                # per launch, averaged over the two sides (2 d and 3 d per pair)
    add_peak = eng.num_sms * 128 * 1.965e9
    out["eval_c2"] = {"workload": w["name"] + ": filtered 1-N ranking, fp32 CUDA-core tile kernel", "value": rnd(10 * 2 * nt / (t_ms * 1e-3)), "unit": "queries/s",
                      "queries_per_pass": 2 * nt, "ms_per_pass": rnd(t_ms / 10), "scaling": "strong", "cuda_graph": True,
                      "sides_as_parallel_graph_branches": bool(plan.fork_sides),
                      "roofline": {"bound": "fp32 add pipe", "kernel": "k_rank_tile3 (packed FADD2, TMA-fed; tail and head launches averaged)",
                                   "kernel_ms": rnd(tile_ms), "achieved": rnd(lane_adds / (tile_ms * 1e-3) / 1e12), "peak": rnd(add_peak / 1e12),
                                   "unit": "T lane-adds/s", "frac": rnd(lane_adds / (tile_ms * 1e-3) / add_peak),
                                   "peak_source": "SMs x 128 fp32 lanes x clocks.max.sm (1965 MHz); FADD2 occupies the pipe two cycles, so the lane rate is the ceiling"}}
    del plan2
    del plan, eng
    return out


def sec_eval_c3(args, rank, world, dev, barrier, max_over_ranks, timed_passes, pk):
    """BASELINE configs[2] shape: ComplEx filtered eval on tcgen05, rows sharded, the pass replayed from a CUDA graph."""
    import torch
    from KGEkeras import _abi as abi
    from KGEkeras import utils as ku
    from KGEkeras._engine import Engine
    out = {}
    c = C3
    trng = np.random.default_rng(11)
    teng = Engine(abi.COMPLEX, c["N"], c["R"], c["W"], c["W"], E=glorot(trng, c["N"], c["W"]), R=glorot(trng, c["R"], c["W"]))
    if world > 1:
        teng.comm_init_from_torch()
    ttri = torch.from_numpy(synth_triples(trng, c["triples"], c["N"], c["R"])).to(dev)
    tct = tuple(torch.from_numpy(a).to(dev) for a in synth_filter_csr(trng, c["triples"], c["N"], 5.4))
    tch = tuple(torch.from_numpy(a).to(dev) for a in synth_filter_csr(trng, c["triples"], c["N"], 23.5))
    trows = ku.shard_rows(c["N"], world, rank)
    ent = {"workload": c["name"] + ": filtered 1-N ranking on tcgen05", "queries_per_pass": 2 * c["triples"], "scaling": "strong"}
    for mname, mode in (("exact", "exact"), ("approx_bf16", "bf16")):
        plan = teng.rank_plan(ttri, tct, tch, row_range=trows, mode=mode)
        forked, plan.fork_sides = plan.fork_sides, False           # per-kernel times: the two sides one after the other, no overlap
        ms, launches, kms = timed_passes(plan, teng, 3, 10)        # un-captured: kernel timing needs plain launches
        ktc = kms["tc"][0] / max(1, kms["tc"][1])
        plan.fork_sides = forked                                   # the timed pass: sides as parallel graph branches on small shards
        plan.capture()
        gms, _, _ = timed_passes(plan, teng, 3, 10)
        t_ms = max_over_ranks(float(np.sum(gms)))
        flops = 2.0 * c["triples"] * (trows[1] - trows[0]) * c["W"]
        ent[mname] = {"value": rnd(10 * 2 * c["triples"] / (t_ms * 1e-3)), "unit": "queries/s", "ms_per_pass": rnd(t_ms / 10),
                      "ms_per_pass_without_graph": rnd(max_over_ranks(float(np.sum(ms))) / 10), "gpu_launches_per_pass": launches // 10,
                      "k_rank_tc_ms_one_side": rnd(ktc), "algorithmic_tflops": rnd(flops / (ktc * 1e-3) / 1e12),
                      "frac_of_bf16_peak": rnd(flops / (ktc * 1e-3) / 1e12 / pk["bf16_tflops"]), "undecided_pairs": int(sum(plan.undecided())),
                      "sides_as_parallel_graph_branches": bool(plan.fork_sides)}
        del plan
    out["eval_c3"] = ent
    del teng
    return out


def sec_train_sharded_c5(args, rank, world, dev, barrier, max_over_ranks, timed_passes, pk):
    """C5 row-sharded training: owner-sharded table + Adam state.  Two exchanges of rows and gradient rows are measured:
    `train_sharded_c5` = peer-mapped table, the fused kernel reads remote rows / reduces remote gradient rows over NVLink itself
    (KGEkeras/_peer.py); `train_sharded_c5_alltoall` = compact table filled and drained by two NCCL all-to-alls (_sharded.py)."""
    import torch
    from KGEkeras import _abi as abi
    out = {}
    comm_of = lambda: None
    if world > 1:      # both run over the library's own NCCL binding (kgeb_comm_*)
        import torch.distributed as dist
        from KGEkeras._engine import comm_unique_id

        def comm_of():
            box = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            return (box[0], rank, world)
    srng = np.random.default_rng(500 + rank)
    ssteps, swarm = 5, 3
    sb = torch.from_numpy(np.stack([srng.integers(0, C5["N"], (ssteps + swarm, C5["B"])), srng.integers(0, C5["R"], (ssteps + swarm, C5["B"])),
                                    srng.integers(0, C5["N"], (ssteps + swarm, C5["B"]))], axis=2).astype(np.int32)).to(dev)
    skw = dict(K=C5["K"], loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, seed=99, opt_kind=abi.LAZY_ADAM, lr=1e-3)
    for key in ("train_sharded_c5", "train_sharded_c5_local_negatives", "train_sharded_c5_alltoall"):
        if key == "train_sharded_c5_local_negatives" and world == 1:
            continue          # on one GPU the local range is the global one
        try:
            if key != "train_sharded_c5_alltoall":
                from KGEkeras._peer import PeerShardedTrainer
                st = PeerShardedTrainer(abi.COMPLEX, C5["N"], C5["R"], C5["W"], C5["W"], seed=5, comm=comm_of(), tag=key,
                                        negatives="global" if key == "train_sharded_c5" else "local")
            else:
                from KGEkeras._sharded import ShardedTrainer
                st = ShardedTrainer(abi.COMPLEX, C5["N"], C5["R"], C5["W"], C5["W"], seed=5, comm=comm_of())
            for i in range(swarm):
                st.step(sb[i], None, offset=i, **skw)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(ssteps):
                st.step(sb[swarm + i], None, offset=swarm + i, **skw)
            e1.record()
            barrier()
            s_ms = max_over_ranks(e0.elapsed_time(e1))
            out[key] = {"workload": "C5 ComplEx dim=256, N=1e6 rows + Adam state sharded by owner; %d x %d triples per rank per step" % (C5["B"], 1 + C5["K"]),
                        "value": rnd(world * ssteps * C5["B"] * (1 + C5["K"]) / (s_ms * 1e-3)), "unit": "triples/s", "ms_per_step": rnd(s_ms / ssteps),
                        "scaling": "weak", "exchange": st.exchange_kind}
            if key.endswith("alltoall"):
                out[key].update(unique_rows_per_rank=st.last.get("unique_rows"),
                                a2a_bytes_per_rank_per_step=int(st.last.get("a2a_bytes_fwd", 0) + st.last.get("a2a_bytes_bwd", 0)))
            else:
                refs = 2 * C5["B"] * (1 + C5["K"])
                if key == "train_sharded_c5":
                    out[key].update(rows_per_shard=st.per, negatives="uniform over all N entities (the reference's distribution)",
                                    nvlink_row_load_bytes_per_rank_per_step=int(refs * C5["W"] * 4 * (world - 1) / world))
                else:
                    out[key].update(rows_per_shard=st.per, negatives="drawn from the rank's own shard only: NOT the reference's sampling distribution (opt-in)",
                                    nvlink_row_load_bytes_per_rank_per_step=int(2 * C5["B"] * C5["W"] * 4 * (world - 1) / world))
                st.close()
            del st
        except Exception as ex:
            out[key] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    return out


def sec_train_large(args, rank, world, dev, barrier, max_over_ranks, timed_passes, pk):
    """C5 unsharded step on one GPU: the table does not fit the L2, so the HBM roofline is the real bound."""
    import torch
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    out = {}
    if world == 1:
        lN, lR, lW, lB, lK = C5["N"], C5["R"], C5["W"], C5["B"], C5["K"]
        for key, gmode in (("train_large_table", "red"), ("train_large_table_sortseg_fused", "sortseg_fused")):
            lrng = np.random.default_rng(21)
            lim = float(np.sqrt(6.0 / (lN + lW)))
            leng = Engine(abi.COMPLEX, lN, lR, lW, lW, E=torch.empty(lN, lW, device=dev).uniform_(-lim, lim), R=glorot(lrng, lR, lW))
            leng.set_grad_mode(gmode)
            lsteps, lwarm = 8, 3
            lb = torch.from_numpy(np.stack([lrng.integers(0, lN, (lsteps + lwarm, lB)), lrng.integers(0, lR, (lsteps + lwarm, lB)),
                                            lrng.integers(0, lN, (lsteps + lwarm, lB))], axis=2).astype(np.int32)).to(dev)
            lkw = dict(K=lK, loss_kind=abi.PAIRWISE_LOGISTIC, pair_mode=abi.PAIRED, margin=1.0, corrupt_mode=abi.CORRUPT_BOTH, seed=99)
            for i in range(lwarm):
                leng.train_step(lb[i], None, offset=i, **lkw, opt_kind=abi.LAZY_ADAM, lr=1e-3)
            lU_E, lU_R = int((leng.tE2 if gmode == "sortseg_fused" else leng.tE).sum().item()), lR
            if gmode == "red":      # flags are cleared by the update: count on a forward pass of the next batch
                leng.forward_backward(lb[lwarm], None, offset=lwarm, **lkw, opt_kind=abi.LAZY_ADAM)
                lU_E, lU_R = int(leng.tE.sum().item()), int(leng.tR.sum().item())
                leng.apply_optimizer(abi.LAZY_ADAM, lr=1e-3)
            leng.set_kernel_timing(True)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(lsteps)]
            torch.cuda.synchronize()
            for i in range(lsteps):
                ev[i][0].record()
                leng.train_step(lb[(lwarm + i) % (lsteps + lwarm)], None, offset=lwarm + 1 + i, **lkw, opt_kind=abi.LAZY_ADAM, lr=1e-3)
                ev[i][1].record()
            torch.cuda.synchronize()
            lms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            kf, ka, ko_ = leng.kernel_ms_sum(abi.TIMED_TRAIN), leng.kernel_ms_sum(abi.TIMED_AUX), leng.kernel_ms_sum(abi.TIMED_OPT)
            bstep = 28.0 * lW * (lU_E + lU_R) + 12.0 * lB
            out[key] = {"workload": "C5 step unsharded on one GPU (table 2 GB + Adam state 4 GB >> L2)", "backward": {
                            "red": "vector reductions into a dense gradient buffer + touched-row update (2 launches)",
                            "sortseg_fused": "records sorted by entity id, a warp per segment sums the row's gradient and applies LazyAdam there "
                                             "(kgeb_train_step_sortseg: no gradient buffer, two table copies)"}[gmode],
                        "batch_positives": lB, "negatives_per_positive": lK, "unique_rows": [lU_E, lU_R], "value": rnd(lB * (1 + lK) / (lms * 1e-3)),
                        "unit": "triples/s", "ms_per_step": rnd(lms),
                        "kernel_ms": {"k_train_ring": rnd(kf[0] / max(1, kf[1])), "records+sort+segments(+optimizer)": rnd(ka[0] / max(1, ka[1])),
                                      "k_opt_update": rnd(ko_[0] / max(1, ko_[1]))},
                        "roofline_step": {"bound": "hbm", "achieved": rnd(bstep / (lms * 1e-3) / 1e9), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                          "frac": rnd(bstep / (lms * 1e-3) / 1e9 / pk["hbm_gbs"]), "traffic": ncu_traffic("c5_step"),
                                          "algorithmic": "28 * W * (U_E + U_R) + 12 * B bytes per step (SURVEY 8d)"}}
            del leng, lb
            torch.cuda.empty_cache()
    if args.full and world == 1:
        out["other_shapes"] = other_shapes(dev, torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev))
    return out


def other_shapes(dev, flush):
    """--full: one train step on each remaining bundled shape (C1, C3, C4) and the reference harness' own configuration."""
    import torch
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    res = {}
    for sname, skind, sN, sR, sW, sB, sK, sloss, fft in (
            ("C1 DistMult dim=100 kinship-shape", abi.DISTMULT, 104, 25, 100, 2048, 10, abi.PAIRWISE_HINGE, False),
            ("C3 ComplEx dim=200 FB15k-237-shape", abi.COMPLEX, 14541, 237, 400, 16384, 10, abi.PAIRWISE_HINGE, False),
            ("C4 HolE dim=256 YAGO3-10-shape, Fourier-domain mode", abi.HOLE, 123182, 37, 256, 2048, 64, abi.POINTWISE_LOGISTIC, True)):
        srng = np.random.default_rng(31)
        seng = Engine(skind, sN, sR, sW, sW, E=glorot(srng, sN, sW), R=glorot(srng, sR, sW))
        if fft:
            seng.set_hole_mode(True)
        ssteps, swarm = 20, 3
        sb = torch.from_numpy(np.stack([srng.integers(0, sN, (ssteps + swarm, sB)), srng.integers(0, sR, (ssteps + swarm, sB)),
                                        srng.integers(0, sN, (ssteps + swarm, sB))], axis=2).astype(np.int32)).to(dev)
        skw = dict(K=sK, loss_kind=sloss, pair_mode=abi.PAIRED, margin=1.0, corrupt_mode=abi.CORRUPT_BOTH, seed=77, opt_kind=abi.LAZY_ADAM, lr=1e-3)
        for i in range(swarm):
            seng.train_step(sb[i], None, offset=i, **skw)
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ssteps)]
        for i in range(ssteps):
            flush.zero_()
            sev[i][0].record()
            seng.train_step(sb[swarm + i], None, offset=swarm + i, **skw)
            sev[i][1].record()
        torch.cuda.synchronize()
        sms = float(np.mean([a.elapsed_time(b) for a, b in sev]))
        res[sname] = {"batch_positives": sB, "negatives_per_positive": sK, "ms_per_step": rnd(sms), "triples_per_s": rnd(sB * (1 + sK) / (sms * 1e-3))}
        del seng, sb
        torch.cuda.empty_cache()
    try:
        import KGEkeras
        hrng = np.random.default_rng(41)
        hN, hR, hn = 14541, 237, 2048 * 24
        htrain = synth_triples(hrng, hn, hN, hR)
        hm = KGEkeras.DistMult(e_dim=100, r_dim=100, num_entities=hN, num_relations=hR, literals=hrng.random((hN, 5)), literal_activation="tanh",
                               loss_function="pairwize_hinge", margin=1, negative_samples=10, batch_size=2048)
        hm.pair_mode = "allpairs"
        hm.compile(optimizer="adam")
        hgen = KGEkeras.DataGenerator(htrain, ns=10, batch_size=2048, engine=hm._engine_())
        hm.fit(hgen, epochs=1, verbose=0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hm.fit(hgen, epochs=2, verbose=0)
        torch.cuda.synchronize()
        hdt = (time.perf_counter() - t0) / (2 * len(hgen))
        res["reference_harness_step"] = {"workload": "tests.py main(): DistMult dim=100, 5 literals + tanh Dense, dp=0.2, bs=2048, ns=10, all-pairs hinge, adam",
                                         "ms_per_batch": rnd(hdt * 1e3), "scored_triples_per_s": rnd(2 * 20480 / hdt)}
    except Exception as ex:
        res["reference_harness_step"] = {"error": repr(ex)[:300]}
    return res


if __name__ == "__main__":
    main()
