#!/usr/bin/env python
"""bench.py -- throughput of the B200 triple-scoring hot path on BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Workload (config.workload): BASELINE.json configs[1] -- TransE-L1, dim 200, WN18RR shape (40 943 entities, 11 relations),
synthetic Glorot tables and uniform synthetic triples.  One "step" = one fused train step over B positives
with K in-kernel Philox negatives each: gather -> score -> pairwise hinge -> gradient scatter -> LazyAdam on touched rows.
`value` = train triples/s incl. negatives, B*(1+K)/t_step, inputs resident in HBM, CUDA-event timed per step with an L2
flush between steps (the 33 MB table + Adam state fit the 126 MB L2, so without the flush every step would be L2-hot).
`e2e` = the same step through the public API with HOST index buffers (pinned H2D of the batch, D2H of the loss) in the
timed region.  `eval` = filtered 1-N ranking queries/s (tail+head), entity rows sharded over the ranks, integer partial
counts all-reduced over NCCL.  Training on this shape is "replicas only" at N > 1 (tables fit one GPU).
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

TENSOR_EVAL = dict(name="C3 ComplEx dim=200 FB15k-237-shape (N=14541, R=237), 1-N filtered ranking on tcgen05", N=14541, R=237, d=200,
                   test_triples=20466)
WORKLOAD = dict(name="C2 TransE-L1 dim=200 WN18RR-shape (N=40943, R=11)", N=40943, R=11, d=200, gamma=12.0, K=10,
                margin=1.0, test_triples=3134)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]), source="MEASURED_PEAKS.json (measured)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="B200_PROFILING.md fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the warm-up and the timed region run.  NVML is opened
    in the constructor (it takes longer than a short timed region); samples taken while `timed` is set are the ones
    reported, the warm-up samples are the fallback when the timed region is shorter than one sampling period."""

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
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/r01_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            return json.load(fh)[key]["bytes"]
    except Exception:
        return None


def glorot(rng, rows, cols):
    lim = np.sqrt(6.0 / (rows + cols))
    return rng.uniform(-lim, lim, size=(rows, cols)).astype(np.float32)


def synth_filter_csr(rng, Q, N, lam):
    """Poisson(lam) known-true ids per query, sorted, unique (WN18RR tail mean 5.4, head 23.5)."""
    lens = rng.poisson(lam, Q)
    offs = np.zeros(Q + 1, np.int64)
    offs[1:] = np.cumsum(lens)
    ids = np.empty(offs[-1], np.int32)
    for q in range(Q):
        ids[offs[q]:offs[q + 1]] = np.sort(rng.choice(N, lens[q], replace=False))
    return offs, ids


def run_reference(args, rank, world):
    """Reference arm: the CPU port of the reference step (TensorFlow is not installable here), all host threads."""
    if rank != 0:
        return
    from oracle import cpu_port, kge_oracle as ko
    import torch
    w = WORKLOAD
    rng = np.random.default_rng(0)
    E, R = glorot(rng, w["N"], w["d"]), glorot(rng, w["R"], w["d"])
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    r = cpu_port.time_train(ko.TRANSE_L1, E, R, args.batch, w["K"], w["gamma"], steps=args.steps, warmup=max(1, min(args.warmup, 2)),
                            threads=threads, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "train_triples_per_s", "value": r["triples_per_s"], "unit": "triples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "batch_positives": args.batch, "negatives_per_positive": w["K"],
                   "loss": "pairwize_hinge (paired)", "optimizer": "lazy sparse Adam (torch.optim.SparseAdam)"},
        "cpu_baseline": {"value": r["triples_per_s"], "unit": "triples/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{r['steps']} full steps of the same workload (PyTorch-CPU fp32 restatement; TensorFlow not installable)"},
        "e2e": {"value": r["triples_per_s"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="positives per step")
    ap.add_argument("--corrupt", default="both", choices=["both", "head", "tail", "either"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-steps", type=int, default=10)
    ap.add_argument("--skip-large", action="store_true", help="skip the 1M-entity (8 GB) secondary train measurement")
    ap.add_argument("--skip-shapes", action="store_true", help="skip the secondary C1 / C3 / C4 shape measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    N, R, d, K, B = w["N"], w["R"], w["d"], w["K"], args.batch
    corrupt = {"both": abi.CORRUPT_BOTH, "head": abi.CORRUPT_HEAD, "tail": abi.CORRUPT_TAIL, "either": abi.CORRUPT_EITHER}[args.corrupt]
    rng = np.random.default_rng(0)  # same tables on every rank (replicated), different batches per rank
    E0, R0 = glorot(rng, N, d), glorot(rng, R, d)
    eng = Engine(abi.TRANSE_L1, N, R, d, d, gamma=w["gamma"], E=E0, R=R0)
    brng = np.random.default_rng(1000 + rank)
    steps, warm = args.steps, max(args.warmup, 3)
    nb = steps + warm
    host_batches = torch.from_numpy(np.stack([brng.integers(0, N, (nb, B)), brng.integers(0, R, (nb, B)),
                                              brng.integers(0, N, (nb, B))], axis=2).astype(np.int32)).pin_memory()
    dev_batches = host_batches.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    step_kw = dict(K=K, loss_kind=abi.PAIRWISE_HINGE, pair_mode=abi.PAIRED, margin=w["margin"], corrupt_mode=corrupt, seed=1234 + rank)
    opt_kw = dict(opt_kind=abi.LAZY_ADAM, lr=1e-3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed diagnostic: unique rows touched by one step (for the algorithmic-bytes figure)
    res = eng.forward_backward(dev_batches[0], None, want_neg=True, offset=0, **step_kw, opt_kind=abi.LAZY_ADAM)
    U_E, U_R = int(eng.tE.sum().item()), int(eng.tR.sum().item())
    eng.apply_optimizer(**opt_kw)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(warm):
        flush.zero_()
        eng.train_step(dev_batches[i], None, offset=i, **step_kw, **opt_kw)
    # ---- timed: K steps, device-resident inputs, per-step CUDA events, L2 flushed between steps ----------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    barrier()
    sampler.timed = True
    l0 = eng.launches
    t_wall0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        eng.forward_backward(dev_batches[warm + i], None, offset=warm + i, **step_kw, opt_kind=abi.LAZY_ADAM)
        ev[i][1].record()
        eng.apply_optimizer(**opt_kw)
        ev[i][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launches - l0
    if sampler.h is not None and not sampler.timed_samples:
        try:
            sampler.sample_once()   # the device has just finished the last timed step
        except Exception:
            pass
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t_fb = np.array([e[0].elapsed_time(e[1]) for e in ev])
    t_op = np.array([e[1].elapsed_time(e[2]) for e in ev])
    t_step_ms = float((t_fb + t_op).sum())
    tt = torch.tensor([t_step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total_ms = float(tt.item())
    value = world * steps * B * (1 + K) / (t_total_ms * 1e-3)

    # ---- e2e: host index buffers, H2D + step + D2H(loss) timed per step on the host clock ---------------------------
    e2e_t = 0.0
    n_e2e = min(steps, 50)
    for i in range(3 + n_e2e):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch = host_batches[(warm + i) % nb].to(dev, non_blocking=True)
        r = eng.train_step(batch, None, offset=10_000 + i, **step_kw, **opt_kw)
        loss_val = r["loss"].item()
        if i >= 3:
            e2e_t += time.perf_counter() - t0
    te = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e * B * (1 + K) / float(te.item())

    # ---- filtered 1-N evaluation: rows sharded over ranks, integer counts all-reduced ------------------------------
    erng = np.random.default_rng(7)
    nt = w["test_triples"]
    tri = np.stack([erng.integers(0, N, nt), erng.integers(0, R, nt), erng.integers(0, N, nt)], axis=1).astype(np.int32)
    csr_t, csr_h = synth_filter_csr(erng, nt, N, 5.4), synth_filter_csr(erng, nt, N, 23.5)
    per = -(-N // world)
    rr = (rank * per, min(N, (rank + 1) * per))
    tri_d = torch.from_numpy(tri).to(dev)
    csr_t = tuple(torch.from_numpy(a).to(dev) for a in csr_t)
    csr_h = tuple(torch.from_numpy(a).to(dev) for a in csr_h)
    eval_ms = []
    l1 = eng.launches
    for i in range(3 + args.eval_steps):
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ct = eng.rank_counts(tri_d, abi.SIDE_TAIL, csr_t[0], csr_t[1], row_range=rr)
        ch = eng.rank_counts(tri_d, abi.SIDE_HEAD, csr_h[0], csr_h[1], row_range=rr)
        if world > 1:
            both = torch.cat([ct, ch])
            dist.all_reduce(both)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            eval_ms.append(e0.elapsed_time(e1))
    eval_launches = (eng.launches - l1) // (3 + args.eval_steps)
    tev = torch.tensor([float(np.sum(eval_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tev, op=dist.ReduceOp.MAX)
    eval_qps = args.eval_steps * 2 * nt / (float(tev.item()) * 1e-3)

    # ---- the same evaluation end to end through the reference-facing call: validate(model, test, N, bs, filtering_triples) with
    #      HOST arrays (filter-CSR build, uploads, ranking, read-back of the ranks, the 15-key metrics dict) ------------------------
    eval_e2e = None
    if world == 1:
        import KGEkeras
        from KGEkeras import utils as ku
        vm = KGEkeras.TransE(e_dim=d, r_dim=d, num_entities=N, num_relations=R, norm=1, gamma=w["gamma"], dp=0.0, init_entities=E0, init_relations=R0)
        vfilt = np.stack([erng.integers(0, N, 89869), erng.integers(0, R, 89869), erng.integers(0, N, 89869)], axis=1)   # |train|+|valid| of WN18RR
        vkey = lambda a: (a[:, 0].astype(np.int64) * R + a[:, 1]) * N + a[:, 2]
        vfilt = vfilt[~np.isin(vkey(vfilt), vkey(tri))]          # the reference crashes when a test triple is in its own filter
        ku.validate(vm, tri, N, 32, vfilt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            vres = ku.validate(vm, tri, N, 32, vfilt)
        torch.cuda.synchronize()
        vdt = (time.perf_counter() - t0) / 3
        eval_e2e = {"api": "KGEkeras.utils.validate(model, test_triples, N, bs, filtering_triples) with host arrays", "queries_per_s": 2 * nt / vdt,
                    "ms_per_call": vdt * 1e3, "filter_triples": int(len(vfilt)), "mrr": float(vres["mrr"])}
        del vm
        torch.cuda.empty_cache()

    # ---- 1-N filtered evaluation of a dot-product model on the tensor cores (BASELINE configs[2] shape) --------------
    te = TENSOR_EVAL
    tN, tR, tW, tnt = te["N"], te["R"], 2 * te["d"], te["test_triples"]
    trng = np.random.default_rng(11)
    teng = Engine(abi.COMPLEX, tN, tR, tW, tW, E=glorot(trng, tN, tW), R=glorot(trng, tR, tW))
    teng.set_kernel_timing(True)
    ttri = torch.from_numpy(np.stack([trng.integers(0, tN, tnt), trng.integers(0, tR, tnt), trng.integers(0, tN, tnt)], axis=1).astype(np.int32)).to(dev)
    tcsr_t = tuple(torch.from_numpy(x).to(dev) for x in synth_filter_csr(trng, tnt, tN, 5.4))
    tcsr_h = tuple(torch.from_numpy(x).to(dev) for x in synth_filter_csr(trng, tnt, tN, 23.5))
    tper = -(-tN // world)
    trr = (rank * tper, min(tN, (rank + 1) * tper))
    tensor_eval = {}
    for mname, mode in (("exact_bf16x3", abi.TC_BF16X3_EXACT), ("approx_bf16", abi.TC_BF16)):
        packed = teng.pack_entities_tc(mode, trr)   # once per table version: outside the timed region, like the table itself
        ms_pass, ms_kern, ms_ref, stats = [], [], [], {}
        lt0 = teng.launches
        for i in range(3 + args.eval_steps):
            flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ct = teng.rank_counts_tc(ttri, abi.SIDE_TAIL, tcsr_t[0], tcsr_t[1], row_range=trr, mode=mode, packed=packed, stats=stats)
            k1, r1 = teng.last_kernel_ms(abi.TIMED_RANK_TC), teng.last_kernel_ms(abi.TIMED_RANK_REFINE)
            ch = teng.rank_counts_tc(ttri, abi.SIDE_HEAD, tcsr_h[0], tcsr_h[1], row_range=trr, mode=mode, packed=packed)
            if world > 1:
                both = torch.cat([ct, ch])
                dist.all_reduce(both)
            e1.record()
            torch.cuda.synchronize()
            k2, r2 = teng.last_kernel_ms(abi.TIMED_RANK_TC), teng.last_kernel_ms(abi.TIMED_RANK_REFINE)
            if i >= 3:
                ms_pass.append(e0.elapsed_time(e1)); ms_kern.append(k1 + k2); ms_ref.append(r1 + r2)
        tl = (teng.launches - lt0) // (3 + args.eval_steps)
        tt2 = torch.tensor([float(np.sum(ms_pass))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
        planes = 3 if mode == abi.TC_BF16X3_EXACT else 1
        flops_alg = 2.0 * (2 * tnt) * (trr[1] - trr[0]) * tW          # this rank's shard, both sides (2*Q*N*Kc)
        kms = float(np.mean(ms_kern))
        tensor_eval[mname] = {
            "queries_per_s": args.eval_steps * 2 * tnt / (float(tt2.item()) * 1e-3), "ms_per_pass": float(np.mean(ms_pass)),
            "kernel_ms": {"k_rank_tc (tail+head)": kms, "exact re-scoring (tail+head)": float(np.mean(ms_ref))},
            "undecided_pairs_tail": stats.get("undecided_pairs"), "gpu_launches_per_pass": int(tl),
            "algorithmic_tflops": flops_alg / (kms * 1e-3) / 1e12, "executed_tflops": planes * flops_alg / (kms * 1e-3) / 1e12,
        }
    del teng

    # ---- the same fused step on a table that does NOT fit the L2 (BASELINE configs[4] shape, one GPU's worth, unsharded):
    #      here the HBM roofline is the real bound, so this is where the step's HBM fraction is read ----------------------
    train_large = None
    if world == 1 and not args.skip_large:
        lN, lR, lW, lB, lK = 1_000_000, 1000, 512, 16384, 64
        lrng = np.random.default_rng(21)
        lim = float(np.sqrt(6.0 / (lN + lW)))
        leng = Engine(abi.COMPLEX, lN, lR, lW, lW, E=torch.empty(lN, lW).uniform_(-lim, lim, generator=torch.Generator().manual_seed(3)),
                      R=glorot(lrng, lR, lW))
        lsteps, lwarm = 10, 3
        lb = torch.from_numpy(np.stack([lrng.integers(0, lN, (lsteps + lwarm, lB)), lrng.integers(0, lR, (lsteps + lwarm, lB)),
                                        lrng.integers(0, lN, (lsteps + lwarm, lB))], axis=2).astype(np.int32)).to(dev)
        lkw = dict(K=lK, loss_kind=abi.PAIRWISE_LOGISTIC, pair_mode=abi.PAIRED, margin=1.0, corrupt_mode=abi.CORRUPT_BOTH, seed=99)
        leng.forward_backward(lb[0], None, offset=0, **lkw, opt_kind=abi.LAZY_ADAM)
        lU_E, lU_R = int(leng.tE.sum().item()), int(leng.tR.sum().item())
        leng.apply_optimizer(abi.LAZY_ADAM, lr=1e-3)
        for i in range(lwarm):
            leng.train_step(lb[i], None, offset=i, **lkw, opt_kind=abi.LAZY_ADAM, lr=1e-3)
        lev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(lsteps)]
        torch.cuda.synchronize()
        for i in range(lsteps):
            lev[i][0].record()
            leng.forward_backward(lb[lwarm + i], None, offset=lwarm + i, **lkw, opt_kind=abi.LAZY_ADAM)
            lev[i][1].record()
            leng.apply_optimizer(abi.LAZY_ADAM, lr=1e-3)
            lev[i][2].record()
        torch.cuda.synchronize()
        lfb = float(np.mean([e[0].elapsed_time(e[1]) for e in lev])); lop = float(np.mean([e[1].elapsed_time(e[2]) for e in lev]))
        train_large = dict(workload="C5 ComplEx dim=256 (stored width 512), N=1e6, R=1e3, one GPU, unsharded; table 2 GB + Adam 4 GB + gradient 2 GB >> 126 MB L2",
                           batch_positives=lB, negatives_per_positive=lK, unique_entity_rows_per_step=lU_E, unique_relation_rows_per_step=lU_R,
                           triples_per_s=lB * (1 + lK) / ((lfb + lop) * 1e-3), kernel_ms={"k_train_ring": lfb, "k_opt_update2": lop},
                           bytes_fwd_bwd=8.0 * lW * (lU_E + lU_R) + 12.0 * lB, bytes_opt=32.0 * lW * (lU_E + lU_R),
                           bytes_step=28.0 * lW * (lU_E + lU_R) + 12.0 * lB)
        del leng, lb
        torch.cuda.empty_cache()

    # ---- the same C2 step with the reference's default Dropout(0.2) on every gathered row (models.py:122; the headline runs
    #      dp = 0 like the harness' params dict leaves it): Philox mask bits per element double the kernel's instruction count ----
    train_dropout = None
    if world == 1:
        dms = []
        for i in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.train_step(dev_batches[i % nb], None, offset=50_000 + i, dropout=0.2, **step_kw, **opt_kw)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                dms.append(e0.elapsed_time(e1))
        train_dropout = {"dropout": 0.2, "ms_per_step": float(np.mean(dms)), "triples_per_s": B * (1 + K) / (float(np.mean(dms)) * 1e-3),
                         "note": "counter-based Philox keep masks regenerated in the backward pass, nothing stored"}

    # ---- the remaining named shapes (BASELINE configs 0, 2, 3): one fused / unfused train step each, and HolE's 1-N eval ------
    other_shapes = None
    if world == 1 and not args.skip_shapes:
        other_shapes = {}
        for sname, skind, sN, sR, sW, sB, sK, sloss, snote in (
                ("C1 DistMult dim=100 kinship-shape (N=104, R=25)", abi.DISTMULT, 104, 25, 100, 2048, 10, abi.PAIRWISE_HINGE,
                 "harness batch 2048 x 10 negatives, pairwize_hinge; launch-latency bound (41 KB table)"),
                ("C3 ComplEx dim=200 FB15k-237-shape (N=14541, R=237)", abi.COMPLEX, 14541, 237, 400, 16384, 10, abi.PAIRWISE_HINGE,
                 "fused ring kernel, pairwize_hinge"),
                ("C4 HolE dim=256 YAGO3-10-shape (N=123182, R=37)", abi.HOLE, 123182, 37, 256, 2048, 64, abi.POINTWISE_LOGISTIC,
                 "64 negatives, pointwize_logistic; unfused score_fwd -> loss -> score_bwd (O(d^2) correlation kernels)")):
            srng = np.random.default_rng(31)
            seng = Engine(skind, sN, sR, sW, sW, E=glorot(srng, sN, sW), R=glorot(srng, sR, sW))
            ssteps, swarm = 20, 3
            sb = torch.from_numpy(np.stack([srng.integers(0, sN, (ssteps + swarm, sB)), srng.integers(0, sR, (ssteps + swarm, sB)),
                                            srng.integers(0, sN, (ssteps + swarm, sB))], axis=2).astype(np.int32)).to(dev)
            skw = dict(K=sK, loss_kind=sloss, pair_mode=abi.PAIRED, margin=1.0, corrupt_mode=abi.CORRUPT_BOTH, seed=77, opt_kind=abi.LAZY_ADAM, lr=1e-3)
            def time_shape():
                for i in range(swarm):
                    seng.train_step(sb[i], None, offset=i, **skw)
                sev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(ssteps)]
                sl0 = seng.launches
                for i in range(ssteps):
                    flush.zero_()
                    sev[i][0].record()
                    seng.train_step(sb[swarm + i], None, offset=swarm + i, **skw)
                    sev[i][1].record()
                torch.cuda.synchronize()
                sms = float(np.mean([e[0].elapsed_time(e[1]) for e in sev]))
                return sms, (seng.launches - sl0) // ssteps

            direct = None
            if skind == abi.HOLE:   # the O(d^2) sequential-contract kernels first, then the Fourier-domain mode (the reference's own formulation)
                dms, _ = time_shape()
                direct = {"ms_per_step": dms, "triples_per_s": sB * (1 + sK) / (dms * 1e-3), "note": "default mode: direct O(d^2) correlation, fp32 sequential contract"}
                seng.set_hole_mode(True)
                snote = "64 negatives, pointwize_logistic; kgeb_set_hole_mode(KGEB_HOLE_FFT): one fused kgeb_train_fwd_bwd kernel on shared-memory radix-4 FFTs + the optimizer"
            sms, slaunch = time_shape()
            entry = {"batch_positives": sB, "negatives_per_positive": sK, "ms_per_step": sms, "triples_per_s": sB * (1 + sK) / (sms * 1e-3),
                     "gpu_launches_per_step": slaunch, "note": snote}
            if direct is not None:
                entry["direct_kernels"] = direct
            if skind == abi.HOLE:   # C4 filtered 1-N eval on the tensor cores (exact mode), 5 000 test triples x 2 sides
                qn = 5000
                qtri = torch.from_numpy(np.stack([srng.integers(0, sN, qn), srng.integers(0, sR, qn), srng.integers(0, sN, qn)], axis=1).astype(np.int32)).to(dev)
                qc_t = tuple(torch.from_numpy(x).to(dev) for x in synth_filter_csr(srng, qn, sN, 5.4))
                qc_h = tuple(torch.from_numpy(x).to(dev) for x in synth_filter_csr(srng, qn, sN, 23.5))
                packed = seng.pack_entities_tc(abi.TC_BF16X3_EXACT)
                qms = []
                for i in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    seng.rank_counts_tc(qtri, abi.SIDE_TAIL, qc_t[0], qc_t[1], mode=abi.TC_BF16X3_EXACT, packed=packed)
                    seng.rank_counts_tc(qtri, abi.SIDE_HEAD, qc_h[0], qc_h[1], mode=abi.TC_BF16X3_EXACT, packed=packed)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 2:
                        qms.append(e0.elapsed_time(e1))
                entry["eval_exact"] = {"queries_per_pass": 2 * qn, "ms_per_pass": float(np.mean(qms)),
                                       "queries_per_s": 2 * qn / (float(np.mean(qms)) * 1e-3)}
            other_shapes[sname] = entry
            del seng, sb
            torch.cuda.empty_cache()

    # ---- the reference harness' own step, as tests.py main() configures it: DistMult dim=100 on the FB15k-237 vocabulary, 5 random
    #      literals through a tanh Dense, Dropout(0.2), batch 2048 x 10 negatives oversampled to m = 20480, all-pairs pairwize_hinge,
    #      Adam -- through the reference-facing API (model.fit on the DataGenerator, device-side batch pipeline) ------------------------
    harness = None
    if world == 1 and not args.skip_shapes:
        try:
            import KGEkeras
            hrng = np.random.default_rng(41)
            hN, hR, hn = 14541, 237, 2048 * 24
            htrain = np.stack([hrng.integers(0, hN, hn), hrng.integers(0, hR, hn), hrng.integers(0, hN, hn)], axis=1).astype(np.int32)
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
            harness = {"workload": "tests.py main(): DistMult dim=100, 5 literals + tanh Dense, dp=0.2, bs=2048, ns=10, all-pairs hinge, adam (FB15k-237 vocabulary)",
                       "ms_per_batch": hdt * 1e3, "scored_triples_per_s": 2 * 20480 / hdt,
                       "api": "model.fit(KGEkeras.DataGenerator(train, ns=10, batch_size=2048, engine=model._engine_()))"}
            del hm, hgen
            torch.cuda.empty_cache()
        except Exception as ex:   # a secondary measurement must never cost the headline line
            harness = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    if train_large is not None:
        tl = train_large
        for key, byt, ms in (("roofline_fwd_bwd", tl["bytes_fwd_bwd"], tl["kernel_ms"]["k_train_ring"]),
                             ("roofline_opt", tl["bytes_opt"], tl["kernel_ms"]["k_opt_update2"]),
                             ("roofline_step", tl["bytes_step"], tl["kernel_ms"]["k_train_ring"] + tl["kernel_ms"]["k_opt_update2"])):
            a_gbs = byt / (ms * 1e-3) / 1e9
            tl[key] = {"bound": "hbm", "achieved": a_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": a_gbs / pk["hbm_gbs"]}
            tkey = {"roofline_fwd_bwd": "k_train_ring_c5", "roofline_opt": "k_opt_update_E_c5"}.get(key)
            if tkey and ncu_traffic(tkey):
                # DRAM bytes really moved (ncu): every one of the 2.1 M row references misses the 126 MB L2 (hit rate 2 %),
                # so the kernel streams ~3.5x its compulsory bytes -- and does so at the HBM peak
                tl[key]["traffic"] = ncu_traffic(tkey)
                tl[key]["traffic_gbs"] = ncu_traffic(tkey) / (ms * 1e-3) / 1e9
                tl[key]["traffic_frac_of_peak"] = tl[key]["traffic_gbs"] / pk["hbm_gbs"]
    for mname, v in tensor_eval.items():
        v["roofline"] = {"bound": "tensor", "kernel": "k_rank_tc", "achieved": v["algorithmic_tflops"], "peak": pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": v["algorithmic_tflops"] / pk["bf16_tflops"],
                         "pipe_frac": v["executed_tflops"] / pk["bf16_tflops"],
                         "traffic": ncu_traffic("k_rank_tc_exact" if mname.startswith("exact") else "k_rank_tc_bf16"), "peak_source": pk["source"] + ", burst",
                         "note": "achieved = algorithmic 2*Q*N*Kc flops / CUDA-event kernel time; pipe_frac counts the MMAs really issued "
                                 "(3 per k-step in the exact hi/lo-split mode)"}
    # algorithmic bytes: fused fwd+bwd kernel reads each unique touched row once and writes its gradient once
    # -> 8 B/element + 12 B of ids per positive; the optimizer reads g,p,m,v and writes p,m,v,g(=0) -> 32 B/element.
    bytes_fb = 8.0 * (d * U_E + d * U_R) + 12.0 * B
    bytes_op = 32.0 * (d * U_E + d * U_R)
    bytes_step = (4 + 24) * (d * U_E + d * U_R) + 12.0 * B  # compulsory bytes of the whole step
    fb_ms, op_ms = float(t_fb.mean()), float(t_op.mean())
    dom = ("k_train_fwd_bwd", bytes_fb, fb_ms) if fb_ms >= op_ms else ("k_opt_update (E and R)", bytes_op, op_ms)
    ach = dom[1] / (dom[2] * 1e-3) / 1e9
    ach_step = bytes_step / ((fb_ms + op_ms) * 1e-3) / 1e9
    # gather traffic the kernel actually asks the memory system for (no reuse): 2 entity rows read + 2 written per triple
    gather_bytes = B * (1 + K) * (4 * d * 4.0) + B * (2 * d * 4.0)
    line = {
        "metric": "train_triples_per_s", "value": value, "unit": "triples/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": t_total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "batch_positives": B, "negatives_per_positive": K, "corruption": args.corrupt,
                   "loss": "pairwize_hinge (paired)", "optimizer": "LazyAdam (touched rows)", "negatives": "in-kernel Philox4x32-10",
                   "l2": "flushed between steps (256 MiB write); table+state 98 MB < 126 MB L2",
                   "multi_gpu": "replicas only for training; eval rows sharded + NCCL all-reduce of int32 counts",
                   "unique_entity_rows_per_step": U_E, "unique_relation_rows_per_step": U_R},
        "e2e": {"value": e2e_value, "unit": "triples/s", "h2d_bytes_per_step": int(B * 3 * 4), "d2h_bytes_per_step": 4,
                "steps": n_e2e, "api": "Engine.train_step(host int32 batch) -> loss.item()"},
        "gpu_launches": int(launches),
        "kernel_ms": {"k_train_fwd_bwd": fb_ms, "k_opt_update2": op_ms, "wall_ms_per_step_incl_flush": 1e3 * t_wall / steps},
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / pk["hbm_gbs"],
                     "traffic": ncu_traffic("k_train_ring" if dom[0] == "k_train_fwd_bwd" else "k_opt_update_E"), "peak_source": pk["source"],
                     "note": "algorithmic = compulsory bytes with perfect on-chip reuse of the %d unique rows; the kernel's "
                             "no-reuse gather+scatter request volume is %.0f MB/launch (L2-resident table)" % (U_E, gather_bytes / 1e6),
                     "l2_path": {"kernel_us": fb_ms * 1e3, "vector_reduction_floor_us": 59.0, "gather_plus_reduction_floor_us": 67.0,
                                 "frac_of_floor": 67.0 / (fb_ms * 1e3),
                                 "source": "profiles/micro/l2_rows.cu on a B200: red.global.add.v4.f32 row reductions sustain 5.1 TB/s "
                                           "whatever the occupancy (cp.reduce.async.bulk the same), row gathers 14 TB/s; floors are for "
                                           "this workload's 378 880 x 800 B row references (profiles/r01_v4_train_bound.md)"}},
        "roofline_step": {"bound": "hbm", "achieved": ach_step, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_step / pk["hbm_gbs"],
                          "bytes_step": bytes_step, "formula": "(4+24)*(We*U_E+Wr*U_R)+12*B (SURVEY 8d, Adam)"},
        "eval": {"metric": "filtered_eval_queries_per_s", "value": eval_qps, "unit": "queries/s", "queries_per_pass": 2 * nt,
                 "ms_per_pass": float(np.mean(eval_ms)), "passes": args.eval_steps, "scaling": "strong",
                 "gpu_launches_per_pass": int(eval_launches), "e2e": eval_e2e,
                 "fp32_alu_tflops": (2 * d * nt * N + 3 * d * nt * N) * 1.0 / (float(np.mean(eval_ms)) * 1e-3) / 1e12,
                 "bound": "fp32 issue: two rounded fp32 operations per (query, entity, k) for the tail side, three for the head side; "
                          "the tile kernel runs at 65 % of the FADD issue ceiling measured by profiles/micro/fadd_peak.cu "
                          "(profiles/r01_v3_summary.md)"},
        "eval_tensor": {"workload": te["name"], "queries_per_pass": 2 * tnt, "scaling": "strong",
                        "filter": "synthetic CSR, Poisson(5.4) tail / Poisson(23.5) head known-true ids per query",
                        "parity": "exact_bf16x3 counts are bit-identical to the fp32 CUDA-core kernel (tests/test_gpu_rank_tc.py)",
                        **tensor_eval},
        "train_large_table": train_large,
        "other_shapes": other_shapes,
        "train_with_dropout": train_dropout,
        "reference_harness_step": harness,
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_port, kge_oracle as ko
        cb = cpu_port.time_train(ko.TRANSE_L1, E0, R0, B, K, w["gamma"], steps=6, warmup=1, threads=os.cpu_count(), budget_s=20.0)
        ce = cpu_port.time_eval(ko.TRANSE_L1, E0, R0, tri, [tuple(int(x) for x in r) for r in tri[:2000]], w["gamma"],
                                max_queries=48, budget_s=10.0)
        line["cpu_baseline"] = {"value": cb["triples_per_s"], "unit": "triples/s", "cores": cb["threads"], "kind": "port",
                                "sample": f"{cb['steps']} full steps (B={B}, K={K}) of the PyTorch-CPU fp32 restatement; TensorFlow not installable",
                                "eval_queries_per_s": ce["queries_per_s"], "eval_sample": f"{ce['queries']} queries, NumPy restatement of validate, 1 process"}
        # the reference's host batch pipeline (DataGenerator.__get_data, tests.py:55-61: generate_negative + oversample_data,
        # restated literally in the oracle: per-row Python lists, random.choices) at the harness' batch of 2048 x 10,
        # beside its device-side replacement (kgeb_negatives + kgeb_oversample through Engine.device_batch)
        import random
        hb = np.asarray(host_batches[0][:2048].numpy(), dtype=np.int64)
        np.random.seed(0); random.seed(0)
        t0 = time.perf_counter(); nrep = 0
        while time.perf_counter() - t0 < 3.0:
            X = ko.oversample_data([hb, ko.generate_negative(hb, N, K)])
            nrep += 1
        host_rate = nrep * len(X[0]) / (time.perf_counter() - t0)
        db = dev_batches[0][:2048]
        for _ in range(3):
            eng.device_batch(db, K, 1, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(200):
            eng.device_batch(db, K, 1, i)
        torch.cuda.synchronize()
        dev_rate = 200 * 2048 * K / (time.perf_counter() - t0)
        line["cpu_baseline"]["host_batch_pipeline"] = {"reference_style_pairs_per_s": host_rate, "device_side_pairs_per_s": dev_rate,
                                                        "batch": "2048 positives x 10 negatives -> (20480,3) positive and negative lists",
                                                        "note": "1 thread; the reference's loop feeds its GPU step from this host code"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
