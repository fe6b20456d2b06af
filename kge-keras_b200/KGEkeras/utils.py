"""Host-side mirror of KGEkeras/utils.py (reference): load_kg, generate_negative, oversample_data, pad, mrr, hits,
gen_tail_data, gen_head_data, validate, KGEValidateCallback, the seven losses and loss_function_lookup -- same names,
argument meaning and error behaviour.  The data-preparation helpers stay host code (they ARE host code in the
reference) but are vectorised and consume the global `np.random` / `random` streams exactly like the reference, so
seeded runs give bit-identical index tensors (pinned by tests/golden/negatives_ref.npz).  validate() and the losses
run on the GPU through libkgeb200.  No TensorFlow, rdflib or spaCy imports (LiteralConverter is out of scope).
"""
from __future__ import annotations

import math
import os
from collections import defaultdict
from random import choice, choices

import numpy as np

from . import _abi as abi

EPSILON = 1e-6  # utils.py:12


# ---- data preparation (host) -----------------------------------------------------------------------------------
def isint(value):
    """utils.py:23-28."""
    try:
        int(value)
        return True
    except ValueError:
        return False


def load_kg(path):
    """utils.py:106-112: one whitespace-split token list per line (no id mapping)."""
    with open(path, 'r') as f:
        return [line.strip().split() for line in f]


def generate_negative(kg, N, negative=2, check_kg=False, corrupt_head=True, corrupt_tail=True):
    """utils.py:114-121.  Both head and tail are redrawn uniformly in [0,N) (fs in full first, then fo), the relation of
    positive i fills rows i*negative .. i*negative+negative-1, result shape (negative*len(kg), 3, 1).  As in the
    reference the three flags are accepted and (beyond the assert) ignored."""
    assert corrupt_head or corrupt_tail
    n = negative * len(kg)
    rel = np.repeat(np.asarray(kg)[:, 1].reshape((-1, 1)), negative, axis=0) if len(kg) else np.zeros((0, 1), dtype=np.int64)
    fs = np.random.randint(0, N, size=(n, 1))
    fo = np.random.randint(0, N, size=(n, 1))
    return np.stack([fs, rel, fo], axis=1)


def oversample_data(kgs, x=None, y=None, testing=False):
    """utils.py:123-147: every list is resampled WITH replacement to the longest length.  `random.choices(range(n), k=m)`
    draws the same indices from the same global stream as the reference's choices(list_of_rows, k=m)."""
    if testing:
        kgs = [np.asarray(kg)[:len(y)] for kg in kgs]
    else:
        kgs = [np.asarray(kg) for kg in kgs]
    m = max(len(kg) for kg in kgs)
    if y is not None:
        m = max(m, len(y))
    out = [kg[np.asarray(choices(range(len(kg)), k=m), dtype=np.int64)] for kg in kgs]
    if x is not None and y is not None:
        k = np.ceil(m / len(y))
        y = np.repeat(y, k, axis=0)[:m]
        x = np.repeat(x, k, axis=0)[:m, :]
        for s in np.split(x, 3, axis=1):
            out.append(s.reshape((-1,)))
        return [np.squeeze(np.asarray(o)) for o in out], np.asarray(y)
    return [np.squeeze(o) for o in out]


def pad(kg, bs):
    """utils.py:149-153: append random existing triples until len % bs == 0."""
    kg = list(kg)
    while len(kg) % bs != 0:
        kg.append(choice(kg))
    return np.asarray(kg)


def mrr(target, scores):
    """utils.py:155-158 (stable descending sort: ties keep list order)."""
    labels = [x for x, _ in sorted(scores, key=lambda x: x[1], reverse=True)]
    return 1 / (1 + labels.index(target))


def hits(target, scores, k=10):
    """utils.py:160-163"""
    labels = [x for x, _ in sorted(scores, key=lambda x: x[1], reverse=True)][:k]
    return int(target in labels)


def _candidates(num_entities, target, filtered):
    keep = np.ones(num_entities, dtype=bool)
    for e in [target] + list(filtered):
        e = int(e)
        if not keep[e]:
            raise ValueError("list.remove(x): x not in list")  # what the reference's list.remove raises (utils.py:169-171)
        keep[e] = False
    return np.nonzero(keep)[0]


def gen_tail_data(test_data, num_entities, bs, filter_t):
    """utils.py:165-179: per test triple a (C+1,3) array, row 0 the target, then every non-filtered tail candidate."""
    for s, p, o in test_data:
        cand = _candidates(num_entities, o, filter_t[(s, p)])
        tri = np.empty((len(cand) + 1, 3), dtype=np.int64)
        tri[:, 0], tri[:, 1] = int(s), int(p)
        tri[0, 2] = int(o)
        tri[1:, 2] = cand
        yield tri


def gen_head_data(test_data, num_entities, bs, filter_h):
    """utils.py:181-196"""
    for s, p, o in test_data:
        cand = _candidates(num_entities, s, filter_h[(p, o)])
        tri = np.empty((len(cand) + 1, 3), dtype=np.int64)
        tri[:, 2], tri[:, 1] = int(o), int(p)
        tri[0, 0] = int(s)
        tri[1:, 0] = cand
        yield tri


# ---- filtered evaluation (GPU) ---------------------------------------------------------------------------------
def build_filter_csr(test, filtering_triples, side):
    """Sorted, de-duplicated known-true ids per query in CSR form (offsets int64[Q+1], ids int32[nnz]).
    Replaces the dict-of-sets of utils.py:201-205 and the per-query list.remove of utils.py:169-171 by one sort of the
    filter triples and two binary searches per query (vectorised; FB15k-237-sized inputs take milliseconds).
    Raises ValueError if a query's own target is in its filter set -- the reference crashes there (utils.py:169-171)."""
    filt = np.asarray(filtering_triples, dtype=np.int64).reshape(-1, 3)
    test = np.asarray(test, dtype=np.int64).reshape(-1, 3)
    kc = (0, 1) if side == "tail" else (1, 2)       # key columns: (s,p) for tail queries, (p,o) for head queries
    vc = 2 if side == "tail" else 0
    Q = len(test)
    if len(filt) == 0 or Q == 0:
        return np.zeros(Q + 1, dtype=np.int64), np.zeros(0, dtype=np.int32)
    # one span per column (how many ids it can hold), so that the packed key stays below 2^63 for N * R * N, not max(N, R)^3
    sp = [int(max(filt[:, c].max(), test[:, c].max())) + 1 for c in range(3)]
    s0, s1, span = sp[kc[0]], sp[kc[1]], sp[vc]
    if s0 * s1 * span >= 2 ** 63:
        raise ValueError("ids too large for the packed filter key")
    # one sortable key per filter triple: (key0, key1, value); sort, then drop adjacent duplicates (np.unique's hash
    # path is 10x slower than the integer sort on these sizes)
    packed = np.sort((filt[:, kc[0]] * s1 + filt[:, kc[1]]) * span + filt[:, vc])
    if len(packed) > 1:
        packed = packed[np.concatenate(([True], packed[1:] != packed[:-1]))]
    qkey = (test[:, kc[0]] * s1 + test[:, kc[1]]) * span
    lo = np.searchsorted(packed, qkey, side="left")
    hi = np.searchsorted(packed, qkey + span, side="left")
    lens = hi - lo
    offs = np.zeros(Q + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    # gather the segments [lo_q, hi_q) back to back
    nnz = int(offs[-1])
    if nnz == 0:
        return offs, np.zeros(0, dtype=np.int32)
    seg = np.repeat(np.arange(Q), lens)
    pos = lo[seg] + (np.arange(nnz) - offs[seg])
    ids = packed[pos] - qkey[seg]
    if np.any(ids == test[seg, vc]):
        raise ValueError("list.remove(x): x not in list")
    return offs, ids.astype(np.int32)


def build_filter_csr_device(test_t, filt_t, side):
    """build_filter_csr restated with torch index ops on the GPU -- kept as a CHECKER of Engine.filter_csr (the library's own
    builder, which validate() uses); the NumPy builder takes 4-23 ms per side on the bundled datasets.
    test_t (Q,3) / filt_t (F,3) int64 CUDA tensors -> (offsets int64[Q+1], ids int32[nnz]) on the same device; same ValueError
    as the host builder when a query's own target is in its filter set."""
    import torch
    kc = (0, 1) if side == "tail" else (1, 2)
    vc = 2 if side == "tail" else 0
    Q = test_t.shape[0]
    dev = test_t.device
    if filt_t.shape[0] == 0 or Q == 0:
        return torch.zeros(Q + 1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)
    span = int(max(int(filt_t.max()), int(test_t.max()))) + 1
    if span ** 3 >= 2 ** 62:
        raise ValueError("ids too large for the packed filter key")
    packed = torch.unique((filt_t[:, kc[0]] * span + filt_t[:, kc[1]]) * span + filt_t[:, vc])   # sorted, de-duplicated
    qkey = (test_t[:, kc[0]] * span + test_t[:, kc[1]]) * span
    lo = torch.searchsorted(packed, qkey)
    hi = torch.searchsorted(packed, qkey + span)
    lens = hi - lo
    offs = torch.zeros(Q + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=offs[1:])
    nnz = int(offs[-1].item())
    if nnz == 0:
        return offs, torch.zeros(0, dtype=torch.int32, device=dev)
    seg = torch.repeat_interleave(torch.arange(Q, device=dev), lens, output_size=nnz)
    pos = lo[seg] + (torch.arange(nnz, device=dev) - offs[seg])
    ids = packed[pos] - qkey[seg]
    if bool((ids == test_t[seg, vc]).any().item()):
        raise ValueError("list.remove(x): x not in list")
    return offs, ids.to(torch.int32)


def shard_rows(num_entities, world, rank):
    """Contiguous row shard [r0, r1) of rank `rank` out of `world`: ceil(N/world) rows each, last may be short."""
    per = -(-int(num_entities) // int(world))
    return min(num_entities, rank * per), min(num_entities, (rank + 1) * per)


def merge_partial_counts(counts, group=None):
    """Sum the integer partial rank counts over the ranks (all-reduce; NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, group=group)
    return counts


def filtered_ranks(model, test_data, filtering_triples, group=None):
    """(tail_ranks, head_ranks) int64 arrays: rank = 1 + #{c not filtered, c != target : score_c > score_target}.
    With torch.distributed initialised (world > 1) the entity rows are split over the ranks and the integer partial
    counts are all-reduced (NCCL over NVLink): the result is identical on every rank and for any world size."""
    import torch
    import torch.distributed as dist
    eng = model._eval_engine_() if hasattr(model, "_eval_engine_") else model._engine_()
    test = np.asarray(test_data, dtype=np.int64).reshape(-1, 3)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    rows = shard_rows(eng.N, world, rank)
    # dot-product models (DistMult, ComplEx, HolE) rank on the tensor cores; "exact" gives the same integers as the fp32
    # CUDA-core kernel ("cuda"), "bf16" is the approximate single-plane mode.  TransE always takes the CUDA-core kernel.
    mode = str(getattr(model, "rank_mode", None) or os.environ.get("KGEB_RANK_MODE", "exact")).lower()
    if mode not in ("exact", "bf16", "cuda"):
        raise ValueError(f"rank_mode must be 'exact', 'bf16' or 'cuda', got {mode!r}")
    test_t = torch.from_numpy(test.astype(np.int32)).to(eng.device)
    filt_t = torch.from_numpy(np.asarray(filtering_triples, dtype=np.int32).reshape(-1, 3)).to(eng.device)
    csr = [eng.filter_csr(test_t, filt_t, code) for code in (abi.SIDE_TAIL, abi.SIDE_HEAD)]     # kgeb_filter_csr_count / _fill
    # one plan = both sides queued back to back, ONE all-reduce of the 2*Q partial counts, one read-back at the end
    plan = eng.rank_plan(test_t, csr[0], csr[1], row_range=rows, mode=mode, group=group)
    return plan.ranks()


def predict_ranks(model, test_data, num_entities, filtering_triples):
    """(tail_ranks, head_ranks) for ANY object with a Keras-style `.predict((n,3)) -> (n,)`, the way the reference drives it
    (utils.py:207-245): per test triple the (C+1,3) candidate array of gen_tail_data / gen_head_data (target in row 0) goes
    through `model.predict` in ONE call, exactly as the reference passes it -- a model whose predictions depend on the call's
    batching (HAKE, models.py:397) therefore ranks as it does there.  The scores are ranked on the GPU (kgeb_count_greater:
    rankdata(res,'max') with the target in row 0 == 1 + #{res > res[0]}); one read-back at the end."""
    import ctypes as C
    import torch
    from collections import defaultdict
    from . import _engine as E_
    eng = model._eval_engine_() if hasattr(model, "_eval_engine_") else _loss_eng()
    filter_h, filter_t = defaultdict(set), defaultdict(set)
    for s, p, o in filtering_triples:          # utils.py:201-205
        filter_h[(int(p), int(o))].add(int(s))
        filter_t[(int(s), int(p))].add(int(o))
    test = [(int(s), int(p), int(o)) for s, p, o in np.asarray(test_data).reshape(-1, 3)]
    Q = len(test)
    counts = torch.zeros(2, max(Q, 1), dtype=torch.int32, device=eng.device)
    keep = []
    for side, gen in ((0, gen_tail_data(test, num_entities, None, filter_t)), (1, gen_head_data(test, num_entities, None, filter_h))):
        for qi, cand in enumerate(gen):
            res = np.ascontiguousarray(np.asarray(model.predict(cand), dtype=np.float32).reshape(-1))
            if res.shape[0] != cand.shape[0]:
                raise ValueError(f"predict returned {res.shape[0]} scores for {cand.shape[0]} candidate triples")
            sc = torch.from_numpy(res).to(eng.device, non_blocking=True)
            eng._check(eng.lib.kgeb_count_greater(eng.ctx, C.c_void_p(sc.data_ptr()), sc.numel(), 0,
                                                  C.c_void_p(counts[side, qi:].data_ptr()), E_._stream()))
            keep.append(sc)
            if len(keep) >= 256:               # the copies are asynchronous: hold the buffers until the stream has used them
                torch.cuda.current_stream().synchronize()
                keep.clear()
    host = counts[:, :Q].cpu().numpy().astype(np.int64) + 1
    return host[0], host[1]


def validate(model, test_data, num_entities, bs, filtering_triples=None):
    """utils.py:199-264: filtered Hits@1/3/10, MR, MRR for tail and head prediction and their means, as the reference's
    15-key dict.  `filtering_triples=None` raises TypeError like the reference; `bs` is unused there and here.
    A KGEkeras model with a 1-N ranking kernel is ranked by it (every candidate of every query in one pass); any other
    object with `.predict` -- a foreign model, or HAKE, whose scores depend on predict's batching -- is driven exactly like
    the reference drives it (predict_ranks)."""
    if filtering_triples is None:
        raise TypeError("'NoneType' object is not iterable")
    n = float(len(test_data))
    fast = hasattr(model, "_engine_") and getattr(model, "_score_kind", None) != abi.HAKE
    if fast:
        tail_ranks, head_ranks = filtered_ranks(model, test_data, filtering_triples)
    elif hasattr(model, "predict"):
        tail_ranks, head_ranks = predict_ranks(model, test_data, num_entities, filtering_triples)
    else:
        raise AttributeError(f"{type(model).__name__!r} object has no attribute 'predict'")   # what utils.py:211 would raise

    def side(r):
        return (np.count_nonzero(r <= 1) / n, np.count_nonzero(r <= 3) / n, np.count_nonzero(r <= 10) / n,
                np.mean(r), np.mean(1.0 / r))   # == np.mean([1 / m for m in r]) bit for bit (same float64 array, same reduction)
    t1, t3, t10, tmr, tmrr = side(tail_ranks)
    h1, h3, h10, hmr, hmrr = side(head_ranks)
    return {'tail_hits@1': t1, 'tail_hits@3': t3, 'tail_hits@10': t10, 'tail_mr': tmr, 'tail_mrr': tmrr,
            'head_hits@1': h1, 'head_hits@3': h3, 'head_hits@10': h10, 'head_mr': hmr, 'head_mrr': hmrr,
            'hits@1': (t1 + h1) / 2, 'hits@3': (t3 + h3) / 2, 'hits@10': (t10 + h10) / 2,
            'mr': (tmr + hmr) / 2, 'mrr': (tmrr + hmrr) / 2}


class KGEValidateCallback:
    """utils.py:267-285 with its wiring fixed (the reference passes train_data as `bs` and never sets the filter):
    every 5th epoch and at train end, writes val_<metric> into the logs."""

    def __init__(self, validation_data, train_data=None, *args, **kwargs):
        self.validation_data = validation_data
        self.train_data = train_data
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        if epoch % 5 == 0:
            logs = logs if logs is not None else {}
            tmp = validate(self.model, self.validation_data, self.model.num_entities, None,
                           self.train_data if self.train_data is not None else [])
            for k in tmp:
                logs['val_' + k] = tmp[k]

    def on_train_end(self, logs=None):
        self.on_epoch_end(100, logs=logs)


# ---- losses (GPU) ----------------------------------------------------------------------------------------------
_loss_engine = None


def _loss_eng():
    global _loss_engine
    if _loss_engine is None:
        from ._engine import Engine
        _loss_engine = Engine(abi.DISTMULT, 1, 1, 4, 4)
    return _loss_engine


def _as_dev(x):
    import torch
    if isinstance(x, torch.Tensor):
        return x.detach().to("cuda", torch.float32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).cuda()


def _pair_layout(true, false, negative_samples):
    """Reproduce the broadcast `false.reshape(-1,k) - true` (utils.py:301-302): returns ('paired'|'allpairs', t, f, shape)."""
    t, f = _as_dev(true), _as_dev(false)
    F = f.reshape(-1, negative_samples)
    shape = tuple(np.broadcast_shapes(tuple(F.shape), tuple(t.shape)))  # raises ValueError exactly when TF/NumPy would
    if t.dim() == 2 and t.shape[1] == 1 and t.shape[0] == F.shape[0]:
        return "paired", t.reshape(-1), f.reshape(-1), shape
    if t.dim() == 1 and negative_samples == 1:
        return "allpairs", t, f.reshape(-1), shape
    if t.numel() == 1:
        return "allpairs", t.reshape(-1), f.reshape(-1), shape
    raise NotImplementedError(f"pairwise loss broadcast {tuple(F.shape)} vs {tuple(t.shape)} is not served by the kernels")


def _loss(kind, true, false, margin, negative_samples, reduce_mean):
    import torch
    eng = _loss_eng()
    if not abi.is_pairwise(kind):
        out, _, _ = eng.loss(_as_dev(true).reshape(-1), _as_dev(false).reshape(-1), 1, kind, abi.PAIRED, float(margin), 1.0, want_grads=False)
        return out[0]
    mode, t, f, shape = _pair_layout(true, false, negative_samples)
    if reduce_mean:
        pm = abi.PAIRED if mode == "paired" else abi.ALLPAIRS
        out, _, _ = eng.loss(t, f, negative_samples if mode == "paired" else 1, kind, pm, float(margin), 1.0, want_grads=False)
        return out[0]
    # un-reduced matrix requested: elementwise on device from the kernel's definition (small; not a hot path)
    X = f.reshape(-1, negative_samples) - (t.reshape(-1, 1) if mode == "paired" else t.reshape(1, -1))
    if kind == abi.PAIRWISE_HINGE:
        return torch.relu(float(margin) + X)
    if kind == abi.PAIRWISE_LOGISTIC:
        return torch.log(torch.tensor(EPSILON + 1, dtype=torch.float32, device=X.device) + torch.exp(X))
    return -X * X


def pointwize_hinge(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:288-289"""
    return _loss(abi.POINTWISE_HINGE, true, false, margin, negative_samples, True)


def pointwize_logistic(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:291-292"""
    return _loss(abi.POINTWISE_LOGISTIC, true, false, margin, negative_samples, True)


def pointwize_square_loss(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:294-295"""
    return _loss(abi.POINTWISE_SQUARE, true, false, margin, negative_samples, True)


def pointwize_cross_entropy(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:297-298: Keras binary_crossentropy on raw scores (clip to [1e-7, 1-1e-7], -log(. + 1e-7)), mean over each list."""
    return _loss(abi.POINTWISE_BCE, true, false, margin, negative_samples, True)


def pairwize_hinge(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:300-305"""
    return _loss(abi.PAIRWISE_HINGE, true, false, margin, negative_samples, reduce_mean)


def pairwize_logistic(true, false, margin=0, negative_samples=1, reduce_mean=True):
    """utils.py:307-312"""
    return _loss(abi.PAIRWISE_LOGISTIC, true, false, margin, negative_samples, reduce_mean)


def pairwize_square_loss(true, false, margin=0, negative_samples=1, reduce_mean=True):
    """utils.py:314-319"""
    return _loss(abi.PAIRWISE_SQUARE, true, false, margin, negative_samples, reduce_mean)


def loss_function_lookup(name):
    """utils.py:321-330 (KeyError on an unknown name, as there)."""
    return {
        'pointwize_hinge': pointwize_hinge,
        'pointwize_logistic': pointwize_logistic,
        'pointwize_cross_entropy': pointwize_cross_entropy,
        'pointwize_square_loss': pointwize_square_loss,
        'pairwize_hinge': pairwize_hinge,
        'pairwize_logistic': pairwize_logistic,
        'pairwize_square_loss': pairwize_square_loss,
    }[name]
