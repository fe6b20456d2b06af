"""Device-side state of one embedding model and the calls into libkgeb200.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only; every number comes out of the
hand-written sm_100a kernels behind the C ABI.  There is no CPU path: constructing an Engine without CUDA raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _abi as abi


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


_OPT_SPLIT = os.environ.get("KGEB_OPT_FUSED", "1") == "0"
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """The caller's current CUDA stream as a raw handle.  torch.cuda.current_stream() builds a Stream object through
    several Python layers (12 us a call, measured: a third of a train step's launch path); the raw getter is ~0.3 us."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def as_idx(x, device):
    """(m,3) ids in any int/float dtype (the harness feeds float32 ids, tests.py:72-73) -> int32 device tensor."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, non_blocking=True)
    else:
        a = np.asarray(x)
        if a.ndim == 3 and a.shape[-1] == 1:  # generate_negative returns (n,3,1), utils.py:120
            a = a[..., 0]
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device=device, non_blocking=True)
    if t.dim() == 3 and t.shape[-1] == 1:
        t = t[..., 0]
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError(f"expected triples of shape (batch, 3), got {tuple(t.shape)}")
    return t.to(torch.int32).contiguous()


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (rank 0 calls this and ships the bytes to the other ranks)."""
    buf = C.create_string_buffer(abi.COMM_ID_BYTES)
    abi.check(abi.lib().kgeb_comm_unique_id(buf))
    return buf.raw


_CHECK_TRAIN_IDS = os.environ.get("KGEB_CHECK_IDS", "0") == "1"


def check_ids(tri, num_entities, num_relations, what="triples"):
    """Raise IndexError for ids outside the tables (the reference's Embedding lookup raises InvalidArgumentError on CPU and reads
    zeros on GPU; the kernels here would read or reduce out of bounds).  One small reduction + read-back: used on the evaluation
    inputs always and on training batches when KGEB_CHECK_IDS=1."""
    if tri.numel() == 0:
        return
    lo = tri.amin(dim=0)
    hi = tri.amax(dim=0)
    v = torch.stack([lo, hi]).cpu().tolist()
    if v[0][0] < 0 or v[0][1] < 0 or v[0][2] < 0 or v[1][0] >= num_entities or v[1][2] >= num_entities or v[1][1] >= num_relations:
        raise IndexError(f"{what}: ids out of range (entities must be in [0, {num_entities}), relations in [0, {num_relations}); "
                         f"column minima {v[0]}, maxima {v[1]})")


class Engine:
    """Tables + gradient/optimizer buffers of one model on one GPU, bound to a kgeb context."""

    def __init__(self, kind, num_entities, num_relations, we, wr, gamma=12.0, device=None, E=None, R=None, rng=None, param=1.0):
        if not torch.cuda.is_available():
            raise RuntimeError("KGEkeras (B200 build) needs a CUDA device: there is no CPU fallback")
        self.lib = abi.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.kind, self.N, self.R_, self.we, self.wr, self.gamma = kind, int(num_entities), int(num_relations), int(we), int(wr), float(gamma)
        # zero-padded columns are neutral for DistMult/TransE, so odd widths are padded to the 128-bit granule
        pad_ok = kind in (abi.DISTMULT, abi.TRANSE_L1, abi.TRANSE_L2, abi.PROTATE, abi.MODE_L1, abi.MODE_L2, abi.HAKE)
        self.we_k = (self.we + 3) // 4 * 4 if pad_ok else self.we
        self.wr_k = (self.wr + 3) // 4 * 4 if pad_ok else self.wr
        rng = rng or np.random.default_rng()
        self.E = self._table(E, self.N, self.we, self.we_k, rng)
        self.Rel = self._table(R, self.R_, self.wr, self.wr_k, rng)
        h = C.c_void_p()
        # HAKE rows are [phase | modulus (| bias)]: the kernels split at the TRUE width, the padding only rounds the stride
        cw = (self.we, self.wr) if kind == abi.HAKE else (self.we_k, self.wr_k)
        abi.check(self.lib.kgeb_create(C.byref(h), self.device.index or 0, self.N, self.R_, cw[0], cw[1],
                                       self.we_k, self.wr_k, kind, self.gamma))
        self.ctx = h
        if kind == abi.HAKE:   # param = (embedding_range / pi, phase_weight, mod_weight), models.py:377-381,390-392
            self.param = tuple(float(x) for x in param)
            self._check(self.lib.kgeb_set_hake_params(self.ctx, *self.param))
        else:
            self.param = float(param)
            if self.param != 1.0:
                self._check(self.lib.kgeb_set_score_param(self.ctx, self.param))
        self.hole_fft = False
        self.grad_mode, self._ss_ws = "red", None
        self.E2 = self.tE2 = None
        self.gE = self.gR = self.tE = self.tR = None
        self.opt_state = None
        self.step = 0
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _table(self, init, rows, w, wk, rng):
        if init is None:  # GlorotUniform, models.py:87-88
            lim = float(np.sqrt(6.0 / (rows + w)))
            init = rng.uniform(-lim, lim, size=(rows, w)).astype(np.float32)
        if isinstance(init, torch.Tensor):      # device tensors are copied on the device (a 2 GB table must not visit the host)
            a = init.detach().to(torch.float32)
        else:
            a = torch.as_tensor(np.asarray(init, dtype=np.float32))
        if tuple(a.shape) != (rows, w):
            raise ValueError(f"initial weights have shape {tuple(a.shape)}, expected {(rows, w)}")
        t = torch.zeros(rows, wk, dtype=torch.float32, device=self.device)
        t[:, :w] = a.to(self.device)
        return t

    def __del__(self):
        try:
            for h in getattr(self, "_samplers", {}).values():
                self.lib.kgeb_destroy(h)
            self.__dict__.pop("_samplers", None)
            if getattr(self, "ctx", None):
                self.lib.kgeb_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- bookkeeping --------------------------------------------------------------------------------------------
    @property
    def num_sms(self):
        """SM count of the context's device (kgeb_num_sms)."""
        return int(self.lib.kgeb_num_sms(self.ctx))

    @property
    def launches(self):
        return int(self.lib.kgeb_launch_count(self.ctx))

    def _check(self, rc):
        abi.check(rc, self.ctx)

    def set_kernel_timing(self, on=True):
        self._check(self.lib.kgeb_set_kernel_timing(self.ctx, 1 if on else 0))

    def last_kernel_ms(self, which):
        ms = C.c_float()
        self._check(self.lib.kgeb_last_kernel_ms(self.ctx, which, C.byref(ms)))
        return float(ms.value)

    def kernel_ms_sum(self, which):
        """(total ms, launches) of that timed kernel since the previous call (kgeb_kernel_ms_sum)."""
        ms, n = C.c_float(), C.c_int()
        self._check(self.lib.kgeb_kernel_ms_sum(self.ctx, which, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def _ensure_train_buffers(self, opt_kind, need_gE=True):
        if self.gR is None:
            self.gR = torch.zeros_like(self.Rel)
            self.tE = torch.zeros(self.N, dtype=torch.uint8, device=self.device)
            self.tR = torch.zeros(self.R_, dtype=torch.uint8, device=self.device)
        if need_gE and self.gE is None:      # the fused sort-segment step never needs it
            self.gE = torch.zeros_like(self.E)
        if self.opt_state is None or self.opt_state[0] != opt_kind:
            if opt_kind == abi.SGD:
                st = (None, None, None, None)
            elif opt_kind == abi.ADAGRAD:  # Keras Adagrad: initial_accumulator_value = 0.1
                st = (torch.full_like(self.E, 0.1), None, torch.full_like(self.Rel, 0.1), None)
            else:
                st = (torch.zeros_like(self.E), torch.zeros_like(self.E), torch.zeros_like(self.Rel), torch.zeros_like(self.Rel))
            self.opt_state = (opt_kind,) + st
            self.step = 0

    def set_hole_mode(self, fft=True):
        """HolE through the Fourier domain (the reference's own formulation) for power-of-two widths; default = direct O(d^2)."""
        self._check(self.lib.kgeb_set_hole_mode(self.ctx, abi.HOLE_FFT if fft else abi.HOLE_DIRECT))
        self.hole_fft = bool(fft)

    def set_grad_mode(self, mode="red"):
        """'red' (default): every row reference adds its gradient row with a 128-bit vector reduction.  'sortseg': the entity
        gradient rows are sort-segmented by entity id and added once per touched row in a reproducible order (kgeb_set_grad_sortseg;
        fused-step kinds, no dropout).  'sortseg_fused': as 'sortseg', and train_step applies the optimizer at segment end
        (kgeb_train_step_sortseg): no entity gradient buffer, two copies of the entity table that swap every step -- the faster
        step for tables far larger than the L2 (C5: 1M x 512), slower on L2-resident ones."""
        if mode not in ("red", "sortseg", "sortseg_fused"):
            raise ValueError(f"grad mode must be 'red', 'sortseg' or 'sortseg_fused', got {mode!r}")
        self.grad_mode = mode
        if mode == "red":
            self._check(self.lib.kgeb_set_grad_sortseg(self.ctx, None, 0))
            self._ss_ws = None
        if mode == "sortseg_fused" and getattr(self, "E2", None) is None:
            self.E2 = self.E.clone()                       # the copy the next step writes; both start identical
            self.tE2 = torch.zeros(self.N, dtype=torch.uint8, device=self.device)
        if mode != "sortseg_fused":
            self.E2 = self.tE2 = None

    def _fused_sortseg_ok(self, neg, B, K, opt_kind, loss_kind, pair_mode, dropout, l3):
        return (self.grad_mode == "sortseg_fused" and self.kind in abi.FUSED_TRAIN_KINDS and opt_kind in (abi.SGD, abi.ADAGRAD, abi.LAZY_ADAM)
                and not dropout and not l3 and not (abi.is_pairwise(loss_kind) and pair_mode == abi.ALLPAIRS)
                and (neg is None or neg.shape[0] == B * K))

    def _arm_sortseg(self, B, K):
        need = int(self.lib.kgeb_sortseg_workspace_bytes(self.ctx, B, K))
        if need < 0:
            self._check(need)
        if self._ss_ws is None or self._ss_ws.numel() < need:
            self._ss_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._check(self.lib.kgeb_set_grad_sortseg(self.ctx, _ptr(self._ss_ws), self._ss_ws.numel()))

    # ---- collectives behind the C ABI (comm.cu) --------------------------------------------------------------------
    comm_rank, comm_world = 0, 1

    def comm_init(self, id_bytes, rank, world):
        """Join an NCCL communicator through the library's own binding (kgeb_comm_init): no torch.distributed involved.
        `id_bytes` = the 128 bytes rank 0 obtained from `comm_unique_id()` and handed to the other ranks."""
        if len(id_bytes) != abi.COMM_ID_BYTES:
            raise ValueError(f"an NCCL unique id has {abi.COMM_ID_BYTES} bytes, got {len(id_bytes)}")
        buf = C.create_string_buffer(bytes(id_bytes), abi.COMM_ID_BYTES)
        self._check(self.lib.kgeb_comm_init(self.ctx, buf, int(rank), int(world)))
        self.comm_rank, self.comm_world = int(rank), int(world)
        return self

    def comm_init_from_torch(self, group=None):
        """Same, with the unique id carried to the other ranks by an existing torch.distributed group (one broadcast)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return self.comm_init(box[0], rank, world)

    def comm_destroy(self):
        """Orderly teardown.  NCCL waits for every CUDA graph that captured one of the communicator's collectives, so the
        captured passes of this engine's RankPlans are dropped first (they are re-captured on demand)."""
        for plan in list(getattr(self, "_plans", ())):
            plan.graph = None
        torch.cuda.synchronize()
        self._check(self.lib.kgeb_comm_destroy(self.ctx))
        self.comm_rank, self.comm_world = 0, 1

    def allreduce_counts(self, counts):
        """In-place sum of int32 partial rank counts over the communicator's ranks, on the current stream."""
        if counts.dtype != torch.int32 or not counts.is_contiguous():
            raise ValueError("allreduce_counts needs a contiguous int32 tensor")
        self._check(self.lib.kgeb_allreduce_counts(self.ctx, _ptr(counts), counts.numel(), _stream()))
        return counts

    # ---- dropout ------------------------------------------------------------------------------------------------
    def set_dropout(self, p=0.0, seed=0, offset=0, stream_id=0):
        """models.py:122,144-150: arm (p > 0) or disarm (p = 0) dropout for the next scoring / training calls."""
        self._check(self.lib.kgeb_set_dropout(self.ctx, float(p), int(seed), int(offset), int(stream_id)))

    # ---- scoring ------------------------------------------------------------------------------------------------
    def score(self, idx, out=None, dropout=None):
        """dropout = (p, seed, offset, stream_id) scores the dropped rows of a training-mode call; None = inference."""
        idx = as_idx(idx, self.device)
        m = idx.shape[0]
        out = torch.empty(m, dtype=torch.float32, device=self.device) if out is None else out
        if dropout is not None:
            self.set_dropout(*dropout)
        try:
            self._check(self.lib.kgeb_score_fwd(self.ctx, _ptr(self.E), _ptr(self.Rel), _ptr(idx), m, _ptr(out), _stream()))
        finally:
            if dropout is not None:
                self.set_dropout(0.0)
        return out

    def score_bwd(self, idx, dscore, dropout=None):
        idx = as_idx(idx, self.device)
        self._ensure_train_buffers(self.opt_state[0] if self.opt_state else abi.SGD)
        dscore = dscore.to(device=self.device, dtype=torch.float32).contiguous()
        if dropout is not None:
            self.set_dropout(*dropout)
        try:
            self._check(self.lib.kgeb_score_bwd(self.ctx, _ptr(self.E), _ptr(self.Rel), _ptr(idx), idx.shape[0], _ptr(dscore),
                                                _ptr(self.gE), _ptr(self.gR), _ptr(self.tE), _ptr(self.tR), _stream()))
        finally:
            if dropout is not None:
                self.set_dropout(0.0)

    def negatives(self, pos, K, seed, offset, corrupt_mode=abi.CORRUPT_BOTH, num_entities=None):
        """Philox negatives.  num_entities bounds the corrupted ids ([0, num_entities), default: the whole vocabulary) -- the
        harness draws below the number of DISTINCT entities of its split (tests.py:34), which can be smaller than the table."""
        pos = as_idx(pos, self.device)
        neg = torch.empty(pos.shape[0] * K, 3, dtype=torch.int32, device=self.device)
        ctx = self.ctx
        if num_entities is not None and int(num_entities) != self.N:
            ctx = self._sampler_ctx(int(num_entities))
        abi.check(self.lib.kgeb_negatives(ctx, _ptr(pos), pos.shape[0], K, seed, offset, corrupt_mode, _ptr(neg), _stream()), ctx)
        return neg

    def _sampler_ctx(self, n):
        """A table-less context whose entity range is [0, n): only kgeb_negatives is ever called on it."""
        if not 0 < n <= self.N:
            raise ValueError(f"negative-sampling range {n} outside (0, {self.N}]")
        cache = self.__dict__.setdefault("_samplers", {})
        if n not in cache:
            h = C.c_void_p()
            abi.check(self.lib.kgeb_create(C.byref(h), self.device.index or 0, n, self.R_, self.we_k, self.wr_k, self.we_k, self.wr_k,
                                           self.kind if self.kind != abi.HAKE else abi.DISTMULT, self.gamma))
            cache[n] = h
        return cache[n]

    def oversample(self, src, m, seed, offset, stream_id=0):
        """oversample_data (utils.py:123-147) on the device: m rows drawn from src with replacement."""
        src = as_idx(src, self.device)
        dst = torch.empty(int(m), 3, dtype=torch.int32, device=self.device)
        self._check(self.lib.kgeb_oversample(self.ctx, _ptr(src), src.shape[0], int(m), seed, offset, stream_id, _ptr(dst), _stream()))
        return dst

    def device_batch(self, kg_batch, negative, seed, step, num_entities=None):
        """The harness' DataGenerator.__get_data (tests.py:55-61) without leaving the GPU: generate_negative (both slots
        corrupted, relation kept) followed by oversample_data of both lists to the common length m = negative*len(batch).
        Returns [pos (m,3), neg (m,3)] int32 device tensors -- the unaligned lists the reference feeds its all-pairs loss."""
        pos = as_idx(kg_batch, self.device)
        neg = self.negatives(pos, int(negative), seed, step, abi.CORRUPT_BOTH, num_entities=num_entities)
        m = max(pos.shape[0], neg.shape[0])
        return [self.oversample(pos, m, seed, step, 0), self.oversample(neg, m, seed, step, 1)]

    def loss(self, t, f, K, loss_kind, pair_mode, margin, loss_weight=1.0, want_grads=True):
        t = t.to(device=self.device, dtype=torch.float32).contiguous().view(-1)
        f = f.to(device=self.device, dtype=torch.float32).contiguous().view(-1)
        out = torch.empty(1, dtype=torch.float32, device=self.device)
        dt = torch.empty_like(t) if want_grads else None
        df = torch.empty_like(f) if want_grads else None
        self._check(self.lib.kgeb_loss_fwd_bwd(self.ctx, _ptr(t), t.numel(), _ptr(f), f.numel(), K, loss_kind, pair_mode,
                                               margin, loss_weight, _ptr(out), _ptr(dt), _ptr(df), _stream()))
        return out, dt, df

    # ---- training -----------------------------------------------------------------------------------------------
    def forward_backward(self, pos, neg=None, K=1, loss_kind=abi.PAIRWISE_HINGE, pair_mode=abi.PAIRED, margin=1.0,
                         loss_weight=1.0, corrupt_mode=abi.CORRUPT_BOTH, seed=0, offset=0, want_scores=False,
                         want_neg=False, force_unfused=False, opt_kind=abi.ADAM_DENSE, dropout=0.0):
        """Scores, loss and gradient scatter of one batch (no parameter update).  Returns a dict of device tensors.
        dropout > 0 applies Dropout(dp) to every gathered row (masks keyed by seed / offset, independent for the positive
        and the negative call, models.py:144-150)."""
        pos = as_idx(pos, self.device)
        B = pos.shape[0]
        if _CHECK_TRAIN_IDS:
            check_ids(pos, self.N, self.R_, "positive triples")
        if neg is not None:
            neg = as_idx(neg, self.device)
            if _CHECK_TRAIN_IDS:
                check_ids(neg, self.N, self.R_, "negative triples")
            if pair_mode == abi.PAIRED or not abi.is_pairwise(loss_kind):
                if neg.shape[0] % B != 0 and abi.is_pairwise(loss_kind):
                    raise ValueError("PAIRED pairwise loss needs len(neg) to be a multiple of len(pos)")
                if neg.shape[0] % B == 0:
                    K = neg.shape[0] // B
        self._ensure_train_buffers(opt_kind)
        pairwise = abi.is_pairwise(loss_kind)
        fused = (not force_unfused and self.kind in abi.FUSED_TRAIN_KINDS and not (pairwise and pair_mode == abi.ALLPAIRS)
                 and (neg is None or neg.shape[0] == B * K))
        if (not force_unfused and self.kind == abi.HOLE and self.hole_fft and not pairwise and (neg is None or neg.shape[0] == B * K)
                and 32 <= self.we <= 1024 and self.we & (self.we - 1) == 0):
            fused = True      # Fourier-domain HolE: one fused kernel for the pointwise losses (hole_fft_kernels.cu)
        res = {}
        if fused:
            ps = torch.empty(B, dtype=torch.float32, device=self.device) if want_scores else None
            ns = torch.empty(B * K, dtype=torch.float32, device=self.device) if want_scores else None
            no = torch.empty(B * K, 3, dtype=torch.int32, device=self.device) if (want_neg and neg is None) else None
            if dropout:
                self.set_dropout(dropout, seed, offset, 0)
            if self.grad_mode == "sortseg" and self.kind in abi.FUSED_TRAIN_KINDS:
                self._arm_sortseg(B, K)
            try:
                self._check(self.lib.kgeb_train_fwd_bwd(
                    self.ctx, _ptr(self.E), _ptr(self.Rel), _ptr(pos), B, _ptr(neg), K, loss_kind, margin, loss_weight,
                    corrupt_mode, seed, offset, _ptr(self._loss), _ptr(ps), _ptr(ns), _ptr(no), _ptr(self.gE), _ptr(self.gR),
                    _ptr(self.tE), _ptr(self.tR), _stream()))
            finally:
                if dropout:
                    self.set_dropout(0.0)
            res.update(loss=self._loss, pos_scores=ps, neg_scores=ns, neg=neg if neg is not None else no)
        else:
            if neg is None:
                neg = self.negatives(pos, K, seed, offset, corrupt_mode)
            dpos = (dropout, seed, offset, 0) if dropout else None
            dneg = (dropout, seed, offset, 1) if dropout else None
            t, f = self.score(pos, dropout=dpos), self.score(neg, dropout=dneg)
            loss, dt, df = self.loss(t, f, K, loss_kind, pair_mode, margin, loss_weight)
            self.score_bwd(pos, dt, dropout=dpos)
            self.score_bwd(neg, df, dropout=dneg)
            res.update(loss=loss, pos_scores=t, neg_scores=f, neg=neg)
        return res

    def apply_optimizer(self, opt_kind=abi.ADAM_DENSE, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self._ensure_train_buffers(opt_kind)
        self.step += 1
        _, m_e, v_e, m_r, v_r = self.opt_state
        if _OPT_SPLIT:   # A/B switch (KGEB_OPT_FUSED=0 at import): one launch per table
            self._check(self.lib.kgeb_opt_update(self.ctx, _ptr(self.E), _ptr(self.gE), _ptr(self.tE), self.N, self.we_k, self.we_k,
                                                 _ptr(m_e), _ptr(v_e), opt_kind, lr, beta1, beta2, eps, self.step, _stream()))
            self._check(self.lib.kgeb_opt_update(self.ctx, _ptr(self.Rel), _ptr(self.gR), _ptr(self.tR), self.R_, self.wr_k, self.wr_k,
                                                 _ptr(m_r), _ptr(v_r), opt_kind, lr, beta1, beta2, eps, self.step, _stream()))
            return
        # entity and relation table in one launch
        self._check(self.lib.kgeb_opt_update2(self.ctx, _ptr(self.E), _ptr(self.gE), _ptr(self.tE), self.N, self.we_k, self.we_k,
                                              _ptr(m_e), _ptr(v_e), _ptr(self.Rel), _ptr(self.gR), _ptr(self.tR), self.R_, self.wr_k,
                                              self.wr_k, _ptr(m_r), _ptr(v_r), opt_kind, lr, beta1, beta2, eps, self.step, _stream()))

    def export_grad_rows(self, table="entity", clear=True):
        """The backward's result as (unique_ids, grad_rows) -- the IndexedSlices a tf.keras Embedding hands its optimizer
        (models.py:89-112) -- instead of an in-library update: ids int32 ascending, rows float32 [len(ids), width].
        clear=True leaves the dense buffer and its flags all-zero for the next step (kgeb_grad_rows_export)."""
        if self.gE is None:
            raise RuntimeError("no gradient buffer yet: run forward_backward / score_bwd first")
        G, t, n, w, wk = (self.gE, self.tE, self.N, self.we, self.we_k) if table == "entity" else (self.gR, self.tR, self.R_, self.wr, self.wr_k)
        ids = torch.empty(n, dtype=torch.int32, device=self.device)
        rows = torch.empty(n, wk, dtype=torch.float32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._check(self.lib.kgeb_grad_rows_export(self.ctx, _ptr(G), _ptr(t), n, wk, wk, _ptr(ids), _ptr(rows), n, _ptr(cnt), 1 if clear else 0, _stream()))
        k = int(cnt.item())
        return ids[:k], rows[:k, :w]

    def l3_reg(self, w, loss):
        """models.py:10-11,79-82: loss += w*sum|E|^3, gE += 3 w E|E| on every entity row (the relation table is not regularised)."""
        self._check(self.lib.kgeb_l3_reg(self.ctx, _ptr(self.E), self.N, self.we_k, self.we_k, float(w), _ptr(self.gE), _ptr(self.tE),
                                         _ptr(loss), _stream()))

    def train_step(self, pos, neg=None, opt_kind=abi.ADAM_DENSE, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, l3=0.0, **kw):
        if self.grad_mode == "sortseg_fused":
            r = self._train_step_sortseg_fused(pos, neg, opt_kind, lr, beta1, beta2, eps, l3, **kw)
            if r is not None:
                return r
            # a request the fused entry does not serve (dense Adam, dropout, all-pairs loss, L3): ordinary in-place step on the
            # current copy, then the other copy is brought up to date wholesale (correct, not fast: leave this mode for such runs)
            self._check(self.lib.kgeb_set_grad_sortseg(self.ctx, None, 0))     # plain reduction path for this step
            self._ss_ws = None
            res = self.forward_backward(pos, neg, opt_kind=opt_kind, **kw)
            if l3:
                self.l3_reg(l3, res["loss"])
            self.apply_optimizer(opt_kind, lr, beta1, beta2, eps)
            self.E2.copy_(self.E)
            self.tE2.zero_()
            return res
        res = self.forward_backward(pos, neg, opt_kind=opt_kind, **kw)
        if l3:
            self.l3_reg(l3, res["loss"])
        self.apply_optimizer(opt_kind, lr, beta1, beta2, eps)
        return res

    def _train_step_sortseg_fused(self, pos, neg, opt_kind, lr, beta1, beta2, eps, l3, K=1, loss_kind=abi.PAIRWISE_HINGE, pair_mode=abi.PAIRED,
                                  margin=1.0, loss_weight=1.0, corrupt_mode=abi.CORRUPT_BOTH, seed=0, offset=0, dropout=0.0, **other):
        """The whole step through kgeb_train_step_sortseg; None when the request is outside what that entry serves (the caller
        then takes the ordinary two-kernel path on the CURRENT table copy)."""
        pos = as_idx(pos, self.device)
        B = pos.shape[0]
        if neg is not None:
            neg = as_idx(neg, self.device)
            if neg.shape[0] % B == 0:
                K = neg.shape[0] // B
        if other.get("want_scores") or other.get("want_neg") or other.get("force_unfused"):
            return None
        if not self._fused_sortseg_ok(neg, B, K, opt_kind, loss_kind, pair_mode, dropout, l3):
            return None
        self._ensure_train_buffers(opt_kind, need_gE=False)
        self._arm_sortseg(B, K)
        self.step += 1
        _, m_e, v_e, m_r, v_r = self.opt_state
        self._check(self.lib.kgeb_train_step_sortseg(
            self.ctx, _ptr(self.E), _ptr(self.E2), _ptr(self.Rel), _ptr(pos), B, _ptr(neg), K, loss_kind, margin, loss_weight, corrupt_mode,
            seed, offset, _ptr(self._loss), _ptr(self.gR), _ptr(self.tR), _ptr(self.tE), _ptr(self.tE2), _ptr(m_e), _ptr(v_e), _ptr(m_r), _ptr(v_r),
            opt_kind, lr, beta1, beta2, eps, self.step, _stream()))
        self.E, self.E2 = self.E2, self.E            # the copy just written is the table now
        self.tE, self.tE2 = self.tE2, self.tE        # this step's flags become "previous"; the cleared array is the next step's
        return {"loss": self._loss, "pos_scores": None, "neg_scores": None, "neg": neg}

    # ---- filter sets ----------------------------------------------------------------------------------------------
    def filter_csr(self, queries, filtering_triples, side):
        """Sorted, de-duplicated known-true ids per query (utils.py:201-205) in CSR form, built on the device by
        kgeb_filter_csr_count / kgeb_filter_csr_fill: (offsets int64[Q+1], ids int32[nnz]).  Raises ValueError when a query's own
        target is in its filter set -- the reference's list.remove crashes there (utils.py:169-171)."""
        q = as_idx(queries, self.device)
        f = as_idx(filtering_triples, self.device) if len(filtering_triples) else None
        check_ids(q, self.N, self.R_, "filter_csr queries")
        if f is not None:
            check_ids(f, self.N, self.R_, "filtering_triples")
        Q = q.shape[0]
        offs = torch.zeros(Q + 1, dtype=torch.int64, device=self.device)
        if f is None or Q == 0:
            return offs, torch.zeros(0, dtype=torch.int32, device=self.device)
        F = f.shape[0]
        nbytes = C.c_int64()
        self._check(self.lib.kgeb_filter_csr_workspace(self.ctx, Q, F, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        st = _stream()
        self._check(self.lib.kgeb_filter_csr_count(self.ctx, _ptr(q), Q, _ptr(f), F, side, _ptr(offs), _ptr(ws), nbytes.value, st))
        nnz = int(offs[-1].item())
        ids = torch.empty(nnz, dtype=torch.int32, device=self.device)
        if nnz:
            hit = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._check(self.lib.kgeb_filter_csr_fill(self.ctx, _ptr(q), Q, F, side, _ptr(offs), _ptr(ids), _ptr(hit), _ptr(ws), st))
            if int(hit.item()):
                raise ValueError("list.remove(x): x not in list")
        return offs, ids

    # ---- 1-N filtered ranking -----------------------------------------------------------------------------------
    def rank_counts(self, triples, side, filt_offsets=None, filt_ids=None, row_range=None, counts=None):
        """int32 partial counts #{c in rows, c not filtered : score > target score} for every query.
        row_range=(r0, r1) restricts the candidates to a row shard of the (replicated) table."""
        tri = as_idx(triples, self.device)
        Q = tri.shape[0]
        qw = self.lib.kgeb_query_width(self.ctx, side)
        qvec = torch.empty(Q, qw, dtype=torch.float32, device=self.device)   # every row is written: the whole table is local
        tscore = torch.empty(Q, dtype=torch.float32, device=self.device)
        counts = torch.zeros(Q, dtype=torch.int32, device=self.device) if counts is None else counts
        st = _stream()
        self._check(self.lib.kgeb_query_prep(self.ctx, _ptr(self.E), 0, self.N, _ptr(self.Rel), _ptr(tri), Q, side, _ptr(qvec), st))
        self._check(self.lib.kgeb_target_scores(self.ctx, _ptr(self.E), 0, self.N, _ptr(qvec), _ptr(tri), Q, side, _ptr(tscore), st))
        r0, r1 = (0, self.N) if row_range is None else row_range
        fo = fi = None
        if filt_offsets is not None:
            fo = (filt_offsets if isinstance(filt_offsets, torch.Tensor) else torch.as_tensor(np.asarray(filt_offsets, dtype=np.int64))).to(self.device, torch.int64).contiguous()
            fi = (filt_ids if isinstance(filt_ids, torch.Tensor) else torch.as_tensor(np.asarray(filt_ids, dtype=np.int32))).to(self.device, torch.int32).contiguous()
            if fi.numel() == 0:
                fo = fi = None
        if r1 > r0:
            e_shard = C.c_void_p(self.E.data_ptr() + r0 * self.we_k * 4)
            self._check(self.lib.kgeb_rank_1n(self.ctx, e_shard, r0, r1 - r0, _ptr(qvec), _ptr(tscore), _ptr(tri), Q, side,
                                              _ptr(fo), _ptr(fi), _ptr(counts), st))
        self._keep = (qvec, tscore, tri, fo, fi)  # keep alive until the stream has consumed them
        return counts

    # ---- 1-N filtered ranking on the tensor cores (DistMult / ComplEx / HolE) -------------------------------------------
    def tc_supported(self, mode="exact"):
        """Kinds the tcgen05 ranking kernel serves: the dot-product scores in both modes, TransE-L2 in the exact mode only (its
        approximate single-plane mode is not built: a request for it is served by the exact CUDA-core kernel)."""
        if self.kind == abi.TRANSE_L2:
            return str(mode).lower() == "exact"
        return self.kind in (abi.DISTMULT, abi.COMPLEX, abi.HOLE, abi.ROTATE)

    def pack_entities_tc(self, mode=abi.TC_BF16X3_EXACT, row_range=None):
        """bf16 (hi/lo) image + row norms of the entity rows [r0, r1): build once per table version, reuse for every query batch."""
        r0, r1 = (0, self.N) if row_range is None else row_range
        nbytes = int(self.lib.kgeb_rank_tc_packed_bytes(self.ctx, r1 - r0, mode))
        if nbytes < 0:
            self._check(nbytes)
        packed = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        if r1 > r0:
            e_shard = C.c_void_p(self.E.data_ptr() + r0 * self.we_k * 4)
            self._check(self.lib.kgeb_rank_tc_pack_entities(self.ctx, e_shard, r1 - r0, mode, _ptr(packed), nbytes, _stream()))
        return packed

    def rank_counts_tc(self, triples, side, filt_offsets=None, filt_ids=None, row_range=None, counts=None,
                       mode=abi.TC_BF16X3_EXACT, packed=None, worklist_cap=None, stats=None):
        """Same contract as rank_counts, computed by the tcgen05 kernel.  In TC_BF16X3_EXACT mode the integer counts are
        identical to rank_counts'; TC_BF16 is approximate.  The undecided-pair worklist is sized by a heuristic and the
        pass is repeated with a larger one if it overflowed (checked on the host: one 8-byte read-back per pass)."""
        tri = as_idx(triples, self.device)
        Q = tri.shape[0]
        qw = self.lib.kgeb_query_width(self.ctx, side)
        qvec = torch.empty(Q, qw, dtype=torch.float32, device=self.device)   # every row is written: the whole table is local
        tscore = torch.empty(Q, dtype=torch.float32, device=self.device)
        st = _stream()
        self._check(self.lib.kgeb_query_prep(self.ctx, _ptr(self.E), 0, self.N, _ptr(self.Rel), _ptr(tri), Q, side, _ptr(qvec), st))
        self._check(self.lib.kgeb_target_scores(self.ctx, _ptr(self.E), 0, self.N, _ptr(qvec), _ptr(tri), Q, side, _ptr(tscore), st))
        r0, r1 = (0, self.N) if row_range is None else row_range
        fo = fi = None
        if filt_offsets is not None:
            fo = (filt_offsets if isinstance(filt_offsets, torch.Tensor) else torch.as_tensor(np.asarray(filt_offsets, dtype=np.int64))).to(self.device, torch.int64).contiguous()
            fi = (filt_ids if isinstance(filt_ids, torch.Tensor) else torch.as_tensor(np.asarray(filt_ids, dtype=np.int32))).to(self.device, torch.int32).contiguous()
            if fi.numel() == 0:
                fo = fi = None
        out = torch.zeros(Q, dtype=torch.int32, device=self.device) if counts is None else counts
        if r1 <= r0 or Q == 0:
            return out
        if packed is None:
            packed = self.pack_entities_tc(mode, (r0, r1))
        cap = int(worklist_cap) if worklist_cap else min(2 ** 31 - 2, max(1 << 20, Q * (64 + (r1 - r0) // 128)))
        e_shard = C.c_void_p(self.E.data_ptr() + r0 * self.we_k * 4)
        while True:
            part = torch.zeros(Q, dtype=torch.int32, device=self.device)
            nbytes = int(self.lib.kgeb_rank_tc_workspace_bytes(self.ctx, Q, mode, cap))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._check(self.lib.kgeb_rank_1n_tc(self.ctx, e_shard, r0, r1 - r0, _ptr(packed), _ptr(qvec), _ptr(tscore), _ptr(tri), Q,
                                                 side, _ptr(fo), _ptr(fi), mode, _ptr(ws), nbytes, cap, _ptr(part), st))
            appended = int(ws[:8].view(torch.int32)[0].item()) & 0xFFFFFFFF
            if stats is not None:
                stats.update(undecided_pairs=appended, worklist_cap=cap)
            if appended <= cap:
                break
            cap = min(2 ** 31 - 2, appended + appended // 4)
        out += part
        self._keep = (qvec, tscore, tri, fo, fi, ws, packed)
        return out

    def rank_plan(self, triples, filt_tail=None, filt_head=None, row_range=None, mode="exact", group=None, worklist_cap=None,
                  packed=None):
        """A reusable filtered-ranking pass over a fixed query set (both sides): see RankPlan."""
        return RankPlan(self, triples, filt_tail, filt_head, row_range, mode, group, worklist_cap, packed)

    # ---- weights ------------------------------------------------------------------------------------------------
    def get_tables(self):
        return self.E[:, :self.we].cpu().numpy(), self.Rel[:, :self.wr].cpu().numpy()

    def set_tables(self, E=None, R=None):
        if E is not None:
            self.E[:, :self.we] = torch.as_tensor(np.asarray(E, dtype=np.float32)).to(self.device)
            if getattr(self, "E2", None) is not None:
                self.E2.copy_(self.E)
        if R is not None:
            self.Rel[:, :self.wr] = torch.as_tensor(np.asarray(R, dtype=np.float32)).to(self.device)


class RankPlan:
    """One filtered 1-N ranking pass (tail AND head side) over a fixed set of test triples, with everything it needs
    allocated once: query vectors, target scores, the packed table image, worklists, the [2, Q] count buffer.

    `run()` only enqueues work on the caller's stream -- query prep, target scores, the ranking kernels of both sides, the
    exact re-scoring -- and, when the entity rows are sharded over a process group, ONE all-reduce of the 2*Q integer
    partial counts.  No host synchronisation, no allocation, no read-back happens inside a pass; whether a
    worklist overflowed is read afterwards (`overflowed()`, two 8-byte headers) and `ranks()` repeats the pass with a larger
    worklist in that (rare) case.  `capture()` records the pass into a CUDA graph, so that the ~14 launches of a small
    shape cost one graph launch.

    mode: "exact" (tcgen05, integer counts identical to the CUDA-core kernel), "bf16" (tcgen05, approximate),
          "cuda" (fp32 CUDA-core kernels; the only mode for TransE and the other norm-based kinds)."""

    def __init__(self, eng, triples, filt_tail=None, filt_head=None, row_range=None, mode="exact", group=None, worklist_cap=None,
                 packed=None):
        self.eng, self.group = eng, group
        self.tri = as_idx(triples, eng.device)
        check_ids(self.tri, eng.N, eng.R_, "test triples")
        self.Q = Q = int(self.tri.shape[0])
        self.r0, self.r1 = (0, eng.N) if row_range is None else (int(row_range[0]), int(row_range[1]))
        mode = str(mode).lower()
        if mode not in ("exact", "bf16", "cuda"):
            raise ValueError(f"rank mode must be 'exact', 'bf16' or 'cuda', got {mode!r}")
        self.use_tc = mode != "cuda" and eng.tc_supported(mode) and self.r1 > self.r0 and Q > 0
        self.mode = mode if self.use_tc else "cuda"
        self.tc_mode = abi.TC_BF16 if mode == "bf16" else abi.TC_BF16X3_EXACT
        dev = eng.device

        def csr(f):
            if f is None:
                return None, None
            fo = (f[0] if isinstance(f[0], torch.Tensor) else torch.as_tensor(np.asarray(f[0], dtype=np.int64))).to(dev, torch.int64).contiguous()
            fi = (f[1] if isinstance(f[1], torch.Tensor) else torch.as_tensor(np.asarray(f[1], dtype=np.int32))).to(dev, torch.int32).contiguous()
            return (None, None) if fi.numel() == 0 else (fo, fi)
        self.filt = [csr(filt_tail), csr(filt_head)]
        self.counts = torch.zeros(2, max(Q, 1), dtype=torch.int32, device=dev)
        self.qvec = [torch.empty(max(Q, 1), eng.lib.kgeb_query_width(eng.ctx, sd), dtype=torch.float32, device=dev) for sd in (0, 1)]
        self.tscore = [torch.empty(max(Q, 1), dtype=torch.float32, device=dev) for _ in (0, 1)]
        self.packed = packed
        self.ws = [None, None]
        self.cap = 0
        if self.use_tc:
            if self.packed is None:
                self.packed = eng.pack_entities_tc(self.tc_mode, (self.r0, self.r1))
            cap = int(worklist_cap) if worklist_cap else min(2 ** 31 - 2, max(1 << 20, Q * (64 + (self.r1 - self.r0) // 128)))
            self._alloc_ws(cap)
        self.graph = None
        # run the two sides as parallel branches when a side is too small to fill the GPU (see _enqueue_sides); KGEB_RANK_FORK=0/1 overrides
        fk = os.environ.get("KGEB_RANK_FORK")
        self.fork_sides = (Q * max(self.r1 - self.r0, 0) < 10 ** 9) if fk is None else fk == "1"
        self._side_streams = None
        if not hasattr(eng, "_plans"):
            import weakref
            eng._plans = weakref.WeakSet()
        eng._plans.add(self)
        self.dist_on = False
        # the partial counts are summed by the library's own NCCL binding when the engine joined a communicator
        # (kgeb_comm_init: the all-reduce is then part of the enqueued pass and of its CUDA graph), else by torch.distributed
        self.kgeb_comm = eng.comm_world > 1
        if not self.kgeb_comm and (group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.dist_on = torch.distributed.get_world_size(group) > 1

    def _alloc_ws(self, cap):
        eng = self.eng
        self.cap = int(cap)
        nbytes = int(eng.lib.kgeb_rank_tc_workspace_bytes(eng.ctx, self.Q, self.tc_mode, self.cap))
        if nbytes < 0:
            eng._check(nbytes)
        self.ws_bytes = nbytes
        self.ws = [torch.empty(nbytes, dtype=torch.uint8, device=eng.device) for _ in (0, 1)]
        self.graph = None

    def _enqueue(self):
        eng, lib, Q = self.eng, self.eng.lib, self.Q
        st = _stream()
        self.counts.zero_()
        if Q == 0:
            return
        self._enqueue_sides(eng, lib, Q, st)
        if self.kgeb_comm:
            eng._check(lib.kgeb_allreduce_counts(eng.ctx, _ptr(self.counts), self.counts.numel(), st))

    def _enqueue_sides(self, eng, lib, Q, st):
        e_shard = C.c_void_p(eng.E.data_ptr() + self.r0 * eng.we_k * 4)
        if self.fork_sides:
            # the two sides share nothing but read-only inputs (own query vectors, target scores, worklist, count row): on small shards,
            # where a side's kernels do not fill the GPU and a pass is mostly launch-to-launch latency, they run as two branches
            # (two streams joined by events; two parallel branches of the CUDA graph when captured)
            main = torch.cuda.current_stream()
            if self._side_streams is None:
                self._side_streams = [torch.cuda.Stream(device=eng.device) for _ in (0, 1)]
            for sd in (abi.SIDE_TAIL, abi.SIDE_HEAD):
                ss = self._side_streams[sd]
                ss.wait_stream(main)
                with torch.cuda.stream(ss):
                    self._enqueue_side(eng, lib, Q, C.c_void_p(ss.cuda_stream), e_shard, sd)
            for ss in self._side_streams:
                main.wait_stream(ss)
            return
        for sd in (abi.SIDE_TAIL, abi.SIDE_HEAD):
            self._enqueue_side(eng, lib, Q, st, e_shard, sd)

    def _enqueue_side(self, eng, lib, Q, st, e_shard, sd):
        qv, ts, (fo, fi), cnt = self.qvec[sd], self.tscore[sd], self.filt[sd], self.counts[sd]
        eng._check(lib.kgeb_query_prep(eng.ctx, _ptr(eng.E), 0, eng.N, _ptr(eng.Rel), _ptr(self.tri), Q, sd, _ptr(qv), st))
        eng._check(lib.kgeb_target_scores(eng.ctx, _ptr(eng.E), 0, eng.N, _ptr(qv), _ptr(self.tri), Q, sd, _ptr(ts), st))
        if self.r1 <= self.r0:
            return
        if self.use_tc:
            eng._check(lib.kgeb_rank_1n_tc(eng.ctx, e_shard, self.r0, self.r1 - self.r0, _ptr(self.packed), _ptr(qv), _ptr(ts),
                                           _ptr(self.tri), Q, sd, _ptr(fo), _ptr(fi), self.tc_mode, _ptr(self.ws[sd]), self.ws_bytes,
                                           self.cap, _ptr(cnt), st))
        else:
            eng._check(lib.kgeb_rank_1n(eng.ctx, e_shard, self.r0, self.r1 - self.r0, _ptr(qv), _ptr(ts), _ptr(self.tri), Q, sd,
                                        _ptr(fo), _ptr(fi), _ptr(cnt), st))

    def capture(self):
        """Record the device work of one pass into a CUDA graph; run() then replays it.  The all-reduce of the partial counts is
        part of the graph when it goes through the library's NCCL binding (kgeb_comm_init), and follows the replay otherwise."""
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._enqueue()          # warm-up outside capture: lazy attribute / module loading must not happen while capturing
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue()
        self.graph = g
        return self

    def run(self):
        """Enqueue one pass; returns the [2, Q] int32 device counts (tail, head), summed over the shards of `group`."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        if self.dist_on:
            torch.distributed.all_reduce(self.counts, group=self.group)
        return self.counts[:, :self.Q]

    def undecided(self):
        """(tail, head) number of pairs the last pass sent to the exact re-scoring (host read-back)."""
        if not self.use_tc:
            return (0, 0)
        return tuple(int(self.ws[sd][:4].view(torch.int32)[0].item()) & 0xFFFFFFFF for sd in (0, 1))

    def overflowed(self):
        return self.use_tc and max(self.undecided()) > self.cap

    def ranks(self):
        """Run to completion: (tail_ranks, head_ranks) int64 host arrays, rank = 1 + count."""
        while True:
            c = self.run()
            host = c.cpu().numpy()          # synchronises: the headers are final once this returns
            if not self.use_tc:
                break
            und = max(self.undecided())
            if self.dist_on:                # every rank must take the same branch
                t = torch.tensor([und > self.cap], dtype=torch.int32, device=self.eng.device)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self.group)
                over = bool(t.item())
            elif self.kgeb_comm:
                t = torch.tensor([int(und > self.cap)], dtype=torch.int32, device=self.eng.device)
                over = bool(self.eng.allreduce_counts(t).item())
            else:
                over = und > self.cap
            if not over:
                break
            self._alloc_ws(min(2 ** 31 - 2, max(und + und // 4, 2 * self.cap)))
        host = host.astype(np.int64) + 1
        return host[0], host[1]
