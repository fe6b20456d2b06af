"""Row-sharded training of one embedding model over the GPUs of a box (BASELINE configs[4], the synthetic
1M-entity table).  The reference has nothing like it (single process, `tests.py:130-134`); what is kept is the step's
arithmetic: the result equals ONE single-GPU step over the concatenation of all ranks' batches.

Layout: rank g owns the contiguous entity rows `shard_rows(N, world, g)` with their optimizer state; the small relation
table is replicated.  One synchronous step on every rank:

  1. negatives for the rank's own positives (kgeb_negatives, Philox, GLOBAL entity range);
  2. exchange plan (index bookkeeping only, torch ops): sorted unique entity ids the rank references, split by owner;
     the id lists travel with all_to_all_single;
  3. owners pack the requested rows (kgeb_rows_gather) -> all-to-all -> every rank holds a compact table of exactly
     the rows it needs, in id order;
  4. the ordinary fused kernel (kgeb_train_fwd_bwd) runs on the compact table with remapped ids;
  5. compact gradient rows -> all-to-all back -> owners add them into their shard gradient (kgeb_rows_scatter_add);
     the relation gradient and its touched flags are all-reduced;
  6. every owner updates its own rows (kgeb_opt_update), the replicated relation table is updated identically everywhere.

NCCL (torch.distributed) moves the bytes; no kernel reads another rank's memory.  The plan and the collectives are
device agnostic, so the CPU suite runs them with gloo at world 2 against the oracle (tests/test_cpu_dist_gloo.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _abi as abi
from .utils import shard_rows


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class TorchExchange:
    """The step's collectives through torch.distributed (NCCL on GPUs, gloo in the CPU suite)."""
    kind = "torch.distributed all_to_all_single"

    def __init__(self, group=None):
        self.group = group
        self.world, self.rank = _world(group)

    def all_to_all(self, out, inp, out_splits, in_splits):
        """out <- rows from every rank (out_splits[g] rows from rank g); inp rows go out (in_splits[g] rows to rank g)."""
        if self.world == 1:
            out.copy_(inp)
            return out
        dist.all_to_all_single(out, inp, output_split_sizes=list(out_splits), input_split_sizes=list(in_splits), group=self.group)
        return out

    def sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def max_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t


class KgebExchange:
    """The same collectives through the library's own NCCL binding (kgeb_alltoall_rows / kgeb_allreduce_*, comm.cu): what a
    host without torch.distributed uses.  `eng` is an Engine that joined a communicator (Engine.comm_init).  Index tensors
    (int64 ids, counts) travel as rows of two floats: NCCL moves bytes, nothing is converted."""
    kind = "libkgeb200 NCCL binding (grouped ncclSend/ncclRecv)"

    def __init__(self, eng):
        self.eng, self.lib = eng, eng.lib
        self.world, self.rank = eng.comm_world, eng.comm_rank

    def all_to_all(self, out, inp, out_splits, in_splits):
        if self.world == 1:
            out.copy_(inp)
            return out
        if out.dtype != inp.dtype or out.dtype not in (torch.float32, torch.int64):
            raise TypeError("KgebExchange moves float32 rows or int64 ids")
        width = (int(inp.shape[1]) if inp.dim() == 2 else 1) * (2 if inp.dtype == torch.int64 else 1)
        snd = (C.c_int64 * self.world)(*[int(x) for x in in_splits])
        rcv = (C.c_int64 * self.world)(*[int(x) for x in out_splits])
        self.eng._check(self.lib.kgeb_alltoall_rows(self.eng.ctx, _ptr(inp.contiguous()), snd, _ptr(out), rcv, width, _stream()))
        return out

    def sum_(self, t):
        if self.world > 1:
            if t.dtype == torch.float32 and t.is_contiguous():
                self.eng._check(self.lib.kgeb_allreduce_f32(self.eng.ctx, _ptr(t), t.numel(), _stream()))
            elif t.dtype == torch.int32 and t.is_contiguous():
                self.eng._check(self.lib.kgeb_allreduce_counts(self.eng.ctx, _ptr(t), t.numel(), _stream()))
            else:
                raise TypeError("KgebExchange.sum_ serves contiguous float32 / int32 tensors")
        return t

    def max_(self, t):
        if self.world > 1:
            if t.dtype == torch.uint8:
                self.eng._check(self.lib.kgeb_allreduce_flags(self.eng.ctx, _ptr(t), t.numel(), _stream()))
            else:
                raise TypeError("KgebExchange.max_ serves uint8 flags")
        return t


def all_to_all_rows(out, inp, out_splits, in_splits, group=None):
    """(kept for callers of the first build) out <- rows from every rank over torch.distributed."""
    return TorchExchange(group).all_to_all(out, inp, out_splits, in_splits)


class ExchangePlan:
    """Which entity rows this rank needs, from whom, and which of its own rows the others need (built once per step).

    uniq        int64 [U]   sorted unique global ids referenced by this rank (ascending => grouped by owner)
    remap       int64 [N]   global id -> position in uniq (valid for referenced ids only)
    req_counts  list[world] how many of uniq live on each owner
    serve_local int32 [S]   shard-local row index of every row the other ranks (and this one) request from this rank,
                            concatenated by requesting rank
    serve_counts list[world]
    """

    def __init__(self, ids, num_entities, group=None, exchange=None):
        ex = exchange if exchange is not None else TorchExchange(group)
        world, rank = ex.world, ex.rank
        dev = ids.device
        N = int(num_entities)
        mark = torch.zeros(N, dtype=torch.int32, device=dev)
        mark[ids.reshape(-1).long()] = 1
        csum = torch.cumsum(mark, 0)
        self.remap = csum - 1
        self.uniq = torch.nonzero(mark).reshape(-1)
        self.U = int(self.uniq.numel())
        bounds = torch.tensor([shard_rows(N, world, g)[1] for g in range(world)], dtype=torch.int64)
        at = csum[(bounds.clamp(min=1) - 1).to(dev)].cpu().tolist()     # one read-back: unique ids below each shard end
        upto = [0] + [int(a) if int(b) > 0 else 0 for a, b in zip(at, bounds.tolist())]
        self.req_counts = [upto[g + 1] - upto[g] for g in range(world)]
        req = torch.tensor(self.req_counts, dtype=torch.int64, device=dev)
        serve = torch.empty(world, dtype=torch.int64, device=dev)
        ex.all_to_all(serve, req, [1] * world, [1] * world)
        self.serve_counts = [int(x) for x in serve.cpu().tolist()]
        serve_ids = torch.empty(sum(self.serve_counts), dtype=torch.int64, device=dev)
        ex.all_to_all(serve_ids, self.uniq, self.serve_counts, self.req_counts)
        self.row0 = shard_rows(N, world, rank)[0]
        self.serve_local = (serve_ids - self.row0).to(torch.int32)

    def remap_triples(self, tri):
        """(m,3) global ids -> ids into the compact table (relation column untouched)."""
        t = tri.long()
        return torch.stack([self.remap[t[:, 0]], t[:, 1], self.remap[t[:, 2]]], dim=1).to(torch.int32).contiguous()


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class ShardedTrainer:
    """One model, entity rows sharded over the ranks of `group` (one process per GPU)."""

    def __init__(self, kind, num_entities, num_relations, we, wr, gamma=12.0, group=None, seed=0, E=None, R=None, comm=None):
        """comm = None: the collectives go through torch.distributed (`group`).  comm = (id_bytes, rank, world): they go through
        the library's own NCCL binding (kgeb_comm_init / kgeb_alltoall_rows / kgeb_allreduce_*), torch.distributed is not used."""
        from ._engine import Engine
        if kind not in abi.FUSED_TRAIN_KINDS:
            # the compact-table engine below carries no per-kind parameters (pRotatE modulus, HAKE weights, HolE mode):
            # refuse instead of silently scoring with defaults
            raise ValueError(f"ShardedTrainer serves the fused-step score kinds {abi.FUSED_TRAIN_KINDS} "
                             f"(DistMult, TransE L1/L2, ComplEx); got kind {kind}")
        self.group = group
        self.world, self.rank = (int(comm[2]), int(comm[1])) if comm is not None else _world(group)
        self.kind, self.N, self.R_, self.we, self.wr, self.gamma = kind, int(num_entities), int(num_relations), int(we), int(wr), float(gamma)
        self.r0, self.r1 = shard_rows(self.N, self.world, self.rank)
        if E is None:   # GlorotUniform of the GLOBAL table shape (models.py:87-88); each rank draws only its rows
            lim = float(np.sqrt(6.0 / (self.N + self.we)))
            E_local = np.random.default_rng([seed, 1, self.rank]).uniform(-lim, lim, size=(self.r1 - self.r0, self.we)).astype(np.float32)
        else:
            E_local = np.asarray(E, dtype=np.float32)[self.r0:self.r1]
        if R is None:   # replicated: the same stream on every rank
            lim = float(np.sqrt(6.0 / (self.R_ + self.wr)))
            R = np.random.default_rng([seed, 2]).uniform(-lim, lim, size=(self.R_, self.wr)).astype(np.float32)
        self.shard = Engine(kind, max(self.r1 - self.r0, 1), self.R_, self.we, self.wr, gamma, E=E_local if self.r1 > self.r0 else None, R=R)
        self.lib, self.device = self.shard.lib, self.shard.device
        self.W = self.shard.we_k
        if comm is not None:
            self.shard.comm_init(comm[0], self.rank, self.world)
            self.ex = KgebExchange(self.shard)
        else:
            self.ex = TorchExchange(group)
        self.exchange_kind = self.ex.kind
        # a table-less context bound to the GLOBAL entity count: the Philox sampler draws ids in [0, N)
        h = C.c_void_p()
        abi.check(self.lib.kgeb_create(C.byref(h), self.device.index or 0, self.N, self.R_, self.shard.we_k, self.shard.wr_k,
                                       self.shard.we_k, self.shard.wr_k, kind, self.gamma))
        self._sampler = h
        self.compact = None
        self._shape_checked = None
        self.last = {}

    def __del__(self):
        try:
            if getattr(self, "_sampler", None):
                self.lib.kgeb_destroy(self._sampler)
                self._sampler = None
        except Exception:
            pass

    @property
    def launches(self):
        n = self.shard.launches + int(self.lib.kgeb_launch_count(self._sampler))
        return n + (self.compact.launches if self.compact is not None else 0)

    def _compact_engine(self, U, opt_kind):
        """Engine over the per-step compact table: capacity grows geometrically, relation table / gradient shared with the shard."""
        from ._engine import Engine
        sh = self.shard
        sh._ensure_train_buffers(opt_kind)
        if self.compact is None or self.compact.N < U:
            cap = max(int(U * 1.25), 1024)
            ce = Engine.__new__(Engine)
            ce.lib, ce.device = sh.lib, sh.device
            ce.kind, ce.N, ce.R_, ce.we, ce.wr, ce.gamma = self.kind, cap, self.R_, self.we, self.wr, self.gamma
            ce.we_k, ce.wr_k = sh.we_k, sh.wr_k
            ce.E = torch.zeros(cap, sh.we_k, dtype=torch.float32, device=self.device)
            ce.gE = torch.zeros_like(ce.E)
            ce.tE = torch.zeros(cap, dtype=torch.uint8, device=self.device)
            h = C.c_void_p()
            abi.check(self.lib.kgeb_create(C.byref(h), self.device.index or 0, cap, self.R_, sh.we_k, sh.wr_k, sh.we_k, sh.wr_k, self.kind, self.gamma))
            ce.ctx = h
            ce.step = 0
            ce.hole_fft, ce.param = False, 1.0
            ce.grad_mode, ce._ss_ws = "red", None
            ce._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
            self.compact = ce
        ce = self.compact
        ce.Rel, ce.gR, ce.tR = sh.Rel, sh.gR, sh.tR          # replicated relation table: one storage, two views
        ce.opt_state = (opt_kind, None, None, None, None)    # the compact table is never updated in place
        return ce

    def negatives(self, pos, K, seed, offset, corrupt_mode=abi.CORRUPT_BOTH):
        neg = torch.empty(pos.shape[0] * K, 3, dtype=torch.int32, device=self.device)
        abi.check(self.lib.kgeb_negatives(self._sampler, _ptr(pos), pos.shape[0], K, seed, offset, corrupt_mode, _ptr(neg), _stream()), self._sampler)
        return neg

    def step(self, pos, neg=None, K=1, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, loss_weight=1.0, corrupt_mode=abi.CORRUPT_BOTH,
             seed=0, offset=0, opt_kind=abi.LAZY_ADAM, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        """One synchronous step over this rank's positives (global ids).  Returns the global mean loss (device scalar)."""
        from ._engine import as_idx
        sh = self.shard
        pos = as_idx(pos, self.device)
        if neg is None:
            neg = self.negatives(pos, K, seed + 7919 * self.rank, offset, corrupt_mode)
        else:
            neg = as_idx(neg, self.device)
            K = neg.shape[0] // pos.shape[0]
        # every rank's mean is folded into the global mean with weight 1/world: exact only for equal batch shapes
        if self.world > 1 and self._shape_checked != (pos.shape[0], K):
            # every rank writes its (B, K) into its own slot of a zero vector; the sum hands everyone the whole list
            shp = torch.zeros(self.world, 2, dtype=torch.int32, device=self.device)
            shp[self.rank, 0], shp[self.rank, 1] = int(pos.shape[0]), int(K)
            allshp = self.ex.sum_(shp).cpu().tolist()
            if any(tuple(x) != (int(pos.shape[0]), int(K)) for x in allshp):
                raise ValueError("ShardedTrainer.step needs the same number of positives and negatives on every rank "
                                 f"(this rank: B={pos.shape[0]}, K={K}; all ranks: {allshp})")
            self._shape_checked = (pos.shape[0], K)
        plan = ExchangePlan(torch.cat([pos[:, 0], pos[:, 2], neg[:, 0], neg[:, 2]]), self.N, exchange=self.ex)
        ce = self._compact_engine(plan.U, opt_kind)
        W, st = self.W, _stream()
        # rows: owners -> requesters
        nserve = int(plan.serve_local.numel())
        send = torch.empty(nserve, W, dtype=torch.float32, device=self.device)
        abi.check(self.lib.kgeb_rows_gather(sh.ctx, _ptr(sh.E), W, W, _ptr(plan.serve_local), nserve, _ptr(send), st), sh.ctx)
        self.ex.all_to_all(ce.E[:plan.U], send, plan.req_counts, plan.serve_counts)
        # the ordinary fused step on the compact table; 1/world folds the ranks' means into the global mean
        res = ce.forward_backward(plan.remap_triples(pos), plan.remap_triples(neg), K=K, loss_kind=loss_kind, margin=margin,
                                  loss_weight=loss_weight / self.world, opt_kind=opt_kind)
        # gradient rows: requesters -> owners
        back = torch.empty(nserve, W, dtype=torch.float32, device=self.device)
        self.ex.all_to_all(back, ce.gE[:plan.U], plan.serve_counts, plan.req_counts)
        ce.gE[:plan.U].zero_()
        ce.tE[:plan.U].zero_()
        abi.check(self.lib.kgeb_rows_scatter_add(sh.ctx, _ptr(sh.gE), W, W, _ptr(sh.tE), _ptr(plan.serve_local), nserve, _ptr(back), st), sh.ctx)
        loss = res["loss"].clone()
        if self.world > 1:
            self.ex.sum_(sh.gR)
            self.ex.max_(sh.tR)
            self.ex.sum_(loss)
        sh.apply_optimizer(opt_kind, lr, beta1, beta2, eps)
        self.last = {"unique_rows": plan.U, "served_rows": nserve, "a2a_bytes_fwd": nserve * W * 4, "a2a_bytes_bwd": plan.U * W * 4}
        self._keep = (send, back, plan, pos, neg)
        return loss

    def gather_full_table(self):
        """The whole entity table on every rank (tests / checkpoint export), [N, we] numpy."""
        local = self.shard.E[: self.r1 - self.r0, : self.we].contiguous()
        if self.world == 1:
            return local.cpu().numpy()
        per = shard_rows(self.N, self.world, 0)[1]
        pad = torch.zeros(per, self.we, dtype=torch.float32, device=self.device)
        pad[: local.shape[0]] = local
        out = torch.empty(self.world * per, self.we, dtype=torch.float32, device=self.device)
        self.ex.all_to_all(out, pad.repeat(self.world, 1), [per] * self.world, [per] * self.world)   # an all-gather as an all-to-all
        return out[: self.N].cpu().numpy()
