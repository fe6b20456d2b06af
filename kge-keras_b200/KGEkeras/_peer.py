"""Row-sharded training with the row exchange INSIDE the scoring kernel (BASELINE configs[4]).

`_sharded.ShardedTrainer` moves the rows a rank needs with two all-to-alls around an unchanged kernel.  Here the entity table and
its gradient buffer are `PeerTable`s instead: shard g lives in the memory of GPU g, and all shards are mapped back to back at one
virtual address range on every rank (kgeb_peer_*, CUDA virtual memory management).  The fused train kernel then runs on GLOBAL
ids with the table's base pointer: its bulk-async row gathers read remote rows over NVLink as they are needed, its vector
reductions add gradient rows into the owner's shard over NVLink -- transfer and arithmetic overlap row by row, nothing is
packed, deduplicated or remapped, and at world = 1 the step IS the single-GPU step.

One step on every rank:
  1. kgeb_train_fwd_bwd on (peer E, replicated R) with in-kernel Philox negatives; gradient rows go to the peer gradient table;
     every rank marks the rows it referenced in a LOCAL flag array over all N rows;
  2. the flags are max-reduced (and the replicated relation gradient / loss summed): this collective is also the barrier after
     which every remote reduction into this rank's gradient shard has landed;
  3. the owner updates its own rows (touched-row LazyAdam / Adagrad / SGD) and the replicated relation table;
  4. a last tiny all-reduce fences the update against the next step's remote reads.
The collectives go through the library's NCCL binding (kgeb_comm_*) -- torch.distributed is not needed."""
from __future__ import annotations

import ctypes as C
import os
import socket
import threading
import time

import numpy as np
import torch

from . import _abi as abi


class _DevView:
    """A raw device range as something torch.as_tensor accepts (no copy, no ownership)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class _RawTable:
    """The base pointer of a peer-mapped range, duck-typed for the engine's `_ptr` (slot 0 is physically another GPU's memory on
    every rank but 0, so no torch tensor is laid over the whole range: torch would attribute it to that device)."""

    def __init__(self, ptr):
        self.ptr = int(ptr)

    def data_ptr(self):
        return self.ptr


def exchange_fds(rank, world, fd, tag, timeout=120.0):
    """Every rank sends `fd` to every other rank's process over UNIX sockets (SCM_RIGHTS); returns {src_rank: fd valid here}.
    `tag` must be the same string on all ranks and unique per exchange (it names the socket files)."""
    if world == 1:
        return {}
    path = lambda r: os.path.join(os.environ.get("KGEB_SOCKET_DIR", "/tmp"), f"kgeb_peer_{tag}_{r}.sock")
    try:
        os.unlink(path(rank))
    except FileNotFoundError:
        pass
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path(rank))
    srv.listen(world)
    srv.settimeout(timeout)
    got, err = {}, []

    def serve():
        try:
            for _ in range(world - 1):
                conn, _a = srv.accept()
                with conn:
                    msg, fds, _f, _ad = socket.recv_fds(conn, 16, 1)
                    got[int(msg.decode())] = fds[0]
        except Exception as e:      # surfaced by the caller
            err.append(e)

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    deadline = time.time() + timeout
    for g in range(world):
        if g == rank:
            continue
        while True:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            try:
                c.connect(path(g))
                break
            except (FileNotFoundError, ConnectionRefusedError):
                c.close()
                if time.time() > deadline:
                    raise TimeoutError(f"rank {rank}: rank {g} never opened {path(g)}")
                time.sleep(0.02)
        with c:
            socket.send_fds(c, [str(rank).encode()], [fd])
    th.join(timeout)
    srv.close()
    try:
        os.unlink(path(rank))
    except FileNotFoundError:
        pass
    if err or len(got) != world - 1:
        raise RuntimeError(f"rank {rank}: descriptor exchange failed ({err or 'timeout'}; received from {sorted(got)})")
    return got


class PeerTable:
    """fp32 [world * rows_per_shard, width] whose shard g is physical memory of GPU g, addressable as ONE tensor on every rank."""

    def __init__(self, rows_per_shard, width, rank, world, tag, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("PeerTable needs CUDA devices: there is no CPU fallback")
        self.lib = abi.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rank, self.world, self.width = int(rank), int(world), int(width)
        dev = self.device.index or 0
        gran = int(self.lib.kgeb_peer_granularity(dev))
        if gran < 0:
            abi.check(gran)
        row_bytes = 4 * self.width
        unit = gran // int(np.gcd(gran, row_bytes))           # rows per shard must make the shard a whole number of granules
        self.rows_per_shard = (int(rows_per_shard) + unit - 1) // unit * unit
        h = C.c_void_p()
        abi.check(self.lib.kgeb_peer_create(C.byref(h), dev, self.rank, self.world, self.rows_per_shard * row_bytes))
        self.h = h
        assert int(self.lib.kgeb_peer_shard_bytes(h)) == self.rows_per_shard * row_bytes
        if self.world > 1:
            fd = C.c_int()
            self._check(self.lib.kgeb_peer_export_fd(h, C.byref(fd)))
            theirs = exchange_fds(self.rank, self.world, fd.value, tag)
            for g, f in sorted(theirs.items()):
                try:
                    self._check(self.lib.kgeb_peer_import_fd(h, g, f))
                finally:
                    os.close(f)
        if int(self.lib.kgeb_peer_complete(h)) != 1:
            raise RuntimeError("PeerTable: not every shard was mapped")
        base = int(self.lib.kgeb_peer_base(h))
        self.all = _RawTable(base)
        self.r0 = self.rank * self.rows_per_shard
        # the own shard is ordinary local memory: a torch view (no copy) for initialisation, export and the optimizer's pointers
        self.local = torch.as_tensor(_DevView(base + self.r0 * row_bytes, (self.rows_per_shard, self.width), "<f4"), device=self.device)
        if self.local.data_ptr() != base + self.r0 * row_bytes:
            raise RuntimeError("PeerTable: torch copied the shard instead of viewing it")
        self.local.zero_()

    def _check(self, rc):
        if rc != 0:
            raise abi.KgebError(rc, (self.lib.kgeb_peer_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.all = self.local = None
            self.lib.kgeb_peer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_rows_per_shard(num_entities, world, width, device_index=0):
    """Rows per shard of an N-row table over `world` GPUs such that a shard is a whole number of allocation granules."""
    gran = int(abi.lib().kgeb_peer_granularity(device_index))
    if gran < 0:
        abi.check(gran)
    unit = gran // int(np.gcd(gran, 4 * int(width)))
    per = (int(num_entities) + world - 1) // world
    return (per + unit - 1) // unit * unit


class PeerShardedTrainer:
    """One model, entity rows sharded over the GPUs of one NVLink box, rows and gradient rows exchanged inside the kernel.

    comm = (id_bytes, rank, world) as for ShardedTrainer; `tag` names this trainer's descriptor exchange (same on all ranks)."""

    def __init__(self, kind, num_entities, num_relations, we, wr, gamma=12.0, seed=0, E=None, R=None, comm=None, tag="t0", grad="remote_red", negatives="global"):
        """grad = "remote_red" (default): the kernel's 16-byte vector reductions go straight to the owner's gradient shard over NVLink.
        grad = "pull": every rank reduces the gradient rows of its batch into a PRIVATE full-size buffer (local vector
        reductions), and after the step's barrier the owner of a row pulls and adds the other ranks' partial rows over NVLink in
        whole coalesced rows (kgeb_peer_pull_add) -- measured slower on 2 x B200 (9.1 vs 8.3 ms per C5 step): the remote row LOADS
        are what saturates NVLink, the reductions ride the other direction.

        negatives = "global" (default): corrupted entities are uniform over all N ids, the reference's distribution
        (utils.py:118-119); the step equals the single-GPU step on the concatenated batch.  negatives = "local": every rank draws
        them from ITS OWN shard only (the usual trick of distributed KGE trainers): only the positives' rows cross NVLink, but the
        sampling distribution is NOT the reference's -- an explicit opt-in, never the default."""
        from ._engine import Engine
        if grad not in ("pull", "remote_red"):
            raise ValueError(f"grad must be 'pull' or 'remote_red', got {grad!r}")
        if negatives not in ("global", "local"):
            raise ValueError(f"negatives must be 'global' or 'local', got {negatives!r}")
        self.grad, self.negatives = grad, negatives
        if kind not in abi.FUSED_TRAIN_KINDS:
            raise ValueError(f"PeerShardedTrainer serves the fused-step score kinds {abi.FUSED_TRAIN_KINDS}; got kind {kind}")
        self.world, self.rank = (int(comm[2]), int(comm[1])) if comm is not None else (1, 0)
        self.kind, self.N, self.R_, self.we, self.wr, self.gamma = kind, int(num_entities), int(num_relations), int(we), int(wr), float(gamma)
        self.device = torch.device("cuda", torch.cuda.current_device())
        # the global engine: context bound to ALL N rows, tables = the peer-mapped ranges
        g = Engine.__new__(Engine)
        g.lib, g.device = abi.lib(), self.device
        g.kind, g.N, g.R_, g.we, g.wr, g.gamma = kind, self.N, self.R_, self.we, self.wr, self.gamma
        pad_ok = kind in (abi.DISTMULT, abi.TRANSE_L1, abi.TRANSE_L2)
        g.we_k = (self.we + 3) // 4 * 4 if pad_ok else self.we
        g.wr_k = (self.wr + 3) // 4 * 4 if pad_ok else self.wr
        self.W = g.we_k
        self.per = peer_rows_per_shard(self.N, self.world, self.W, self.device.index or 0)
        self.r0 = min(self.N, self.rank * self.per)
        self.r1 = min(self.N, self.r0 + self.per)
        mport = os.environ.get("MASTER_PORT", "0")
        self.tE_ = PeerTable(self.per, self.W, self.rank, self.world, f"{mport}_{tag}_E", self.device)
        if grad == "pull":      # slot p = rank p's whole private gradient buffer
            self.tG_ = PeerTable(self.world * self.per, self.W, self.rank, self.world, f"{mport}_{tag}_G", self.device)
            assert self.tG_.rows_per_shard == self.world * self.per
            g.E, g.gE = self.tE_.all, _RawTable(self.tG_.local.data_ptr())
            self.G_loc = self.tG_.local[self.r0:self.r0 + self.per]
        else:                   # slot p = rank p's shard of ONE gradient table
            self.tG_ = PeerTable(self.per, self.W, self.rank, self.world, f"{mport}_{tag}_G", self.device)
            g.E, g.gE = self.tE_.all, self.tG_.all
            self.G_loc = self.tG_.local
        nloc = self.r1 - self.r0
        if E is None:   # GlorotUniform of the GLOBAL table shape (models.py:87-88); each rank draws only its rows
            lim = float(np.sqrt(6.0 / (self.N + self.we)))
            if nloc:
                self.tE_.local[:nloc, :self.we].uniform_(-lim, lim, generator=torch.Generator(device=self.device).manual_seed(1000003 * seed + self.rank))
        elif nloc:
            self.tE_.local[:nloc, :self.we] = torch.as_tensor(np.asarray(E, dtype=np.float32)[self.r0:self.r1]).to(self.device)
        if R is None:   # replicated: the same stream on every rank
            lim = float(np.sqrt(6.0 / (self.R_ + self.wr)))
            R = np.random.default_rng([seed, 2]).uniform(-lim, lim, size=(self.R_, self.wr)).astype(np.float32)
        g.Rel = torch.zeros(self.R_, g.wr_k, dtype=torch.float32, device=self.device)
        g.Rel[:, :self.wr] = torch.as_tensor(np.asarray(R, dtype=np.float32)).to(self.device)
        h = C.c_void_p()
        abi.check(g.lib.kgeb_create(C.byref(h), self.device.index or 0, self.N, self.R_, g.we_k, g.wr_k, g.we_k, g.wr_k, kind, self.gamma))
        g.ctx = h
        g.hole_fft, g.param, g.grad_mode, g._ss_ws = False, 1.0, "red", None
        g.step = 0
        g._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        g.tE = torch.zeros(self.N, dtype=torch.uint8, device=self.device)        # LOCAL marks over all N rows
        g.gR = torch.zeros_like(g.Rel)
        g.tR = torch.zeros(self.R_, dtype=torch.uint8, device=self.device)
        g.opt_state = None
        g._ensure_train_buffers = lambda opt_kind: None     # buffers above are final; optimizer state lives with the owner below
        self.g = g
        self.lib = g.lib
        if comm is not None:
            g.comm_init(comm[0], self.rank, self.world)
        self._sampler = None
        if negatives == "local":   # a table-less context whose entity range is this rank's shard: kgeb_negatives draws ids in [0, nloc)
            hs = C.c_void_p()
            abi.check(g.lib.kgeb_create(C.byref(hs), self.device.index or 0, max(self.r1 - self.r0, 1), self.R_, g.we_k, g.wr_k, g.we_k, g.wr_k, kind, self.gamma))
            self._sampler = hs
        self.opt_state = None
        self.steps = 0
        self.exchange_kind = ("in-kernel NVLink peer row loads over one virtual table; gradient rows reduced locally, pulled and added by their owner (kgeb_peer_pull_add)"
                              if grad == "pull" else "in-kernel NVLink peer row loads and 16-byte vector reductions over one virtual table (kgeb_peer_*)")
        self._sync()

    # -- collectives (library binding; no-ops at world 1) -----------------------------------------------------------
    def _sum(self, t):
        if self.world > 1:
            fn = self.lib.kgeb_allreduce_f32 if t.dtype == torch.float32 else self.lib.kgeb_allreduce_counts
            self.g._check(fn(self.g.ctx, C.c_void_p(t.data_ptr()), t.numel(), _stream()))
        return t

    def _max(self, t):
        if self.world > 1:
            self.g._check(self.lib.kgeb_allreduce_flags(self.g.ctx, C.c_void_p(t.data_ptr()), t.numel(), _stream()))
        return t

    def _sync(self):
        """All ranks' earlier device work is complete when this returns on the stream (a 4-byte all-reduce)."""
        if self.world > 1:
            if not hasattr(self, "_tok"):
                self._tok = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._sum(self._tok)

    def _owner_state(self, opt_kind):
        if self.opt_state is None or self.opt_state[0] != opt_kind:
            n = max(self.r1 - self.r0, 1)
            z = lambda: torch.zeros(n, self.W, dtype=torch.float32, device=self.device)
            zr = lambda: torch.zeros_like(self.g.Rel)
            if opt_kind == abi.SGD:
                st = (None, None, None, None)
            elif opt_kind == abi.ADAGRAD:
                st = (z() + 0.1, None, zr() + 0.1, None)
            else:
                st = (z(), z(), zr(), zr())
            self.opt_state = (opt_kind,) + st
            self.steps = 0
        return self.opt_state

    @property
    def launches(self):
        return self.g.launches

    def step(self, pos, neg=None, K=1, loss_kind=abi.PAIRWISE_HINGE, margin=1.0, loss_weight=1.0, corrupt_mode=abi.CORRUPT_BOTH,
             seed=0, offset=0, opt_kind=abi.LAZY_ADAM, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        """One synchronous step over this rank's positives (global ids; every rank the same B and K).  Returns the global mean
        loss (device scalar)."""
        g, lib = self.g, self.lib
        _, m_e, v_e, m_r, v_r = self._owner_state(opt_kind)
        if neg is None and self.negatives == "local":
            if corrupt_mode != abi.CORRUPT_BOTH:
                raise ValueError("negatives='local' serves the reference's corruption of both slots (CORRUPT_BOTH)")
            from ._engine import as_idx
            pos = as_idx(pos, self.device)
            neg = torch.empty(pos.shape[0] * K, 3, dtype=torch.int32, device=self.device)
            abi.check(lib.kgeb_negatives(self._sampler, _ptr(pos), pos.shape[0], K, seed + 7919 * self.rank, offset, corrupt_mode, _ptr(neg), _stream()), self._sampler)
            neg[:, 0] += self.r0
            neg[:, 2] += self.r0
        res = g.forward_backward(pos, neg, K=K, loss_kind=loss_kind, margin=margin, loss_weight=loss_weight / self.world,
                                 corrupt_mode=corrupt_mode, seed=seed + 7919 * self.rank, offset=offset, opt_kind=opt_kind)
        loss = res["loss"].clone()
        # marks, relation gradient, loss: summed / max-ed over the ranks.  Completing these also means every rank's kernel --
        # and with it every remote reduction into this rank's gradient shard -- has finished.
        self._max(g.tE)
        self._max(g.tR)
        self._sum(g.gR)
        self._sum(loss)
        self.steps += 1
        st = _stream()
        nloc = self.r1 - self.r0
        E_loc, G_loc = self.tE_.local, self.G_loc
        if self.grad == "pull" and self.world > 1 and nloc > 0:
            g._check(lib.kgeb_peer_pull_add(g.ctx, C.c_void_p(self.tG_.all.data_ptr()), self.world * self.per * self.W, self.world, self.rank,
                                            self.r0, nloc, self.W, self.W, C.c_void_p(g.tE.data_ptr() + self.r0), st))
        g._check(lib.kgeb_opt_update2(g.ctx, _ptr(E_loc), _ptr(G_loc), C.c_void_p(g.tE.data_ptr() + self.r0), max(nloc, 0), self.W, self.W,
                                      _ptr(m_e), _ptr(v_e), _ptr(g.Rel), _ptr(g.gR), _ptr(g.tR), self.R_, g.wr_k, g.wr_k, _ptr(m_r), _ptr(v_r),
                                      opt_kind, lr, beta1, beta2, eps, self.steps, st))
        # marks of rows other ranks own were only needed for the max above
        if self.r0 > 0:
            g.tE[:self.r0].zero_()
        if self.r1 < self.N:
            g.tE[self.r1:].zero_()
        self._sync()        # nobody reads a row before its owner has updated it
        self.last = {"rows_per_shard": self.per}
        return loss

    def gather_full_table(self):
        """The whole entity table on every rank (tests / checkpoint export), [N, we] numpy."""
        local = self.tE_.local[:, :self.we].contiguous()
        if self.world == 1:
            return local[:self.N].cpu().numpy()
        out = torch.empty(self.world * self.per, self.we, dtype=torch.float32, device=self.device)
        n = (C.c_int64 * self.world)(*([self.per] * self.world))
        self.g._check(self.lib.kgeb_alltoall_rows(self.g.ctx, _ptr(local.repeat(self.world, 1)), n, _ptr(out), n, self.we, _stream()))
        return out[:self.N].cpu().numpy()

    def relation_table(self):
        return self.g.Rel[:, :self.wr].cpu().numpy()

    def close(self):
        torch.cuda.synchronize()
        if self.world > 1 and self.g.comm_world > 1:
            self._sync()
            torch.cuda.synchronize()
            self.g.comm_destroy()
        self.g.E = self.g.gE = self.G_loc = None
        if self._sampler:
            self.lib.kgeb_destroy(self._sampler)
            self._sampler = None
        self.tE_.close()
        self.tG_.close()


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
