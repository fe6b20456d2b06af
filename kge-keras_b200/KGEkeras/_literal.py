"""The literal branch of EmbeddingModel (models.py:114-120,152-158; the reference harness trains it, tests.py:113,125-126):

    s' = act([Dropout(E[s]) | Dropout(L[s])] W + b),  o' likewise (ONE shared Dense(e_dim)),  score = func(s', Dropout(R[p]), o')

L is the non-trainable literal table, W / b a Keras Dense (glorot_uniform kernel, zero bias).  Dropout draws a fresh mask per
gathered row, so s' depends on the reference, not on the entity: the step materialises the transformed rows as a scratch
"entity table" (and the dropped relation rows as a scratch relation table), scores them with the ordinary kernels through
identity ids, and sends the gradients back through the Dense (two small fp32 GEMMs) and the gathers.  For inference (no dropout)
the transformed table E' = act([E | L] W + b) is a function of the entity and is cached per weight version.
No torch op computes anything here: gathers, GEMMs, activation and scatters are libkgeb200 kernels (literal_kernels.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi as abi
from ._engine import Engine, _ptr, _stream, as_idx


class LiteralHead:
    def __init__(self, main: Engine, literals, activation, rng=None):
        if activation not in abi.ACT_KINDS:
            raise NotImplementedError(f"literal_activation={activation!r}: the Dense kernel serves {sorted(k for k in abi.ACT_KINDS if k)}")
        self.main, self.lib, self.dev = main, main.lib, main.device
        self.act = abi.ACT_KINDS[activation]
        lit = np.asarray(literals, dtype=np.float32)
        if lit.ndim != 2 or lit.shape[0] != main.N:
            raise ValueError(f"literals must have shape (num_entities, n_literals), got {lit.shape}")
        self.wl = lit.shape[1]
        self.L = torch.from_numpy(np.ascontiguousarray(lit)).to(self.dev)
        self.we, self.wek = main.we, main.we_k
        self.kin = self.we + self.wl                       # Dense input width
        rng = rng or np.random.default_rng()
        lim = float(np.sqrt(6.0 / (self.kin + self.we)))   # Keras Dense: glorot_uniform kernel, zeros bias
        self.W = torch.zeros(self.kin, self.wek, dtype=torch.float32, device=self.dev)
        self.W[:, :self.we] = torch.from_numpy(rng.uniform(-lim, lim, size=(self.kin, self.we)).astype(np.float32)).to(self.dev)
        self.b = torch.zeros(1, self.wek, dtype=torch.float32, device=self.dev)
        self.gW, self.gb = torch.zeros_like(self.W), torch.zeros_like(self.b)
        self.state = None
        self.version = 0          # bumped by every update of E, W or b: invalidates the cached inference table
        self._tmp = None          # scratch engine: 2m transformed entity rows, m dropped relation rows
        self._eff, self._eff_version = None, -1
        self._X = self._gX = None

    # ---- small wrappers -----------------------------------------------------------------------------------------------
    def _chk(self, rc):
        abi.check(rc, self.main.ctx)

    def _concat(self, table, w, ld, lit, wl, ldl, ids, out, ldo):
        self._chk(self.lib.kgeb_rows_concat(self.main.ctx, _ptr(table), w, ld, _ptr(lit), wl, ldl, _ptr(ids), ids.numel(), _ptr(out), ldo, _stream()))

    def _concat_bwd(self, g, ldg, w, ids, G, ldG, touched):
        self._chk(self.lib.kgeb_rows_concat_bwd(self.main.ctx, _ptr(g), ldg, w, _ptr(ids), ids.numel(), _ptr(G), ldG, _ptr(touched), _stream()))

    def _gemm(self, ta, tb, M, N, K, A, lda, B, ldb, Cm, ldc, acc=0):
        self._chk(self.lib.kgeb_sgemm(self.main.ctx, ta, tb, M, N, K, _ptr(A), lda, _ptr(B), ldb, _ptr(Cm), ldc, acc, _stream()))

    def _dense_fwd(self, X, rows, out):
        """out[:rows] = act(X W + b): one GEMM with the bias and activation in its epilogue."""
        self._chk(self.lib.kgeb_dense_fwd(self.main.ctx, _ptr(X), self.kin, rows, self.kin, _ptr(self.W), self.wek, self.we, _ptr(self.b), self.act,
                                          _ptr(out), self.wek, _stream()))

    def _scratch(self, m):
        if self._tmp is None or self._tmp.R_ < m:
            cap = max(m, 1024)
            mn = self.main
            self._tmp = Engine(mn.kind, 2 * cap, cap, mn.we, mn.wr, mn.gamma, E=torch.zeros(2 * cap, mn.we), R=torch.zeros(cap, mn.wr), param=mn.param)
            if mn.hole_fft:
                self._tmp.set_hole_mode(True)
            self._X = torch.empty(2 * cap, self.kin, dtype=torch.float32, device=self.dev)
            self._gX = torch.empty(2 * cap, self.kin, dtype=torch.float32, device=self.dev)
        return self._tmp

    # ---- inference: E' as an ordinary engine --------------------------------------------------------------------------
    def effective_engine(self) -> Engine:
        if self._eff is None or self._eff_version != self.version:
            mn = self.main
            if self._eff is None:
                self._eff = Engine(mn.kind, mn.N, mn.R_, mn.we, mn.wr, mn.gamma, E=torch.zeros(1, mn.we).expand(mn.N, mn.we), R=torch.zeros(mn.R_, mn.wr),
                                   param=mn.param)
                if mn.hole_fft:
                    self._eff.set_hole_mode(True)
            ids = torch.arange(mn.N, dtype=torch.int32, device=self.dev)
            X = torch.empty(mn.N, self.kin, dtype=torch.float32, device=self.dev)
            mn.set_dropout(0.0)
            self._concat(mn.E, self.we, mn.we_k, self.L, self.wl, self.wl, ids, X, self.kin)
            self._dense_fwd(X, mn.N, self._eff.E)
            self._eff.Rel.copy_(mn.Rel)
            self._eff_version = self.version
        return self._eff

    def score(self, idx, dropout=None):
        """Scores of (m,3) triples: inference through the cached E'; dropout = (p, seed, offset) draws training-mode masks."""
        if dropout is None:
            return self.effective_engine().score(idx)
        mn = self.main
        tri = as_idx(idx, self.dev)
        m = tri.shape[0]
        tmp = self._scratch(m)
        p, seed, offset = dropout
        refs = torch.cat([tri[:, 0], tri[:, 2]]).contiguous()
        rel = tri[:, 1].contiguous()
        mn.set_dropout(p, seed, offset, 0)
        self._concat(mn.E, self.we, mn.we_k, self.L, self.wl, self.wl, refs, self._X, self.kin)
        mn.set_dropout(p, seed, offset, 1)
        self._concat(mn.Rel, mn.wr, mn.wr_k, None, 0, 0, rel, tmp.Rel, mn.wr_k)
        mn.set_dropout(0.0)
        self._dense_fwd(self._X, 2 * m, tmp.E)
        ar = torch.arange(m, dtype=torch.int32, device=self.dev)
        return tmp.score(torch.stack([ar, ar, ar + m], dim=1))

    # ---- one training step ---------------------------------------------------------------------------------------------
    def forward_backward(self, pos, neg=None, K=1, loss_kind=abi.PAIRWISE_HINGE, pair_mode=abi.PAIRED, margin=1.0, loss_weight=1.0,
                         corrupt_mode=abi.CORRUPT_BOTH, seed=0, offset=0, opt_kind=abi.ADAM_DENSE, dropout=0.0, want_scores=False):
        mn = self.main
        pos = as_idx(pos, self.dev)
        B = pos.shape[0]
        if neg is None:
            neg = mn.negatives(pos, K, seed, offset, corrupt_mode)
        else:
            neg = as_idx(neg, self.dev)
        tri = torch.cat([pos, neg], dim=0)
        m = tri.shape[0]
        tmp = self._scratch(m)
        mn._ensure_train_buffers(opt_kind)
        tmp._ensure_train_buffers(abi.SGD)
        refs = torch.cat([tri[:, 0], tri[:, 2]]).contiguous()
        rel = tri[:, 1].contiguous()
        X, gX = self._X, self._gX

        def arm(stream_id):
            mn.set_dropout(dropout, seed, offset, stream_id) if dropout else mn.set_dropout(0.0)

        # forward: dropped rows -> Dense -> scratch tables
        arm(0)
        self._concat(mn.E, self.we, mn.we_k, self.L, self.wl, self.wl, refs, X, self.kin)
        arm(1)
        self._concat(mn.Rel, mn.wr, mn.wr_k, None, 0, 0, rel, tmp.Rel, mn.wr_k)
        mn.set_dropout(0.0)
        self._dense_fwd(X, 2 * m, tmp.E)
        ar = torch.arange(m, dtype=torch.int32, device=self.dev)
        idsT = torch.stack([ar, ar, ar + m], dim=1)
        tmp.gE[:2 * m].zero_(); tmp.gR[:m].zero_(); tmp.tE[:2 * m].zero_(); tmp.tR[:m].zero_()
        res = tmp.forward_backward(idsT[:B], idsT[B:], K=K, loss_kind=loss_kind, pair_mode=pair_mode, margin=margin, loss_weight=loss_weight,
                                   opt_kind=abi.SGD, want_scores=want_scores)
        # backward through the Dense: gZ = gY * act'(Y);  gb += colsum(gZ);  gW = X^T gZ;  gX = gZ W^T
        self._chk(self.lib.kgeb_dense_act_bwd(mn.ctx, _ptr(tmp.E), _ptr(tmp.gE), 2 * m, self.we, self.wek, self.act, _ptr(self.gb), _stream()))
        self._gemm(1, 0, self.kin, self.we, 2 * m, X, self.kin, tmp.gE, self.wek, self.gW, self.wek, 1)
        self._gemm(0, 1, 2 * m, self.we, self.we, tmp.gE, self.wek, self.W, self.wek, gX, self.kin)   # literal columns get no gradient
        arm(0)
        self._concat_bwd(gX, self.kin, self.we, refs, mn.gE, mn.we_k, mn.tE)
        arm(1)
        self._concat_bwd(tmp.gR, mn.wr_k, mn.wr, rel, mn.gR, mn.wr_k, mn.tR)
        mn.set_dropout(0.0)
        res["neg"] = neg
        return res

    def apply_optimizer(self, opt_kind=abi.ADAM_DENSE, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        """E and R through the engine, then the Dense kernel and bias (every row 'touched': Keras updates them densely)."""
        mn = self.main
        mn.apply_optimizer(opt_kind, lr, beta1, beta2, eps)
        if self.state is None or self.state[0] != opt_kind:
            if opt_kind == abi.SGD:
                st = (None, None, None, None)
            elif opt_kind == abi.ADAGRAD:
                st = (torch.full_like(self.W, 0.1), None, torch.full_like(self.b, 0.1), None)
            else:
                st = (torch.zeros_like(self.W), torch.zeros_like(self.W), torch.zeros_like(self.b), torch.zeros_like(self.b))
            self.state = (opt_kind,) + st
            self._ones = torch.ones(self.kin, dtype=torch.uint8, device=self.dev)
            self._one_b = torch.ones(1, dtype=torch.uint8, device=self.dev)
        _, mW, vW, mb, vb = self.state
        self._ones.fill_(1)
        self._one_b.fill_(1)
        self._chk(self.lib.kgeb_opt_update2(mn.ctx, _ptr(self.W), _ptr(self.gW), _ptr(self._ones), self.kin, self.wek, self.wek, _ptr(mW), _ptr(vW),
                                            _ptr(self.b), _ptr(self.gb), _ptr(self._one_b), 1, self.wek, self.wek, _ptr(mb), _ptr(vb), opt_kind, lr, beta1,
                                            beta2, eps, mn.step, _stream()))
        self.version += 1

    def train_step(self, pos, neg=None, opt_kind=abi.ADAM_DENSE, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, l3=0.0, **kw):
        res = self.forward_backward(pos, neg, opt_kind=opt_kind, **kw)
        if l3:
            self.main.l3_reg(l3, res["loss"])
        self.apply_optimizer(opt_kind, lr, beta1, beta2, eps)
        return res

    # ---- weights -------------------------------------------------------------------------------------------------------
    def get_weights(self):
        return [self.W[:, :self.we].cpu().numpy(), self.b[0, :self.we].cpu().numpy()]

    def set_weights(self, W, b):
        self.W[:, :self.we] = torch.as_tensor(np.asarray(W, dtype=np.float32)).to(self.dev)
        self.b[0, :self.we] = torch.as_tensor(np.asarray(b, dtype=np.float32)).to(self.dev)
        self.version += 1
