"""Timed CPU port of the reference hot path (TEST / BASELINE INFRASTRUCTURE, see oracle/__init__.py).

TensorFlow cannot be installed in this image, so the "reference arm" of bench.py is this restatement:
  * train step: fp32 PyTorch-CPU, following tests.py:55-87 -- host negatives exactly like generate_negative
    (utils.py:114-121: both sides uniform), Embedding lookups (sparse gradients = IndexedSlices), score function
    (models.py:171-237), pairwise hinge paired per positive, autograd, lazy sparse Adam on the touched rows;
  * filtered evaluation: the vectorised NumPy restatement of validate (utils.py:199-264).
It is labelled kind="port" everywhere it is reported; it is never on the product path."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import kge_oracle as ko


class TorchCpuTrainer:
    def __init__(self, kind, E, R, gamma=12.0, lr=1e-3, margin=1.0, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.kind, self.gamma, self.margin = kind, gamma, margin
        self.E = torch.nn.Embedding.from_pretrained(torch.from_numpy(np.array(E, np.float32)), freeze=False, sparse=True)
        self.R = torch.nn.Embedding.from_pretrained(torch.from_numpy(np.array(R, np.float32)), freeze=False, sparse=True)
        self.opt = torch.optim.SparseAdam(list(self.E.parameters()) + list(self.R.parameters()), lr=lr, eps=1e-7)
        self.N = E.shape[0]

    def score(self, idx):
        s, p, o = self.E(idx[:, 0]), self.R(idx[:, 1]), self.E(idx[:, 2])
        k = self.kind
        if k == ko.DISTMULT:
            return (s * p * o).sum(-1)
        if k in (ko.TRANSE_L1, ko.TRANSE_L2):
            r = s + p - o
            n = r.abs().sum(-1) if k == ko.TRANSE_L1 else r.pow(2).sum(-1).sqrt()
            return self.gamma - n if self.gamma > 0 else n
        if k == ko.COMPLEX:
            d = s.shape[-1] // 2
            sr, si, pr, pi, orr, oi = s[:, :d], s[:, d:], p[:, :d], p[:, d:], o[:, :d], o[:, d:]
            return (sr * pr * orr + pr * si * oi + pi * sr * oi - pi * si * orr).sum(-1)
        if k == ko.HOLE:
            x = torch.fft.ifft(torch.conj(torch.fft.fft(s.to(torch.complex64))) * torch.fft.fft(o.to(torch.complex64))).real
            return (p * x).sum(-1)
        raise KeyError(k)

    def step(self, pos, K):
        """pos: (B,3) int64 numpy.  Returns the loss value (python float)."""
        B = pos.shape[0]
        fs = np.random.randint(0, self.N, size=(K * B,))
        fo = np.random.randint(0, self.N, size=(K * B,))
        neg = np.stack([fs, np.repeat(pos[:, 1], K), fo], axis=1)
        t = self.score(torch.from_numpy(pos))
        f = self.score(torch.from_numpy(neg)).view(B, K)
        loss = torch.relu(self.margin + f - t.view(B, 1)).mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return float(loss.item())


def time_train(kind, E, R, B, K, gamma, steps, warmup=1, threads=None, seed=0, budget_s=25.0):
    """triples/s (incl. negatives) of the CPU port on `threads` cores; bounded by budget_s seconds."""
    tr = TorchCpuTrainer(kind, E, R, gamma, threads=threads)
    rng = np.random.default_rng(seed)
    N, Rn = E.shape[0], R.shape[0]
    np.random.seed(seed)

    def batch():
        return np.stack([rng.integers(0, N, B), rng.integers(0, Rn, B), rng.integers(0, N, B)], axis=1).astype(np.int64)
    for _ in range(warmup):
        tr.step(batch(), K)
    times = []
    t_all = time.perf_counter()
    for _ in range(steps):
        pos = batch()
        t0 = time.perf_counter()
        tr.step(pos, K)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    dt = float(np.sum(times))
    return dict(triples_per_s=len(times) * B * (1 + K) / dt, ms_per_step=1e3 * dt / len(times), steps=len(times),
                threads=torch.get_num_threads())


def time_eval(kind, E, R, triples, filt_known, gamma, max_queries=64, budget_s=15.0):
    """queries/s of the vectorised NumPy restatement of validate (tail + head), on a bounded sample of queries."""
    fh, ft = ko.build_filters(filt_known)
    tri = np.asarray(triples)[:max_queries]
    t0 = time.perf_counter()
    done = 0
    for i in range(0, len(tri), 8):
        part = tri[i:i + 8]
        for side, filt in (("tail", ft), ("head", fh)):
            sc = ko.all_entity_scores(kind, E, R, part, side, gamma, np.float32, hole_fft=True)
            ko.ranks_from_scores(sc, part, side, filt)
        done += 2 * len(part)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(queries_per_s=done / dt, queries=done, seconds=dt)


def time_eval_dot(kind, E, R, triples, csr_tail, csr_head, max_triples=32, chunk=8, budget_s=20.0):
    """queries/s of the NumPy restatement of validate for the dot-product models on a BOUNDED sample of test triples, written
    the way that is kindest to the CPU: the candidate scores of a chunk of queries are ONE BLAS sgemm q.E^T (all host threads;
    SURVEY Appendix A regrouping of models.py:171-172 / 212-222) instead of the reference's (C+1, 3) gather-multiply-sum per
    query, then rank = 1 + #{c not filtered, c != target : s_c > s_t} (utils.py:210-215).  csr_* = (offsets, ids) known-true
    ids per query."""
    tri = np.asarray(triples)[:max_triples].astype(np.int64)
    E = np.asarray(E, np.float32)
    t0 = time.perf_counter()
    done = 0
    ranks = []
    for i in range(0, len(tri), chunk):
        part = tri[i:i + chunk]
        for side, (offs, ids) in (("tail", csr_tail), ("head", csr_head)):
            rows, inv = np.unique(part[:, [0, 2]], return_inverse=True)     # query_vectors copies its table: give it the few rows it reads
            loc = np.stack([inv.reshape(-1, 2)[:, 0], part[:, 1], inv.reshape(-1, 2)[:, 1]], axis=1)
            q = ko.query_vectors(kind, E[rows], R, loc, side)
            sc = q @ E.T
            for j, (s, p, o) in enumerate(part):
                tgt = o if side == "tail" else s
                row = sc[j]
                t = row[tgt]
                cnt = int(np.count_nonzero(row > t))
                fl = np.asarray(ids[offs[i + j]:offs[i + j + 1]], dtype=np.int64)
                fl = fl[fl != tgt]
                cnt -= int(np.count_nonzero(row[fl] > t))
                ranks.append(1 + cnt)
        done += 2 * len(part)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(queries_per_s=done / dt, queries=done, seconds=dt, mean_rank=float(np.mean(ranks)))
