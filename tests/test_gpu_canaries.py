"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this pool): every buffer a kernel writes is carved out
of a larger allocation whose margins hold a sentinel pattern; after the call the margins must be untouched.  Row counts and
widths are chosen odd / non-multiples of the tile sizes so the tail predicates of every kernel are exercised."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu
GUARD = 4096   # elements of margin on each side


class Guarded:
    def __init__(self, shape, dtype, fill=0):
        n = int(np.prod(shape))
        self.raw = torch.empty(n + 2 * GUARD + 64, dtype=dtype, device="cuda")
        self.sentinel = {torch.float32: 1.2345e30, torch.int32: 0x5A5A5A5A, torch.uint8: 0xA5}[dtype]
        self.raw.fill_(self.sentinel)
        # keep the view 16-byte aligned (float4 / red.v4 kernels require it)
        off = GUARD
        while (self.raw.data_ptr() + off * self.raw.element_size()) % 16:
            off += 1
        self.off, self.n = off, n
        self.view = self.raw[off:off + n].view(*shape)
        self.view.fill_(fill)

    def intact(self):
        lo, hi = self.raw[:self.off], self.raw[self.off + self.n:]
        return bool((lo == self.sentinel).all()) and bool((hi == self.sentinel).all())


def _engine(kind, N, R, W, seed):
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(seed)
    Wr = W // 2 if kind == ko.ROTATE else W
    return Engine(kind, N, R, W, Wr, 12.0, E=ko.glorot_uniform(rng, N, W) * 3, R=ko.glorot_uniform(rng, R, Wr) * 3)


@pytest.mark.parametrize("kind,W", [(ko.DISTMULT, 100), (ko.TRANSE_L1, 200), (ko.TRANSE_L2, 36), (ko.COMPLEX, 200), (ko.COMPLEX, 512),
                                    (ko.HOLE, 100), (ko.ROTATE, 64), (ko.PROTATE, 52), (ko.MODE_L1, 24)])
@pytest.mark.parametrize("fused_neg", [True, False])
def test_train_step_and_optimizer_write_only_inside_their_buffers(kind, W, fused_neg):
    N, R, B, K = 1003, 7, 301, 5
    eng = _engine(kind, N, R, W, seed=W)
    eng._ensure_train_buffers(ko.OPT_LAZY_ADAM)
    g = {"gE": Guarded(tuple(eng.E.shape), torch.float32), "gR": Guarded(tuple(eng.Rel.shape), torch.float32),
         "tE": Guarded((N,), torch.uint8), "tR": Guarded((R,), torch.uint8)}
    eng.gE, eng.gR, eng.tE, eng.tR = g["gE"].view, g["gR"].view, g["tE"].view, g["tR"].view
    st = [Guarded(tuple(eng.E.shape), torch.float32), Guarded(tuple(eng.E.shape), torch.float32),
          Guarded(tuple(eng.Rel.shape), torch.float32), Guarded(tuple(eng.Rel.shape), torch.float32)]
    eng.opt_state = (ko.OPT_LAZY_ADAM, st[0].view, st[1].view, st[2].view, st[3].view)
    rng = np.random.default_rng(1)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    pos[-1] = [N - 1, R - 1, N - 1]       # the last row of every table is referenced
    neg = None if fused_neg else ko.philox_negatives(pos, N, K, 3, 1, ko.CORRUPT_EITHER).astype(np.int32)
    res = eng.train_step(pos, neg, opt_kind=ko.OPT_LAZY_ADAM, lr=1e-3, K=K, loss_kind=ko.PAIRWISE_LOGISTIC, margin=0.0, seed=3, offset=1,
                         want_scores=True, want_neg=True, dropout=0.25 if kind not in (ko.HOLE,) else 0.0)
    torch.cuda.synchronize()
    assert np.isfinite(float(res["loss"].item()))
    for name, buf in list(g.items()) + [(f"state{i}", b) for i, b in enumerate(st)]:
        assert buf.intact(), name
    assert float(eng.gE.abs().max()) == 0.0 and int(eng.tE.sum()) == 0   # handed back clean


@pytest.mark.parametrize("kind,W", [(ko.TRANSE_L1, 200), (ko.COMPLEX, 200), (ko.HOLE, 36), (ko.PROTATE, 52), (ko.ROTATE, 64)])
def test_rank_counts_write_only_inside_their_buffers(kind, W):
    from KGEkeras import _abi as abi
    N, R, Q = 1003, 7, 333      # neither a multiple of the 64 / 128 / 256 tile sizes
    eng = _engine(kind, N, R, W, seed=W + 1)
    rng = np.random.default_rng(2)
    tri = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int32)
    tri[-1] = [N - 1, R - 1, N - 1]
    for side in (0, 1):
        c = Guarded((Q,), torch.int32)
        eng.rank_counts(tri, side, counts=c.view, row_range=(17, N))
        torch.cuda.synchronize()
        assert c.intact() and int(c.view.min()) >= 0 and int(c.view.max()) < N
        if eng.tc_supported():
            c2 = Guarded((Q,), torch.int32)
            eng.rank_counts_tc(tri, side, counts=c2.view, row_range=(17, N), mode=abi.TC_BF16X3_EXACT)
            torch.cuda.synchronize()
            assert c2.intact()
            assert torch.equal(c.view, c2.view)


def test_score_and_negatives_outputs_stay_inside():
    eng = _engine(ko.COMPLEX, 517, 5, 136, seed=9)
    rng = np.random.default_rng(3)
    M = 1001
    idx = np.stack([rng.integers(0, 517, M), rng.integers(0, 5, M), rng.integers(0, 517, M)], 1).astype(np.int32)
    out = Guarded((M,), torch.float32)
    eng.score(idx, out=out.view)
    torch.cuda.synchronize()
    assert out.intact() and bool(torch.isfinite(out.view).all())
    neg = eng.negatives(idx, 7, 5, 9)
    assert tuple(neg.shape) == (7 * M, 3) and int(neg[:, [0, 2]].max()) < 517 and int(neg.min()) >= 0


@pytest.mark.parametrize("case", ["hole_fft_fused", "hole_fft_unfused", "hole_fft_dropout", "hake", "hake_dropout"])
def test_fourier_hole_and_hake_write_only_inside_their_buffers(case):
    """The Fourier-domain HolE kernels (fused step and the score_fwd / score_bwd pair) and the HAKE kernels, on tables whose last
    row is referenced and with a batch that does not divide the warp ranges."""
    from KGEkeras._engine import Engine
    N, R, B, K = 1003, 7, 301, 5
    rng = np.random.default_rng(4)
    if case.startswith("hole"):
        W = 64
        eng = Engine(ko.HOLE, N, R, W, W, 12.0, E=ko.glorot_uniform(rng, N, W) * 3, R=ko.glorot_uniform(rng, R, W) * 3)
        eng.set_hole_mode(True)
        loss = ko.POINTWISE_LOGISTIC
    else:
        d = 18
        eng = Engine(ko.HAKE, N, R, 2 * d, 3 * d, 12.0, E=ko.glorot_uniform(rng, N, 2 * d), R=ko.glorot_uniform(rng, R, 3 * d),
                     param=ko.hake_params(12.0, 2.0, 2 * d, 0.5, 1.0))
        loss = ko.PAIRWISE_LOGISTIC
    eng._ensure_train_buffers(ko.OPT_LAZY_ADAM)
    g = {"gE": Guarded(tuple(eng.E.shape), torch.float32), "gR": Guarded(tuple(eng.Rel.shape), torch.float32),
         "tE": Guarded((N,), torch.uint8), "tR": Guarded((R,), torch.uint8)}
    eng.gE, eng.gR, eng.tE, eng.tR = g["gE"].view, g["gR"].view, g["tE"].view, g["tR"].view
    st = [Guarded(tuple(eng.E.shape), torch.float32), Guarded(tuple(eng.E.shape), torch.float32),
          Guarded(tuple(eng.Rel.shape), torch.float32), Guarded(tuple(eng.Rel.shape), torch.float32)]
    eng.opt_state = (ko.OPT_LAZY_ADAM, st[0].view, st[1].view, st[2].view, st[3].view)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    pos[-1] = [N - 1, R - 1, N - 1]
    res = eng.train_step(pos, None, opt_kind=ko.OPT_LAZY_ADAM, lr=1e-3, K=K, loss_kind=loss, margin=0.0, seed=3, offset=1, want_scores=True,
                         want_neg=True, dropout=0.25 if case.endswith("dropout") else 0.0, force_unfused=case == "hole_fft_unfused")
    torch.cuda.synchronize()
    assert np.isfinite(float(res["loss"].item()))
    assert res["neg"].shape == (B * K, 3) and int(res["neg"].max()) < N
    for name, buf in list(g.items()) + [(f"state{i}", b) for i, b in enumerate(st)]:
        assert buf.intact(), name
    assert float(eng.gE.abs().max()) == 0.0 and int(eng.tE.sum()) == 0
