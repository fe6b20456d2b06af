"""CPU restatement of the KGE-Keras triple-scoring hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to /root/reference).
Two precisions are offered everywhere: `np.float32` mirrors the reference op by op (TF computes in
fp32, NumPy broadcasting == TF broadcasting), `np.float64` is the "truth" the fp32 numbers are
judged against.

Parity status: the NumPy functions are pinned by tests/golden/ (made by running the reference's own
code through oracle/ref_ast.py); the TF-typed functions (scores, losses, regulariser, literal branch,
gradients) by tests/golden/fp_ref.npz (the reference's models.py / utils.py executed unmodified over
the torch-backed `tensorflow` stand-in of oracle/ref_tf_exec.py); the optimizer arithmetic is a
restatement of Keras' published update rules, cross-checked against torch.optim only.
"""
from __future__ import annotations

import numpy as np

# score kinds -- numeric values are shared with include/kgeb200.h (kgeb_score_kind)
DISTMULT, TRANSE_L1, TRANSE_L2, COMPLEX, HOLE = 0, 1, 2, 3, 4
# sibling score functions: RotatE models.py:421-459, pRotatE 461-489, ModE 403-418 (norm 1 | 2)
ROTATE, PROTATE, MODE_L1, MODE_L2 = 5, 6, 7, 8
SCORE_KINDS = {"DistMult": DISTMULT, "TransE_L1": TRANSE_L1, "TransE_L2": TRANSE_L2,
               "ComplEx": COMPLEX, "HolE": HOLE}
SIBLING_KINDS = {"RotatE": ROTATE, "pRotatE": PROTATE, "ModE_L1": MODE_L1, "ModE_L2": MODE_L2}
# HAKE (models.py:362-400): entity rows [phase | modulus] (2d), relation rows [phase | modulus | bias] (3d).  Its modulus term is
# ONE Frobenius norm over the whole call (tf.norm without axis, models.py:397), so the scores of a call are coupled: they are
# restated by hake_scores / hake_row_grads below, not by the per-triple score / score_grads
HAKE = 9

# loss kinds -- shared with kgeb_loss_kind
POINTWISE_HINGE, POINTWISE_LOGISTIC, POINTWISE_SQUARE, PAIRWISE_HINGE, PAIRWISE_LOGISTIC, PAIRWISE_SQUARE, POINTWISE_BCE = range(7)
LOSS_NAMES = {  # utils.py:321-330 (spelling "wize" is the reference's)
    "pointwize_hinge": POINTWISE_HINGE, "pointwize_logistic": POINTWISE_LOGISTIC,
    "pointwize_square_loss": POINTWISE_SQUARE, "pairwize_hinge": PAIRWISE_HINGE,
    "pairwize_logistic": PAIRWISE_LOGISTIC, "pairwize_square_loss": PAIRWISE_SQUARE,
    "pointwize_cross_entropy": POINTWISE_BCE,
}
# pair modes -- shared with kgeb_pair_mode
PAIRED, ALLPAIRS = 0, 1
# optimizers -- shared with kgeb_opt_kind
OPT_SGD, OPT_ADAGRAD, OPT_LAZY_ADAM, OPT_ADAM_DENSE = range(4)
# corruption modes of the Philox sampler -- shared with kgeb_corrupt_mode
CORRUPT_BOTH, CORRUPT_HEAD, CORRUPT_TAIL, CORRUPT_EITHER = range(4)

EPSILON = 1e-6  # utils.py:12


# --------------------------------------------------------------------------------------------
# tables
# --------------------------------------------------------------------------------------------
def glorot_uniform(rng: np.random.Generator, rows: int, cols: int, dtype=np.float32):
    """models.py:87-88 GlorotUniform: U(+-sqrt(6/(fan_in+fan_out))) with fan = (rows, cols)."""
    lim = np.sqrt(6.0 / (rows + cols))
    return rng.uniform(-lim, lim, size=(rows, cols)).astype(dtype)


def dyadic_table(rng: np.random.Generator, rows: int, cols: int, lo=-4, hi=4, denom=4.0):
    """Small dyadic rationals k/denom: every product/sum the score functions form is exact in
    fp32 (and in bf16 x bf16 -> fp32), so scores and ranks are independent of summation order."""
    return (rng.integers(lo, hi + 1, size=(rows, cols)) / denom).astype(np.float32)


# --------------------------------------------------------------------------------------------
# score functions  (models.py:134-237)
# --------------------------------------------------------------------------------------------
def gather(E, R, idx):
    """EmbeddingModel.call, models.py:140-150: s,p,o = unstack(inputs,axis=1); Embedding lookups."""
    idx = np.asarray(idx).astype(np.int64)
    return E[idx[:, 0]], R[idx[:, 1]], E[idx[:, 2]]


def circular_correlation_fft(a, b):
    """HolE.func, models.py:231-234: real(ifft(conj(fft(a)) * fft(b))) in complex64 (or complex128)."""
    ctype = np.complex64 if a.dtype == np.float32 else np.complex128
    fa = np.fft.fft(a.astype(ctype), axis=-1)
    fb = np.fft.fft(b.astype(ctype), axis=-1)
    return np.real(np.fft.ifft(np.conj(fa) * fb, axis=-1)).astype(a.dtype)


def circular_correlation_direct(a, b):
    """x_k = sum_i a_i b_{(i+k) mod d}  (equals the FFT form)."""
    d = a.shape[-1]
    out = np.zeros_like(a)
    for k in range(d):
        out[..., k] = np.sum(a * np.roll(b, -k, axis=-1), axis=-1)
    return out


def score(kind, s, p, o, gamma=12.0, hole_fft=True, param=1.0):
    """func(s,p,o) of the four hot-path models; dtype follows the inputs.

    DistMult models.py:171-172, TransE 186-190, ComplEx 212-222, HolE 231-237."""
    dt = s.dtype
    if kind == DISTMULT:
        return np.sum(s * p * o, axis=-1)
    if kind in (TRANSE_L1, TRANSE_L2):
        r = s + p - o
        n = np.sum(np.abs(r), axis=-1) if kind == TRANSE_L1 else np.sqrt(np.sum(r * r, axis=-1))
        if gamma > 0:  # models.py:187-190
            return (dt.type(gamma) - n).astype(dt)
        return n
    if kind == COMPLEX:
        d = s.shape[-1] // 2
        sr, si = s[..., :d], s[..., d:]
        pr, pi = p[..., :d], p[..., d:]
        orr, oi = o[..., :d], o[..., d:]
        s1 = sr * pr * orr
        s2 = pr * si * oi
        s3 = pi * sr * oi
        s4 = pi * si * orr
        return np.sum(s1 + s2 + s3 - s4, axis=-1)
    if kind == HOLE:
        x = circular_correlation_fft(s, o) if hole_fft else circular_correlation_direct(s, o)
        return np.sum(p * x, axis=-1)
    if kind == ROTATE:   # models.py:439-459: rows [re | im], p = phases (width d); SIGNED sum of the residuals, no norm
        d = s.shape[-1] // 2
        ph = np.arctan2(np.sin(p), np.cos(p))
        c, sn = np.cos(ph), np.sin(ph)
        re = s[..., :d] * c - s[..., d:] * sn - o[..., :d]
        im = s[..., :d] * sn + s[..., d:] * c - o[..., d:]
        a = np.sum(np.concatenate([re, im], axis=-1), axis=-1)
        return (dt.type(gamma) - a).astype(dt) if gamma > 0 else a
    if kind == PROTATE:  # models.py:479-489; `param` = modulus * embedding_range
        ph = lambda x: np.arctan2(np.sin(x), np.cos(x))
        a = np.sum(np.abs(np.sin((ph(s) + ph(p) - ph(o)) / 2)), axis=-1) * dt.type(param)
        return (dt.type(gamma) - a).astype(dt) if gamma > 0 else a
    if kind in (MODE_L1, MODE_L2):  # models.py:417-418 (with self.gamma assigned, Appendix D)
        r = s * p - o
        n = np.sum(np.abs(r), axis=-1) if kind == MODE_L1 else np.sqrt(np.sum(r * r, axis=-1))
        return (dt.type(gamma) - n).astype(dt)
    raise KeyError(kind)


def score_triples(kind, E, R, idx, gamma=12.0, dtype=np.float32, hole_fft=True, param=1.0):
    """EmbeddingModel.call with dp inactive (models.py:134-162): (B,3) ids -> (B,) scores."""
    s, p, o = gather(E.astype(dtype), R.astype(dtype), idx)
    return score(kind, s, p, o, gamma, hole_fft, param)


def score_abs_terms(kind, E, R, idx, gamma=12.0):
    """sum of |terms| entering each score, fp64: the scale an fp32 summation error is relative to."""
    s, p, o = gather(E.astype(np.float64), R.astype(np.float64), idx)
    if kind == DISTMULT:
        return np.sum(np.abs(s * p * o), axis=-1)
    if kind in (TRANSE_L1, TRANSE_L2):
        return np.abs(gamma) + np.sum(np.abs(s) + np.abs(p) + np.abs(o), axis=-1)
    if kind == COMPLEX:
        d = s.shape[-1] // 2
        a = np.abs(s[..., :d]) + np.abs(s[..., d:])
        b = np.abs(p[..., :d]) + np.abs(p[..., d:])
        c = np.abs(o[..., :d]) + np.abs(o[..., d:])
        return np.sum(a * b * c, axis=-1)
    if kind == HOLE:
        return np.sum(np.abs(p) * circular_correlation_direct(np.abs(s), np.abs(o)), axis=-1)
    if kind == ROTATE:    # |re_s| + |im_s| bounds both rotated components; plus the candidate's own entries
        return np.abs(gamma) + 2.0 * np.sum(np.abs(s), axis=-1) + np.sum(np.abs(o), axis=-1)
    if kind == PROTATE:   # d|sin(y/2)|/dy <= 1/2: the argument's rounding scale bounds the term's
        return np.abs(gamma) + np.sum(np.abs(s) + np.abs(p) + np.abs(o), axis=-1)
    if kind in (MODE_L1, MODE_L2):
        return np.abs(gamma) + np.sum(np.abs(s * p) + np.abs(o), axis=-1)
    raise KeyError(kind)


def score_grads(kind, s, p, o, gamma=12.0, param=1.0):
    """Analytic d score / d (s, p, o); checked against torch autograd in tests."""
    if kind == DISTMULT:
        return p * o, s * o, s * p
    if kind in (TRANSE_L1, TRANSE_L2):
        r = s + p - o
        if kind == TRANSE_L1:
            g = np.sign(r)
        else:
            g = r / np.sqrt(np.sum(r * r, axis=-1, keepdims=True))
        sg = -1.0 if gamma > 0 else 1.0
        g = (sg * g).astype(s.dtype)
        return g, g, -g
    if kind == COMPLEX:
        d = s.shape[-1] // 2
        sr, si = s[..., :d], s[..., d:]
        pr, pi = p[..., :d], p[..., d:]
        orr, oi = o[..., :d], o[..., d:]
        ds = np.concatenate([pr * orr + pi * oi, pr * oi - pi * orr], axis=-1)
        dp = np.concatenate([sr * orr + si * oi, sr * oi - si * orr], axis=-1)
        do = np.concatenate([sr * pr - pi * si, pr * si + pi * sr], axis=-1)
        return ds, dp, do
    if kind == HOLE:
        d = s.shape[-1]
        dp = circular_correlation_direct(s, o)           # dp_k = sum_i s_i o_{i+k}
        ds = np.zeros_like(s)
        do = np.zeros_like(o)
        for k in range(d):
            ds += p[..., k:k + 1] * np.roll(o, -k, axis=-1)   # ds_i = sum_k p_k o_{i+k}
            do += p[..., k:k + 1] * np.roll(s, k, axis=-1)    # do_j = sum_k p_k s_{j-k}
        return ds, dp, do
    sg = -1.0 if gamma > 0 else 1.0
    if kind == ROTATE:
        d = s.shape[-1] // 2
        c, sn = np.cos(p), np.sin(p)
        sr, si = s[..., :d], s[..., d:]
        ds = sg * np.concatenate([c + sn, c - sn], axis=-1)
        dp = sg * (sr * (c - sn) - si * (c + sn))
        do = sg * -np.ones_like(o)
        return ds, dp, do
    if kind == PROTATE:
        y = (s + p - o) / 2
        g = sg * param * np.sign(np.sin(y)) * np.cos(y) / 2
        return g, g, -g
    if kind in (MODE_L1, MODE_L2):
        r = s * p - o
        g = np.sign(r) if kind == MODE_L1 else r / np.sqrt(np.sum(r * r, axis=-1, keepdims=True))
        g = -g
        return g * p, g * s, -g
    raise KeyError(kind)


# --------------------------------------------------------------------------------------------
# losses  (utils.py:288-319) -- NumPy broadcasting == TF broadcasting
# --------------------------------------------------------------------------------------------
def _c(dt):
    # EPSILON+1 is a Python float added to an fp32 tensor -> fp32(1.000001) = 1.00000095367431640625
    # (the fp64 "truth" uses the same fp32-rounded constant: it is the reference's constant, not an error to judge)
    return dt.type(np.float32(EPSILON + 1))


def pointwize_hinge(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:288-289"""
    dt = true.dtype
    return np.mean(np.maximum(dt.type(margin) - true, 0)) + np.mean(np.maximum(dt.type(margin) + false, 0))


def pointwize_logistic(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:291-292"""
    c = _c(true.dtype)
    with np.errstate(over="ignore"):
        return np.mean(np.log(c + np.exp(-true))) + np.mean(np.log(c + np.exp(false)))


def pointwize_square_loss(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:294-295"""
    dt = true.dtype
    return np.mean(np.square(dt.type(margin) - true)) + np.mean(np.square(dt.type(margin) + false))


def pointwize_cross_entropy(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:297-298; Keras binary_crossentropy on raw scores (third-party, recalled: clip to
    [1e-7, 1-1e-7], -(y log(p+eps) + (1-y) log(1-p+eps)), mean over the last axis)."""
    dt = true.dtype
    eps = dt.type(1e-7)

    def bce(y, pred):
        pred = np.clip(pred, eps, dt.type(1) - eps)
        return np.mean(-(y * np.log(pred + eps) + (1 - y) * np.log(1 - pred + eps)), axis=-1)
    return bce(dt.type(1), true) + bce(dt.type(0), false)


def pairwize_hinge(true, false, margin=1, negative_samples=1, reduce_mean=True):
    """utils.py:300-305"""
    false = np.reshape(false, (-1, negative_samples))
    tmp = np.maximum(true.dtype.type(margin) + false - true, 0)
    return np.mean(tmp) if reduce_mean else tmp


def pairwize_logistic(true, false, margin=0, negative_samples=1, reduce_mean=True):
    """utils.py:307-312"""
    false = np.reshape(false, (-1, negative_samples))
    with np.errstate(over="ignore"):
        tmp = np.log(_c(true.dtype) + np.exp(false - true))
    return np.mean(tmp) if reduce_mean else tmp


def pairwize_square_loss(true, false, margin=0, negative_samples=1, reduce_mean=True):
    """utils.py:314-319 (sign as written in the reference)"""
    false = np.reshape(false, (-1, negative_samples))
    tmp = -np.square(false - true)
    return np.mean(tmp) if reduce_mean else tmp


def loss_function_lookup(name):
    """utils.py:321-330"""
    return {
        "pointwize_hinge": pointwize_hinge,
        "pointwize_logistic": pointwize_logistic,
        "pointwize_cross_entropy": pointwize_cross_entropy,
        "pointwize_square_loss": pointwize_square_loss,
        "pairwize_hinge": pairwize_hinge,
        "pairwize_logistic": pairwize_logistic,
        "pairwize_square_loss": pairwize_square_loss,
    }[name]


def loss_and_coefs(loss_kind, pair_mode, t, f, K, margin):
    """Loss value and dL/dt (B,), dL/df (Bf,) in the dtype of t.

    PAIRED   : true is (B,1), false (B*K,) -> (B,K)            
    ALLPAIRS : true is (m,), false (m',) reshaped (m',1): the harness' (m',m) broadcast, tests.py:81
    pointwise kinds ignore pair_mode."""
    dt = t.dtype
    M = dt.type(margin)
    c = _c(dt)
    B, Bf = t.shape[0], f.shape[0]
    with np.errstate(over="ignore", invalid="ignore"):
        if loss_kind == POINTWISE_HINGE:
            L = np.mean(np.maximum(M - t, 0)) + np.mean(np.maximum(M + f, 0))
            return L, -(M - t > 0).astype(dt) / B, (M + f > 0).astype(dt) / Bf
        if loss_kind == POINTWISE_LOGISTIC:
            et, ef = np.exp(-t), np.exp(f)
            L = np.mean(np.log(c + et)) + np.mean(np.log(c + ef))
            return L, -(et / (c + et)) / B, (ef / (c + ef)) / Bf
        if loss_kind == POINTWISE_SQUARE:
            L = np.mean(np.square(M - t)) + np.mean(np.square(M + f))
            return L, -2 * (M - t) / B, 2 * (M + f) / Bf
        if loss_kind == POINTWISE_BCE:   # utils.py:297-298 (Keras clip to [eps, 1-eps], gradient passes inside the interval)
            eps, hi = dt.type(1e-7), dt.type(1) - dt.type(1e-7)
            ct, cf = np.clip(t, eps, hi), np.clip(f, eps, hi)
            L = np.mean(-np.log(ct + eps)) + np.mean(-np.log(dt.type(1) - cf + eps))
            it, jf = ((t >= eps) & (t <= hi)).astype(dt), ((f >= eps) & (f <= hi)).astype(dt)
            return L, -it / (ct + eps) / B, jf / (dt.type(1) - cf + eps) / Bf
        if pair_mode == PAIRED:
            F = f.reshape(B, K)
            T = t.reshape(B, 1)
        else:
            F = f.reshape(Bf, 1)
            T = t.reshape(1, B)
        X = F - T
        cnt = dt.type(X.size)
        if loss_kind == PAIRWISE_HINGE:
            A = (M + X > 0).astype(dt)
            L = np.mean(np.maximum(M + X, 0))
        elif loss_kind == PAIRWISE_LOGISTIC:
            ex = np.exp(X)
            A = ex / (c + ex)
            L = np.mean(np.log(c + ex))
        elif loss_kind == PAIRWISE_SQUARE:
            A = -2 * X
            L = np.mean(-np.square(X))
        else:
            raise KeyError(loss_kind)
        if pair_mode == PAIRED:
            return L, -A.sum(axis=1) / cnt, (A / cnt).reshape(-1)
        return L, -A.sum(axis=0) / cnt, A.sum(axis=1) / cnt


# --------------------------------------------------------------------------------------------
# backward + optimizer (third-party Keras/TF semantics, recalled)
# --------------------------------------------------------------------------------------------
def dense_grads(kind, E, R, idx, dscore, gamma=12.0, dtype=np.float64, param=1.0):
    """IndexedSlices -> dense: sum of dscore_m * dscore/drow over every reference to a row."""
    s, p, o = gather(E.astype(dtype), R.astype(dtype), idx)
    ds, dp, do = score_grads(kind, s, p, o, gamma, param)
    w = np.asarray(dscore, dtype=dtype)[:, None]
    gE = np.zeros(E.shape, dtype)
    gR = np.zeros(R.shape, dtype)
    idx = np.asarray(idx).astype(np.int64)
    np.add.at(gE, idx[:, 0], w * ds)
    np.add.at(gE, idx[:, 2], w * do)
    np.add.at(gR, idx[:, 1], w * dp)
    tE = np.zeros(E.shape[0], bool)
    tR = np.zeros(R.shape[0], bool)
    tE[idx[:, 0]] = True
    tE[idx[:, 2]] = True
    tR[idx[:, 1]] = True
    return gE, gR, tE, tR


def dense_grads_abs(kind, E, R, idx, dscore, gamma=12.0, param=1.0):
    """sum over the references to a row of |dscore_m * dscore/drow| in fp64: the magnitude the parity bar of a gradient entry is
    stated against (1e-5 of the summed magnitudes of its terms, like the scores; BASELINE north star)."""
    s, p, o = gather(E.astype(np.float64), R.astype(np.float64), idx)
    if kind == COMPLEX:        # an entry is a signed sum of two products (pr*or + pi*oi, pr*oi - pi*or, ...): its terms are the products
        d = s.shape[-1] // 2
        (sr, si), (pr, pi), (orr, oi) = [(np.abs(x[..., :d]), np.abs(x[..., d:])) for x in (s, p, o)]
        ds = np.concatenate([pr * orr + pi * oi, pr * oi + pi * orr], axis=-1)
        dp = np.concatenate([sr * orr + si * oi, sr * oi + si * orr], axis=-1)
        do = np.concatenate([sr * pr + pi * si, pr * si + pi * sr], axis=-1)
    elif kind == HOLE:         # every term of the correlation carries a plus sign: the same formulas on |inputs|
        ds, dp, do = score_grads(kind, np.abs(s), np.abs(p), np.abs(o), gamma, param)
    else:
        ds, dp, do = score_grads(kind, s, p, o, gamma, param)
        ds, dp, do = np.abs(ds), np.abs(dp), np.abs(do)
    if kind == TRANSE_L2:      # an entry is r_k / ||r|| with r_k = s_k + p_k - o_k: its terms are the three operands, not their (cancelling) sum
        nr = np.maximum(np.linalg.norm(s + p - o, axis=1, keepdims=True), 1e-300)
        ds = dp = do = (np.abs(s) + np.abs(p) + np.abs(o)) / nr
    w = np.abs(np.asarray(dscore, dtype=np.float64))[:, None]
    aE, aR = np.zeros(E.shape, np.float64), np.zeros(R.shape, np.float64)
    idx = np.asarray(idx).astype(np.int64)
    np.add.at(aE, idx[:, 0], w * ds)
    np.add.at(aE, idx[:, 2], w * do)
    np.add.at(aR, idx[:, 1], w * dp)
    return aE, aR


def l3_reg(E, w):
    """models.py:10-11: w * ||E||_3^3 = w * sum |e|^3 and its dense gradient 3 w e |e| (entity table only, models.py:92)."""
    return w * np.sum(np.abs(E) ** 3), 3.0 * w * E * np.abs(E)


def opt_update(opt_kind, P, G, touched, state, lr, step, b1=0.9, b2=0.999, eps=1e-7):
    """One optimizer step on a table, in the dtype of P.  state = (m, v) for Adam, (acc,) for Adagrad.

    SGD: p -= lr g.  Adagrad (Keras: initial accumulator 0.1): acc += g^2; p -= lr g/(sqrt(acc)+eps).
    Adam (Keras OptimizerV2): lr_t = lr sqrt(1-b2^t)/(1-b1^t); m = b1 m+(1-b1) g; v = b2 v+(1-b2) g^2;
    p -= lr_t m/(sqrt(v)+eps).  LAZY_ADAM applies it to touched rows only; ADAM_DENSE to every row with
    g = 0 for untouched ones (what Keras' sparse apply does: dense decay of m and v)."""
    dt = P.dtype
    P = P.copy()
    state = [x.copy() for x in state]
    rows = np.ones(P.shape[0], bool) if opt_kind == OPT_ADAM_DENSE else np.asarray(touched, bool)
    g = np.where(np.asarray(touched, bool)[:, None], G, 0).astype(dt)[rows]
    if opt_kind == OPT_SGD:
        P[rows] -= dt.type(lr) * g
    elif opt_kind == OPT_ADAGRAD:
        acc = state[0]
        acc[rows] += g * g
        P[rows] -= dt.type(lr) * g / (np.sqrt(acc[rows]) + dt.type(eps))
    else:
        m, v = state
        lr_t = dt.type(lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step))
        m[rows] = dt.type(b1) * m[rows] + dt.type(1 - b1) * g
        v[rows] = dt.type(b2) * v[rows] + dt.type(1 - b2) * g * g
        P[rows] -= lr_t * m[rows] / (np.sqrt(v[rows]) + dt.type(eps))
    return P, state


def train_step(kind, loss_kind, pair_mode, E, R, pos, neg, K, margin, gamma, opt_kind, state_E, state_R,
               lr, step, dtype=np.float64, b1=0.9, b2=0.999, eps=1e-7):
    """One full step of the harness' loop (tests.py:72-87,130-134) with dp=0, regularization=0:
    scores -> loss -> gradients (dedup by id) -> optimizer.  Returns dict."""
    E = E.astype(dtype)
    R = R.astype(dtype)
    t = score_triples(kind, E, R, pos, gamma, dtype, hole_fft=False)
    f = score_triples(kind, E, R, neg, gamma, dtype, hole_fft=False)
    L, dt_, df_ = loss_and_coefs(loss_kind, pair_mode, t, f, K, margin)
    idx = np.concatenate([np.asarray(pos), np.asarray(neg)], axis=0)
    gE, gR, tE, tR = dense_grads(kind, E, R, idx, np.concatenate([dt_, df_]), gamma, dtype)
    E2, sE = opt_update(opt_kind, E, gE, tE, [x.astype(dtype) for x in state_E], lr, step, b1, b2, eps)
    R2, sR = opt_update(opt_kind, R, gR, tR, [x.astype(dtype) for x in state_R], lr, step, b1, b2, eps)
    return dict(loss=L, pos_scores=t, neg_scores=f, dpos=dt_, dneg=df_, gE=gE, gR=gR, touched_E=tE,
                touched_R=tR, E=E2, R=R2, state_E=sE, state_R=sR)


# --------------------------------------------------------------------------------------------
# Philox4x32-10 counter-based sampler (new in the B200 build; the host restatement pins the kernel)
# --------------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (...,4) uint32, key: (...,2) uint32 -> (...,4) uint32.  Random123 Philox4x32-10."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _PH_M0 * c[0].astype(np.uint64)
            p1 = _PH_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _U32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _U32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            if r < 9:
                k0 = (k0 + _PH_W0).astype(np.uint32)
                k1 = (k1 + _PH_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_negatives(pos, N, K, seed, offset, mode=CORRUPT_BOTH):
    """Negatives exactly as the CUDA sampler draws them (kgeb_negatives / fused train step).

    For positive i and j < K, g = i*K + j: (x,y,z,w) = Philox(ctr=(g_lo, g_hi, off_lo, off_hi), key=seed);
    head' = (x*N)>>32, tail' = (y*N)>>32; BOTH replaces both (reference behaviour, utils.py:117-120),
    HEAD/TAIL one side, EITHER picks the side from the top bit of z (1 -> head).  Relation is kept."""
    pos = np.asarray(pos, dtype=np.int64)
    B = pos.shape[0]
    g = np.arange(B * K, dtype=np.uint64)
    ctr = np.stack([(g & _U32), (g >> np.uint64(32)),
                    np.full_like(g, np.uint64(offset) & _U32), np.full_like(g, np.uint64(offset) >> np.uint64(32))],
                   axis=-1).astype(np.uint32)
    key = np.empty((B * K, 2), np.uint32)
    key[:, 0] = np.uint64(seed) & _U32
    key[:, 1] = np.uint64(seed) >> np.uint64(32)
    r = philox4x32_10(ctr, key).astype(np.uint64)
    nh = ((r[:, 0] * np.uint64(N)) >> np.uint64(32)).astype(np.int64)
    nt = ((r[:, 1] * np.uint64(N)) >> np.uint64(32)).astype(np.int64)
    head_side = (r[:, 2] >> np.uint64(31)).astype(bool)
    rep = np.repeat(pos, K, axis=0)
    out = rep.copy()
    if mode == CORRUPT_BOTH:
        out[:, 0], out[:, 2] = nh, nt
    elif mode == CORRUPT_HEAD:
        out[:, 0] = nh
    elif mode == CORRUPT_TAIL:
        out[:, 2] = nt
    else:
        out[:, 0] = np.where(head_side, nh, rep[:, 0])
        out[:, 2] = np.where(head_side, rep[:, 2], nt)
    return out


# --------------------------------------------------------------------------------------------
# reference-style host negatives (utils.py:114-147) -- restated; pinned by tests/golden/negatives_ref.npz
# --------------------------------------------------------------------------------------------
def generate_negative(kg, N, negative=2, check_kg=False, corrupt_head=True, corrupt_tail=True):
    """utils.py:114-121: both s and o uniform in [0,N) (fs drawn first in full, then fo); relation
    repeated `negative` times consecutively; output (negative*len(kg), 3, 1); flags ignored."""
    assert corrupt_head or corrupt_tail
    R = np.repeat(np.asarray([p for _, p, _ in kg]).reshape((-1, 1)), negative, axis=0)
    fs = np.random.randint(0, N, size=(negative * len(kg), 1))
    fo = np.random.randint(0, N, size=(negative * len(kg), 1))
    return np.stack([fs, R, fo], axis=1)


def oversample_data(kgs):
    """utils.py:123-147 without the unused x/y branch: each list resampled WITH replacement to the
    longest length via random.choices, then squeezed."""
    from random import choices
    kgs = [list(kg) for kg in kgs]
    m = max(map(len, kgs))
    return [np.squeeze(np.asarray(choices(kg, k=m))) for kg in kgs]


# --------------------------------------------------------------------------------------------
# filtered ranking (utils.py:165-264)
# --------------------------------------------------------------------------------------------
def build_filters(filtering_triples):
    """utils.py:201-205"""
    from collections import defaultdict
    filter_h, filter_t = defaultdict(set), defaultdict(set)
    for s, p, o in filtering_triples:
        filter_h[(p, o)].add(s)
        filter_t[(s, p)].add(o)
    return filter_h, filter_t


def all_entity_scores(kind, E, R, triples, side, gamma=12.0, dtype=np.float32, hole_fft=False, param=1.0):
    """(Q,N) matrix: score of every candidate entity in the `side` slot of each test triple, computed
    exactly like call() would on the (C+1,3) candidate array of gen_tail_data/gen_head_data."""
    triples = np.asarray(triples, dtype=np.int64)
    N = E.shape[0]
    out = np.empty((len(triples), N), dtype)
    cand = np.arange(N, dtype=np.int64)
    for qi, (s, p, o) in enumerate(triples):
        idx = np.empty((N, 3), np.int64)
        idx[:, 0] = cand if side == "head" else s
        idx[:, 1] = p
        idx[:, 2] = cand if side == "tail" else o
        out[qi] = score_triples(kind, E, R, idx, gamma, dtype, hole_fft, param)
    return out


def ranks_from_scores(scores, triples, side, filters):
    """rank = 1 + #{c not in F u {target} : score_c > score_target}  (utils.py:211-215).
    `filters` is the dict of sets from build_filters for this side."""
    triples = np.asarray(triples, dtype=np.int64)
    ranks = np.empty(len(triples), np.int64)
    for qi, (s, p, o) in enumerate(triples):
        tgt = o if side == "tail" else s
        key = (s, p) if side == "tail" else (p, o)
        keep = np.ones(scores.shape[1], bool)
        fl = list(filters.get(key, ()))
        if fl:
            keep[np.asarray(fl, dtype=np.int64)] = False
        keep[tgt] = False
        ranks[qi] = 1 + int(np.count_nonzero(scores[qi][keep] > scores[qi][tgt]))
    return ranks


def metrics_from_ranks(tail_ranks, head_ranks):
    """utils.py:216-262: the 15-key dict; denominators len(test_data); combined = (tail+head)/2."""
    def side(r):
        r = np.asarray(r, dtype=np.int64)
        n = float(len(r))
        return (np.count_nonzero(r <= 1) / n, np.count_nonzero(r <= 3) / n, np.count_nonzero(r <= 10) / n,
                np.mean(r), np.mean([1 / m for m in r]))
    t1, t3, t10, tmr, tmrr = side(tail_ranks)
    h1, h3, h10, hmr, hmrr = side(head_ranks)
    return {'tail_hits@1': t1, 'tail_hits@3': t3, 'tail_hits@10': t10, 'tail_mr': tmr, 'tail_mrr': tmrr,
            'head_hits@1': h1, 'head_hits@3': h3, 'head_hits@10': h10, 'head_mr': hmr, 'head_mrr': hmrr,
            'hits@1': (t1 + h1) / 2, 'hits@3': (t3 + h3) / 2, 'hits@10': (t10 + h10) / 2,
            'mr': (tmr + hmr) / 2, 'mrr': (tmrr + hmrr) / 2}


def validate(kind, E, R, test_data, filtering_triples, gamma=12.0, dtype=np.float32, chunk=256):
    """Vectorised restatement of validate (utils.py:199-264) for the four hot-path models."""
    filter_h, filter_t = build_filters(filtering_triples)
    test = np.asarray(test_data, dtype=np.int64)
    tr, hr = [], []
    for i in range(0, len(test), chunk):
        part = test[i:i + chunk]
        tr.append(ranks_from_scores(all_entity_scores(kind, E, R, part, "tail", gamma, dtype), part, "tail", filter_t))
        hr.append(ranks_from_scores(all_entity_scores(kind, E, R, part, "head", gamma, dtype), part, "head", filter_h))
    tr, hr = np.concatenate(tr), np.concatenate(hr)
    return metrics_from_ranks(tr, hr), tr, hr


# --------------------------------------------------------------------------------------------
# 1-N query vectors and the sequential-k arithmetic of the CUDA-core rank kernel
# --------------------------------------------------------------------------------------------
def query_vectors(kind, E, R, triples, side, dtype=np.float32):
    """Per-query vector q such that score(candidate c) = op(q, E[c]):
    DistMult tail s*p / head p*o; ComplEx tail [sr pr - si pi | si pr + sr pi], head
    [pr or + pi oi | pr oi - pi or]; HolE tail q_j = sum_i s_i p_{(j-i)}, head q_i = sum_k p_k o_{(i+k)};
    TransE tail a = s+p (dist |a - e|), head: (p, o) kept separately so (e+p)-o keeps the reference's order."""
    triples = np.asarray(triples, dtype=np.int64)
    s, p, o = gather(E.astype(dtype), R.astype(dtype), triples)
    if kind == DISTMULT:
        return s * p if side == "tail" else p * o
    if kind == COMPLEX:
        d = s.shape[-1] // 2
        sr, si, pr, pi, orr, oi = s[:, :d], s[:, d:], p[:, :d], p[:, d:], o[:, :d], o[:, d:]
        if side == "tail":
            return np.concatenate([sr * pr - si * pi, si * pr + sr * pi], axis=1)
        return np.concatenate([pr * orr + pi * oi, pr * oi - pi * orr], axis=1)
    if kind == HOLE:
        d = s.shape[-1]
        q = np.zeros_like(s)
        if side == "tail":
            for i in range(d):   # q_j = sum_i s_i p_{(j-i) mod d}, sequential in i
                q = q + s[:, i:i + 1] * np.roll(p, i, axis=1)
        else:
            for k in range(d):   # q_i = sum_k p_k o_{(i+k) mod d}, sequential in k
                q = q + p[:, k:k + 1] * np.roll(o, -k, axis=1)
        return q
    if kind in (TRANSE_L1, TRANSE_L2, PROTATE):
        return s + p if side == "tail" else np.concatenate([p, o], axis=1)
    if kind in (MODE_L1, MODE_L2):   # tail r = s*p - e, head r = e*p - o
        return s * p if side == "tail" else np.concatenate([p, o], axis=1)
    if kind == ROTATE:
        # the score is linear in the candidate: tail gamma - A = const + sum_j e_j, head = const - e . [(c+s)|(c-s)]
        # (constants cancel inside one query's ranking); sign flips when gamma <= 0 -- the caller passes gamma_sign
        c, sn = np.cos(p), np.sin(p)
        return np.ones_like(s) if side == "tail" else -np.concatenate([c + sn, c - sn], axis=1)
    raise KeyError(kind)


def seqk_scores(kind, q, Erows, side, gamma=12.0, param=1.0):
    """Scores op(q, e) accumulated sequentially over k in fp32 with separate multiply and add roundings --
    bit-for-bit what kgeb_rank_1n's CUDA-core kernels compute (they use __fmul_rn/__fadd_rn, no FMA)."""
    q = q.astype(np.float32)
    Er = Erows.astype(np.float32)
    Q, Nn, W = q.shape[0], Er.shape[0], Er.shape[1]
    acc = np.zeros((Q, Nn), np.float32)
    if kind in (DISTMULT, COMPLEX, HOLE, ROTATE):
        if kind == ROTATE and gamma <= 0:
            q = -q
        for k in range(W):
            acc = acc + q[:, k:k + 1] * Er[None, :, k]
        return acc
    for k in range(W):
        if side == "tail":
            r = q[:, k:k + 1] - Er[None, :, k]
        elif kind in (MODE_L1, MODE_L2):
            r = Er[None, :, k] * q[:, k:k + 1] - q[:, W + k:W + k + 1]
        else:
            r = (Er[None, :, k] + q[:, k:k + 1]) - q[:, W + k:W + k + 1]
        if kind == PROTATE:
            acc = acc + np.abs(np.sin(r * np.float32(0.5)))
        else:
            acc = acc + (np.abs(r) if kind in (TRANSE_L1, MODE_L1) else r * r)
    if kind in (TRANSE_L2, MODE_L2):
        acc = np.sqrt(acc)
    if kind == PROTATE:
        acc = acc * np.float32(param)
    if kind in (MODE_L1, MODE_L2):
        return np.float32(gamma) - acc
    return (np.float32(gamma) - acc) if gamma > 0 else acc


def filter_csr(triples, side, filters):
    """Sorted per-query filter lists in CSR form (offsets int64[Q+1], ids int32[nnz]); target removed."""
    triples = np.asarray(triples, dtype=np.int64)
    offs = np.zeros(len(triples) + 1, np.int64)
    ids = []
    for qi, (s, p, o) in enumerate(triples):
        key = (s, p) if side == "tail" else (p, o)
        tgt = o if side == "tail" else s
        fl = sorted(int(x) for x in filters.get(key, ()) if int(x) != tgt)
        ids.extend(fl)
        offs[qi + 1] = len(ids)
    return offs, np.asarray(ids, dtype=np.int32)


# --------------------------------------------------------------------------------------------
# Dropout masks exactly as the kernels draw them (csrc/frag.cuh: dp_bits / dp_keep_elem); models.py:122,144-150
# --------------------------------------------------------------------------------------------
def dropout_keep(kind, width, n, slot, stream, p, seed, offset):
    """(n, width) bool keep mask of row-slot `slot` (0/1/2 = s/p/o) for the n triples of call `stream` (0 = positive call,
    1 = negative call).  16 Philox bits per element, keep iff bits >= round(p*65536).  `width` = stored row width."""
    thr = int(p * 65536.0 + 0.5)
    split = kind == COMPLEX or (kind == ROTATE and slot != 1)   # RotatE: entity rows are [re | im], the relation row is plain
    if split:
        d = width // 2
        nch = d // 4
    else:
        d = width
        nch = (width + 3) // 4
    NC = (nch + 31) // 32
    x = np.arange(width)
    half = (x >= d).astype(np.int64) if split else np.zeros(width, np.int64)
    xx = x - half * d
    chunk, e = xx // 4, xx % 4
    lane, i = chunk % 32, chunk // 32
    t = i + half * NC
    call, h = t // 2, t % 2
    g = np.arange(n, dtype=np.uint64)
    ctr = np.zeros((n, width, 4), np.uint32)
    ctr[..., 0] = (32 * call + lane).astype(np.uint32)[None, :]
    ctr[..., 1] = np.uint32(((int(offset) << 3) & 0xFFFFFFFF) | slot | (stream << 2))
    ctr[..., 2] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    ctr[..., 3] = ((g >> np.uint64(32)).astype(np.uint32) ^ np.uint32((int(offset) >> 29) & 0xFFFFFFFF))[:, None]
    key = np.zeros((n, width, 2), np.uint32)
    key[..., 0] = np.uint32((int(seed) & 0xFFFFFFFF) ^ 0x44524F50)
    key[..., 1] = np.uint32((int(seed) >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key)
    word_idx = (2 * h + (e >= 2)).astype(np.int64)        # h ? (e<2 ? z : w) : (e<2 ? x : y)
    word = np.take_along_axis(r, np.broadcast_to(word_idx[None, :, None], (n, width, 1)), axis=2)[..., 0]
    u = (word >> (16 * (e % 2)).astype(np.uint32)[None, :]) & np.uint32(0xFFFF)
    return u >= thr


def dropout_mult(kind, width, n, slot, stream, p, seed, offset, dtype=np.float64):
    """keep * fp32(1/(1-p)) as an (n, width) multiplier: forward x*mult, backward grad*mult."""
    keep = dropout_keep(kind, width, n, slot, stream, p, seed, offset)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return keep.astype(dtype) * dtype(scale)


def train_step_dropout(kind, loss_kind, E, R, pos, neg, K, margin, gamma, p, seed, offset, dtype=np.float64):
    """Forward + dense gradients of one PAIRED step with Dropout(p) on every gathered row (independent masks for the
    positive call, stream 0, and the negative call, stream 1).  Returns loss, scores, gE, gR."""
    E, R = E.astype(dtype), R.astype(dtype)
    We, Wr = E.shape[1], R.shape[1]
    out = {}
    rows, mults = [], []
    for stream, idx in ((0, np.asarray(pos)), (1, np.asarray(neg))):
        s, pr, o = gather(E, R, idx)
        ms = dropout_mult(kind, We, len(idx), 0, stream, p, seed, offset, dtype)
        mp = dropout_mult(kind, Wr, len(idx), 1, stream, p, seed, offset, dtype)
        mo = dropout_mult(kind, We, len(idx), 2, stream, p, seed, offset, dtype)
        rows.append((s * ms, pr * mp, o * mo))
        mults.append((ms, mp, mo))
    t = score(kind, *rows[0], gamma, hole_fft=False)
    f = score(kind, *rows[1], gamma, hole_fft=False)
    L, dt_, df_ = loss_and_coefs(loss_kind, PAIRED, t, f, K, margin)
    gE, gR = np.zeros(E.shape, dtype), np.zeros(R.shape, dtype)
    for (s, pr, o), (ms, mp, mo), idx, dsc in ((rows[0], mults[0], np.asarray(pos), dt_), (rows[1], mults[1], np.asarray(neg), df_)):
        ds, dp_, do = score_grads(kind, s, pr, o, gamma)
        w = np.asarray(dsc, dtype)[:, None]
        np.add.at(gE, idx[:, 0].astype(np.int64), w * ds * ms)
        np.add.at(gE, idx[:, 2].astype(np.int64), w * do * mo)
        np.add.at(gR, idx[:, 1].astype(np.int64), w * dp_ * mp)
    return dict(loss=L, pos_scores=t, neg_scores=f, gE=gE, gR=gR)


# --------------------------------------------------------------------------------------------
# literal branch (models.py:114-120,152-158; trained by the reference harness, tests.py:113,125-126)
# --------------------------------------------------------------------------------------------
ACT_LINEAR, ACT_TANH, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3


def _act(act, z):
    if act == ACT_TANH:
        return np.tanh(z)
    if act == ACT_RELU:
        return np.maximum(z, 0)
    if act == ACT_SIGMOID:
        return 1.0 / (1.0 + np.exp(-z))
    return z


def _act_grad(act, y):
    if act == ACT_TANH:
        return 1.0 - y * y
    if act == ACT_RELU:
        return (y > 0).astype(y.dtype)
    if act == ACT_SIGMOID:
        return y * (1.0 - y)
    return np.ones_like(y)


def literal_table(E, L, W, b, act, dtype=np.float64):
    """E' = act([E | L] W + b): what every entity row becomes without dropout (inference)."""
    X = np.concatenate([E.astype(dtype), L.astype(dtype)], axis=1)
    return _act(act, X @ W.astype(dtype) + b.astype(dtype))


def literal_step(kind, loss_kind, pair_mode, E, R, L, W, b, act, pos, neg, K, margin, gamma, p=0.0, seed=0, offset=0, dtype=np.float64):
    """Forward + gradients of one step of the literal model: s' = act([dp(E[s]) | dp(L[s])] W + b), o' likewise, relation rows
    dp(R[p]); masks as the kernels draw them (row index of the materialised list and column as the counter; stream 0 = entity
    rows, the 2m references ordered [all s | all o]; stream 1 = the m relation rows).  Returns loss, scores, gE, gR, gW, gb."""
    E, R, L, W, b = (np.asarray(a, dtype=dtype) for a in (E, R, L, W, b))
    tri = np.concatenate([np.asarray(pos), np.asarray(neg)], axis=0).astype(np.int64)
    m, B = len(tri), len(pos)
    we, wl, wr = E.shape[1], L.shape[1], R.shape[1]
    refs = np.concatenate([tri[:, 0], tri[:, 2]])
    if p:
        mX = dropout_mult(DISTMULT, we + wl, 2 * m, 0, 0, p, seed, offset, dtype)
        mP = dropout_mult(DISTMULT, wr, m, 0, 1, p, seed, offset, dtype)
    else:
        mX, mP = np.ones((2 * m, we + wl), dtype), np.ones((m, wr), dtype)
    X = np.concatenate([E[refs], L[refs]], axis=1) * mX
    Y = _act(act, X @ W + b)
    P = R[tri[:, 1]] * mP
    S, O = Y[:m], Y[m:]
    sc = score(kind, S, P, O, gamma, hole_fft=False)
    t, f = sc[:B], sc[B:]
    Lval, dt_, df_ = loss_and_coefs(loss_kind, pair_mode, t, f, K, margin)
    w = np.concatenate([dt_, df_])[:, None]
    ds, dp_, do = score_grads(kind, S, P, O, gamma)
    gY = np.concatenate([w * ds, w * do], axis=0)
    gZ = gY * _act_grad(act, Y)
    gb = gZ.sum(0)
    gW = X.T @ gZ
    gX = (gZ @ W.T) * mX
    gE, gR = np.zeros(E.shape, dtype), np.zeros(R.shape, dtype)
    np.add.at(gE, refs, gX[:, :we])
    np.add.at(gR, tri[:, 1], w * dp_ * mP)
    return dict(loss=Lval, pos_scores=t, neg_scores=f, gE=gE, gR=gR, gW=gW, gb=gb)


def philox_oversample_indices(n, m, seed, offset, stream_id=0):
    """Indices exactly as kgeb_oversample draws them (device-side oversample_data, utils.py:123-147)."""
    i = np.arange(m, dtype=np.uint64)
    ctr = np.stack([(i & _U32).astype(np.uint32), (i >> np.uint64(32)).astype(np.uint32),
                    np.full(m, int(offset) & 0xFFFFFFFF, np.uint32), np.full(m, (int(offset) >> 32) & 0xFFFFFFFF, np.uint32)], axis=-1)
    key = np.stack([np.full(m, (int(seed) & 0xFFFFFFFF) ^ 0x4F565253, np.uint32),
                    np.full(m, ((int(seed) >> 32) & 0xFFFFFFFF) ^ int(stream_id), np.uint32)], axis=-1)
    r = philox4x32_10(ctr, key)
    return ((r[:, 0].astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


# --------------------------------------------------------------------------------------------
# HAKE (models.py:362-400) -- batch-coupled score, restated as written
# --------------------------------------------------------------------------------------------
def hake_params(gamma=12.0, epsilon=2.0, e_dim_stored=None, phase_weight=0.5, mod_weight=1.0):
    """(phase_div, phase_weight, mod_weight): phase_div = embedding_range / pi -- what models.py:390-392 DIVIDE the phases by -- with embedding_range =
    (gamma + epsilon) / e_dim / 2 and e_dim the STORED entity width 2d (models.py:371,381: e_dim is doubled before super().__init__)."""
    rng_ = (float(gamma) + float(epsilon)) / float(e_dim_stored) / 2.0
    return (float(rng_ / np.pi), float(phase_weight), float(mod_weight))


def _hake_parts(s, p, o, phase_div):
    d = s.shape[-1] // 2
    assert p.shape[-1] == 3 * d and o.shape[-1] == 2 * d
    a = (s[..., :d] / phase_div + p[..., :d] / phase_div - o[..., :d] / phase_div) / 2     # models.py:390-392,398: three divisions, op for op
    ms, mo = s[..., d:], o[..., d:]
    mp, beta = p[..., d:2 * d], p[..., 2 * d:]
    b1 = np.minimum(beta, 1.0)                                     # K.clip(bias, -inf, 1), models.py:394
    cond = b1 < -np.abs(mp)                                        # tf.where(bias < -|mod_p|, -|mod_p|, bias), models.py:395
    b = np.where(cond, -np.abs(mp), b1)
    u = ms * (mp + b) - np.abs(mo) * (1 - b)                       # models.py:397
    return d, a, ms, mo, mp, beta, b, cond, u


def hake_scores(s, p, o, gamma, phase_div, phase_weight, mod_weight):
    """gamma - (phase_weight * sum_k |sin a_k|  +  mod_weight * ||U||_F): the modulus norm runs over EVERY row of the call."""
    _, a, *_, u = _hake_parts(s, p, o, phase_div)
    r = mod_weight * np.sqrt(np.sum(u * u))
    return gamma - (phase_weight * np.sum(np.abs(np.sin(a)), axis=-1) + r)


def hake_row_grads(s, p, o, dscore, gamma, phase_div, phase_weight, mod_weight):
    """d(sum_i dscore_i * score_i) / d(s, p, o rows) with TensorFlow's conventions: d|x| = sign(x) (0 at 0), the clip passes the
    gradient where bias <= 1, the where() routes it to -|mod_p| or to the bias.  A zero modulus norm contributes no gradient
    (TensorFlow would produce NaN there)."""
    d, a, ms, mo, mp, beta, b, cond, u = _hake_parts(s, p, o, phase_div)
    c = np.asarray(dscore, dtype=s.dtype)[:, None]
    T = np.sum(u * u)
    dphi = -phase_weight * c * np.sign(np.sin(a)) * np.cos(a) * 0.5 / phase_div
    gu = (-(np.sum(c) * mod_weight) / np.sqrt(T)) * u if T > 0 else np.zeros_like(u)
    db = gu * (ms + np.abs(mo))
    g_ms = gu * (mp + b)
    g_mo = gu * (-np.sign(mo) * (1 - b))
    g_mp = gu * ms + np.where(cond, db * (-np.sign(mp)), 0.0)
    g_beta = np.where(cond, 0.0, np.where(beta <= 1.0, db, 0.0))
    return (np.concatenate([dphi, g_ms], axis=-1), np.concatenate([dphi, g_mp, g_beta], axis=-1), np.concatenate([-dphi, g_mo], axis=-1))


def hake_step(loss_kind, pair_mode, E, R, pos, neg, K, margin, gamma, hp, p=0.0, seed=0, offset=0, dtype=np.float64):
    """One training step's forward + dense gradients: the positive and the negative call are two separate func() calls
    (models.py:144-150), each with its own modulus norm and, with p > 0, its own dropout masks (plain-row counter layout)."""
    E, R = E.astype(dtype), R.astype(dtype)
    calls = []
    for stream, idx in ((0, np.asarray(pos)), (1, np.asarray(neg))):
        s, pr, o = gather(E, R, idx)
        if p > 0:
            ms_ = dropout_mult(HAKE, E.shape[1], len(idx), 0, stream, p, seed, offset, dtype)
            mp_ = dropout_mult(HAKE, R.shape[1], len(idx), 1, stream, p, seed, offset, dtype)
            mo_ = dropout_mult(HAKE, E.shape[1], len(idx), 2, stream, p, seed, offset, dtype)
        else:
            ms_ = mp_ = mo_ = None
        rows = (s, pr, o) if ms_ is None else (s * ms_, pr * mp_, o * mo_)
        calls.append((idx, rows, (ms_, mp_, mo_)))
    t = hake_scores(*calls[0][1], gamma, *hp)
    f = hake_scores(*calls[1][1], gamma, *hp)
    L, dt_, df_ = loss_and_coefs(loss_kind, pair_mode, t, f, K, margin)
    gE, gR = np.zeros(E.shape, dtype), np.zeros(R.shape, dtype)
    for (idx, rows, mults), dsc in zip(calls, (dt_, df_)):
        ds, dp_, do = hake_row_grads(*rows, dsc, gamma, *hp)
        if mults[0] is not None:
            ds, dp_, do = ds * mults[0], dp_ * mults[1], do * mults[2]
        np.add.at(gE, idx[:, 0].astype(np.int64), ds)
        np.add.at(gE, idx[:, 2].astype(np.int64), do)
        np.add.at(gR, idx[:, 1].astype(np.int64), dp_)
    return dict(loss=L, pos_scores=t, neg_scores=f, gE=gE, gR=gR)

