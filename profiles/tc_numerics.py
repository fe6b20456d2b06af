#!/usr/bin/env python
"""Known-answer experiments for profiles/micro/tc_numerics.cu: how does tcgen05.mma kind::f16 (fp32 accumulate) align, sum
and round the 16 products of one K=16 instruction and the incoming accumulator?

    python profiles/tc_numerics.py gen  <dir>      # writes <dir>/<exp>.in + <dir>/manifest.json            (anywhere)
    for e in <dir>/*.in; do profiles/micro/tc_numerics $e ${e%.in}.out; done                                 (on a B200)
    python profiles/tc_numerics.py analyze <dir>   # prints / writes <dir>/summary.json                      (anywhere)

Case c of an experiment lives in row c of A and row c of B; its result is D[c][c].  Every value is exactly representable in
the operand format, so the EXACT sum of products is known in rational arithmetic and the returned fp32 bits tell which
low-order bits the hardware kept.

Families
  align    p0 = +-1, then n products of +-2^-s in the SAME instruction (s = 20..40): how far below ulp(max) does an addend survive?
  accal    the 1 arrives through the accumulator (previous instruction), the small products in the next one
  cancel   +1, -1 and 2^-s: is the sum exact before the final rounding (the 2^-s must survive)?
  round    1 + m * 2^-26 (m = 1..15 eighths of an ulp), all signs: truncation, round-to-nearest, sticky?
  random   dense random operands, K = 256 x nrep: observed |D - exact| against sum|terms| and ||a|| ||b|| (what a bound has to cover)
"""
import json
import os
import sys
from fractions import Fraction

import numpy as np

M, N = 128, 256


def enc(x, fmt):
    """python float / array -> operand bits (uint16); asserts exact representability."""
    x = np.asarray(x, dtype=np.float64)
    if fmt == 0:
        h = x.astype(np.float16)
        assert np.all(h.astype(np.float64) == x), "not fp16-representable"
        return h.view(np.uint16)
    f = x.astype(np.float32)
    assert np.all(f.astype(np.float64) == x)
    b = f.view(np.uint32)
    assert np.all((b & 0xFFFF) == 0), "not bf16-representable"
    return (b >> 16).astype(np.uint16)


def rnd_to(x, fmt):
    """round float64 array to the operand format (for the random family)."""
    if fmt == 0:
        return x.astype(np.float16).astype(np.float64)
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
    return b.astype(np.uint32).view(np.float32).astype(np.float64)


def write_exp(path, A, B, fmt, nseg=1, nrep=1):
    ktot = A.shape[1]
    with open(path, "wb") as fh:
        np.asarray([ktot, fmt, nseg, nrep], np.int32).tofile(fh)
        enc(A, fmt).tofile(fh)
        enc(B, fmt).tofile(fh)


def small_pair(s, fmt):
    """(a, b) with a*b = 2^-s, both representable and normal in fmt (fp16: exponents >= -14)."""
    ea = s // 2
    eb = s - ea
    if fmt == 0:
        assert max(ea, eb) <= 14, s
    return 2.0 ** -ea, 2.0 ** -eb


def cases_align(fmt, via_acc):
    """list of dicts(case) + A, B.  via_acc: K = 32, the big product sits in instruction 0, the small ones in instruction 1."""
    smax = 28 if fmt == 0 else 40
    ktot = 32 if via_acc else 16
    A, B = np.zeros((M, ktot)), np.zeros((N, ktot))
    cases = []
    c = 0
    for big_sign in (1, -1):
        for small_sign in (1, -1):
            for n in (1, 3, 15):
                for s in range(20, smax + 1, 1 if n == 15 else 2):
                    if c >= M:
                        break
                    A[c, 0], B[c, 0] = big_sign, 1.0
                    a, b = small_pair(s, fmt)
                    off = 16 if via_acc else 1
                    for k in range(n):
                        A[c, off + k], B[c, off + k] = small_sign * a, b
                    cases.append(dict(case=c, big=big_sign, small=small_sign, n=n, s=s))
                    c += 1
    return cases, A, B


def cases_cancel(fmt):
    smax = 28 if fmt == 0 else 60
    A, B = np.zeros((M, 32)), np.zeros((N, 32))
    cases = []
    c = 0
    for layout in ("same", "acc"):
        for s in range(2, smax + 1, 1 if fmt == 0 else 2):
            if c >= M:
                break
            a, b = small_pair(s, fmt)
            if layout == "same":    # +1, -1, 2^-s in one instruction
                A[c, 0], B[c, 0] = 1.0, 1.0
                A[c, 1], B[c, 1] = -1.0, 1.0
                A[c, 2], B[c, 2] = a, b
            else:                   # +1 first, then (-1, 2^-s) against the accumulator
                A[c, 0], B[c, 0] = 1.0, 1.0
                A[c, 16], B[c, 16] = -1.0, 1.0
                A[c, 17], B[c, 17] = a, b
            cases.append(dict(case=c, layout=layout, s=s))
            c += 1
    return cases, A, B


def cases_round(fmt):
    A, B = np.zeros((M, 32)), np.zeros((N, 32))
    cases = []
    c = 0
    for layout in ("same", "acc"):
        for big_sign in (1, -1):
            for small_sign in (1, -1):
                for m in range(1, 16):
                    # small = m * 2^-26 = (m * 2^-13) * 2^-13 : m / 8 of ulp(1) = 2^-23
                    A[c, 0], B[c, 0] = big_sign, 1.0
                    k = 1 if layout == "same" else 16
                    A[c, k], B[c, k] = small_sign * m * 2.0 ** -13, 2.0 ** -13
                    cases.append(dict(case=c, layout=layout, big=big_sign, small=small_sign, m=m))
                    c += 1
    assert c <= M
    return cases, A, B


def cases_random(fmt, seed, dist):
    rng = np.random.default_rng(seed)
    K = 256
    if dist == "uniform":
        A, B = rng.uniform(-1, 1, (M, K)), rng.uniform(-1, 1, (N, K))
    elif dist == "positive":      # no cancellation: the accumulator grows steadily, every truncation has the same sign
        A, B = rng.uniform(0.5, 1, (M, K)), rng.uniform(0.5, 1, (N, K))
    else:                          # wide dynamic range: lognormal magnitudes
        A = rng.choice([-1, 1], (M, K)) * np.exp(rng.normal(0, 3, (M, K)))
        B = rng.choice([-1, 1], (N, K)) * np.exp(rng.normal(0, 3, (N, K)))
        A, B = np.clip(A, -6e4, 6e4), np.clip(B, -6e4, 6e4)
        A[np.abs(A) < 1e-4] = 0
        B[np.abs(B) < 1e-4] = 0
    return rnd_to(A, fmt), rnd_to(B, fmt)


def gen(d):
    os.makedirs(d, exist_ok=True)
    man = {}
    for fmt, fname in ((0, "fp16"), (1, "bf16")):
        for fam, (cases, A, B) in (("align", cases_align(fmt, False)), ("accal", cases_align(fmt, True)),
                                    ("cancel", cases_cancel(fmt)), ("round", cases_round(fmt))):
            name = f"{fam}_{fname}"
            write_exp(os.path.join(d, name + ".in"), A, B, fmt)
            np.savez_compressed(os.path.join(d, name + ".npz"), A=A, B=B)
            man[name] = dict(fmt=fmt, family=fam, cases=cases, nrep=1)
        for dist in ("uniform", "positive", "lognormal"):
            for nrep in (1, 6):
                A, B = cases_random(fmt, 5, dist)
                name = f"random_{dist}_{fname}_x{nrep}"
                write_exp(os.path.join(d, name + ".in"), A, B, fmt, nrep=nrep)
                np.savez_compressed(os.path.join(d, name + ".npz"), A=A, B=B)
                man[name] = dict(fmt=fmt, family="random", dist=dist, nrep=nrep)
    with open(os.path.join(d, "manifest.json"), "w") as fh:
        json.dump(man, fh)
    print("wrote", len(man), "experiments to", d)


def exact_diag(A, B, c):
    return sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(A[c], B[c]) if a != 0 and b != 0)


def f32_round(fr, mode):
    """Fraction -> nearest / truncated fp32 as float."""
    x = float(fr)                      # double: exact enough to locate the binade
    if x == 0:
        return 0.0
    e = int(np.floor(np.log2(abs(x))))
    ulp = Fraction(2) ** (e - 23)
    q = fr / ulp
    if mode == "rz":
        n = int(q) if q >= 0 else -int(-q)
    elif mode == "rn":
        n = int(np.floor(q))
        r = q - n
        if r > Fraction(1, 2) or (r == Fraction(1, 2) and n % 2 == 1):
            n += 1
    elif mode == "rd":
        n = int(np.floor(q))
    return float(n * ulp)


def analyze(d):
    with open(os.path.join(d, "manifest.json")) as fh:
        man = json.load(fh)
    summary = {}
    for name, info in man.items():
        out = os.path.join(d, name + ".out")
        if not os.path.isfile(out):
            continue
        D = np.fromfile(out, dtype=np.float32).reshape(M, N)
        z = np.load(os.path.join(d, name + ".npz"))
        A, B = z["A"], z["B"]
        if info["family"] == "random":
            nrep = info["nrep"]
            ex = nrep * (A @ B.T)                                   # float64: error ~1e-16 relative to sum|terms|
            sabs = nrep * (np.abs(A) @ np.abs(B).T)
            nn = nrep * np.outer(np.linalg.norm(A, axis=1), np.linalg.norm(B, axis=1))
            err = np.abs(D.astype(np.float64) - ex)
            summary[name] = dict(products=int(256 * nrep), max_err_over_sumabs=float((err / sabs).max()), max_err_over_norms=float((err / nn).max()),
                                 max_err_over_result=float((err / np.maximum(np.abs(ex), 1e-300)).max()),
                                 mean_signed_err_over_sumabs=float(((D.astype(np.float64) - ex) / sabs).mean()),
                                 u24=2.0 ** -24)
            continue
        rows = []
        for cs in info["cases"]:
            c = cs["case"]
            ex = exact_diag(A, B, c)
            got = float(D[c, c])
            rows.append(dict(cs, got=got.hex(), exact_rn=f32_round(ex, "rn").hex(), exact_rz=f32_round(ex, "rz").hex(),
                             is_rn=got == f32_round(ex, "rn"), is_rz=got == f32_round(ex, "rz"),
                             err_ulp1=float((Fraction(got) - ex) / Fraction(2) ** -23)))
        summary[name] = rows
    with open(os.path.join(d, "summary.json"), "w") as fh:
        json.dump(summary, fh, indent=0)
    # ---- readable digest ----
    for name, rows in summary.items():
        if isinstance(rows, dict):
            print(f"{name}: products={rows['products']} max|err|/sum|terms|={rows['max_err_over_sumabs']:.3e} (= {rows['max_err_over_sumabs'] / 2 ** -24:.2f} u24) "
                  f"max|err|/(|a||b|)={rows['max_err_over_norms']:.3e} mean signed={rows['mean_signed_err_over_sumabs']:.2e}")
            continue
        fam = man[name]["family"]
        print(f"== {name}")
        if fam in ("align", "accal"):
            for big in (1, -1):
                for small in (1, -1):
                    for n in (1, 3, 15):
                        sel = [r for r in rows if r["big"] == big and r["small"] == small and r["n"] == n]
                        print(f"  big={big:+d} small={small:+d} n={n:2d}: " + " ".join(f"s{r['s']}:{r['err_ulp1']:+.3f}" for r in sel))
        elif fam == "cancel":
            for lay in ("same", "acc"):
                sel = [r for r in rows if r["layout"] == lay]
                print(f"  {lay}: " + " ".join(f"s{r['s']}:{'ok' if r['is_rn'] else r['got']}" for r in sel))
        elif fam == "round":
            for lay in ("same", "acc"):
                for big in (1, -1):
                    for small in (1, -1):
                        sel = [r for r in rows if r["layout"] == lay and r["big"] == big and r["small"] == small]
                        print(f"  {lay} big={big:+d} small={small:+d}: " + " ".join(f"m{r['m']}:{r['err_ulp1']:+.3f}" for r in sel))


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("gen", "analyze"):
        print(__doc__)
        sys.exit(1)
    (gen if sys.argv[1] == "gen" else analyze)(sys.argv[2])
