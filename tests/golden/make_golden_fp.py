"""Freeze floating-point golden vectors made by the REFERENCE'S OWN SOURCE (run in the build container only):

    python tests/golden/make_golden_fp.py          ->  tests/golden/fp_ref.npz

`oracle/ref_tf_exec.py` loads /root/reference/KGEkeras/models.py and utils.py unmodified over a torch-backed `tensorflow`
stand-in; this script drives them the way the reference harness does (tests.py:63-87: `embedding_model(triple)`,
`embedding_model(ftriple)`, `loss_function(score, fscore, margin, 1)`, `model.add_loss(loss)`) and stores, for every case of
tests/golden/fp_cases.py, the inputs and

  * `score32` / `score64`: `EmbeddingModel.call` (models.py:134-162) -> `func`, in fp32 (the reference's dtype) and in fp64
    (the same source, the truth the parity bars refer to);
  * `gE*`, `gR*`: autograd of sum_i c_i * score_i with a random c -- the dense form of the IndexedSlices gradient;
  * losses: the seven functions of utils.py:288-319 on seeded (true, false) vectors, value and d/dtrue, d/dfalse;
  * steps: model + loss composed like the harness, loss value and table gradients (plus Dense kernel / bias for the
    literal branch, models.py:114-120,152-158, and the l3 regulariser of models.py:10-11).

Everything is seeded; re-running reproduces the file bit for bit (tests/test_cpu_fp_goldens.py re-checks that live when
/root/reference is mounted)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import fp_cases as fc                      # noqa: E402
from oracle import ref_tf_exec as rx      # noqa: E402


def tables_for(name, kind, ew, rw, d, rng):
    """fp32 tables.  HAKE: small phases/moduli and biases on both sides of the clip at 1 and of -|mod_p| (models.py:394-395)."""
    N, R = fc.N_ENT, fc.N_REL
    if kind == "HAKE":
        E = (rng.normal(size=(N, ew * d)) * 0.05).astype(np.float32)
        Rm = (rng.normal(size=(R, rw * d)) * 0.6).astype(np.float32)
        Rm[:, 2 * d:] += 0.5
        return E, Rm
    lim_e, lim_r = np.sqrt(6.0 / (N + ew * d)), np.sqrt(6.0 / (R + rw * d))
    E = (rng.uniform(-lim_e, lim_e, size=(N, ew * d)) * 3).astype(np.float32)
    Rm = (rng.uniform(-lim_r, lim_r, size=(R, rw * d)) * 3).astype(np.float32)
    return E, Rm


def triples(rng, n):
    t = np.stack([rng.integers(0, fc.N_ENT, n), rng.integers(0, fc.N_REL, n), rng.integers(0, fc.N_ENT, n)], 1).astype(np.int32)
    t[:3, 2] = t[:3, 0]          # self loops: one row gathered as subject and object
    t[5] = t[4]                  # a repeated triple: two references to the same rows
    return t


def build_model(rm, cls_name, kwargs, d, E, Rm, extra=None):
    """Construct the reference class as the harness does (tests.py:65-71: e_dim = r_dim = dim); returns (model, note)."""
    kw = dict(e_dim=d, r_dim=d, num_entities=E.shape[0], num_relations=Rm.shape[0], init_entities=E, init_relations=Rm,
              dp=0.0, name="embedding_model")
    kw.update(kwargs)
    kw.update(extra or {})
    cls = getattr(rm, cls_name)
    if cls_name != "ModE":
        return cls(**kw), ""
    # ModE.__init__ reads self.gamma before anything assigned it (models.py:415), and stores relation rows 3/2 as wide as
    # entity rows (409-410) which `s * p` (418) cannot broadcast: the class cannot be constructed or called in the reference.
    # Its `func` (models.py:417-418) is still the contract: run it on a stand-in carrying gamma and norm, with the lookups
    # of EmbeddingModel.call done by a DistMult instance over equally wide tables.
    try:
        cls(e_dim=d, r_dim=d, num_entities=4, num_relations=2, norm=kw["norm"])
        note = "ModE() constructed"
    except AttributeError as e:
        note = "ModE() raises %s: %s; func run on a stand-in" % (type(e).__name__, e)
    host = rm.DistMult(**{k: v for k, v in kw.items() if k not in ("norm", "gamma")})
    host.gamma, host.norm = kw.get("gamma", 12), kw["norm"]
    host.func = types.MethodType(cls.func, host)
    return host, note


def run_score_case(rm, case, prec):
    import torch
    name, cls_name, kwargs, kind, ew, rw, d = case
    rng = np.random.default_rng(abs(hash_name(name)))
    E, Rm = tables_for(name, kind, ew, rw, d, rng)
    idx = triples(rng, fc.N_TRIPLES)
    c = rng.normal(size=len(idx)).astype(np.float32)
    rx.set_precision(prec)
    model, note = build_model(rm, cls_name, kwargs, d, E, Rm)
    sc = model(idx)
    w = torch.as_tensor(c.astype(np.float64)).to(sc.dtype)
    torch.sum(w * sc).backward()
    out = dict(E=E, R=Rm, idx=idx, c=c, score=sc.detach().numpy(), gE=model.entity_embedding.embeddings.grad.numpy(),
               gR=model.relational_embedding.embeddings.grad.numpy(), note=note)
    if cls_name == "pRotatE":
        out["param"] = np.array([model.modulus], np.float64)
    if cls_name == "HAKE":
        out["param"] = np.array([model.embedding_range / model.pi, model.phase_weight, model.mod_weight], np.float64)
    out["gamma"] = np.array([float(model.gamma)] if hasattr(model, "gamma") else [12.0])
    return out


def hash_name(name):
    import zlib
    return zlib.crc32(name.encode())


def run_loss_case(ru, lname, mode, margin, prec):
    import torch
    rng = np.random.default_rng(hash_name(lname + mode) + int(margin * 8))
    B, K = fc.LOSS_B, fc.LOSS_K
    t = (rng.normal(size=B) * 2).astype(np.float32)
    f = (rng.normal(size=B * K) * 2).astype(np.float32)
    if lname == "pointwize_cross_entropy":      # half of the scores inside Keras' clip interval, the rest clipped
        t[::2], f[::2] = rng.uniform(0.01, 0.99, len(t[::2])), rng.uniform(0.01, 0.99, len(f[::2]))
    rx.set_precision(prec)
    dt = torch.float32 if prec == 32 else torch.float64
    tt = torch.as_tensor(t.astype(np.float64)).to(dt).requires_grad_(True)
    ff = torch.as_tensor(f.astype(np.float64)).to(dt).requires_grad_(True)
    fn = ru.loss_function_lookup(lname)
    if not lname.startswith("pairwize"):
        L = fn(tt, ff, margin, K)                   # pointwise: two independent means, the shapes do not interact
    elif mode == "PAIRED":
        L = fn(tt[:, None], ff, margin, K)          # true (B,1) against false (B*K,) -> (B,K)
    else:
        L = fn(tt, ff, margin, 1)                   # the harness call, tests.py:81: (B*K,1) against (B,) -> (B*K,B)
    L.backward()
    return dict(t=t, f=f, loss=np.array([L.item()]), dt=tt.grad.numpy(), df=ff.grad.numpy())


def run_step_case(rm, ru, case, prec):
    import torch
    name, base, lname, mode, extra = case
    _, cls_name, kwargs, kind, ew, rw, d = next(c for c in fc.SCORE_CASES if c[0] == base)
    rng = np.random.default_rng(hash_name(name))
    E, Rm = tables_for(name, kind, ew, rw, d, rng)
    B, K = fc.STEP_B, fc.STEP_K
    pos = triples(rng, B)
    if mode == "PAIRED":
        neg = np.repeat(pos, K, axis=0)
        m = len(neg)
        side = rng.integers(0, 2, m)
        repl = rng.integers(0, fc.N_ENT, m).astype(np.int32)
        neg[side == 0, 0] = repl[side == 0]
        neg[side == 1, 2] = repl[side == 1]
    else:                                            # two equally long lists, as oversample_data leaves them
        neg = triples(rng, B)
        neg[:, 2] = rng.integers(0, fc.N_ENT, B)
    extra = dict(extra)
    lit = W = b = None
    if "literals" in extra:
        wl = extra["literals"]
        lit = rng.random((fc.N_ENT, wl)).astype(np.float32)                   # tests.py:113: np.random.rand(len(E), 5)
        extra["literals"] = lit
        W = (rng.normal(size=(ew * d + wl, ew * d)) * 0.3).astype(np.float32)
        b = (rng.normal(size=ew * d) * 0.1).astype(np.float32)
    rx.set_precision(prec)
    model, _ = build_model(rm, cls_name, kwargs, d, E, Rm, extra)
    if lit is not None:
        model.literal_activation.set_weights(W, b)
    t, f = model(pos), model(neg)
    fn = ru.loss_function_lookup(lname)
    margin = 1.0
    L = fn(t[:, None], f, margin, K) if (mode == "PAIRED" and lname.startswith("pairwize")) else fn(t, f, margin, K if mode == "PAIRED" else 1)
    reg = rx.model_regularization_loss(model)        # Keras adds the regulariser's value to the loss `fit` minimises
    total = L if reg is None else L + reg
    total.backward()
    out = dict(E=E, R=Rm, pos=pos, neg=neg, K=np.array([K]), margin=np.array([margin]), loss=np.array([L.item()]),
               total=np.array([total.item()]), pos_scores=t.detach().numpy(), neg_scores=f.detach().numpy(),
               gE=model.entity_embedding.embeddings.grad.numpy(), gR=model.relational_embedding.embeddings.grad.numpy())
    if reg is not None:
        out["reg"] = np.array([reg.item()])
        out["reg_w"] = np.array([extra["regularization"]])
    if lit is not None:
        out.update(L=lit, W=W, b=b, gW=model.literal_activation.kernel.grad.numpy(), gb=model.literal_activation.bias.grad.numpy())
    if cls_name == "HAKE":
        out["param"] = np.array([model.embedding_range / model.pi, model.phase_weight, model.mod_weight], np.float64)
    return out


INPUT_KEYS = {"E", "R", "idx", "c", "t", "f", "pos", "neg", "K", "margin", "L", "W", "b", "param", "gamma", "reg_w", "note"}


def generate():
    """dict key -> array: inputs once, outputs with a 32 / 64 suffix."""
    assert rx.available(), "needs /root/reference (build container)"
    rm, ru, _ = rx.load(32)
    out = {}

    def merge(prefix, r32, r64):
        for k, v in r32.items():
            if k in INPUT_KEYS:
                out[f"{prefix}/{k}"] = np.asarray(v)
            else:
                out[f"{prefix}/{k}32"] = np.asarray(v, dtype=np.float32)
                out[f"{prefix}/{k}64"] = np.asarray(r64[k], dtype=np.float64)

    for case in fc.SCORE_CASES:
        merge("score/" + case[0], run_score_case(rm, case, 32), run_score_case(rm, case, 64))
    for lname in fc.LOSS_NAMES:
        for mode in (("PAIRED", "ALLPAIRS") if lname.startswith("pairwize") else ("PAIRED",)):
            for margin in fc.LOSS_MARGINS:
                merge(f"loss/{lname}/{mode}/m{margin}", run_loss_case(ru, lname, mode, margin, 32), run_loss_case(ru, lname, mode, margin, 64))
    for case in fc.STEP_CASES:
        merge("step/" + case[0], run_step_case(rm, ru, case, 32), run_step_case(rm, ru, case, 64))
    rx.set_precision(32)
    return out


if __name__ == "__main__":
    data = generate()
    np.savez_compressed(fc.FIXTURE, **data)
    print("wrote", fc.FIXTURE, len(data), "arrays,", os.path.getsize(fc.FIXTURE), "bytes")
