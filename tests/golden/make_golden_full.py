"""Generate the BASELINE-config golden fixtures under tests/golden/full/ (run in the build container: needs /root/reference).

    python tests/golden/make_golden_full.py                  # every config of full_fixtures.CONFIGS, one after the other
    python tests/golden/make_golden_full.py WN18RR_TransE_L1_d200 ...   # the named ones (they are independent: run in parallel)

Per config:
  1. the tables are TRAINED: oracle/cpu_port.TorchCpuTrainer (the PyTorch-CPU restatement of the harness step, tests.py:55-87)
     runs `epochs` passes over the dataset's train split (FB15k-237: valid.txt, its train.txt is not in the mount) with
     reference-style negatives and the pairwise hinge;
  2. large tables are put on an int8 grid (+ seeded jitter, full_fixtures.expand_table) so that the fixture stays a few MB;
     the fixture's table is what the goldens are computed on -- it is trained, then quantised, never re-trained;
  3. the reference's OWN validate (utils.py:199-264, AST-loaded, unmodified) ranks the FULL test split against every entity,
     filtered by the dataset's other splits.  Its `model.predict` is the sequential-k fp32 score arithmetic of the rank
     contract (oracle.kge_oracle.seqk_scores), so the GPU path must reproduce every rank EXACTLY; per-query ranks are
     captured by handing the reference a recording wrapper around scipy.stats.rankdata (the reference's code is untouched:
     only the callable it finds under the name `rankdata` also remembers what it returned);
  4. small configs also store the ranks validate produces when predict sums in NumPy's own order (`alt_*`): the size of the
     near-tie noise between two fp32 summation orders of the same model.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import kge_oracle as ko  # noqa: E402
from oracle import ref_ast  # noqa: E402
import full_fixtures as ff  # noqa: E402

REF_DATA = "/root/reference/data"
KINDS = {"DistMult": ko.DISTMULT, "TransE_L1": ko.TRANSE_L1, "TransE_L2": ko.TRANSE_L2, "ComplEx": ko.COMPLEX, "HolE": ko.HOLE}


def ensure_fb15k(ref):
    """datasets/FB15k-237.npz from valid.txt + test.txt (sorted id mapping over what the mount has)."""
    path = os.path.join(HERE, "datasets", "FB15k-237.npz")
    if os.path.isfile(path):
        return
    splits = {sp: ref["load_kg"](os.path.join(REF_DATA, "FB15k-237", sp + ".txt")) for sp in ("valid", "test")}
    ents = sorted({t[0] for sp in splits.values() for t in sp} | {t[2] for sp in splits.values() for t in sp})
    rels = sorted({t[1] for sp in splits.values() for t in sp})
    em, rm = {e: i for i, e in enumerate(ents)}, {r: i for i, r in enumerate(rels)}
    arrs = {sp: np.asarray([(em[a], rm[b], em[c]) for a, b, c in rows], dtype=np.int32) for sp, rows in splits.items()}
    np.savez_compressed(path, num_entities=len(ents), num_relations=len(rels), **arrs)
    print("FB15k-237 (valid u test vocabulary):", len(ents), "entities", len(rels), "relations")


def train_tables(cfg, ds, seed=0):
    import torch
    from oracle.cpu_port import TorchCpuTrainer
    torch.manual_seed(seed)
    torch.set_num_threads(2)
    kind = KINDS[cfg["kind"]]
    W = 2 * cfg["d"] if kind == ko.COMPLEX else cfg["d"]
    rng = np.random.default_rng(seed)
    E, Rm = ko.glorot_uniform(rng, ds["N"], W), ko.glorot_uniform(rng, ds["R"], W)
    tr = TorchCpuTrainer(kind, E, Rm, gamma=cfg["gamma"], lr=cfg["lr"], margin=1.0)
    train = ds[cfg.get("train_split", "train")].astype(np.int64)
    np.random.seed(seed)
    B, K = cfg["B"], cfg["K"]
    t0 = time.time()
    for ep in range(cfg["epochs"]):
        perm = rng.permutation(len(train))
        losses = [tr.step(train[perm[i:i + B]], K) for i in range(0, len(train), B)]
        if ep % 10 == 0 or ep == cfg["epochs"] - 1:
            print(f"  epoch {ep}: loss {np.mean(losses):.4f} ({time.time() - t0:.0f}s)", flush=True)
    return tr.E.weight.detach().numpy().copy(), tr.R.weight.detach().numpy().copy()


def make_predict(kind, E, Rm, gamma, seqk=True):
    def predict(t):
        t = np.asarray(t, np.int64).reshape(-1, 3)
        if not seqk:
            return ko.score_triples(kind, E, Rm, t, gamma, np.float32, hole_fft=False)
        # gen_tail_data keeps (s,p) and varies o; gen_head_data keeps (p,o) and varies s (utils.py:165-197)
        tail = len(t) == 1 or bool(np.all(t[:, 0] == t[0, 0]))
        side = "tail" if tail else "head"
        q = ko.query_vectors(kind, E, Rm, t[:1], side)
        return ko.seqk_scores(kind, q, E[t[:, 2] if tail else t[:, 0]], side, gamma)[0]
    return predict


def reference_ranks(ref, predict, test_l, N, filt_l):
    """The reference's validate, unmodified, with a recording rankdata: -> (metrics dict, tail ranks, head ranks)."""
    from scipy.stats import rankdata as real_rankdata
    rec = []

    def rankdata(a, method="average", **kw):
        r = real_rankdata(a, method, **kw)
        rec.append(int(len(a) - r[0] + 1))      # real_rank, utils.py:215
        return r
    g = ref["validate"].__globals__
    old = g["rankdata"]
    g["rankdata"] = rankdata
    try:
        metrics = ref["validate"](ref_ast.PredictModel(predict), test_l, N, 32, filt_l)
    finally:
        g["rankdata"] = old
    n = len(test_l)
    assert len(rec) == 2 * n
    return metrics, np.asarray(rec[:n], np.int64), np.asarray(rec[n:], np.int64)


def run(name, ref):
    cfg = ff.CONFIGS[name]
    ds = ff.load_dataset(cfg["dataset"])
    kind = KINDS[cfg["kind"]]
    print(f"[{name}] N={ds['N']} R={ds['R']} test={len(ds['test'])}", flush=True)
    E, Rm = train_tables(cfg, ds)
    store = {}
    meta = dict(cfg, N=ds["N"], R=ds["R"], ntest=int(len(ds["test"])), trained_by="oracle/cpu_port.TorchCpuTrainer (PyTorch CPU)")
    if cfg.get("quant"):
        meta["jitter_seed"] = 12345
        store["E_q"] = ff.quantize_table(E, cfg["quant"])
        E = ff.expand_table(store["E_q"], cfg["quant"], meta["jitter_seed"])
        store["R"] = Rm.astype(np.float32)
    else:
        store["E"], store["R"] = E.astype(np.float32), Rm.astype(np.float32)
    Rm = store["R"]
    import hashlib
    meta["E_sha256"] = hashlib.sha256(np.ascontiguousarray(E.astype(np.float32)).tobytes()).hexdigest()   # guards expand_table's reproducibility
    filt = np.concatenate([ds[sp] for sp in cfg.get("filter_splits", ("train", "valid"))], axis=0)
    test_l = [tuple(int(x) for x in r) for r in ds["test"]]
    filt_l = [tuple(int(x) for x in r) for r in filt]
    t0 = time.time()
    metrics, tr, hr = reference_ranks(ref, make_predict(kind, E, Rm, cfg["gamma"]), test_l, ds["N"], filt_l)
    print(f"  reference validate: {time.time() - t0:.0f}s  mrr={metrics['mrr']:.4f} hits@10={metrics['hits@10']:.4f}", flush=True)
    # the vectorised restatement of validate must reproduce the reference's ranks on the same predict arithmetic
    fh, ft = ko.build_filters(filt_l)
    sample = np.arange(len(ds["test"]))[:: max(1, len(ds["test"]) // 200)]
    part = ds["test"][sample]
    for side, got, f in (("tail", tr, ft), ("head", hr, fh)):
        sc = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, part, side), E, side, cfg["gamma"])
        assert np.array_equal(ko.ranks_from_scores(sc, part, side, f), got[sample]), (name, side)
    store["tail_ranks"], store["head_ranks"] = tr.astype(np.int32), hr.astype(np.int32)
    meta["metrics_ref"] = {k: float(v) for k, v in metrics.items()}           # as returned here (SciPy 1.18: fp32 means)
    meta["metrics_f64"] = {k: float(v) for k, v in ko.metrics_from_ranks(tr, hr).items()}   # the pinned scipy==1.10.0 arithmetic
    for k in metrics:
        assert abs(metrics[k] - meta["metrics_f64"][k]) <= 1e-5 * max(1.0, abs(metrics[k])), k
    if not cfg.get("quant"):
        m2, tr2, hr2 = reference_ranks(ref, make_predict(kind, E, Rm, cfg["gamma"], seqk=False), test_l, ds["N"], filt_l)
        store["alt_tail_ranks"], store["alt_head_ranks"] = tr2.astype(np.int32), hr2.astype(np.int32)
        meta["alt_metrics_f64"] = {k: float(v) for k, v in ko.metrics_from_ranks(tr2, hr2).items()}
        print(f"  other summation order: {int(np.sum(tr != tr2) + np.sum(hr != hr2))} of {2 * len(tr)} ranks differ", flush=True)
    os.makedirs(ff.FULL, exist_ok=True)
    np.savez_compressed(os.path.join(ff.FULL, name + ".npz"), **store)
    with open(os.path.join(ff.FULL, name + ".json"), "w") as fh_:
        json.dump(meta, fh_, indent=1, sort_keys=True)
    print(f"[{name}] done: {os.path.getsize(os.path.join(ff.FULL, name + '.npz')) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    assert ref_ast.available(), "run in the build container (needs /root/reference)"
    ref = ref_ast.load()
    ensure_fb15k(ref)
    for nm in (sys.argv[1:] or list(ff.CONFIGS)):
        run(nm, ref)
