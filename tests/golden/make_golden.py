"""Generate the committed golden fixtures (run in the build container, where /root/reference exists).

    python tests/golden/make_golden.py

Outputs (all small, committed):
  datasets/<name>.npz     int32 triples of the COMPLETE bundled datasets (kinship, UMLS, nations, WN18RR) under a
                          deterministic sorted id mapping (the reference maps by set-iteration order, tests.py:102-111)
  negatives_ref.npz       inputs/outputs of the reference's own generate_negative + oversample_data
                          (utils.py:114-147, executed via oracle/ref_ast.py) under fixed np.random / random seeds
  validate_ref.npz/.json  per-query filtered ranks and the 15-key metrics dict produced by the reference's own
                          validate/gen_tail_data/gen_head_data (utils.py:165-264) for seeded tables
  score_kat.npz           known-answer vectors of the restated score / loss / gradient / optimizer functions (fp64 truth
                          and fp32 op-by-op), guarding the oracle against drift
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import kge_oracle as ko  # noqa: E402
from oracle import ref_ast  # noqa: E402

REF_DATA = "/root/reference/data"
DATASETS = ("kinship", "UMLS", "nations", "WN18RR")
KINDS = (("DistMult", ko.DISTMULT), ("TransE_L1", ko.TRANSE_L1), ("TransE_L2", ko.TRANSE_L2),
         ("ComplEx", ko.COMPLEX), ("HolE", ko.HOLE))


def make_datasets(ref):
    os.makedirs(os.path.join(HERE, "datasets"), exist_ok=True)
    out = {}
    for name in DATASETS:
        splits = {sp: ref["load_kg"](os.path.join(REF_DATA, name, sp + ".txt")) for sp in ("train", "valid", "test")}
        ents = sorted({t[0] for sp in splits.values() for t in sp} | {t[2] for sp in splits.values() for t in sp})
        rels = sorted({t[1] for sp in splits.values() for t in sp})
        em = {e: i for i, e in enumerate(ents)}
        rm = {r: i for i, r in enumerate(rels)}
        arrs = {sp: np.asarray([(em[a], rm[b], em[c]) for a, b, c in rows], dtype=np.int32) for sp, rows in splits.items()}
        np.savez_compressed(os.path.join(HERE, "datasets", name + ".npz"), num_entities=len(ents),
                            num_relations=len(rels), **arrs)
        out[name] = dict(N=len(ents), R=len(rels), **arrs)
        print(name, len(ents), len(rels), {k: v.shape for k, v in arrs.items()})
    return out


def make_negatives(ref):
    cases = {}
    for ci, (B, N, K, seed) in enumerate([(8, 14, 2, 0), (64, 104, 10, 1), (2048, 40943, 10, 2), (33, 1000, 3, 3)]):
        rng = np.random.default_rng(100 + ci)
        kg = np.stack([rng.integers(0, N, B), rng.integers(0, 7, B), rng.integers(0, N, B)], axis=1).astype(np.int64)
        np.random.seed(seed)
        random.seed(seed)
        neg = ref["generate_negative"](kg, N=N, negative=K)
        X = ref["oversample_data"](kgs=[kg, neg])
        cases[f"c{ci}_kg"] = kg
        cases[f"c{ci}_params"] = np.asarray([B, N, K, seed], dtype=np.int64)
        cases[f"c{ci}_neg"] = neg
        cases[f"c{ci}_pos_over"] = X[0]
        cases[f"c{ci}_neg_over"] = X[1]
        assert neg.shape == (K * B, 3, 1) and X[0].shape == (K * B, 3) and X[1].shape == (K * B, 3)
    np.savez_compressed(os.path.join(HERE, "negatives_ref.npz"), **cases)
    print("negatives_ref:", len(cases), "arrays")


def tables_for(kind, N, R, d, seed, dyadic):
    rng = np.random.default_rng(seed)
    W = 2 * d if kind == ko.COMPLEX else d
    if dyadic:
        return ko.dyadic_table(rng, N, W), ko.dyadic_table(rng, R, W)
    return ko.glorot_uniform(rng, N, W), ko.glorot_uniform(rng, R, W)


def make_validate(ref, data):
    arrays, meta = {}, {}
    plan = [("kinship", None, 8), ("UMLS", 160, 8), ("nations", None, 4), ("WN18RR", 24, 4)]
    for name, ntest, d in plan:
        ds = data[name]
        test = ds["test"] if ntest is None else ds["test"][:ntest]
        filt = np.concatenate([ds["train"], ds["valid"]], axis=0)
        test_l = [tuple(int(x) for x in r) for r in test]
        filt_l = [tuple(int(x) for x in r) for r in filt]
        for kname, kind in KINDS:
            for dyadic in (True, False):
                if name == "WN18RR" and not dyadic:
                    continue
                seed = 7
                E, Rm = tables_for(kind, ds["N"], ds["R"], d, seed, dyadic)
                model = ref_ast.PredictModel(lambda t: ko.score_triples(kind, E, Rm, t, 12.0, np.float32, hole_fft=False))
                metrics = ref["validate"](model, test_l, ds["N"], 32, filt_l)
                # per-query ranks straight from the reference: validate on one triple at a time (mr == rank)
                tr, hr = [], []
                per_query = test_l if name != "WN18RR" else test_l[:8]
                if name in ("kinship", "UMLS") and kind not in (ko.DISTMULT, ko.TRANSE_L1):
                    per_query = test_l[:64]
                for trip in per_query:
                    m1 = ref["validate"](model, [trip], ds["N"], 32, filt_l)
                    tr.append(int(m1["tail_mr"]))
                    hr.append(int(m1["head_mr"]))
                # the vectorised restatement must agree with the reference on everything
                om, otr, ohr = ko.validate(kind, E, Rm, test, filt, 12.0, np.float32)
                assert otr[:len(tr)].tolist() == tr and ohr[:len(hr)].tolist() == hr, (name, kname)
                # SciPy 1.18 (this image) returns float32 ranks for float32 scores, so the reference's np.mean runs in
                # fp32 here; the pinned scipy==1.10.0 (requirements.txt:1) returns int64 -> float64 means.  Ranks are
                # exact either way; the metrics are compared at fp32 resolution and the float64 values are stored too.
                for k in metrics:
                    assert abs(metrics[k] - om[k]) <= 2e-6 * max(1.0, abs(metrics[k])), (name, kname, k)
                key = f"{name}/{kname}/{'dyadic' if dyadic else 'glorot'}"
                arrays[key + "/tail_ranks"] = otr.astype(np.int32)
                arrays[key + "/head_ranks"] = ohr.astype(np.int32)
                arrays[key + "/ref_tail_ranks"] = np.asarray(tr, np.int32)
                arrays[key + "/ref_head_ranks"] = np.asarray(hr, np.int32)
                meta[key] = dict(dataset=name, kind=kname, dyadic=dyadic, d=d, seed=seed, ntest=len(test_l),
                                 n_ref_per_query=len(tr), gamma=12.0, metrics={k: float(v) for k, v in metrics.items()},
                                 metrics_f64={k: float(v) for k, v in om.items()})
                print(key, "mrr=%.4f" % metrics["mrr"], "hits@10=%.4f" % metrics["hits@10"])
    np.savez_compressed(os.path.join(HERE, "validate_ref.npz"), **arrays)
    with open(os.path.join(HERE, "validate_ref.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


def make_kat():
    out = {}
    rng = np.random.default_rng(42)
    N, R, B, K = 50, 7, 12, 3
    for kname, kind in KINDS:
        d = 10
        W = 2 * d if kind == ko.COMPLEX else d
        E = ko.glorot_uniform(rng, N, W) * 3
        Rm = ko.glorot_uniform(rng, R, W) * 3
        pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
        neg = np.stack([rng.integers(0, N, B * K), np.repeat(pos[:, 1], K), rng.integers(0, N, B * K)], 1).astype(np.int32)
        out[f"{kname}/E"], out[f"{kname}/R"], out[f"{kname}/pos"], out[f"{kname}/neg"] = E, Rm, pos, neg
        out[f"{kname}/score64"] = ko.score_triples(kind, E, Rm, pos, 12.0, np.float64, hole_fft=False)
        out[f"{kname}/score32"] = ko.score_triples(kind, E, Rm, pos, 12.0, np.float32)
        for lname, lk in ko.LOSS_NAMES.items():
            for pm in (ko.PAIRED, ko.ALLPAIRS):
                st = ko.train_step(kind, lk, pm, E, Rm, pos, neg, K, 1.0, 12.0, ko.OPT_LAZY_ADAM,
                                   [np.zeros_like(E), np.zeros_like(E)], [np.zeros_like(Rm), np.zeros_like(Rm)],
                                   1e-3, 1)
                tag = f"{kname}/{lname}/{pm}"
                out[tag + "/loss"] = np.float64(st["loss"])
                out[tag + "/gE"] = st["gE"]
                out[tag + "/gR"] = st["gR"]
                out[tag + "/E1"] = st["E"]
    np.savez_compressed(os.path.join(HERE, "score_kat.npz"), **out)
    print("score_kat:", len(out), "arrays")


if __name__ == "__main__":
    assert ref_ast.available(), "run in the build container (needs /root/reference)"
    ref = ref_ast.load()
    data = make_datasets(ref)
    make_negatives(ref)
    make_validate(ref, data)
    make_kat()
