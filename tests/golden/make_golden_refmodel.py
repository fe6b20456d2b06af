"""The reference pipeline END TO END on the small BASELINE-config fixtures (run in the build container):

    python tests/golden/make_golden_refmodel.py     ->  tests/golden/full_refmodel.npz

For every fixture of tests/golden/full/ whose tables are stored as fp32 (kinship, UMLS, nations: C1 and the gamma <= 0 cases), the
reference's OWN model class (models.py, loaded unmodified through oracle/ref_tf_exec.py, tables injected with `init_entities` /
`init_relations` as models.py:98-112 provides) is handed to the reference's OWN `validate` (utils.py:199-264, AST-loaded through
oracle/ref_ast.py): `model.predict` scores every `(C+1, 3)` batch in Keras' 32-row batches with the reference's `func`.  Nothing of
this repository computes a score on that path.  The per-query ranks are stored next to the fixtures' golden ranks (which used the
rank contract's sequential-k fp32 arithmetic as `predict`): the two differ only where two fp32 summation orders of the same score
fall on different sides of a near-tie (tests/test_cpu_full_goldens.py bounds and explains every difference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_ast, ref_tf_exec as rx  # noqa: E402
import full_fixtures as ff  # noqa: E402
from make_golden_full import reference_ranks  # noqa: E402

OUT = os.path.join(HERE, "full_refmodel.npz")
CLASS = {"DistMult": ("DistMult", {}), "ComplEx": ("ComplEx", {}), "HolE": ("HolE", {}),
         "TransE_L1": ("TransE", {"norm": 1}), "TransE_L2": ("TransE", {"norm": 2})}


def small_configs():
    return [n for n in sorted(ff.CONFIGS) if not ff.CONFIGS[n].get("quant")]


def reference_model(rm, cfg, E, R):
    cls_name, kw = CLASS[cfg["kind"]]
    kw = dict(kw)
    if cls_name == "TransE":
        kw["gamma"] = cfg["gamma"]
    return getattr(rm, cls_name)(e_dim=cfg["d"], r_dim=cfg["d"], num_entities=E.shape[0], num_relations=R.shape[0],
                                 init_entities=E, init_relations=R, dp=0.0, **kw)


def run(name, rm, ref):
    fx = ff.load(name)
    cfg, meta = ff.CONFIGS[name], fx["meta"]
    ds = ff.load_dataset(meta["dataset"])
    model = reference_model(rm, cfg, fx["E"], fx["R"])
    filt = np.concatenate([ds[sp] for sp in meta.get("filter_splits", ("train", "valid"))], axis=0)
    test_l = [tuple(int(x) for x in r) for r in ds["test"]]
    filt_l = [tuple(int(x) for x in r) for r in filt]
    metrics, tr, hr = reference_ranks(ref, model.predict, test_l, ds["N"], filt_l)
    return metrics, tr, hr


def generate(names=None):
    assert rx.available() and ref_ast.available()
    rm, _, _ = rx.load(32)
    ref = ref_ast.load()
    out = {}
    for name in names or small_configs():
        metrics, tr, hr = run(name, rm, ref)
        fx = ff.load(name)
        nd = int(np.sum(tr != fx["tail_ranks"]) + np.sum(hr != fx["head_ranks"]))
        print(f"[{name}] reference model + reference validate: mrr={metrics['mrr']:.6f}; {nd} of {2 * len(tr)} ranks differ from the "
              f"sequential-k goldens (mrr {fx['meta']['metrics_f64']['mrr']:.6f})", flush=True)
        out[name + "/tail_ranks"], out[name + "/head_ranks"] = tr.astype(np.int32), hr.astype(np.int32)
        out[name + "/metrics"] = np.array([float(metrics[k]) for k in sorted(metrics)], np.float64)
        out[name + "/metric_keys"] = np.array(sorted(metrics))
    return out


if __name__ == "__main__":
    # no arguments: the small configurations -> full_refmodel.npz;  names: those configurations (the quantised large ones take
    # 15-40 minutes each: 8-19 million 32-row predict calls) -> full_refmodel_<name>.npz
    if len(sys.argv) == 1:
        data = generate()
        np.savez_compressed(OUT, **data)
        print("wrote", OUT, os.path.getsize(OUT), "bytes")
    else:
        for nm in sys.argv[1:]:
            data = generate([nm])
            path = os.path.join(HERE, "full_refmodel_" + nm + ".npz")
            np.savez_compressed(path, **data)
            print("wrote", path, os.path.getsize(path), "bytes")
