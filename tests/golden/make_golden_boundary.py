"""Freeze the reference's public Python surface on the hot path (SURVEY 8b "upper boundary") as data:

    python tests/golden/make_golden_boundary.py     ->  tests/golden/boundary_ref.json

models.py / utils.py are loaded unmodified through oracle/ref_tf_exec.py; recorded are (1) the signature (parameter names,
order, defaults) of every constructor and function the drop-in keeps, (2) what each model class exposes after construction
with the harness' keywords (tests.py:65-71): scalar attributes, stored table shapes, unknown keywords becoming attributes
(models.py:129), and (3) the key set of `loss_function_lookup`.  tests/test_cpu_boundary_vs_reference.py holds the repo's
KGEkeras package to this file."""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_tf_exec as rx  # noqa: E402

OUT = os.path.join(HERE, "boundary_ref.json")
MODEL_CLASSES = ["EmbeddingModel", "DistMult", "TransE", "ComplEx", "HolE", "HAKE", "ModE", "RotatE", "pRotatE"]
UTIL_FUNCS = ["load_kg", "generate_negative", "oversample_data", "pad", "mrr", "hits", "gen_tail_data", "gen_head_data", "validate",
              "pointwize_hinge", "pointwize_logistic", "pointwize_square_loss", "pointwize_cross_entropy", "pairwize_hinge",
              "pairwize_logistic", "pairwize_square_loss", "loss_function_lookup", "isint"]
CTOR_KW = dict(e_dim=8, r_dim=8, num_entities=10, num_relations=3, custom_keyword=5)
EXTRA_KW = {"TransE": [dict(norm=2, gamma=0)], "HAKE": [dict(gamma=6, epsilon=1, phase_weight=0.25, mod_weight=2)],
            "pRotatE": [dict(modulus=0.25)], "RotatE": [dict(gamma=24)]}


def sig(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        if p.name == "self":
            continue
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([p.name, str(p.kind), d])
    return out


def scalars(obj):
    keep = {}
    for k, v in vars(obj).items():
        if isinstance(v, (bool, int, float, str)) or v is None:
            keep[k] = v
    return keep


def generate():
    rm, ru, _ = rx.load(32)
    out = {"model_signatures": {}, "util_signatures": {}, "instances": {}, "func_signatures": {}}
    for c in MODEL_CLASSES:
        out["model_signatures"][c] = sig(getattr(rm, c).__init__)
        if hasattr(getattr(rm, c), "func"):
            out["func_signatures"][c] = sig(getattr(rm, c).func)
    out["call_signature"] = sig(rm.EmbeddingModel.call)
    out["l3_reg_signature"] = sig(rm.l3_reg)
    for f in UTIL_FUNCS:
        out["util_signatures"][f] = sig(getattr(ru, f))
    out["callback_signature"] = sig(ru.KGEValidateCallback.__init__)
    out["callback_methods"] = sorted(n for n in vars(ru.KGEValidateCallback) if not n.startswith("__"))
    candidates = ("pointwize_hinge", "pointwize_logistic", "pointwize_cross_entropy", "pointwize_square_loss", "pairwize_hinge",
                  "pairwize_logistic", "pairwize_square_loss", "pointwise_hinge", "pairwise_hinge", "hinge", "logistic")
    out["loss_names"] = sorted(k for k in candidates if _has(ru, k))       # the spelling "wize" is the reference's
    out["loss_names_rejected"] = sorted(k for k in candidates if not _has(ru, k))
    out["EPSILON"] = ru.EPSILON
    for c in MODEL_CLASSES[1:]:
        for extra in [{}] + EXTRA_KW.get(c, []):
            kw = dict(CTOR_KW, **extra)
            key = c + ("" if not extra else "|" + json.dumps(extra, sort_keys=True))
            try:
                m = getattr(rm, c)(**kw)
            except Exception as e:          # ModE: models.py:415 reads self.gamma before it exists
                out["instances"][key] = {"raises": type(e).__name__, "kwargs": kw}
                continue
            out["instances"][key] = {"kwargs": kw, "attrs": scalars(m),
                                     "entity_table": list(m.entity_embedding.embeddings.shape),
                                     "relation_table": list(m.relational_embedding.embeddings.shape),
                                     "dropout_rate": m.dp.rate}
    return out


def _has(ru, name):
    try:
        ru.loss_function_lookup(name)
        return True
    except KeyError:
        return False


if __name__ == "__main__":
    data = generate()
    with open(OUT, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
