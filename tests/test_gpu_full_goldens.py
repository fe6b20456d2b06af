"""Parity on the configurations BASELINE.json names, against the reference's OWN validate (utils.py:199-264).

tests/golden/full/<config>.npz holds TRAINED tables and the per-query filtered ranks / 15-key metrics dict the reference's
validate produced for them on the dataset's FULL test split (tests/golden/make_golden_full.py: the reference code is executed
unmodified, its predict() computes the rank contract's sequential-k fp32 scores).  The GPU path must reproduce
  * every rank EXACTLY (integer equality, ties included) in the default exact tensor-core mode and in the CUDA-core mode,
  * MRR / MR / Hits@1/3/10 to 1e-12 of the float64 evaluation of the reference's formulas on its ranks,
  * in the approximate bf16 mode: MRR and Hits@k within 2e-3 absolute (stated tolerance of the single-plane mode).
Covers kinship (DistMult d=100 = configs[0], ComplEx), UMLS (TransE-L2, HolE), nations (TransE with gamma <= 0:
models.py:189-190; DistMult), WN18RR (TransE-L1 d=200 = configs[1], all 3 134 test triples x 40 943 candidates) and
FB15k-237 (ComplEx d=200 = configs[2]; test.txt filtered by valid.txt, vocabulary valid u test: train.txt is not in the mount)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import full_fixtures as ff  # noqa: E402

pytestmark = pytest.mark.gpu


def build_model(meta, E, Rm):
    from KGEkeras import models as km
    kind, d = meta["kind"], meta["d"]
    args = dict(e_dim=d, r_dim=d, num_entities=meta["N"], num_relations=meta["R"], init_entities=E, init_relations=Rm, dp=0.0)
    if kind == "DistMult":
        return km.DistMult(**args)
    if kind == "ComplEx":
        return km.ComplEx(**args)
    if kind == "HolE":
        m = km.HolE(**args)
        return m
    return km.TransE(norm=1 if kind == "TransE_L1" else 2, gamma=meta["gamma"], **args)


@pytest.mark.parametrize("name", sorted(ff.CONFIGS))
def test_ranks_and_metrics_equal_reference_validate(name):
    from KGEkeras import utils as ku
    assert ff.available(name), f"fixture tests/golden/full/{name}.npz missing (run tests/golden/make_golden_full.py)"
    fx = ff.load(name)
    meta = fx["meta"]
    ds = ff.load_dataset(meta["dataset"])
    test = ds["test"]
    filt = np.concatenate([ds[sp] for sp in meta.get("filter_splits", ("train", "valid"))], axis=0)
    model = build_model(meta, fx["E"], fx["R"])
    modes = ["exact", "cuda"] if model._engine_().tc_supported() else ["cuda"]
    for mode in modes:
        model.rank_mode = mode
        tr, hr = ku.filtered_ranks(model, test, filt)
        np.testing.assert_array_equal(tr, fx["tail_ranks"], err_msg=f"{name} tail ranks, mode {mode}")
        np.testing.assert_array_equal(hr, fx["head_ranks"], err_msg=f"{name} head ranks, mode {mode}")
        metrics = ku.validate(model, [tuple(int(x) for x in r) for r in test] if len(test) < 2000 else test, meta["N"], 32, filt)
        assert set(metrics) == set(meta["metrics_f64"])
        for k, v in meta["metrics_f64"].items():
            assert abs(metrics[k] - v) <= 1e-12 * max(1.0, abs(v)), (name, mode, k, metrics[k], v)
        for k, v in meta["metrics_ref"].items():      # what the reference returned here (SciPy 1.18: its means run in fp32)
            assert abs(metrics[k] - v) <= 1e-5 * max(1.0, abs(v)), (name, mode, k)
    if "exact" in modes:
        model.rank_mode = "bf16"
        approx = ku.validate(model, test, meta["N"], 32, filt)
        for k in ("mrr", "hits@1", "hits@3", "hits@10"):
            assert abs(approx[k] - meta["metrics_f64"][k]) <= 2e-3, (name, k, approx[k], meta["metrics_f64"][k])


def test_other_summation_order_stays_within_near_tie_noise():
    """The reference's ranks when its predict sums in NumPy's order instead of sequentially over k (same fp32 model): the two
    differ only at near-ties; the GPU path follows the sequential contract, so it inherits exactly this distance."""
    for name in sorted(ff.CONFIGS):
        if not ff.available(name):
            continue
        fx = ff.load(name)
        if fx["alt_tail_ranks"] is None:
            continue
        for a, b in ((fx["tail_ranks"], fx["alt_tail_ranks"]), (fx["head_ranks"], fx["alt_head_ranks"])):
            assert np.mean(a != b) <= 0.02 and np.max(np.abs(a - b)) <= 3, name
        for k in ("mrr", "hits@10"):
            assert abs(fx["meta"]["metrics_f64"][k] - fx["meta"]["alt_metrics_f64"][k]) <= 2e-3, (name, k)
