"""Loader for the BASELINE-config golden fixtures (tests/golden/full/), shared by the generator and the tests.

A fixture is one <name>.npz (+ <name>.json): the TRAINED tables of a model (small tables as fp32; large ones as an int8
grid plus a seeded sub-grid jitter, see `expand_table`), and what the reference's own `validate`
(/root/reference/KGEkeras/utils.py:199-264, executed unmodified through oracle/ref_ast.py) returned for them on the
dataset's FULL test split: per-query filtered ranks of both sides and the 15-key metrics dict.
"""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FULL = os.path.join(HERE, "full")

# name -> (dataset, model, stored width d, gamma, hyper-parameters of the CPU training run that produced the tables)
CONFIGS = {
    # BASELINE.json configs[0]: DistMult dim=100 on bundled Kinship, 10 negatives, pairwise hinge
    "kinship_DistMult_d100": dict(dataset="kinship", kind="DistMult", d=100, gamma=12.0, epochs=200, B=512, K=10, lr=0.01),
    "kinship_ComplEx_d100": dict(dataset="kinship", kind="ComplEx", d=100, gamma=12.0, epochs=200, B=512, K=10, lr=0.01),
    "UMLS_TransE_L2_d100": dict(dataset="UMLS", kind="TransE_L2", d=100, gamma=12.0, epochs=200, B=512, K=10, lr=0.01),
    "UMLS_HolE_d64": dict(dataset="UMLS", kind="HolE", d=64, gamma=12.0, epochs=150, B=512, K=10, lr=0.01),
    # gamma <= 0: the score is +||s+p-o|| and validate does NOT flip it (models.py:189-190)
    "nations_TransE_L1_gamma0_d50": dict(dataset="nations", kind="TransE_L1", d=50, gamma=0.0, epochs=100, B=256, K=10, lr=0.01),
    "nations_TransE_L2_gammaNeg_d50": dict(dataset="nations", kind="TransE_L2", d=50, gamma=-1.0, epochs=100, B=256, K=10, lr=0.01),
    "nations_DistMult_d100": dict(dataset="nations", kind="DistMult", d=100, gamma=12.0, epochs=200, B=256, K=10, lr=0.01),
    # BASELINE.json configs[1]: TransE L1 dim=200 on bundled WN18RR (all 3 134 test triples, 40 943 candidates each)
    "WN18RR_TransE_L1_d200": dict(dataset="WN18RR", kind="TransE_L1", d=200, gamma=12.0, epochs=30, B=2048, K=10, lr=0.01,
                                  quant=32),
    # BASELINE.json configs[2]: ComplEx dim=200 on bundled FB15k-237.  train.txt is missing from the mount
    # (.MISSING_LARGE_BLOBS), so the vocabulary is valid u test, the table is trained on valid.txt ("surrogate train")
    # and test.txt is ranked filtered by valid.txt.
    "FB15k-237_ComplEx_d200": dict(dataset="FB15k-237", kind="ComplEx", d=200, gamma=12.0, epochs=40, B=2048, K=10, lr=0.01,
                                   quant=32, train_split="valid", filter_splits=("valid",)),
}


def load_dataset(name):
    d = np.load(os.path.join(HERE, "datasets", name + ".npz"))
    out = dict(N=int(d["num_entities"]), R=int(d["num_relations"]))
    for sp in ("train", "valid", "test"):
        if sp in d.files:
            out[sp] = d[sp]
    return out


def expand_table(q, scale, seed):
    """int8 grid -> fp32 table: (q + u - 0.5) / scale with u from the frozen legacy MT19937 stream (bit-stable across NumPy
    versions).  The jitter stays inside the grid cell, so the trained structure survives, while the values get full fp32
    mantissas and the scores realistic near-ties instead of the grid's exact ones."""
    u = np.random.RandomState(int(seed)).random_sample(q.shape)
    return ((q.astype(np.float64) + (u - 0.5)) / float(scale)).astype(np.float32)


def quantize_table(T, scale):
    return np.clip(np.rint(np.asarray(T, np.float64) * scale), -127, 127).astype(np.int8)


def available(name):
    return os.path.isfile(os.path.join(FULL, name + ".npz")) and os.path.isfile(os.path.join(FULL, name + ".json"))


def load(name):
    """-> dict(E, R (fp32), tail_ranks, head_ranks (int64, from the reference), meta (json dict))."""
    with open(os.path.join(FULL, name + ".json")) as fh:
        meta = json.load(fh)
    z = np.load(os.path.join(FULL, name + ".npz"))
    if "E_q" in z.files:
        E = expand_table(z["E_q"], meta["quant"], meta["jitter_seed"])
        R = z["R"]
    else:
        E, R = z["E"], z["R"]
    return dict(E=E, R=R, tail_ranks=z["tail_ranks"].astype(np.int64), head_ranks=z["head_ranks"].astype(np.int64),
                alt_tail_ranks=z["alt_tail_ranks"].astype(np.int64) if "alt_tail_ranks" in z.files else None,
                alt_head_ranks=z["alt_head_ranks"].astype(np.int64) if "alt_head_ranks" in z.files else None, meta=meta)
