"""Live partial oracle: run the reference's OWN NumPy/SciPy functions (TEST INFRASTRUCTURE).

`/root/reference/KGEkeras/utils.py` cannot be imported (it imports tensorflow, rdflib and spacy at module
top level, utils.py:7-20, none installed).  Its pure NumPy/SciPy/stdlib functions are nevertheless
executable: this module parses the file, extracts the wanted `FunctionDef` nodes UNMODIFIED and execs them
with exactly the names they need in scope.  Nothing is copied into the repository; the source is read where
it lies.  Only available in the build container (`/root/reference` does not exist on the GPU box), so
callers must check `available()`; the frozen outputs live in tests/golden/.
"""
from __future__ import annotations

import ast
import os

REF_UTILS = "/root/reference/KGEkeras/utils.py"
WANTED = ("load_kg", "generate_negative", "oversample_data", "pad", "mrr", "hits",
          "gen_tail_data", "gen_head_data", "validate")


def available() -> bool:
    return os.path.isfile(REF_UTILS)


def load(quiet_tqdm: bool = True):
    """Return a dict name -> reference function object (executed from the reference source)."""
    import numpy as np
    from collections import defaultdict
    from random import choice, choices
    from scipy.stats import rankdata

    with open(REF_UTILS, "r") as fh:
        tree = ast.parse(fh.read(), filename=REF_UTILS)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    mod = ast.Module(body=picked, type_ignores=[])
    if quiet_tqdm:
        def tqdm(it, **kw):
            return it
    else:
        from tqdm import tqdm
    scope = {"np": np, "defaultdict": defaultdict, "choice": choice, "choices": choices,
             "rankdata": rankdata, "tqdm": tqdm, "EPSILON": 1e-6}
    exec(compile(mod, REF_UTILS, "exec"), scope)
    return {name: scope[name] for name in WANTED}


class PredictModel:
    """Stand-in for the Keras model `validate` drives (utils.py:211,231): only `.predict((n,3))`."""

    def __init__(self, fn):
        self.fn = fn

    def predict(self, t):
        return self.fn(t)
