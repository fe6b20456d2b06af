"""Dataset ingestion for the hot path: what the reference's harness does inline in tests.py:98-117
(load_kg on three TSV splits, entity / relation id mappings, id triples) and tests.py:29-61 (DataGenerator), as reusable
host code.

Differences from the reference, all deliberate:
  * the id mapping is SORTED (the reference enumerates Python sets, so its ids change from run to run with the string
    hash seed, tests.py:102-111); sorted ids make runs, caches and the committed fixtures reproducible;
  * the parsed splits are cached as one .npz next to nothing the user owns (a cache directory), keyed by the files'
    size and mtime, so re-runs skip the TSV parse;
  * DataGenerator can hand out device-resident batches whose negatives and oversampling are drawn on the GPU
    (Engine.device_batch: kgeb_negatives + kgeb_oversample), besides the reference's host path (generate_negative +
    oversample_data, same global RNG consumption as the reference).
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from .utils import generate_negative, load_kg, oversample_data

SPLITS = ("train", "valid", "test")


def build_mappings(*splits):
    """Sorted entity and relation vocabularies over every given split (lists of [s, p, o] string rows)."""
    ents, rels = set(), set()
    for rows in splits:
        for row in rows:
            if len(row) != 3:
                raise ValueError(f"expected 3 tab/space separated fields per line, got {row!r}")
            ents.add(row[0])
            ents.add(row[2])
            rels.add(row[1])
    ents, rels = sorted(ents), sorted(rels)
    return {e: i for i, e in enumerate(ents)}, {r: i for i, r in enumerate(rels)}


def map_triples(rows, entity_mapping, relation_mapping):
    """(n,3) int32 id triples; KeyError (as in the reference's list comprehension, tests.py:114-116) on an unknown name."""
    out = np.empty((len(rows), 3), dtype=np.int32)
    for i, (a, b, c) in enumerate(rows):
        out[i, 0], out[i, 1], out[i, 2] = entity_mapping[a], relation_mapping[b], entity_mapping[c]
    return out


def _cache_key(paths):
    h = hashlib.sha1()
    for p in paths:
        st = os.stat(p)
        h.update(f"{os.path.abspath(p)}:{st.st_size}:{st.st_mtime_ns};".encode())
    return h.hexdigest()[:16]


def load_dataset(directory, cache_dir=None, splits=SPLITS):
    """Reads <directory>/{train,valid,test}.txt -> dict(num_entities, num_relations, train, valid, test (int32 (n,3)),
    entities, relations (sorted name lists)).  Missing split files raise FileNotFoundError (three of the reference's bundled
    datasets ship without train.txt).  cache_dir (default: $KGEB_CACHE_DIR, no caching when unset) holds the
    parsed arrays keyed by the source files' size and mtime."""
    paths = [os.path.join(directory, sp + ".txt") for sp in splits]
    for p in paths:
        if not os.path.isfile(p):
            raise FileNotFoundError(p)
    cache_dir = cache_dir if cache_dir is not None else os.environ.get("KGEB_CACHE_DIR")
    cpath = None
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        cpath = os.path.join(cache_dir, f"{os.path.basename(os.path.normpath(directory))}-{_cache_key(paths)}.npz")
        if os.path.isfile(cpath):
            z = np.load(cpath, allow_pickle=False)
            out = {sp: z[sp] for sp in splits}
            out.update(num_entities=int(z["num_entities"]), num_relations=int(z["num_relations"]),
                       entities=[str(x) for x in z["entities"]], relations=[str(x) for x in z["relations"]])
            return out
    raw = {sp: load_kg(p) for sp, p in zip(splits, paths)}
    em, rm = build_mappings(*raw.values())
    out = {sp: map_triples(rows, em, rm) for sp, rows in raw.items()}
    ents, rels = sorted(em, key=em.get), sorted(rm, key=rm.get)
    out.update(num_entities=len(em), num_relations=len(rm), entities=ents, relations=rels)
    if cpath:
        tmp = cpath + f".{os.getpid()}.tmp.npz"
        np.savez_compressed(tmp, num_entities=len(em), num_relations=len(rm), entities=np.asarray(ents), relations=np.asarray(rels),
                            **{sp: out[sp] for sp in splits})
        os.replace(tmp, cpath)
    return out


class DataGenerator:
    """The harness' keras.utils.Sequence (tests.py:29-61): shuffled batches of `batch_size` positives, each turned into
    ([pos (m,3), neg (m,3)], None) with m = ns * batch_size by generate_negative + oversample_data.

    device=None reproduces the reference on the host (same global np.random / random consumption).  With an Engine
    (`engine=model._engine_()`) the negatives and the oversampling are drawn on the GPU by counter-based Philox
    (seed, step) and the batch never leaves the device -- the reference's host pipeline caps any GPU step at ~1 M triples/s.
    Both paths corrupt with ids below `num_e`, the number of distinct entities of the split (tests.py:34)."""

    def __init__(self, kg, ns=10, batch_size=32, shuffle=True, engine=None, seed=0):
        self.batch_size = min(batch_size, len(kg))
        self.kg = np.asarray(kg)
        self.ns = ns
        # tests.py:34: the number of DISTINCT entities of this split (not max id + 1): negatives are drawn below it
        self.num_e = int(len(np.unique(self.kg[:, [0, 2]]))) if len(self.kg) else 0
        self.shuffle = shuffle
        self.engine = engine
        self.seed = int(seed)
        self.indices = np.arange(len(self.kg), dtype=np.int64)
        self._step = 0
        self._kg_dev = self._index_dev = None
        if engine is not None and len(self.kg):
            from ._engine import as_idx
            self._kg_dev = as_idx(self.kg, engine.device)      # the split lives on the GPU: a batch is a device gather
        self.on_epoch_end()

    def __len__(self):
        return len(self.kg) // self.batch_size

    def __getitem__(self, index):
        lo, hi = index * self.batch_size, (index + 1) * self.batch_size
        if self._kg_dev is not None:
            self._step += 1
            return self.engine.device_batch(self._kg_dev.index_select(0, self._index_dev[lo:hi]), self.ns, self.seed, self._step,
                                            num_entities=self.num_e), None
        return self._get_data(self.indices[self.index[lo:hi]])

    def on_epoch_end(self):
        self.index = np.arange(len(self.indices))
        if self.shuffle:
            np.random.shuffle(self.index)
        if self._kg_dev is not None:
            import torch
            self._index_dev = torch.from_numpy(self.index).to(self._kg_dev.device)   # one upload per epoch

    def _get_data(self, batch):
        tmp_kg = self.kg[batch]
        if self.engine is not None:
            self._step += 1
            return self.engine.device_batch(tmp_kg, self.ns, self.seed, self._step, num_entities=self.num_e), None
        negative_kg = generate_negative(tmp_kg, N=self.num_e, negative=self.ns)
        X = oversample_data(kgs=[tmp_kg, negative_kg])
        return X, None
