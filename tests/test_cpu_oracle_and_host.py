"""CPU suite (-m "not gpu"): the oracle against the golden vectors made from the reference's own code, the host-side
mirror's data-preparation logic, the C-ABI library's symbols, and the world_size-2 (gloo) shard/merge logic.
No CUDA compute is called here."""
import ctypes
import json
import os
import random
import re

import numpy as np
import pytest

from oracle import kge_oracle as ko
from oracle import ref_ast

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")
KINDS = {"DistMult": ko.DISTMULT, "TransE_L1": ko.TRANSE_L1, "TransE_L2": ko.TRANSE_L2, "ComplEx": ko.COMPLEX, "HolE": ko.HOLE}


def load_dataset(name):
    d = np.load(os.path.join(GOLD, "datasets", name + ".npz"))
    return dict(N=int(d["num_entities"]), R=int(d["num_relations"]), train=d["train"], valid=d["valid"], test=d["test"])


def golden_tables(kind, N, R, d, seed, dyadic):
    rng = np.random.default_rng(seed)
    W = 2 * d if kind == ko.COMPLEX else d
    if dyadic:
        return ko.dyadic_table(rng, N, W), ko.dyadic_table(rng, R, W)
    return ko.glorot_uniform(rng, N, W), ko.glorot_uniform(rng, R, W)


# ---- Philox ---------------------------------------------------------------------------------------------------
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ([0xffffffff] * 4, [0xffffffff] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        r = ko.philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        assert " ".join("%08x" % x for x in r) == want


def test_philox_negatives_properties():
    pos = np.stack([np.arange(50), np.arange(50) % 7, np.arange(50)[::-1]], 1)
    for mode in range(4):
        neg = ko.philox_negatives(pos, 1000, 6, 42, 3, mode)
        assert neg.shape == (300, 3) and neg.min() >= 0 and neg[:, [0, 2]].max() < 1000
        np.testing.assert_array_equal(neg[:, 1], np.repeat(pos[:, 1], 6))
        rep = np.repeat(pos, 6, axis=0)
        if mode == ko.CORRUPT_HEAD:
            np.testing.assert_array_equal(neg[:, 2], rep[:, 2])
        if mode == ko.CORRUPT_TAIL:
            np.testing.assert_array_equal(neg[:, 0], rep[:, 0])
        if mode == ko.CORRUPT_EITHER:
            assert np.all((neg[:, 0] == rep[:, 0]) | (neg[:, 2] == rep[:, 2]))
    assert not np.array_equal(ko.philox_negatives(pos, 1000, 6, 42, 3, 0), ko.philox_negatives(pos, 1000, 6, 42, 4, 0))


# ---- oracle vs golden fixtures made by the reference's own code --------------------------------------------------
def test_oracle_negatives_match_reference_golden():
    g = np.load(os.path.join(GOLD, "negatives_ref.npz"))
    for ci in range(4):
        B, N, K, seed = (int(x) for x in g[f"c{ci}_params"])
        kg = g[f"c{ci}_kg"]
        np.random.seed(seed)
        random.seed(seed)
        neg = ko.generate_negative(kg, N, K)
        X = ko.oversample_data([kg, neg])
        np.testing.assert_array_equal(neg, g[f"c{ci}_neg"])
        np.testing.assert_array_equal(X[0], g[f"c{ci}_pos_over"])
        np.testing.assert_array_equal(X[1], g[f"c{ci}_neg_over"])


def test_host_mirror_negatives_match_reference_golden():
    """KGEkeras.utils (the shipped host code) consumes np.random / random exactly like the reference."""
    from KGEkeras import utils as ku
    g = np.load(os.path.join(GOLD, "negatives_ref.npz"))
    for ci in range(4):
        B, N, K, seed = (int(x) for x in g[f"c{ci}_params"])
        kg = g[f"c{ci}_kg"]
        np.random.seed(seed)
        random.seed(seed)
        neg = ku.generate_negative(kg, N=N, negative=K)
        X = ku.oversample_data(kgs=[kg, neg])
        assert neg.shape == (K * B, 3, 1)
        np.testing.assert_array_equal(neg, g[f"c{ci}_neg"])
        np.testing.assert_array_equal(X[0], g[f"c{ci}_pos_over"])
        np.testing.assert_array_equal(X[1], g[f"c{ci}_neg_over"])
    with pytest.raises(AssertionError):
        ku.generate_negative(kg, N, corrupt_head=False, corrupt_tail=False)


with open(os.path.join(GOLD, "validate_ref.json")) as _fh:
    VALIDATE_META = json.load(_fh)


@pytest.mark.parametrize("key", [k for k in VALIDATE_META if k.split("/")[0] in ("kinship", "nations", "UMLS") and k.split("/")[1] in ("DistMult", "TransE_L1", "ComplEx")])
def test_oracle_validate_matches_reference_golden(key):
    meta = VALIDATE_META[key]
    arrs = np.load(os.path.join(GOLD, "validate_ref.npz"))
    ds = load_dataset(meta["dataset"])
    kind = KINDS[meta["kind"]]
    E, Rm = golden_tables(kind, ds["N"], ds["R"], meta["d"], meta["seed"], meta["dyadic"])
    test = ds["test"][:meta["ntest"]]
    filt = np.concatenate([ds["train"], ds["valid"]])
    m, tr, hr = ko.validate(kind, E, Rm, test, filt, meta["gamma"], np.float32)
    n = meta["n_ref_per_query"]
    np.testing.assert_array_equal(tr[:n], arrs[key + "/ref_tail_ranks"])   # ranks produced by the reference's validate
    np.testing.assert_array_equal(hr[:n], arrs[key + "/ref_head_ranks"])
    np.testing.assert_array_equal(tr, arrs[key + "/tail_ranks"])
    for k, v in meta["metrics"].items():                                   # the reference's own 15-key dict
        assert abs(m[k] - v) <= 2e-6 * max(1.0, abs(v)), k


@pytest.mark.skipif(not ref_ast.available(), reason="/root/reference not mounted (GPU box)")
def test_live_reference_functions_agree_with_mirror():
    from KGEkeras import utils as ku
    ref = ref_ast.load()
    ds = load_dataset("nations")
    test = [tuple(int(x) for x in r) for r in ds["test"][:20]]
    filt = [tuple(int(x) for x in r) for r in np.concatenate([ds["train"], ds["valid"]])]
    fh, ft = ko.build_filters(filt)
    for a, b in zip(ref["gen_tail_data"](test, ds["N"], 32, ft), ku.gen_tail_data(test, ds["N"], 32, ft)):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(ref["gen_head_data"](test, ds["N"], 32, fh), ku.gen_head_data(test, ds["N"], 32, fh)):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):   # target inside its own filter set: both crash the same way
        list(ref["gen_tail_data"](test[:1], ds["N"], 32, {(test[0][0], test[0][1]): {test[0][2]}}))
    with pytest.raises(ValueError):
        list(ku.gen_tail_data(test[:1], ds["N"], 32, {(test[0][0], test[0][1]): {test[0][2]}}))
    sc = [("a", 0.5), ("b", 0.9), ("c", 0.5), ("d", 0.1)]
    for t in "abcd":
        assert ref["mrr"](t, sc) == ku.mrr(t, sc) and ref["hits"](t, sc, 2) == ku.hits(t, sc, 2)
    path = "/root/reference/data/nations/valid.txt"
    assert ref["load_kg"](path) == ku.load_kg(path)
    random.seed(5)
    a = ref["pad"]([(1, 2, 3), (4, 5, 6), (7, 8, 9)], 4)
    random.seed(5)
    b = ku.pad([(1, 2, 3), (4, 5, 6), (7, 8, 9)], 4)
    np.testing.assert_array_equal(a, b)


def test_score_kat_frozen():
    g = np.load(os.path.join(GOLD, "score_kat.npz"))
    for kname, kind in KINDS.items():
        E, Rm, pos, neg = g[f"{kname}/E"], g[f"{kname}/R"], g[f"{kname}/pos"], g[f"{kname}/neg"]
        np.testing.assert_allclose(ko.score_triples(kind, E, Rm, pos, 12.0, np.float64, hole_fft=False), g[f"{kname}/score64"], rtol=1e-13)
        np.testing.assert_allclose(ko.score_triples(kind, E, Rm, pos, 12.0, np.float32), g[f"{kname}/score32"], rtol=2e-6, atol=1e-6)
        # fp32 op-by-op restatement vs fp64 truth: 1e-5 of the magnitude that enters the sum
        bound = 1e-5 * ko.score_abs_terms(kind, E, Rm, pos, 12.0)
        assert np.all(np.abs(g[f"{kname}/score32"].astype(np.float64) - g[f"{kname}/score64"]) <= bound)
        st = ko.train_step(kind, ko.PAIRWISE_HINGE, ko.PAIRED, E, Rm, pos, neg, 3, 1.0, 12.0, ko.OPT_LAZY_ADAM,
                           [np.zeros_like(E), np.zeros_like(E)], [np.zeros_like(Rm), np.zeros_like(Rm)], 1e-3, 1)
        np.testing.assert_allclose(st["loss"], g[f"{kname}/pairwize_hinge/0/loss"], rtol=1e-12)
        np.testing.assert_allclose(st["gE"], g[f"{kname}/pairwize_hinge/0/gE"], rtol=1e-10, atol=1e-14)


# ---- oracle internal consistency: analytic gradients vs torch autograd ------------------------------------------
@pytest.mark.parametrize("kname", list(KINDS))
def test_score_grads_match_autograd(kname):
    import torch
    kind = KINDS[kname]
    rng = np.random.default_rng(3)
    d = 6
    W = 2 * d if kind == ko.COMPLEX else d
    s, p, o = (rng.normal(size=(5, W)) for _ in range(3))
    ts, tp, to = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (s, p, o))
    if kind == ko.DISTMULT:
        f = (ts * tp * to).sum(-1)
    elif kind == ko.TRANSE_L1:
        f = 12.0 - (ts + tp - to).abs().sum(-1)
    elif kind == ko.TRANSE_L2:
        f = 12.0 - (ts + tp - to).pow(2).sum(-1).sqrt()
    elif kind == ko.COMPLEX:
        sr, si, pr, pi, orr, oi = ts[:, :d], ts[:, d:], tp[:, :d], tp[:, d:], to[:, :d], to[:, d:]
        f = (sr * pr * orr + pr * si * oi + pi * sr * oi - pi * si * orr).sum(-1)
    else:
        x = torch.fft.ifft(torch.conj(torch.fft.fft(ts)) * torch.fft.fft(to)).real
        f = (tp * x).sum(-1)
    np.testing.assert_allclose(f.detach().numpy(), ko.score(kind, s, p, o, 12.0, hole_fft=False), rtol=1e-10, atol=1e-12)
    w = torch.tensor(rng.normal(size=5))
    (f * w).sum().backward()
    ds, dp, do = ko.score_grads(kind, s, p, o, 12.0)
    for got, want in ((ds, ts.grad), (dp, tp.grad), (do, to.grad)):
        np.testing.assert_allclose(got * w.numpy()[:, None], want.numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("kname", list(ko.SIBLING_KINDS))
@pytest.mark.parametrize("gamma", [12.0, 0.0])
def test_sibling_scores_and_grads_match_torch_restatement(kname, gamma):
    """RotatE / pRotatE / ModE (models.py:403-489) written in torch op for op like the reference's TF code (atan2 of
    sin and cos included) against the oracle's closed forms and analytic gradients."""
    import torch
    kind = ko.SIBLING_KINDS[kname]
    rng = np.random.default_rng(8)
    d, modulus = 6, 0.37
    We, Wr = (2 * d, d) if kind == ko.ROTATE else (d, d)
    s, o = (rng.normal(size=(7, We)) * 2.5 for _ in range(2))   # |x| > pi occurs: the phase wrap is exercised
    p = rng.normal(size=(7, Wr)) * 2.5
    ts, tp, to = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (s, p, o))
    ph = lambda x: torch.atan2(torch.sin(x), torch.cos(x))
    if kind == ko.ROTATE:
        re_s, im_s, re_o, im_o = ts[:, :d], ts[:, d:], to[:, :d], to[:, d:]
        phase_r = ph(tp)
        re_r, im_r = torch.cos(phase_r), torch.sin(phase_r)
        sc = torch.cat([re_s * re_r - im_s * im_r - re_o, re_s * im_r + im_s * re_r - im_o], dim=1).sum(1)
        f = gamma - sc if gamma > 0 else sc
    elif kind == ko.PROTATE:
        sc = torch.abs(torch.sin((ph(ts) + ph(tp) - ph(to)) / 2)).sum(1) * modulus
        f = gamma - sc if gamma > 0 else sc
    else:
        r = ts * tp - to
        f = gamma - (r.abs().sum(-1) if kind == ko.MODE_L1 else r.pow(2).sum(-1).sqrt())
    np.testing.assert_allclose(f.detach().numpy(), ko.score(kind, s, p, o, gamma, param=modulus), rtol=1e-10, atol=1e-12)
    w = torch.tensor(rng.normal(size=7))
    (f * w).sum().backward()
    ds, dp, do = ko.score_grads(kind, s, p, o, gamma, param=modulus)
    for got, want in ((ds, ts.grad), (dp, tp.grad), (do, to.grad)):
        np.testing.assert_allclose(got * w.numpy()[:, None], want.numpy(), rtol=1e-9, atol=1e-12)
    # the ranking restatement (query vector + sequential-k scores) orders candidates like the score itself
    if gamma > 0:
        E, Rm = rng.normal(size=(40, We)).astype(np.float32), rng.normal(size=(3, Wr)).astype(np.float32)
        tri = np.stack([rng.integers(0, 40, 9), rng.integers(0, 3, 9), rng.integers(0, 40, 9)], axis=1)
        for side in ("tail", "head"):
            full = ko.all_entity_scores(kind, E, Rm, tri, side, gamma, np.float64, param=modulus)
            sk = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, tri, side), E, side, gamma, param=modulus).astype(np.float64)
            delta = full - sk        # RotatE drops a per-query constant; the others restate the score itself
            assert np.max(np.abs(delta - delta[:, :1])) < 2e-4


@pytest.mark.parametrize("lname", list(ko.LOSS_NAMES))
@pytest.mark.parametrize("pm", [ko.PAIRED, ko.ALLPAIRS])
def test_loss_coefs_match_autograd(lname, pm):
    import torch
    lk = ko.LOSS_NAMES[lname]
    rng = np.random.default_rng(lk + 10 * pm)
    B, K = 7, 3
    t, f = rng.normal(size=B), rng.normal(size=B * K)
    tt, tf_ = torch.tensor(t, requires_grad=True), torch.tensor(f, requires_grad=True)
    c = torch.tensor(np.float64(np.float32(1 + 1e-6)))
    M = 1.0
    if lk == ko.POINTWISE_HINGE:
        L = torch.relu(M - tt).mean() + torch.relu(M + tf_).mean()
    elif lk == ko.POINTWISE_LOGISTIC:
        L = torch.log(c + torch.exp(-tt)).mean() + torch.log(c + torch.exp(tf_)).mean()
    elif lk == ko.POINTWISE_SQUARE:
        L = ((M - tt) ** 2).mean() + ((M + tf_) ** 2).mean()
    elif lk == ko.POINTWISE_BCE:   # Keras binary_crossentropy(1, true) + binary_crossentropy(0, false) (utils.py:297-298)
        t[:3], f[:4] = [0.2, 0.9, 0.5], [0.1, 0.7, 0.999, 0.3]          # some scores inside the clip interval
        tt, tf_ = torch.tensor(t, requires_grad=True), torch.tensor(f, requires_grad=True)
        eps = 1e-7
        L = (-torch.log(torch.clamp(tt, eps, 1 - eps) + eps)).mean() + (-torch.log(1 - torch.clamp(tf_, eps, 1 - eps) + eps)).mean()
    else:
        X = (tf_.reshape(B, K) - tt.reshape(B, 1)) if pm == ko.PAIRED else (tf_.reshape(-1, 1) - tt.reshape(1, -1))
        L = {ko.PAIRWISE_HINGE: lambda: torch.relu(M + X).mean(), ko.PAIRWISE_LOGISTIC: lambda: torch.log(c + torch.exp(X)).mean(),
             ko.PAIRWISE_SQUARE: lambda: (-(X ** 2)).mean()}[lk]()
    L.backward()
    Lo, dt, df = ko.loss_and_coefs(lk, pm, t, f, K, M)
    np.testing.assert_allclose(Lo, L.item(), rtol=1e-12)
    np.testing.assert_allclose(dt, tt.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(df, tf_.grad.numpy(), rtol=1e-10, atol=1e-14)
    # and the reference-named function (restated) agrees on the value
    ref = ko.loss_function_lookup(lname)
    Lr = ref(t.reshape(-1, 1) if (pm == ko.PAIRED and ko.PAIRWISE_HINGE <= lk <= ko.PAIRWISE_SQUARE) else t, f, M, K if pm == ko.PAIRED else 1)
    np.testing.assert_allclose(Lr, Lo, rtol=1e-12)


def test_rank_identities_property():
    """rank = 1 + #greater (optimistic ties) == the reference's rankdata('max') formula, and sequential-k fp32 scores
    rank like the fp64 truth away from ties."""
    from scipy.stats import rankdata
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(2, 40))
        sc = rng.integers(0, 6, n).astype(np.float32)  # many ties
        tgt = int(rng.integers(0, n))
        res = np.concatenate([[sc[tgt]], np.delete(sc, tgt)])
        r = rankdata(res, 'max')
        assert len(res) - r[0] + 1 == 1 + np.count_nonzero(np.delete(sc, tgt) > sc[tgt])


# ---- host mirror: error behaviour and CSR build -------------------------------------------------------------------
def test_mirror_error_behaviour_and_csr():
    from KGEkeras import utils as ku
    from KGEkeras import models as km
    with pytest.raises(KeyError):
        ku.loss_function_lookup("hinge")
    with pytest.raises(TypeError):
        ku.validate(object(), [(0, 0, 1)], 5, 32)                 # filtering_triples=None, as the reference
    with pytest.raises(AttributeError):
        ku.validate(object(), [(0, 0, 1)], 5, 32, [(0, 0, 2)])    # no .predict: what utils.py:211 raises in the reference
    test = np.array([[0, 0, 1], [2, 1, 3]])
    filt = [(0, 0, 4), (0, 0, 2), (0, 0, 2), (9, 1, 3), (5, 1, 3)]
    offs, ids = ku.build_filter_csr(test, filt, "tail")
    assert offs.tolist() == [0, 2, 2] and ids.tolist() == [2, 4]
    offs, ids = ku.build_filter_csr(test, filt, "head")
    assert offs.tolist() == [0, 0, 2] and ids.tolist() == [5, 9]
    o2, i2 = ko.filter_csr(test, "head", ko.build_filters(filt)[0])
    assert o2.tolist() == offs.tolist() and i2.tolist() == ids.tolist()
    with pytest.raises(ValueError):
        ku.build_filter_csr(test, filt + [(0, 0, 1)], "tail")     # test triple inside the filter: the reference crashes too
    m = km.ComplEx(e_dim=8, r_dim=8, num_entities=10, num_relations=3, some_extra=7)
    assert (m.e_dim, m.r_dim, m.some_extra, m.num_entities, m.margin, m.dp.rate) == (16, 16, 7, 10, 1, 0.2)
    t = km.TransE(e_dim=8, r_dim=8, num_entities=10, num_relations=3, norm=2, gamma=5)
    assert (t.gamma, t.norm) == (5, 2)
    with pytest.raises(NameError):
        km.CosinE(e_dim=4, r_dim=4, num_entities=3, num_relations=2).func(None, None, None)
    with pytest.raises(NotImplementedError):
        km.ConvE(e_dim=4, r_dim=4, num_entities=3, num_relations=2)
    with pytest.raises(RuntimeError):                              # no GPU here -> loud failure, never a CPU path
        import torch
        if torch.cuda.is_available():
            raise RuntimeError("gpu present")
        m.predict(np.zeros((1, 3)))


# ---- C ABI: the library loads and exports everything the header declares ---------------------------------------------
def test_abi_library_exports_header_symbols():
    from KGEkeras import _abi
    with open(os.path.join(ROOT, "include", "kgeb200.h")) as fh:
        declared = set(re.findall(r"KGEB_API[^;(]*?\b(kgeb_[a-z0-9_]+)\s*\(", fh.read()))
    assert declared == set(_abi.PROTOTYPES), declared ^ set(_abi.PROTOTYPES)
    if not os.path.isfile(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _abi.lib().kgeb_version() == 110
    import torch
    if not torch.cuda.is_available():  # kgeb_create must fail cleanly (negative status + message), not crash
        h = ctypes.c_void_p()
        rc = _abi.lib().kgeb_create(ctypes.byref(h), 0, 10, 3, 8, 8, 8, 8, 0, 12.0)
        assert rc < 0 and not h.value and b"CUDA" in _abi.lib().kgeb_last_error(None)


# ---- dataset ingestion ----------------------------------------------------------------------------
def _write_split(path, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write("\t".join(r) + "\n")


def test_load_dataset_sorted_mapping_and_cache(tmp_path):
    from KGEkeras import data as kd
    d = tmp_path / "toy"
    d.mkdir()
    train = [("b", "likes", "a"), ("c", "knows", "b"), ("a", "likes", "c")]
    valid = [("d", "likes", "a")]
    test = [("a", "knows", "d")]
    for name, rows in (("train", train), ("valid", valid), ("test", test)):
        _write_split(d / f"{name}.txt", rows)
    ds = kd.load_dataset(str(d), cache_dir=str(tmp_path / "cache"))
    assert ds["entities"] == ["a", "b", "c", "d"] and ds["relations"] == ["knows", "likes"]
    assert ds["num_entities"] == 4 and ds["num_relations"] == 2
    np.testing.assert_array_equal(ds["train"], np.asarray([[1, 1, 0], [2, 0, 1], [0, 1, 2]], np.int32))
    np.testing.assert_array_equal(ds["valid"], [[3, 1, 0]])
    np.testing.assert_array_equal(ds["test"], [[0, 0, 3]])
    # second call is served from the cache and is identical
    assert len(os.listdir(tmp_path / "cache")) == 1
    ds2 = kd.load_dataset(str(d), cache_dir=str(tmp_path / "cache"))
    for k in ("train", "valid", "test"):
        np.testing.assert_array_equal(ds[k], ds2[k])
    assert ds2["entities"] == ds["entities"] and ds2["num_relations"] == 2
    # a changed file invalidates the cache key
    _write_split(d / "test.txt", test + [("e", "knows", "a")])
    os.utime(d / "test.txt", ns=(10 ** 18, 10 ** 18))
    ds3 = kd.load_dataset(str(d), cache_dir=str(tmp_path / "cache"))
    assert ds3["num_entities"] == 5
    # error behaviour: a missing split and a malformed line
    os.remove(d / "valid.txt")
    with pytest.raises(FileNotFoundError):
        kd.load_dataset(str(d))
    with pytest.raises(ValueError):
        kd.build_mappings([["a", "b"]])
    with pytest.raises(KeyError):
        kd.map_triples([("zz", "likes", "a")], {"a": 0}, {"likes": 0})


@pytest.mark.skipif(not ref_ast.available(), reason="/root/reference not mounted (GPU box)")
def test_load_dataset_matches_committed_fixture_on_kinship():
    from KGEkeras import data as kd
    ds = kd.load_dataset("/root/reference/data/kinship")
    gold = load_dataset("kinship")
    assert ds["num_entities"] == gold["N"] and ds["num_relations"] == gold["R"]
    for k in ("train", "valid", "test"):
        np.testing.assert_array_equal(ds[k], gold[k])


def test_data_generator_host_path_consumes_rng_like_the_harness():
    """tests.py:29-61 restated literally (set-based num_e, index shuffling, generate_negative + oversample_data) against
    KGEkeras.data.DataGenerator under the same global seeds."""
    import random
    from KGEkeras import data as kd
    from KGEkeras import utils as ku
    rng = np.random.default_rng(3)
    kg = [tuple(int(x) for x in r) for r in np.stack([rng.integers(0, 50, 200), rng.integers(0, 4, 200), rng.integers(0, 50, 200)], 1)]

    def harness_batches(seed, bs, ns, nb):
        np.random.seed(seed); random.seed(seed)
        num_e = len(set([s for s, _, _ in kg]) | set([o for _, _, o in kg]))
        indices = list(range(len(kg)))
        index = np.arange(len(indices))
        np.random.shuffle(index)
        out = []
        for b in range(nb):
            sel = index[b * bs:(b + 1) * bs]
            tmp = np.asarray([kg[indices[k]] for k in sel])
            neg = ku.generate_negative(tmp, N=num_e, negative=ns)
            out.append(ku.oversample_data(kgs=[tmp, neg]))
        return out

    want = harness_batches(11, 32, 5, 3)
    np.random.seed(11); random.seed(11)
    gen = kd.DataGenerator(kg, ns=5, batch_size=32)
    assert len(gen) == 200 // 32
    for b in range(3):
        (pos, neg), y = gen[b]
        assert y is None
        np.testing.assert_array_equal(pos, want[b][0])
        np.testing.assert_array_equal(neg, want[b][1])


# ---- optimizer restatement against independent implementations (third-party arithmetic) ----------
@pytest.mark.parametrize("opt", ["adam_dense", "lazy_adam", "adagrad", "sgd"])
def test_oracle_optimizer_matches_torch_optim(opt):
    """Keras/TF are not installable here, so the recalled update rules are pinned against torch.optim's implementations of
    the same published algorithms.  Keras Adam folds the bias corrections into the step size (epsilon 'hat' form,
    p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)); torch divides m and v first -- identical for eps -> 0, so the
    comparison runs with eps = 1e-12 on gradients of order 1.  Lazy variant: torch.optim.SparseAdam (touched rows only)."""
    import torch
    rng = np.random.default_rng(5)
    N, W, lr, steps = 30, 6, 0.05, 6
    P0 = rng.normal(size=(N, W))
    kind = {"adam_dense": ko.OPT_ADAM_DENSE, "lazy_adam": ko.OPT_LAZY_ADAM, "adagrad": ko.OPT_ADAGRAD, "sgd": ko.OPT_SGD}[opt]
    P = P0.copy()
    state = [np.zeros_like(P), np.zeros_like(P)] if "adam" in opt else ([np.full_like(P, 0.1)] if opt == "adagrad" else [])
    tp = torch.nn.Parameter(torch.tensor(P0.copy()))
    if opt == "adam_dense":
        topt = torch.optim.Adam([tp], lr=lr, betas=(0.9, 0.999), eps=1e-12)
    elif opt == "lazy_adam":
        topt = torch.optim.SparseAdam([tp], lr=lr, betas=(0.9, 0.999), eps=1e-12)
    elif opt == "adagrad":
        topt = torch.optim.Adagrad([tp], lr=lr, initial_accumulator_value=0.1, eps=1e-12)
    else:
        topt = torch.optim.SGD([tp], lr=lr)
    for step in range(1, steps + 1):
        touched = rng.random(N) < 0.5
        G = rng.normal(size=(N, W)) * touched[:, None]
        P, state = ko.opt_update(kind, P, G, touched, state, lr, step, eps=1e-12)
        if opt == "lazy_adam":
            rows = torch.tensor(np.nonzero(touched)[0])
            tp.grad = torch.sparse_coo_tensor(rows.unsqueeze(0), torch.tensor(G[touched]), (N, W))
        else:
            tp.grad = torch.tensor(G)
        topt.step()
        if opt in ("adagrad", "sgd"):
            # dense torch optimizers also "update" untouched rows with a zero gradient: a no-op for SGD / Adagrad
            pass
        np.testing.assert_allclose(P, tp.detach().numpy(), rtol=1e-7, atol=1e-10)


# ---- bench.py contract (reference arm runs on CPU, so it is testable here) ------------------------------------------
def test_bench_reference_arm_prints_contract_line():
    import json
    import subprocess
    import sys
    root = ROOT
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "filtered_eval_queries_per_s" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["scaling"] == "strong"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.headline_config(None)      # both arms print the same config keys and values


# ---- literal branch (models.py:114-120,152-158) -----------------------------------------------------------------------
@pytest.mark.parametrize("act,aname", [(ko.ACT_TANH, "tanh"), (ko.ACT_RELU, "relu"), (ko.ACT_SIGMOID, "sigmoid"), (ko.ACT_LINEAR, None)])
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_literal_step_matches_torch_restatement(act, aname, p):
    """call() of a model with literals written in torch op for op like the reference (concat, Dense, activation, func) with
    the oracle's dropout masks as fixed multipliers, against the oracle's closed-form forward and gradients."""
    import torch
    rng = np.random.default_rng(4)
    N, Rn, we, wl, B, K = 30, 4, 6, 3, 5, 2
    E, Rm, L = rng.normal(size=(N, we)), rng.normal(size=(Rn, we)), rng.random((N, wl))
    W, b = rng.normal(size=(we + wl, we)) * 0.5, rng.normal(size=we) * 0.1
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, Rn, B), rng.integers(0, N, B)], 1)
    neg = ko.philox_negatives(pos, N, K, 1, 0, ko.CORRUPT_BOTH)
    got = ko.literal_step(ko.DISTMULT, ko.PAIRWISE_LOGISTIC, ko.PAIRED, E, Rm, L, W, b, act, pos, neg, K, 0.0, 12.0, p=p, seed=9, offset=2)
    tE, tR, tW, tb = (torch.tensor(a, requires_grad=True) for a in (E, Rm, W, b))
    tL = torch.tensor(L)
    tri = np.concatenate([pos, neg]).astype(np.int64)
    m = len(tri)
    mX = torch.tensor(ko.dropout_mult(ko.DISTMULT, we + wl, 2 * m, 0, 0, p, 9, 2)) if p else torch.ones(2 * m, we + wl, dtype=torch.float64)
    mP = torch.tensor(ko.dropout_mult(ko.DISTMULT, we, m, 0, 1, p, 9, 2)) if p else torch.ones(m, we, dtype=torch.float64)
    refs = torch.tensor(np.concatenate([tri[:, 0], tri[:, 2]]))
    X = torch.cat([tE[refs], tL[refs]], dim=1) * mX
    Z = X @ tW + tb
    Y = {ko.ACT_TANH: torch.tanh, ko.ACT_RELU: torch.relu, ko.ACT_SIGMOID: torch.sigmoid, ko.ACT_LINEAR: lambda z: z}[act](Z)
    P = tR[torch.tensor(tri[:, 1])] * mP
    sc = (Y[:m] * P * Y[m:]).sum(-1)
    t, f = sc[:B], sc[B:]
    c = torch.tensor(np.float64(np.float32(1 + 1e-6)))
    Lt = torch.log(c + torch.exp(f.reshape(B, K) - t.reshape(B, 1))).mean()
    Lt.backward()
    np.testing.assert_allclose(got["loss"], Lt.item(), rtol=1e-12)
    for name, g in (("gE", tE.grad), ("gR", tR.grad), ("gW", tW.grad), ("gb", tb.grad)):
        np.testing.assert_allclose(got[name], g.numpy(), rtol=1e-9, atol=1e-13, err_msg=name)
    if not p:   # without dropout the transformed row is a function of the entity: the inference table
        Ep = ko.literal_table(E, L, W, b, act)
        np.testing.assert_allclose(Ep[tri[:, 0]], Y[:m].detach().numpy(), rtol=1e-12)


def test_hake_oracle_matches_torch_restatement():
    """HAKE.func (models.py:383-400) typed op for op in torch (chunk / clamp / where / Frobenius norm without axis / L1 norm
    of the sine) against oracle.hake_scores and, through autograd, oracle.hake_row_grads -- including biases above the clip,
    below -|mod_p|, and the coupling of every score of a call through the one modulus norm."""
    import torch
    rng = np.random.default_rng(0)
    d, N, R, m = 6, 20, 4, 48
    E = rng.normal(size=(N, 2 * d)) * 0.5
    Rm = rng.normal(size=(R, 3 * d)) * 0.8
    Rm[:, 2 * d:] += 0.5
    idx = np.stack([rng.integers(0, N, m), rng.integers(0, R, m), rng.integers(0, N, m)], 1)
    gamma, eps, pw, mw = 12.0, 2.0, 0.5, 1.5
    hp = ko.hake_params(gamma, eps, 2 * d, pw, mw)
    s, p, o = ko.gather(E, Rm, idx)
    assert (p[:, 2 * d:] > 1).any() and (np.minimum(p[:, 2 * d:], 1) < -np.abs(p[:, d:2 * d])).any()

    def func(s, p, o):
        e_dim = s.shape[-1]
        embedding_range = (gamma + eps) / e_dim / 2
        phase_s, mod_s = torch.chunk(s, 2, -1)
        phase_o, mod_o = torch.chunk(o, 2, -1)
        phase_p, mod_p, bias_p = torch.chunk(p, 3, -1)
        phase_s, phase_p, phase_o = (x / (embedding_range / np.pi) for x in (phase_s, phase_p, phase_o))
        bias_p = torch.clamp(bias_p, max=1.0)
        bias_p = torch.where(bias_p < -mod_p.abs(), -mod_p.abs(), bias_p)
        r_score = mw * torch.linalg.norm(mod_s * (mod_p + bias_p) - mod_o.abs() * (1 - bias_p))
        p_score = pw * torch.sin((phase_s + phase_p - phase_o) / 2).abs().sum(-1)
        return gamma - (p_score + r_score)

    ts, tp, to = (torch.tensor(x, requires_grad=True) for x in (s, p, o))
    out = func(ts, tp, to)
    np.testing.assert_allclose(out.detach().numpy(), ko.hake_scores(s, p, o, gamma, *hp), rtol=1e-12, atol=1e-12)
    c = rng.normal(size=m)
    (out * torch.tensor(c)).sum().backward()
    gs, gp, go = ko.hake_row_grads(s, p, o, c, gamma, *hp)
    for got, want in ((ts.grad, gs), (tp.grad, gp), (to.grad, go)):
        np.testing.assert_allclose(got.numpy(), want, rtol=1e-10, atol=1e-12)
    # hake_step == autograd through loss_and_coefs' loss on the two calls
    pos, neg = idx[:12], idx[12:]
    want = ko.hake_step(ko.PAIRWISE_LOGISTIC, ko.PAIRED, E, Rm, pos, neg, 3, 0.0, gamma, hp)
    tE, tR = torch.tensor(E, requires_grad=True), torch.tensor(Rm, requires_grad=True)
    tpos, tneg = torch.tensor(pos), torch.tensor(neg)
    t = func(tE[tpos[:, 0]], tR[tpos[:, 1]], tE[tpos[:, 2]])
    f = func(tE[tneg[:, 0]], tR[tneg[:, 1]], tE[tneg[:, 2]])
    L, dt_, df_ = ko.loss_and_coefs(ko.PAIRWISE_LOGISTIC, ko.PAIRED, t.detach().numpy(), f.detach().numpy(), 3, 0.0)
    ((t * torch.tensor(dt_)).sum() + (f * torch.tensor(df_)).sum()).backward()
    np.testing.assert_allclose(tE.grad.numpy(), want["gE"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(tR.grad.numpy(), want["gR"], rtol=1e-9, atol=1e-12)
    assert abs(L - want["loss"]) < 1e-12


def test_hake_and_hole_constructors_without_gpu():
    """Width multipliers and derived constants of the sibling constructors (models.py:371-381) are host logic: no GPU needed."""
    import KGEkeras
    m = KGEkeras.HAKE(e_dim=10, r_dim=10, num_entities=5, num_relations=2, gamma=12, epsilon=2, phase_weight=0.25, mod_weight=2.0)
    assert (m.e_dim, m.r_dim) == (20, 30)
    assert abs(m.embedding_range - (12 + 2) / 20 / 2) < 1e-12
    ps, pw, mw = m._param_()
    assert abs(ps - m.embedding_range / np.pi) < 1e-12 and (pw, mw) == (0.25, 2.0)      # the DIVISOR of models.py:390-392
    assert ko.hake_params(12, 2, 20, 0.25, 2.0) == pytest.approx((ps, pw, mw))
    h = KGEkeras.HolE(e_dim=64, r_dim=64, num_entities=5, num_relations=2, fft=False)
    assert h.fft is False and KGEkeras.HolE(e_dim=64, r_dim=64, num_entities=5, num_relations=2).fft is True
    assert KGEkeras.isint("12") and not KGEkeras.isint("1.5x")


def test_tf_shim_type_checks_against_the_header():
    """The TensorFlow op shim cannot be built here (no TensorFlow), but its calls into the C ABI can be type-checked: compile it with
    -fsyntax-only against tests/tf_stub (a stand-in for the few TensorFlow classes it uses) and the REAL include/kgeb200.h."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no CUDA headers")
    src = os.path.join(ROOT, "kge-keras_b200", "tf_shim", "kgeb_tf_ops.cc")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "tf_stub"),
                        "-I" + os.path.join(ROOT, "include"), "-I" + cuda_inc, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(src) as fh:
        text = fh.read()
    for op in ("KgebScore", "KgebScoreGrad", "KgebScoreGradSlices", "KgebTrainStep", "KgebRank1N"):
        assert f'REGISTER_OP("{op}")' in text and f'Name("{op}")' in text, op


@pytest.mark.parametrize("kind", [ko.DISTMULT, ko.TRANSE_L1, ko.TRANSE_L2, ko.COMPLEX, ko.HOLE])
def test_dense_grads_abs_bounds_the_gradient_and_the_fp32_pass(kind):
    """The magnitude the gradient parity bar is stated against (sum of |terms| of an entry) dominates the entry itself, and the
    oracle's own fp32 pass sits inside 1e-5 of it: the bar the GPU tests apply is attainable in fp32 for every headline kind."""
    rng = np.random.default_rng(5 + kind)
    N, R, d, m = 40, 4, 16, 200
    W = 2 * d if kind == ko.COMPLEX else d
    E, Rm = ko.glorot_uniform(rng, N, W) * 4.0, ko.glorot_uniform(rng, R, W) * 4.0
    idx = np.stack([rng.integers(0, N, m), rng.integers(0, R, m), rng.integers(0, N, m)], axis=1).astype(np.int32)
    dscore = rng.normal(size=m)
    g64 = ko.dense_grads(kind, E, Rm, idx, dscore, 2.0, np.float64)
    g32 = ko.dense_grads(kind, E, Rm, idx, dscore.astype(np.float32), 2.0, np.float32)
    aE, aR = ko.dense_grads_abs(kind, E, Rm, idx, dscore, 2.0)
    assert np.all(np.abs(g64[0]) <= aE * (1 + 1e-12) + 1e-300) and np.all(np.abs(g64[1]) <= aR * (1 + 1e-12) + 1e-300)
    assert np.all(np.abs(g32[0].astype(np.float64) - g64[0]) <= 1e-5 * aE + 1e-30)
    assert np.all(np.abs(g32[1].astype(np.float64) - g64[1]) <= 1e-5 * aR + 1e-30)
