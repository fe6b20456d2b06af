"""GPU tests of the reference-facing API (KGEkeras.models / KGEkeras.utils mirror) against the oracle and the golden
fixtures produced by the reference's own validate / generate_negative code."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")
KINDS = {"DistMult": ko.DISTMULT, "TransE_L1": ko.TRANSE_L1, "TransE_L2": ko.TRANSE_L2, "ComplEx": ko.COMPLEX, "HolE": ko.HOLE}
with open(os.path.join(GOLD, "validate_ref.json")) as _fh:
    VALIDATE_META = json.load(_fh)


def load_dataset(name):
    d = np.load(os.path.join(GOLD, "datasets", name + ".npz"))
    return dict(N=int(d["num_entities"]), R=int(d["num_relations"]), train=d["train"], valid=d["valid"], test=d["test"])


def golden_tables(kind, N, R, d, seed, dyadic):
    rng = np.random.default_rng(seed)
    W = 2 * d if kind == ko.COMPLEX else d
    if dyadic:
        return ko.dyadic_table(rng, N, W), ko.dyadic_table(rng, R, W)
    return ko.glorot_uniform(rng, N, W), ko.glorot_uniform(rng, R, W)


def make_model(kname, N, R, d, E=None, Rm=None, **kw):
    from KGEkeras import models as km
    args = dict(e_dim=d, r_dim=d, num_entities=N, num_relations=R, init_entities=E, init_relations=Rm, dp=0.0, **kw)
    if kname == "DistMult":
        return km.DistMult(**args)
    if kname == "TransE_L1":
        return km.TransE(norm=1, gamma=12, **args)
    if kname == "TransE_L2":
        return km.TransE(norm=2, gamma=12, **args)
    if kname == "ComplEx":
        return km.ComplEx(**args)
    return km.HolE(**args)


@pytest.mark.parametrize("kname", list(KINDS))
def test_model_call_predict_func(kname):
    kind = KINDS[kname]
    N, R, d = 60, 5, 8
    E, Rm = golden_tables(kind, N, R, d, 1, False)
    m = make_model(kname, N, R, d, E, Rm, some_unknown_kw=3)
    assert m.some_unknown_kw == 3 and m.num_entities == N
    rng = np.random.default_rng(0)
    idx = np.stack([rng.integers(0, N, 40), rng.integers(0, R, 40), rng.integers(0, N, 40)], 1)
    want = ko.score_triples(kind, E, Rm, idx, 12.0, np.float64, hole_fft=False)
    bound = 1e-5 * ko.score_abs_terms(kind, E, Rm, idx, 12.0)
    for got in (m(idx).cpu().numpy(), m.call(idx.astype(np.float32)).cpu().numpy(), m.predict(idx), m.predict(idx.reshape(-1, 3, 1)[..., 0])):
        assert got.shape == (40,) and np.all(np.abs(got - want) <= bound)
    s, p, o = ko.gather(E, Rm, idx)
    assert np.all(np.abs(m.func(s, p, o).cpu().numpy() - want) <= bound)
    w = m.get_weights()
    np.testing.assert_array_equal(w[0], E)
    np.testing.assert_array_equal(m.entity_embedding.get_weights()[0], E)
    m.set_weights([E * 2, Rm])
    np.testing.assert_array_equal(m.get_weights()[0], E * 2)


@pytest.mark.parametrize("key", sorted(VALIDATE_META))
def test_validate_matches_reference_golden(key):
    """Filtered ranks / the 15-key metrics dict against what the reference's own validate produced (tests/golden)."""
    from KGEkeras import utils as ku
    meta = VALIDATE_META[key]
    arrs = np.load(os.path.join(GOLD, "validate_ref.npz"))
    ds = load_dataset(meta["dataset"])
    kind = KINDS[meta["kind"]]
    E, Rm = golden_tables(kind, ds["N"], ds["R"], meta["d"], meta["seed"], meta["dyadic"])
    model = make_model(meta["kind"], ds["N"], ds["R"], meta["d"], E, Rm)
    test = ds["test"][:meta["ntest"]]
    filt = np.concatenate([ds["train"], ds["valid"]])
    tr, hr = ku.filtered_ranks(model, test, filt)
    metrics = ku.validate(model, [tuple(int(x) for x in r) for r in test], ds["N"], 32, filt)
    assert list(metrics) == ['tail_hits@1', 'tail_hits@3', 'tail_hits@10', 'tail_mr', 'tail_mrr', 'head_hits@1', 'head_hits@3',
                             'head_hits@10', 'head_mr', 'head_mrr', 'hits@1', 'hits@3', 'hits@10', 'mr', 'mrr']  # utils.py:247-262
    assert set(metrics) == set(meta["metrics"])
    if meta["dyadic"]:
        # every score is exact in fp32 -> bit-exact ranks, ties included, against the reference's output
        np.testing.assert_array_equal(tr, arrs[key + "/tail_ranks"])
        np.testing.assert_array_equal(hr, arrs[key + "/head_ranks"])
        n = meta["n_ref_per_query"]
        np.testing.assert_array_equal(tr[:n], arrs[key + "/ref_tail_ranks"])
        np.testing.assert_array_equal(hr[:n], arrs[key + "/ref_head_ranks"])
        for k, v in meta["metrics_f64"].items():
            assert abs(metrics[k] - v) <= 1e-12 * max(1.0, abs(v)), k
    else:
        # real-valued tables: bit-exact against the sequential-k fp32 emulation of the kernel's arithmetic ...
        fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in filt])
        for side, got, f in (("tail", tr, ft), ("head", hr, fh)):
            sc = ko.seqk_scores(kind, ko.query_vectors(kind, E, Rm, test, side), E, side, meta["gamma"])
            np.testing.assert_array_equal(got, ko.ranks_from_scores(sc, test, side, f))
        # ... and within near-tie noise of the reference's ranks (fp32, another summation order)
        gt, gh = arrs[key + "/tail_ranks"], arrs[key + "/head_ranks"]
        assert np.mean(tr != gt) <= 0.02 and np.mean(hr != gh) <= 0.02 and np.max(np.abs(tr - gt)) <= 2
        for k, v in meta["metrics_f64"].items():
            assert abs(metrics[k] - v) <= 2e-3 * max(1.0, abs(v)), k


def test_validate_error_behaviour():
    from KGEkeras import utils as ku
    ds = load_dataset("nations")
    m = make_model("DistMult", ds["N"], ds["R"], 4)
    test = [tuple(int(x) for x in r) for r in ds["test"][:5]]
    with pytest.raises(TypeError):
        ku.validate(m, test, ds["N"], 32)
    with pytest.raises(ValueError):  # filter contains the test triples themselves
        ku.validate(m, test, ds["N"], 32, test)


@pytest.mark.parametrize("lname", ["pointwize_hinge", "pointwize_logistic", "pointwize_square_loss", "pointwize_cross_entropy",
                                   "pairwize_hinge", "pairwize_logistic", "pairwize_square_loss"])
def test_loss_functions_like_reference(lname):
    from KGEkeras import utils as ku
    fn, ref = ku.loss_function_lookup(lname), ko.loss_function_lookup(lname)
    rng = np.random.default_rng(4)
    B, K = 50, 4
    t = rng.normal(size=B).astype(np.float32)
    f = rng.normal(size=B * K).astype(np.float32)
    if lname == "pointwize_cross_entropy":
        t, f = rng.random(B).astype(np.float32), rng.random(B * K).astype(np.float32)
    pair = lname.startswith("pairwize")
    # paired layout: true (B,1) against false (B*K,) with negative_samples = K
    got = float(fn(t.reshape(-1, 1) if pair else t, f, 1.0, K))
    want = float(ref((t.reshape(-1, 1) if pair else t).astype(np.float64), f.astype(np.float64), 1.0, K))
    assert abs(got - want) <= 1e-5 * max(1.0, abs(want))
    # harness layout (tests.py:81): negative_samples = 1 and equal lengths -> (m,m) broadcast
    f1 = f[:B]
    got = float(fn(t, f1, 1.0, 1))
    want = float(ref(t.astype(np.float64), f1.astype(np.float64), 1.0, 1))
    assert abs(got - want) <= 1e-5 * max(1.0, abs(want))
    if pair:
        mat = fn(t.reshape(-1, 1), f, 1.0, K, reduce_mean=False).cpu().numpy()
        np.testing.assert_allclose(mat, ref(t.reshape(-1, 1), f, 1.0, K, reduce_mean=False), rtol=1e-5, atol=1e-6)
        with pytest.raises(ValueError):  # (B,) true against (B,K) false does not broadcast, in TF/NumPy and here
            fn(t, f, 1.0, K)


class _Gen:
    """The harness' DataGenerator (tests.py:29-61) re-typed without Keras: yields ([pos, neg], None)."""

    def __init__(self, kg, num_e, ns=10, batch_size=32):
        from KGEkeras import utils as ku
        self.ku, self.kg, self.ns, self.bs, self.num_e = ku, np.asarray(kg), ns, min(batch_size, len(kg)), num_e
        self.on_epoch_end()

    def __len__(self):
        return len(self.kg) // self.bs

    def on_epoch_end(self):
        self.index = np.random.permutation(len(self.kg))

    def __getitem__(self, i):
        tmp = self.kg[self.index[i * self.bs:(i + 1) * self.bs]]
        neg = self.ku.generate_negative(tmp, N=self.num_e, negative=self.ns)
        return self.ku.oversample_data(kgs=[tmp, neg]), None


@pytest.mark.parametrize("kname,lname,pair_mode", [("DistMult", "pairwize_hinge", "allpairs"), ("TransE_L1", "pairwize_hinge", "paired"),
                                                   ("ComplEx", "pointwize_logistic", "paired"), ("HolE", "pairwize_logistic", "paired")])
def test_fit_learns_kinship(kname, lname, pair_mode):
    """End to end through the mirror: fit() on bundled kinship drives the loss down and filtered MRR up."""
    from KGEkeras import utils as ku
    np.random.seed(0)
    random.seed(0)
    ds = load_dataset("kinship")
    d = 16 if kname == "HolE" else 32
    m = make_model(kname, ds["N"], ds["R"], d, loss_function=lname, negative_samples=8, margin=1, batch_size=512)
    m.pair_mode = pair_mode
    m.corrupt = "either"
    m.compile(optimizer=type("Adam", (), {"learning_rate": 0.02})())
    test = ds["test"][:200]
    filt = np.concatenate([ds["train"], ds["valid"]])
    before = ku.validate(m, test, ds["N"], 32, filt)
    if pair_mode == "allpairs":   # the reference harness' own input form: a Sequence of ([pos, neg], None)
        hist = m.fit(_Gen(ds["train"], ds["N"], ns=4, batch_size=512), epochs=6, verbose=0)
    else:
        cb = ku.KGEValidateCallback(test, train_data=filt)
        hist = m.fit(ds["train"], epochs=12, batch_size=512, verbose=0, callbacks=[cb], validation_data=ds["valid"])
        assert "val_loss" in hist.history
    after = ku.validate(m, test, ds["N"], 32, filt)
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    assert after["mrr"] > before["mrr"] * 1.5, (before["mrr"], after["mrr"])
    assert isinstance(m.evaluate(ds["valid"], verbose=0), float)


def test_sharded_validate_equals_single_gpu_emulated():
    """Row shards of the replicated table counted one after the other sum to the unsharded ranks (what every NCCL rank does)."""
    from KGEkeras import utils as ku, _abi as abi
    ds = load_dataset("UMLS")
    E, Rm = golden_tables(ko.COMPLEX, ds["N"], ds["R"], 8, 3, False)
    m = make_model("ComplEx", ds["N"], ds["R"], 8, E, Rm)
    test = ds["test"][:100]
    filt = np.concatenate([ds["train"], ds["valid"]])
    tr, hr = ku.filtered_ranks(m, test, filt)
    eng = m._engine_()
    for world in (2, 4, 8):
        for side, code, want in (("tail", abi.SIDE_TAIL, tr), ("head", abi.SIDE_HEAD, hr)):
            offs, ids = ku.build_filter_csr(test, filt, side)
            acc = torch.zeros(len(test), dtype=torch.int32, device="cuda")
            for r in range(world):
                acc += eng.rank_counts(test, code, offs, ids, row_range=ku.shard_rows(ds["N"], world, r))
            np.testing.assert_array_equal(acc.cpu().numpy() + 1, want)


def test_fit_with_device_side_data_generator():
    """The harness' own loop (tests.py:130-134) with KGEkeras.DataGenerator: host batches (reference RNG consumption) and
    device batches (Philox negatives + oversampling on the GPU, nothing crosses PCIe) both train the all-pairs hinge."""
    import KGEkeras
    from KGEkeras import utils as ku
    np.random.seed(0)
    random.seed(0)
    ds = load_dataset("kinship")
    test = ds["test"][:200]
    filt = np.concatenate([ds["train"], ds["valid"]])
    for device_side in (False, True):
        m = make_model("DistMult", ds["N"], ds["R"], 32, loss_function="pairwize_hinge", negative_samples=4, margin=1, batch_size=512)
        m.pair_mode = "allpairs"
        m.compile(optimizer=type("Adam", (), {"learning_rate": 0.02})())
        gen = KGEkeras.DataGenerator(ds["train"], ns=4, batch_size=512, engine=m._engine_() if device_side else None, seed=5)
        (pos, neg), y = gen[0]
        assert y is None and tuple(pos.shape) == (2048, 3) and tuple(neg.shape) == (2048, 3)
        if device_side:
            assert pos.is_cuda and neg.is_cuda
            # every oversampled positive is a row of the batch, every negative keeps a relation of the batch
            assert int(neg[:, [0, 2]].max()) < ds["N"]
        before = ku.validate(m, test, ds["N"], 32, filt)
        hist = m.fit(gen, epochs=6, verbose=0)
        after = ku.validate(m, test, ds["N"], 32, filt)
        assert hist.history["loss"][-1] < hist.history["loss"][0]
        assert after["mrr"] > before["mrr"] * 1.5, (device_side, before["mrr"], after["mrr"])


@pytest.mark.parametrize("side", ["tail", "head"])
def test_device_filter_csr_builder_equals_host_builder(side):
    from KGEkeras import utils as ku
    rng = np.random.default_rng(12)
    N, R, F, Q = 300, 5, 4000, 500
    filt = np.stack([rng.integers(0, N, F), rng.integers(0, R, F), rng.integers(0, N, F)], 1).astype(np.int64)
    test = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int64)
    # remove the test triples themselves from the filter (the reference crashes on them, and so do both builders)
    key = lambda a: (a[:, 0] * R + a[:, 1]) * N + a[:, 2]
    filt = filt[~np.isin(key(filt), key(test))]
    ho, hi = ku.build_filter_csr(test, filt, side)
    do, di = ku.build_filter_csr_device(torch.from_numpy(test).cuda(), torch.from_numpy(filt).cuda(), side)
    np.testing.assert_array_equal(do.cpu().numpy(), ho)
    np.testing.assert_array_equal(di.cpu().numpy(), hi)
    with pytest.raises(ValueError):
        ku.build_filter_csr_device(torch.from_numpy(test).cuda(), torch.from_numpy(np.concatenate([filt, test[:1]])).cuda(), side)
    eo, ei = ku.build_filter_csr_device(torch.from_numpy(test).cuda(), torch.zeros(0, 3, dtype=torch.int64).cuda(), side)
    assert eo.numel() == Q + 1 and ei.numel() == 0 and int(eo.sum()) == 0


@pytest.mark.gpu
def test_module_level_l3_reg():
    """models.l3_reg(weight_matrix, w) (models.py:10-11) on arrays of awkward widths, CUDA tensors and a model's Embedding stand-in."""
    import KGEkeras
    rng = np.random.default_rng(3)
    for rows, width in ((7, 5), (100, 100), (33, 202)):
        W = rng.normal(size=(rows, width)).astype(np.float32)
        want = 0.01 * float(np.sum(np.abs(W.astype(np.float64)) ** 3))
        assert abs(KGEkeras.l3_reg(W) - want) <= 2e-6 * want
        assert abs(KGEkeras.l3_reg(torch.from_numpy(W).cuda(), w=0.5) - 50 * want) <= 2e-6 * 50 * want
    m = KGEkeras.DistMult(e_dim=12, r_dim=12, num_entities=40, num_relations=3)
    E = m.entity_embedding.get_weights()[0]
    want = 0.01 * float(np.sum(np.abs(E.astype(np.float64)) ** 3))
    assert abs(KGEkeras.l3_reg(m.entity_embedding) - want) <= 2e-6 * want


@pytest.mark.gpu
@pytest.mark.parametrize("side", ["tail", "head"])
@pytest.mark.parametrize("N,R,F,Q", [(300, 5, 4000, 500), (17, 40, 900, 33), (40943, 11, 93003, 3134), (50, 3, 1, 1)])
def test_library_filter_csr_equals_host_builder(side, N, R, F, Q):
    """kgeb_filter_csr_count / _fill (Engine.filter_csr, what validate() uses) against the NumPy builder: duplicates in the
    filter, relations outnumbering entities, WN18RR-sized inputs, queries without any filter entry, the target-in-filter error."""
    from KGEkeras import utils as ku, _abi as abi
    from KGEkeras._engine import Engine
    rng = np.random.default_rng(N + F)
    eng = Engine(ko.DISTMULT, N, R, 4, 4)
    filt = np.stack([rng.integers(0, N, F), rng.integers(0, R, F), rng.integers(0, N, F)], 1).astype(np.int64)
    filt = np.concatenate([filt, filt[: F // 3]])                       # duplicates
    test = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int64)
    key = lambda a: (a[:, 0] * R + a[:, 1]) * N + a[:, 2]
    filt = filt[~np.isin(key(filt), key(test))]
    if len(filt) == 0:
        filt = np.array([[(test[0, 0] + 1) % N, test[0, 1], test[0, 2]]], dtype=np.int64)
    code = abi.SIDE_TAIL if side == "tail" else abi.SIDE_HEAD
    ho, hi = ku.build_filter_csr(test, filt, side)
    do, di = eng.filter_csr(test, filt, code)
    np.testing.assert_array_equal(do.cpu().numpy(), ho)
    np.testing.assert_array_equal(di.cpu().numpy(), hi)
    with pytest.raises(ValueError):
        eng.filter_csr(test, np.concatenate([filt, test[:1]]), code)
    eo, ei = eng.filter_csr(test, np.zeros((0, 3), np.int64), code)
    assert eo.numel() == Q + 1 and ei.numel() == 0 and int(eo.sum()) == 0


@pytest.mark.parametrize("N", [37, 1024, 5000])
def test_grad_rows_export_is_the_indexed_slices_of_the_backward(N):
    """kgeb_grad_rows_export: (unique_ids, grad_rows) == the touched rows of the dense gradient buffer, ids ascending; with clear
    the buffer is all-zero afterwards, and feeding the pair to an external optimizer reproduces the in-library SGD step."""
    from KGEkeras import _abi as abi
    from KGEkeras._engine import Engine
    R, d, B, K = 5, 16, 64, 3
    rng = np.random.default_rng(N)
    E0, R0 = ko.glorot_uniform(rng, N, d), ko.glorot_uniform(rng, R, d)
    pos = np.stack([rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)], 1).astype(np.int32)
    a, b = Engine(abi.DISTMULT, N, R, d, d, 12.0, E=E0, R=R0), Engine(abi.DISTMULT, N, R, d, d, 12.0, E=E0, R=R0)
    kw = dict(K=K, loss_kind=abi.PAIRWISE_LOGISTIC, margin=1.0, seed=4, offset=1)
    a.forward_backward(pos, None, opt_kind=abi.SGD, **kw)
    dense, flags = a.gE.clone(), a.tE.clone()
    ids, rows = a.export_grad_rows("entity", clear=True)
    want = torch.nonzero(flags).reshape(-1)
    assert torch.equal(ids.long(), want) and rows.shape == (len(want), d)
    assert torch.equal(rows, dense[want][:, :d])
    assert float(a.gE.abs().sum().item()) == 0.0 and int(a.tE.sum().item()) == 0
    rid, rrows = a.export_grad_rows("relation", clear=True)
    # external SGD on the exported slices == the library's own touched-row SGD
    a.E[ids.long(), :d] -= 0.1 * rows
    a.Rel[rid.long(), :d] -= 0.1 * rrows
    b.train_step(pos, None, opt_kind=abi.SGD, lr=0.1, **kw)
    # a and b are two runs of the backward: their fp32 reductions into a row may be applied in a different order (last-bit differences
    # on rows with many references; seen once in ~10 runs at N = 37), so the two steps agree to rounding, not bit for bit
    np.testing.assert_allclose(a.get_tables()[0], b.get_tables()[0], rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(a.get_tables()[1], b.get_tables()[1], rtol=2e-6, atol=1e-8)


def test_device_generator_draws_negatives_below_the_splits_entity_count():
    """the host path corrupts with ids in [0, num_e), num_e = distinct entities of the split (tests.py:34); the device path
    must use the same bound even when the model's vocabulary is larger."""
    import KGEkeras
    rng = np.random.default_rng(3)
    kg = np.stack([rng.integers(0, 40, 600), rng.integers(0, 3, 600), rng.integers(0, 40, 600)], 1)      # ids < 40 ...
    m = KGEkeras.DistMult(e_dim=8, r_dim=8, num_entities=500, num_relations=3, dp=0.0)                    # ... vocabulary of 500
    gen = KGEkeras.DataGenerator(kg, ns=5, batch_size=128, engine=m._engine_(), seed=1)
    assert gen.num_e == len(np.unique(kg[:, [0, 2]])) <= 40
    seen = 0
    for i in range(len(gen)):
        (pos, neg), _ = gen[i]
        neg = neg.cpu().numpy()
        assert neg[:, [0, 2]].max() < gen.num_e and neg[:, [0, 2]].min() >= 0
        seen = max(seen, int(neg[:, [0, 2]].max()))
    assert seen >= gen.num_e - 3          # the whole range is used
    # full-vocabulary bound: unchanged default
    neg = m._engine_().negatives(kg[:128], 8, 1, 1).cpu().numpy()
    assert neg[:, [0, 2]].max() > 60


def test_out_of_range_ids_raise_on_the_evaluation_path():
    import KGEkeras
    from KGEkeras import utils as ku
    m = KGEkeras.DistMult(e_dim=8, r_dim=8, num_entities=50, num_relations=3, dp=0.0)
    good = np.asarray([[1, 0, 2], [3, 1, 4]])
    with pytest.raises(IndexError, match="out of range"):
        ku.validate(m, np.asarray([[1, 0, 50]]), 50, 32, good)
    with pytest.raises(IndexError, match="out of range"):
        ku.validate(m, good, 50, 32, np.asarray([[1, 3, 2]]))
    with pytest.raises(IndexError, match="out of range"):
        ku.validate(m, np.asarray([[-1, 0, 2]]), 50, 32, good)


def test_filter_csr_serves_tables_beyond_the_first_builds_key_limit():
    """the packed filter key cubed max(N, R) and refused N > 1.66 M; per-column spans (N * R * N < 2^63) lift that.
    A context with N = 5 M entities (no table is needed for the filter build) must build the same CSR as the host mirror."""
    import ctypes as C
    from KGEkeras import _abi as abi, utils as ku
    from KGEkeras._engine import Engine
    lib = abi.lib()
    N, R, Q, F = 5_000_000, 40, 300, 5000
    rng = np.random.default_rng(8)
    q = np.stack([rng.integers(0, N, Q), rng.integers(0, R, Q), rng.integers(0, N, Q)], 1).astype(np.int32)
    f = np.stack([rng.integers(0, N, F), rng.integers(0, R, F), rng.integers(0, N, F)], 1).astype(np.int32)
    f[:Q, :2] = q[:, :2]                       # some filter triples share (s, p) ...
    f[Q:2 * Q, 1:] = q[:, 1:]                  # ... and (p, o) with the queries
    f[:2 * Q][np.all(f[:2 * Q] == np.concatenate([q, q]), axis=1), 2] ^= 1   # never the query itself
    eng = Engine.__new__(Engine)               # filter_csr touches nothing but the context and the device
    eng.lib, eng.device, eng.N, eng.R_ = lib, torch.device("cuda", torch.cuda.current_device()), N, R
    h = C.c_void_p()
    abi.check(lib.kgeb_create(C.byref(h), 0, N, R, 8, 8, 8, 8, abi.DISTMULT, 12.0))
    eng.ctx = h
    fh, ft = ko.build_filters([tuple(int(x) for x in r) for r in f])      # the reference's dict-of-sets (utils.py:201-205), no packing
    for side, code, filt in (("tail", abi.SIDE_TAIL, ft), ("head", abi.SIDE_HEAD, fh)):
        offs, ids = eng.filter_csr(q, f, code)
        want_o, want_i = ko.filter_csr(q, side, filt)
        np.testing.assert_array_equal(offs.cpu().numpy(), np.asarray(want_o, np.int64))
        np.testing.assert_array_equal(ids.cpu().numpy(), np.asarray(want_i, np.int32))
