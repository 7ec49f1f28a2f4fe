"""CPU tests: host-side logic, ABI surface, fail-loud behaviour.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import kge_oracle as ko

import mkb_b200
from mkb_b200 import _native, datasets, models, sampling
from mkb_b200.utils import build_filter_csr


def test_library_exports_every_declared_symbol():
    """include/kge_b200.h <-> libkge_b200.so <-> the ctypes prototypes agree symbol by symbol."""
    header = open(os.path.join(ROOT, "include", "kge_b200.h")).read()
    declared = set(re.findall(r"\b(kge_[a-z_0-9]+)\s*\(", header))
    declared -= {"kge_filter_csr", "kge_tables"}
    assert len(declared) >= 15
    lib = _native.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    assert lib.kge_abi_version() == _native.ABI_VERSION == 2
    assert b"NULL" in lib.kge_strerror(-1)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_native.KgeTables) == 56 and _native.KgeTables.modulus.offset == 48
    assert ctypes.sizeof(_native.KgeFilterCsr) == 32
    assert _native.KgeTables.hidden_dim.offset == 32 and _native.KgeTables.gamma.offset == 40


def test_argument_validation_without_gpu():
    lib = _native.load()
    t = _native.KgeTables(None, None, 10, 2, 4, 0, 1.0, 0.5)
    assert lib.kge_score_fwd(ctypes.byref(t), 0, None, 1, None, 0, None, None) == -1  # NULL tables
    t = _native.KgeTables(16, 16, 10, 2, 4, 9, 1.0, 0.5)
    assert lib.kge_score_fwd(ctypes.byref(t), 0, 16, 1, None, 0, 16, None) == -3  # bad model
    t = _native.KgeTables(16, 16, 10, 2, 4, 0, 1.0, 0.5)
    assert lib.kge_score_fwd(ctypes.byref(t), 7, 16, 1, None, 0, 16, None) == -4  # bad mode
    assert lib.kge_adam_step(None, None, None, None, 4, 1, 0.1, 0.9, 0.999, 1e-8, 0, None) == -1
    assert lib.kge_loss_workspace_bytes(1024) == 3 * 1024 * 4 + 16


def test_model_init_matches_reference_under_seed(eval_doctest):
    """Same parameter creation order and init calls as mkb/models/base.py:86-100 => identical tables
    for a given torch seed (the doctest constructs Dataset(seed=42) first, then the model)."""
    g = eval_doctest
    torch.manual_seed(42)
    train = [(0, 0, 1), (0, 1, 1), (2, 0, 3), (2, 1, 3)]
    entities = {"e0": 0, "e1": 1, "e2": 2, "e3": 3}
    relations = {"r0": 0, "r1": 1}
    ds = datasets.Dataset(train=train, valid=train[:1], test=train[:1], entities=entities, relations=relations,
                          batch_size=2, seed=42, shuffle=False)
    m = models.RotatE(hidden_dim=3, entities=ds.entities, relations=ds.relations, gamma=1)
    np.testing.assert_array_equal(m.entity_embedding.detach().numpy(), g["ent0"])
    np.testing.assert_array_equal(m.relation_embedding.detach().numpy(), g["rel0"])
    assert m.entity_dim == 6 and m.relation_dim == 3 and m.modulus.shape == (1, 1)
    assert [n for n, p in m.named_parameters() if p.requires_grad] == [
        "entity_embedding", "relation_embedding", "modulus"]
    # the batches the doctest's loop saw
    for step, data in enumerate(ds):
        np.testing.assert_array_equal(data["sample"].numpy(), g[f"step{step}/sample"])
        np.testing.assert_allclose(data["weight"].numpy(), g[f"step{step}/weight"], rtol=1e-7)
        assert data["mode"] == str(g[f"step{step}/mode"])


def test_dimensions_and_repr():
    e, r = {i: i for i in range(5)}, {i: i for i in range(2)}
    assert models.TransE(4, e, r, 6).entity_embedding.shape == (5, 4)
    assert models.DistMult(4, e, r, 6).relation_embedding.shape == (2, 4)
    cx = models.ComplEx(4, e, r, 6)
    assert cx.entity_embedding.shape == (5, 8) and cx.relation_embedding.shape == (2, 8)
    assert "ComplEx model" in repr(cx)
    assert abs(cx.embedding_range.item() - 2.0) < 1e-7
    emb = cx.embeddings
    assert set(emb) == {"entities", "relations"} and len(emb["entities"]) == 5
    s, shape = cx.format_sample(torch.zeros(2, 3, 3, dtype=torch.long))
    assert s.shape == (6, 3) and shape == (2, 3)


def test_cpu_call_fails_loudly():
    m = models.TransE(4, {0: 0, 1: 1}, {0: 0}, 3)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.tensor([[0, 0, 1]]))
    if not torch.cuda.is_available():
        ns = sampling.NegativeSampling(2, [(0, 0, 1)], {0: 0, 1: 1}, {0: 0})
        with pytest.raises(RuntimeError, match="CUDA"):
            ns.generate(torch.tensor([[0, 0, 1]]), "tail-batch")


def test_filter_csr_matches_oracle(sampler_cases):
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    for side in ("head", "tail"):
        a = build_filter_csr(triples, int(g["N"]), side)
        b = ko.build_filter_csr(triples, int(g["N"]), side)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    ns = sampling.NegativeSampling(4, triples, range(int(g["N"])), range(int(g["R"])))
    ref_keys = [tuple(int(x) for x in k) for k in g["true_head_keys"]]
    assert set(ns.true_head) == set(ref_keys)
    k0 = ref_keys[0]
    np.testing.assert_array_equal(ns.true_head[k0], g["true_head_members"][: int(g["true_head_sizes"][0])])


def test_dataset_weights_and_iteration(sampler_cases):
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    ds = datasets.Dataset(train=triples, entities={i: i for i in range(60)}, relations={i: i for i in range(4)},
                          batch_size=64, shuffle=True, seed=42)
    np.testing.assert_allclose(ds._weights.numpy(), g["weights"], rtol=1e-7)
    batches = list(ds)
    assert len(batches) == 2 * -(-400 // 64) and len(ds) == int(800 / 64)
    assert [b["mode"] for b in batches[:4]] == ["head-batch", "tail-batch", "head-batch", "tail-batch"]
    assert batches[-1]["sample"].shape == (400 % 64, 3)
    seen = torch.cat([b["sample"] for b in batches[0::2]])
    assert sorted(map(tuple, seen.tolist())) == sorted(triples)  # one epoch visits every triple once
    nxt = next(ds)
    assert nxt["mode"] == "tail-batch" and next(ds)["mode"] == "head-batch"


def test_dataset_builds_mappings_like_reference():
    train = [("mkb", "is_a", "library"), ("github", "is_a", "tool"), ("mkb", "is_on", "github")]
    ds = datasets.Dataset(train=train, batch_size=1, shuffle=False)
    assert ds.entities == {"mkb": 0, "github": 1, "library": 2, "tool": 3}
    assert ds.relations == {"is_a": 0, "is_on": 1}
    assert ds.train == [(0, 0, 2), (1, 0, 3), (0, 1, 1)]


@pytest.mark.parametrize("num_workers", (1, 0))
def test_shuffled_batches_match_the_reference_loader(num_workers):
    """datasets.Dataset(shuffle=True) consumes torch's RNG like the reference's two DataLoaders, so a
    seeded run sees the same batches in the same order (fixture: tests/golden/make_golden.py loader)."""
    from conftest import load_golden

    g = load_golden("loader_order.npz")
    tri = [tuple(int(x) for x in r) for r in g["triples"]]
    torch.manual_seed(123)
    ds = datasets.Dataset(train=tri, entities={i: i for i in range(50)}, relations={i: i for i in range(3)},
                          batch_size=32, shuffle=True, seed=42, num_workers=num_workers)
    for epoch in range(2):
        batches = list(ds)
        assert len(batches) == int(g[f"nw{num_workers}/epoch{epoch}/n"])
        for i, d in enumerate(batches):
            np.testing.assert_array_equal(d["sample"].numpy(), g[f"nw{num_workers}/epoch{epoch}/b{i}"])
            assert d["mode"] == str(g[f"nw{num_workers}/epoch{epoch}/m{i}"])


def test_shard_helpers_round_trip_and_layout():
    """K7 host logic: block-cyclic split/merge, shard sizes, struct layout, argument validation."""
    from mkb_b200 import ops

    assert ctypes.sizeof(_native.KgeShards) == 2 * 8 * _native.MAX_SHARDS + 8
    assert _native.KgeShards.n_shards.offset == 2 * 8 * _native.MAX_SHARDS
    full = torch.arange(11 * 3, dtype=torch.float32).view(11, 3)
    for G in (1, 2, 3, 4, 16):
        shards = ops.split_rows(full, G)
        assert len(shards) == G and all(s.shape == (ops.shard_rows(11, G), 3) for s in shards)
        assert sum(ops.shard_rows(11, G, s) for s in range(G)) == 11
        for e in range(11):  # entity e lives on shard e % G at local row e // G
            assert torch.equal(shards[e % G][e // G], full[e])
        assert torch.equal(ops.merge_rows(shards, 11), full)
    lib = _native.load()
    t = _native.KgeTables(None, None, 10, 2, 4, 0, 1.0, 0.5)
    sh = _native.KgeShards()
    sh.n_shards = 2
    assert lib.kge_score_fwd_sharded(ctypes.byref(t), ctypes.byref(sh), 0, None, 1, None, 0, None, None) == -1
    t.relation = 256  # never dereferenced: the checks below fail first
    assert lib.kge_score_fwd_sharded(ctypes.byref(t), ctypes.byref(sh), 0, None, 1, None, 0, None, None) == -1  # shard ptr NULL
    sh.n_shards = 17
    assert lib.kge_score_fwd_sharded(ctypes.byref(t), ctypes.byref(sh), 0, None, 1, None, 0, None, None) == -2
    sh.n_shards = 2
    t.hidden_dim = 6
    assert lib.kge_fused_bwd_sharded(ctypes.byref(t), ctypes.byref(sh), 0, None, 1, None, 1, None, None, None, None,
                                     None, None) == -6


# --------------------------------------------------------------------------------------------------
# evaluation host logic driven by an oracle-backed stand-in model: everything
# around the kernels — TestDataset / TestDatasetRelation items, the stream API, the position count
# that replaces argsort, relation categories, the detail_eval frame — checked against fixtures made by
# executing the reference.  The stand-in only replaces the CUDA scoring calls.
# --------------------------------------------------------------------------------------------------
class _OracleModel:
    def __init__(self, name, ent, rel, gamma):
        self.name, self.ent, self.rel, self.gamma = name, ent, rel, gamma
        self.entity_embedding = torch.zeros(1)
        self.training = True

    def eval(self):
        self.training = False
        return self

    def train(self):
        self.training = True
        return self

    def __call__(self, sample, negative_sample=None, mode=None):
        s = sample.numpy()
        if s.ndim == 3:
            flat = ko.score(self.name, self.ent, self.rel, s.reshape(-1, 3), gamma=self.gamma)
            return torch.from_numpy(flat.reshape(s.shape[0], s.shape[1]).astype(np.float32))
        neg = None if negative_sample is None else negative_sample.numpy()
        return torch.from_numpy(ko.score(self.name, self.ent, self.rel, s, neg, mode, gamma=self.gamma).astype(np.float32))


@pytest.fixture(scope="module")
def next_rows():
    from conftest import load_golden

    return load_golden("next_rows.npz")


def _next_setup(g, name):
    from mkb_b200 import evaluation

    entities = {f"e{i}": i for i in range(40)}
    relations = {f"r{i}": i for i in range(4)}
    true = [tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]]
    test = [tuple(int(x) for x in r) for r in g["test"]]
    ev = evaluation.Evaluation(entities=entities, relations=relations, batch_size=4, true_triples=true)
    m = _OracleModel(name, g[f"{name}/ent"].astype(np.float64), g[f"{name}/rel"].astype(np.float64), float(g[f"{name}/gamma"]))
    return ev, m, test, true, entities, relations


def test_test_datasets_match_reference_items(eval_cases, next_rows):
    g = eval_cases
    true = [tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]]
    test = [tuple(int(x) for x in r) for r in g["test"]]
    ents, rels = {f"e{i}": i for i in range(50)}, {f"r{i}": i for i in range(3)}
    for mode in ("head-batch", "tail-batch"):
        td = datasets.TestDataset(triples=test, true_triples=true, entities=ents, relations=rels, mode=mode)
        assert len(td) == len(test)
        items = [td[i] for i in range(len(test))]
        np.testing.assert_array_equal(np.stack([c.numpy() for _, c, _, _ in items]), g[f"{mode}/cand"])
        np.testing.assert_array_equal(np.stack([b.numpy() for _, _, b, _ in items]), g[f"{mode}/bias"])
        batch = td.collate_fn(items[:3])
        assert batch["sample"].shape == (3, 3) and batch["negative_sample"].shape == (3, 50) and batch["mode"] == mode
        assert batch["filter_bias"].dtype == torch.float32 and batch["negative_sample"].dtype == torch.int64
    g = next_rows
    ev, _, test, true, ents, rels = _next_setup(g, "RotatE")
    tdr = datasets.TestDatasetRelation(triples=test, true_triples=true, entities=ents, relations=rels)
    items = [tdr[i] for i in range(len(test))]
    np.testing.assert_array_equal(np.stack([c.numpy() for _, c, _, _ in items]), g["rel/cand"])
    np.testing.assert_array_equal(np.stack([b.numpy() for _, _, b, _ in items]), g["rel/bias"])
    assert items[0][3] == "relation-batch" and (g["rel/bias"] != 0).sum() > 0
    stream = ev.get_relation_stream(test)
    assert len(stream) == -(-len(test) // 4) and sum(b["sample"].shape[0] for b in stream) == len(test)


@pytest.mark.parametrize("name", ("TransE", "DistMult", "ComplEx", "RotatE"))
def test_stream_api_and_relation_metrics_match_reference(next_rows, name):
    from mkb_b200.evaluation.evaluation import _Mean

    g = next_rows
    ev, m, test, *_ = _next_setup(g, name)
    keys = ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")
    got = ev.eval_relations(model=m, dataset=test)
    np.testing.assert_allclose([got[f"{k}_relations"] for k in keys], g[f"{name}/rel_metrics"], atol=1e-4)
    metrics = ev.compute_score(model=m, test_set=ev.get_relation_stream(test), metrics={k: _Mean() for k in keys},
                               device="cpu")
    np.testing.assert_allclose([round(metrics[k].get(), 4) for k in keys], g[f"{name}/rel_metrics"], atol=1e-4)
    assert m.training


@pytest.mark.parametrize("name", ("TransE", "RotatE"))
def test_types_relations_and_detail_frame_match_reference(next_rows, name, monkeypatch):
    g = next_rows
    ev, m, test, true, *_ = _next_setup(g, name)
    types = ev.types_relations(model=m, dataset=test)
    assert [types[f"r{i}"] for i in range(4)] == list(g[f"{name}/types"])
    hc, tc = ko.build_filter_csr(true, 40, "head"), ko.build_filter_csr(true, 40, "tail")

    def oracle_ranks(model, dataset, mode):
        r, _ = ko.rank_all(name, m.ent, m.rel, np.asarray(dataset, dtype=np.int64), mode, hc, tc, gamma=m.gamma)
        return torch.from_numpy(np.asarray(r, dtype=np.int64))

    monkeypatch.setattr(ev, "ranks", oracle_ranks)
    frame = ev.detail_eval(model=m, dataset=test)
    assert ["|".join(c) for c in frame.columns] == list(g[f"{name}/detail_cols"])
    assert list(frame.index) == list(g[f"{name}/detail_index"])
    np.testing.assert_allclose(frame.to_numpy(dtype=np.float64), g[f"{name}/detail"], atol=1e-4)
    # the reference's own driver for the same table
    from mkb_b200.evaluation.evaluation import _TYPES, _new_metrics

    by_id = {int(k[1:]): v for k, v in types.items()}
    metrics = {mode: {t: _new_metrics() for t in _TYPES} for mode in ("head-batch", "tail-batch")}
    for stream in ev.get_entity_stream(test):
        metrics = ev.compute_detailled_score(model=m, test_set=stream, metrics=metrics, types_relations=by_id,
                                             device="cpu")
    import pandas as pd

    frame2 = ev._detail_frame(pd, metrics, by_id)
    np.testing.assert_allclose(frame2.to_numpy(dtype=np.float64), g[f"{name}/detail"], atol=1e-4)


def test_classification_threshold_matches_sklearn_and_reference(next_rows):
    """evaluation.find_threshold / accuracy (mkb/evaluation/classif.py): the numpy threshold search equals
    sklearn's roc_curve + argmax on random scores with ties, and the reference's outputs on the golden graph."""
    from sklearn import metrics

    from mkb_b200 import evaluation
    from mkb_b200.evaluation.classif import _accuracy, best_threshold

    rng = np.random.RandomState(0)
    for trial in range(200):
        n = rng.randint(2, 40)
        y = np.where(rng.rand(n) < 0.5, 1, -1)
        if abs(y.sum()) == n:
            y[0] = -y[0]
        score = np.round(rng.normal(size=n), 1 if trial % 2 else 3).astype(np.float32)
        fpr, tpr, thr = metrics.roc_curve(y_true=y, y_score=score)
        assert best_threshold(y, score) == thr[np.argmax(tpr - fpr)], (y, score)
    assert best_threshold([-1, -1, -1, -1, -1, 1, 1, 1, 1, 1], [1, 2, 3, 4, 5, 5, 6, 7, 8, 9]) == 6  # classif.py:95-107
    assert _accuracy(np.array([1, 2, 3, 4, 5, 5, 6, 7, 8, 9]), np.array([-1] * 5 + [1] * 5), 5) == 0.9  # :130-133
    g = next_rows
    X = [tuple(int(v) for v in r) for r in g["clf/X"]]
    y = [int(v) for v in g["clf/y"]]
    for name in ("TransE", "RotatE"):
        m = _OracleModel(name, g[f"{name}/ent"].astype(np.float64), g[f"{name}/rel"].astype(np.float64),
                         float(g[f"{name}/gamma"]))
        thr = evaluation.find_threshold(model=m, X=X, y=y, batch_size=8, device="cpu")
        assert abs(thr - float(g[f"{name}/threshold"])) <= 1e-4 * abs(thr)
        acc = evaluation.accuracy(model=m, X=X, y=y, threshold=float(g[f"{name}/threshold"]), batch_size=8, device="cpu")
        assert abs(acc - float(g[f"{name}/accuracy"])) <= 1.0 / len(X) + 1e-9


# ---------------------------------------------------------------------------------------------------
# bundled-dataset wrappers (they read mkb's own data directory; nothing is redistributed here)
# ---------------------------------------------------------------------------------------------------
_REF_DATA = "/root/reference/mkb/datasets"


@pytest.mark.skipif(not os.path.exists(os.path.join(_REF_DATA, "wn18rr", "train.csv")), reason="mkb data not present")
def test_bundled_dataset_wrappers(monkeypatch):
    from mkb_b200 import datasets

    monkeypatch.setenv("MKB_DATASETS", _REF_DATA)
    ds = datasets.Wn18rr(batch_size=256, shuffle=False, seed=42)
    assert (len(ds.entities), len(ds.relations)) == (40943, 11)  # mkb/datasets/wn18rr.py:44-49
    assert (len(ds.train), len(ds.valid), len(ds.test)) == (86835, 3034, 3134)
    assert len(ds.classification_valid["X"]) == len(ds.classification_valid["y"]) == 2 * len(ds.valid)
    assert ds.name == "Wn18rr" and "Train triples  86835" in repr(ds)
    b = next(iter(ds))
    assert b["sample"].shape == (256, 3) and b["mode"] == "head-batch"
    fb = datasets.Fb15k237(batch_size=1024, path=os.path.join(_REF_DATA, "fb15k237"))
    assert (len(fb.entities), len(fb.relations), len(fb.train)) == (14541, 237, 272115)
    monkeypatch.delenv("MKB_DATASETS")
    with pytest.raises(FileNotFoundError, match="MKB_DATASETS"):
        datasets.Yago310(batch_size=8, path="/nonexistent")


# ---------------------------------------------------------------------------------------------------
# distillation.TopKSampling host logic (candidate maps, batching / chunking, RNG consumption): the kernels
# are replaced by the oracle scorer and a stable argsort, the outputs are the reference's own
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ("RotatE", "TransE", "ComplEx"))
def test_top_k_sampling_host_logic_matches_reference(name, monkeypatch):
    from conftest import load_golden
    from mkb_b200 import distillation, ops

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    teacher = _OracleModel(name, d[f"{name}/ent"].astype(np.float64), d[f"{name}/rel"].astype(np.float64), 6.0)
    teacher.entity_embedding = torch.zeros(1)  # only its .device is read
    monkeypatch.setattr(ops, "topk_rows", lambda s, k: torch.from_numpy(
        np.argsort(-s.numpy().astype(np.float64), axis=1, kind="stable")[:, :k].copy()))
    sample = torch.from_numpy(d["sample"])
    for tag, (ke, kr, ne, nr) in {"a": (4, 2, 2, 1), "b": (7, 3, 0, 0)}.items():
        smp = distillation.TopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                        student_relations=rel_s, batch_size_entity=ke, batch_size_relation=kr,
                                        n_random_entities=ne, n_random_relations=nr, seed=42)
        for call in range(2):
            got = smp.get(sample=sample, teacher=teacher, max_ids_per_call=100 if call else 1 << 24)
            for k, t in zip(("ht", "rt", "tt", "hs", "rs", "ts"), got):
                np.testing.assert_array_equal(t.numpy(), d[f"{name}/{tag}/{call}/{k}"], err_msg=f"{tag}/{call}/{k}")
        assert teacher.training


def test_batch_accessor_and_transe_top_k():
    """BaseModel.batch / TransE._top_k (mkb/models/base.py:153-207, transe.py:78-84): pure row indexing, usable
    without CUDA; shapes and values as the reference returns them."""
    ents, rels = {f"e{i}": i for i in range(9)}, {f"r{i}": i for i in range(3)}
    m = models.ComplEx(hidden_dim=4, entities=ents, relations=rels, gamma=6)
    E, R = m.entity_embedding.detach().numpy(), m.relation_embedding.detach().numpy()
    s, n = torch.tensor([[0, 1, 2], [3, 0, 4]]), torch.tensor([[1, 2, 3], [5, 6, 7]])
    h, r, t, shape = m.batch(s)
    assert tuple(shape) == (2, 1) and h.shape == (2, 1, 8) and r.shape == (2, 1, 8)
    np.testing.assert_array_equal(h[:, 0].detach().numpy(), E[[0, 3]])
    np.testing.assert_array_equal(r[:, 0].detach().numpy(), R[[1, 0]])
    h, r, t, shape = m.batch(s, n, "head-batch")
    assert tuple(shape) == (2, 3) and h.shape == (2, 3, 8) and t.shape == (2, 1, 8)
    np.testing.assert_array_equal(h.detach().numpy(), E[n.numpy()])
    h, r, t, shape = m.batch(s, n, "tail-batch")
    np.testing.assert_array_equal(t.detach().numpy(), E[n.numpy()])
    np.testing.assert_array_equal(h[:, 0].detach().numpy(), E[[0, 3]])
    h, r, t, shape = m.batch(torch.stack([s, s, s]))
    assert tuple(shape) == (3, 2) and h.shape == (6, 1, 8)
    te = models.TransE(hidden_dim=4, entities=ents, relations=rels, gamma=6)
    eh, er, et = te._top_k(s)
    E, R = te.entity_embedding.detach().numpy(), te.relation_embedding.detach().numpy()
    np.testing.assert_allclose(eh[:, 0].detach().numpy(), E[[2, 4]] - R[[1, 0]])
    np.testing.assert_allclose(er[:, 0].detach().numpy(), E[[2, 4]] - E[[0, 3]])
    np.testing.assert_allclose(et[:, 0].detach().numpy(), E[[0, 3]] + R[[1, 0]])


def test_pipeline_recognises_only_a_stock_torch_adam():
    """Only the optimizer of the reference's quick-start (README.md:123-126) is taken over by the device step."""
    from mkb_b200 import compose

    p = [torch.nn.Parameter(torch.zeros(3))]
    ok = compose.Pipeline._plain_torch_adam
    assert ok(torch.optim.Adam(p, lr=5e-5))
    assert not ok(torch.optim.Adam(p, lr=5e-5, weight_decay=1e-3))
    assert not ok(torch.optim.Adam(p, lr=5e-5, amsgrad=True))
    assert not ok(torch.optim.Adam(p, lr=torch.tensor(5e-5)))
    assert not ok(torch.optim.AdamW(p, lr=5e-5))
    assert not ok(torch.optim.SGD(p, lr=0.1))
    assert not ok(torch.optim.Adam([{"params": p}, {"params": [torch.nn.Parameter(torch.zeros(2))]}], lr=1e-3))
    # CPU model: never adopted, the generic route raises from the kernels ("CUDA only"), not from the trainer
    pipe = compose.Pipeline(epochs=1)
    assert pipe.adopt_torch_adam and compose.Pipeline(epochs=1, adopt_torch_adam=False).adopt_torch_adam is False


def test_top_k_sampling_transe_host_logic(monkeypatch):
    """TopKSamplingTransE on CPU tensors (row accessor + torch.cdist; only the top-k kernel is replaced by a stable
    argsort): exact L2 neighbours of t - r, t - h, h + r among the shared rows, as faiss.IndexFlatL2 specifies."""
    from conftest import load_golden
    from mkb_b200 import distillation, ops

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    teacher = models.TransE(hidden_dim=8, entities=ent_t, relations=rel_t, gamma=6)
    teacher._set_params(torch.from_numpy(d["TransE/ent"].copy()), torch.from_numpy(d["TransE/rel"].copy()))
    monkeypatch.setattr(ops, "topk_rows", lambda s, k: torch.from_numpy(
        np.argsort(-s.numpy().astype(np.float64), axis=1, kind="stable")[:, :k].copy()))
    smp = distillation.TopKSamplingTransE(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                          student_relations=rel_s, batch_size_entity=5, batch_size_relation=2,
                                          n_random_entities=0, n_random_relations=0, seed=1, teacher=teacher)
    sample = d["sample"]
    ht, rt, tt, hs, rs, ts = (x.numpy() for x in smp.get(sample=torch.from_numpy(sample), teacher=teacher))
    E, R = d["TransE/ent"].astype(np.float64), d["TransE/rel"].astype(np.float64)
    se = np.array([i for e, i in ent_t.items() if e in ent_s])
    sr = np.array([i for r, i in rel_t.items() if r in rel_s])
    h, r, t = sample[:, 0], sample[:, 1], sample[:, 2]

    def near(q, rows, ids, k):
        return ids[np.argsort(((q[:, None, :] - rows[None, :, :]) ** 2).sum(-1), axis=1, kind="stable")[:, :k]]

    np.testing.assert_array_equal(ht, near(E[t] - R[r], E[se], se, 5))
    np.testing.assert_array_equal(rt, near(E[t] - E[h], R[sr], sr, 2))
    np.testing.assert_array_equal(tt, near(E[h] + R[r], E[se], se, 5))
    to_s = {i: ent_s[e] for e, i in ent_t.items() if e in ent_s}
    assert all(to_s[a] == b for a, b in zip(tt.ravel(), ts.ravel()))


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under mkb_b200/ may import or execute it (a product path through
    the oracle or any CPU fallback would void every parity claim)."""
    pkg = os.path.join(ROOT, "mkb_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "kge_oracle" in text or "torch_port" in text:
                    offenders.append(os.path.join(dirpath, name))
    assert not offenders, offenders


def test_top_k_sampling_host_logic_reproduces_the_reference_doctest(monkeypatch):
    """The tensors printed at mkb/distillation/top_k_sampling.py:389-413, from the oracle scorer + a stable argsort
    through this package's TopKSampling (candidate maps, chunking, RNG stream)."""
    from conftest import load_golden
    from mkb_b200 import distillation, ops

    d = load_golden("distill_doctests.npz")

    def label_map(prefix):
        return {str(k): int(v) for k, v in zip(d[f"{prefix}/labels"], d[f"{prefix}/ids"])}

    teacher = _OracleModel("RotatE", d["topk/ent"].astype(np.float64), d["topk/rel"].astype(np.float64), 3.0)
    monkeypatch.setattr(ops, "topk_rows", lambda s, k: torch.from_numpy(
        np.argsort(-s.numpy().astype(np.float64), axis=1, kind="stable")[:, :k].copy()))
    smp = distillation.TopKSampling(teacher_relations=label_map("topk/rel_t"), teacher_entities=label_map("topk/ent_t"),
                                    student_entities=label_map("topk/ent_s"), student_relations=label_map("topk/rel_s"),
                                    batch_size_entity=4, batch_size_relation=1, n_random_entities=1,
                                    n_random_relations=0, seed=42)
    ht, rt, tt, hs, rs, ts = (t.tolist() for t in smp.get(sample=torch.tensor([[0, 0, 266], [1, 1, 56]]), teacher=teacher))
    assert ht == [[197, 50, 75, 176, 30], [10, 240, 251, 3, 30]]
    assert tt == [[269, 210, 270, 261, 30], [120, 160, 212, 244, 30]]
    assert hs == [[186, 47, 70, 166, 28], [10, 229, 240, 3, 28]]
    assert ts == [[269, 198, 270, 256, 28], [111, 149, 201, 234, 28]]
    assert rt == rs == [[0], [1]]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU operator sequence on the host cores, no GPU involved):
    one JSON line with the own arm's metric / unit / config plus impl, cpu_baseline and a zero-copy e2e object."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1",
                        "--steps", "2", "--warmup", "1", "--cpu-budget", "3"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"].startswith("training triples/sec")
    assert line["unit"] == "triples/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["dtype"] == "f32"
    assert "Wn18rr" in line["config"]["workload"] and "TransE" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    # "reference": the unmodified mkb of baseline/_ref ran (installed by baseline/install_ref.sh); "port": the
    # oracle's restatement of its operator sequence (fallback when baseline/_ref is absent)
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "mkb"))
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert ("unmodified mkb" in cb["sample"]) == have_ref
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "positives" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit 0 without work (the driver launches the arm under torchrun for N > 1)
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""


@pytest.mark.parametrize("variant", (("RotatE", "scatter", "0", "independent", "0"),  # the default step
                                     ("ComplEx", "scatter", "0", "reference", "1")))  # pooled GEMM flow
def test_bench_own_arm_line_contract_on_the_emulation(variant):
    """bench.py's own arm end to end (DeviceTrainer loop, Pipeline.learn e2e arm, JSON line) on the CPU emulation
    of the kernels with a toy config — tests/emu/bench_dryrun.py; timings are fake, the plumbing and the line are
    real: every key of the bench contract is present and the launch count is the step's kernel count."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "bench_dryrun.py"), *variant],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["unit"] == "triples/s" and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["data"] == "synthetic" and line["dtype"] == "f32" and line["n_gpus"] == 1 and line["warmup"] >= 3
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["unit"] == "GB/s"
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] == 8 * 3 * 8 + 8 * 4 and line["e2e"]["d2h_bytes_per_step"] == 16
    assert "error" not in line["e2e"] and line["e2e"]["value"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(line["clocks"])
    assert "workload" in line["config"] and "model" not in line["config"]
    if variant[1] == "scatter" and variant[4] == "0":
        assert line["gpu_launches"] == 5 * line["steps"]  # sampler, fused fwd, fused bwd, adam(entity), adam(relation)
    else:
        assert line["gpu_launches"] > 0


def test_graft_entry_smoke_on_the_emulation():
    """__graft_entry__.smoke() (one fused training step per model and mode + filtered ranking, checked against the
    oracle) through the unchanged package on the emulated kernels — tests/emu/smoke_dryrun.py."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "smoke_dryrun.py")], capture_output=True,
                       text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_pipeline_rank_slices_partition_the_global_batch():
    """compose.Pipeline._rank_slice (multi-rank learn): the ranks' blocks cover every row of the dataset's batch
    exactly once; short blocks are padded with weight-0 rows so that every rank keeps the same local batch size."""
    import types

    import torch

    from mkb_b200.compose import Pipeline

    for B, world in ((12, 4), (10, 4), (3, 4), (7, 2), (1024, 8), (5, 1)):
        sample = torch.arange(B * 3).view(B, 3)
        weight = torch.arange(1, B + 1, dtype=torch.float32)
        per = -(-B // world)
        seen, wsum = [], 0.0
        for rank in range(world):
            s, w = Pipeline._rank_slice(sample, weight, types.SimpleNamespace(world=world, rank=rank))
            assert s.shape == (per, 3) and w.shape == (per,) and s.is_contiguous()
            real = w > 0
            seen += s[real][:, 0].tolist()
            wsum += float(w.sum())
            assert torch.equal(s[~real], sample[:1].expand(int((~real).sum()), 3))  # padding = row 0 at weight 0
        assert sorted(seen) == sample[:, 0].tolist()
        assert wsum == float(weight.sum())
