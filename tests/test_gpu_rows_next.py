"""GPU parity for SURVEY §8(f) rows 3-4 — the other consumers of the score / rank kernels — against
fixtures produced by executing the reference (tests/golden/next_rows.npz, make_golden.py gen_next_rows):
relation prediction, the reference's stream-based compute_score, relation categories + detail_eval,
utils.TopK / make_prediction on the exact top-k kernel, the KL-divergence distillation loss."""
import numpy as np
import pytest
import torch

from conftest import MODELS, load_golden, score_tol

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from mkb_b200 import evaluation, losses, models, ops, utils

from conftest import DEV  # "cuda" (or "cpu" under the KGE_TEST_EMU developer shim)
N_ENT, N_REL = 40, 4
KEYS = ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")


@pytest.fixture(scope="module")
def g():
    return load_golden("next_rows.npz")


def _setup(g, name):
    entities = {f"e{i}": i for i in range(N_ENT)}
    relations = {f"r{i}": i for i in range(N_REL)}
    m = getattr(models, name)(hidden_dim=8, entities=entities, relations=relations, gamma=float(g[f"{name}/gamma"]))
    m._set_params(torch.from_numpy(g[f"{name}/ent"].copy()), torch.from_numpy(g[f"{name}/rel"].copy()))
    m = m.to(DEV)
    true = [tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]]
    test = [tuple(int(x) for x in r) for r in g["test"]]
    ev = evaluation.Evaluation(entities=entities, relations=relations, batch_size=4, true_triples=true)
    return m, ev, test, entities, relations


@pytest.mark.parametrize("model", MODELS)
def test_relation_prediction_and_stream_api(g, model):
    m, ev, test, _, _ = _setup(g, model)
    got = ev.eval_relations(model=m, dataset=test)
    np.testing.assert_allclose([got[f"{k}_relations"] for k in KEYS], g[f"{model}/rel_metrics"], atol=2e-3)
    # the reference's own driver: compute_score over get_relation_stream / get_entity_stream
    metrics = {k: evaluation.evaluation._Mean() for k in KEYS}
    metrics = ev.compute_score(model=m, test_set=ev.get_relation_stream(test), metrics=metrics, device=DEV)
    np.testing.assert_allclose([round(metrics[k].get(), 4) for k in KEYS], g[f"{model}/rel_metrics"], atol=2e-3)
    metrics = {k: evaluation.evaluation._Mean() for k in KEYS}
    for stream in ev.get_entity_stream(test):
        metrics = ev.compute_score(model=m, test_set=stream, metrics=metrics, device=DEV)
    fast = ev.eval(model=m, dataset=test)
    np.testing.assert_allclose([round(metrics[k].get(), 4) for k in KEYS], [fast[k] for k in KEYS], atol=2e-3)
    assert m.training  # compute_score restores train mode


@pytest.mark.parametrize("model", MODELS)
def test_detail_eval_matches_reference_frame(g, model):
    m, ev, test, _, _ = _setup(g, model)
    types = ev.types_relations(model=m, dataset=test)
    assert [types[f"r{i}"] for i in range(N_REL)] == list(g[f"{model}/types"]) == ["1_1", "1_M", "M_1", "M_M"]
    frame = ev.detail_eval(model=m, dataset=test)
    assert ["|".join(c) for c in frame.columns] == list(g[f"{model}/detail_cols"])
    assert list(frame.index) == list(g[f"{model}/detail_index"])
    ref = g[f"{model}/detail"]
    got = frame.to_numpy(dtype=np.float64)
    assert got.shape == ref.shape
    # MR columns are means of integer ranks over a handful of queries: one contested rank moves them by < 1
    assert np.all(np.abs(got - ref) <= 2e-3 + 0.02 * np.abs(ref)), (got, ref)


@pytest.mark.parametrize("model", MODELS)
def test_topk_and_make_prediction_match_reference(g, model):
    m, _, test, entities, relations = _setup(g, model)
    topk = utils.TopK(entities=entities, relations=relations)
    assert [entities[e] for e in topk.top_heads(k=7, model=m, relation="r1", tail="e3")] == list(g[f"{model}/top_heads"])
    assert [entities[e] for e in topk.top_tails(k=7, model=m, head="e5", relation=2)] == list(g[f"{model}/top_tails"])
    assert [relations[r] for r in topk.top_relations(k=2, model=m, head=4, tail="e9")] == list(g[f"{model}/top_relations"])
    assert len(topk.top_relations(k=10, model=m, head=4, tail="e9")) == N_REL  # k larger than the candidate list
    pred = utils.make_prediction(model=m, dataset=test[:9], batch_size=4, num_workers=0, device=DEV)
    ref = g[f"{model}/prediction"]
    assert pred.shape == (9,)
    assert np.all(np.abs(pred.cpu().numpy() - ref) <= score_tol(ref))


@pytest.mark.parametrize("T", (1, 3))
def test_kl_divergence_matches_reference(g, T):
    s = torch.from_numpy(g[f"kl_T{T}/f32/student"].copy()).to(DEV).requires_grad_()
    t = torch.from_numpy(g[f"kl_T{T}/f32/teacher"].copy()).to(DEV).requires_grad_()
    loss = losses.KlDivergence()(s, t, T=T)
    (2.0 * loss).backward()
    ref = float(g[f"kl_T{T}/f32/loss"])
    assert loss.dim() == 0 and abs(loss.item() - ref) <= 1e-5 * abs(ref)
    gref = 2.0 * g[f"kl_T{T}/f32/grad"].astype(np.float64)
    assert np.abs(s.grad.cpu().numpy() - gref).max() <= 1e-4 * np.abs(gref).max()
    # teacher gradient against torch autograd on the same device
    s2, t2 = s.detach().clone().requires_grad_(), t.detach().clone().requires_grad_()
    ref_loss = torch.mean(torch.nn.functional.kl_div(torch.log_softmax(s2 / T, 1), torch.softmax(t2 / T, 1),
                                                     reduction="none"))
    (2.0 * ref_loss).backward()
    assert (t.grad - t2.grad).abs().max().item() <= 1e-4 * t2.grad.abs().max().item()
    # teacher without grad: no teacher gradient is computed
    s3 = s.detach().clone().requires_grad_()
    losses.KlDivergence()(s3, t.detach(), T=T).backward()
    assert (s3.grad * 2.0 - s.grad).abs().max().item() <= 1e-6


@pytest.mark.parametrize("rows,cols,k", [(1, 40943, 10), (64, 40943, 100), (7, 14541, 1024), (3, 5, 5), (5, 123182, 1)])
def test_topk_rows_equals_stable_argsort_full_size(rows, cols, k):
    gen = torch.Generator(device=DEV).manual_seed(rows * 1000 + k)
    x = torch.randn(rows, cols, device=DEV, generator=gen)
    x[0] = torch.round(x[0] * 4) / 4  # heavy ties in one row
    idx, val = ops.topk_rows(x, k, return_values=True)
    ref = torch.argsort(x, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(idx, ref)
    assert torch.equal(val, x.gather(1, ref))
    # a strided view (row stride > cols) and a 1-D input
    wide = torch.randn(rows, cols + 8, device=DEV, generator=gen)
    assert torch.equal(ops.topk_rows(wide[:, :cols], k),
                       torch.argsort(wide[:, :cols], dim=1, descending=True, stable=True)[:, :k])
    assert torch.equal(ops.topk_rows(x[0], k)[0], ref[0])
    with pytest.raises(ops.N.KgeError):
        ops.topk_rows(x, cols + 1)


# --------------------------------------------------------------------------------------------------
# pRotatE (mkb/models/protate.py) on the shared kernel template
# --------------------------------------------------------------------------------------------------
def _protate(g, k, D):
    entities = {f"e{i}": i for i in range(N_ENT)}
    relations = {f"r{i}": i for i in range(N_REL)}
    m = models.pRotatE(hidden_dim=D, entities=entities, relations=relations, gamma=9.0)
    m._set_params(torch.from_numpy(g[f"{k}/ent"].copy()), torch.from_numpy(g[f"{k}/rel"].copy()),
                  modulus=torch.from_numpy(g[f"{k}/modulus"].copy()))
    return m.to(DEV)


@pytest.mark.parametrize("fused", (False, True))
@pytest.mark.parametrize("D", (8, 5))
@pytest.mark.parametrize("mode", ("tail-batch", "head-batch"))
def test_protate_matches_reference(g, D, mode, fused):
    k = f"pRotatE_D{D}_{mode}"
    m = _protate(g, k, D)
    s = torch.from_numpy(g[f"{k}/sample"]).to(DEV)
    n = torch.from_numpy(g[f"{k}/neg"]).to(DEV)
    w = torch.from_numpy(g[f"{k}/weight"]).to(DEV)
    if fused:
        loss, ps, ns = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, s, n, w, mode, 0.5,
                                                  return_scores=True, modulus=m.modulus)
    else:
        ps, ns = m(s), m(s, n, mode)
        loss = losses.Adversarial(alpha=0.5)(ps, ns, w)
    for got, name in ((ps, "pos"), (ns, "neg_score")):
        ref = g[f"{k}/f32/{name}"]
        assert np.all(np.abs(got.detach().cpu().numpy() - ref) <= score_tol(ref))
        ref = g[f"{k}/f64/{name}"]
        assert np.all(np.abs(got.detach().cpu().numpy() - ref) <= score_tol(ref))
    loss.backward()
    ref = float(g[f"{k}/f64/loss"])
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    for p, name in ((m.entity_embedding, "grad_ent"), (m.relation_embedding, "grad_rel"), (m.modulus, "grad_modulus")):
        ref = g[f"{k}/f64/{name}"]
        assert p.grad is not None and p.grad.shape == ref.shape
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max(), name
    s3 = m(torch.stack([s[:4], s[2:6]]))
    ref = g[f"{k}/f32/score3d"]
    assert s3.shape == ref.shape and np.all(np.abs(s3.detach().cpu().numpy() - ref) <= score_tol(ref))


def test_protate_device_trainer_and_ranks():
    """DeviceTrainer (sampler -> fused fwd -> fused bwd -> modulus grad -> Adam on tables AND modulus)
    tracks the autograd + DenseAdam route; filtered ranks match the oracle."""
    from mkb_b200 import optim, sampling
    from mkb_b200.compose import DeviceTrainer
    from oracle import kge_oracle as ko

    Nn, R, D, B, K, gamma = 300, 5, 32, 24, 16, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=3000), rng.randint(R, size=3000), rng.randint(Nn, size=3000)], 1), axis=0)
    w_all = torch.from_numpy(rng.uniform(0.1, 0.5, len(tri)).astype(np.float32)).to(DEV)
    T = torch.from_numpy(tri).to(DEV)
    runs = []
    for device_loop in (True, False):
        torch.manual_seed(2)
        m = models.pRotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                           gamma=gamma).to(DEV)
        with torch.no_grad():
            m.entity_embedding.mul_(3.0)
        ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=7)
        if device_loop:
            tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B)
        else:
            opt = optim.DenseAdam(filter(lambda p: p.requires_grad, m.parameters()), lr=1e-3)
        for step in range(5):
            idx = torch.arange(step * B, (step + 1) * B, device=DEV)
            mode = "head-batch" if step % 2 == 0 else "tail-batch"
            if device_loop:
                tr.step(T[idx], w_all[idx], mode)
                last = tr.loss()
            else:
                neg = ns.generate(T[idx], mode)
                loss = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, T[idx], neg,
                                                  w_all[idx], mode, 0.5, modulus=m.modulus)
                opt.zero_grad()
                loss.backward()
                opt.step()
                last = loss.item()
        runs.append((m, last))
    (m0, l0), (m1, l1) = runs
    assert abs(l0 - l1) <= 1e-4 * abs(l1)
    assert abs(m0.modulus.item() - m1.modulus.item()) <= 1e-5
    assert m0.modulus.item() != pytest.approx(0.5 * m0.embedding_range.item(), abs=1e-6)  # it trained
    upd = (m1.entity_embedding - m0.entity_embedding).abs()
    assert (upd > 3e-4).float().mean().item() < 1e-3
    # ranks
    ev = evaluation.Evaluation(entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, batch_size=8,
                               true_triples=[tuple(map(int, r)) for r in tri])
    q = [tuple(map(int, r)) for r in tri[:60]]
    hc, tc = ko.build_filter_csr(tri, Nn, "head"), ko.build_filter_csr(tri, Nn, "tail")
    ent, rel = m0.entity_embedding.detach().cpu().numpy(), m0.relation_embedding.detach().cpu().numpy()
    for mode in ("head-batch", "tail-batch"):
        ref, contested = ko.rank_all("pRotatE", ent, rel, np.array(q), mode, hc, tc, gamma=gamma, tie_margin=2e-5,
                                     modulus=m0.modulus.item())
        got = ev.ranks(m0, q, mode).cpu().numpy()
        assert np.all(np.abs(got - ref) <= contested), (got, ref)


def test_pipeline_falls_back_to_three_calls_when_k_exceeds_the_fused_kernel(capsys):
    """K beyond one CTA's shared memory: the fused kernel reports KGE_E_UNSUPPORTED and Pipeline.learn
    continues on the three-call route (model(sample); model(sample, neg, mode); loss)."""
    from mkb_b200 import compose, datasets, optim, sampling

    Nn, R, D, K = 120, 3, 8, 52000
    rng = np.random.RandomState(0)
    tri = [tuple(map(int, r)) for r in np.unique(np.stack([rng.randint(Nn, size=4), rng.randint(R, size=4),
                                                          rng.randint(Nn, size=4)], 1), axis=0)]
    ents, rels = {i: i for i in range(Nn)}, {i: i for i in range(R)}
    torch.manual_seed(0)
    m = models.TransE(hidden_dim=D, entities=ents, relations=rels, gamma=6.0).to(DEV)
    before = m.entity_embedding.detach().clone()
    ds = datasets.Dataset(train=tri, entities=ents, relations=rels, batch_size=4, seed=1)
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=ents, relations=rels, seed=1)
    with pytest.raises(ops.N.KgeError) as err:
        s = torch.tensor(tri[:2]).to(DEV)
        ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, s, ns.generate(s, "tail-batch"),
                                   torch.ones(2, device=DEV), "tail-batch")
    assert err.value.code == ops.N.E_UNSUPPORTED
    pipe = compose.Pipeline(epochs=1, device=DEV)
    pipe.learn(model=m, dataset=ds, sampling=ns, optimizer=optim.DenseAdam(m.parameters(), lr=1e-3),
               loss=losses.Adversarial(0.5))
    assert np.isfinite(pipe.metric_loss.get()) and not torch.equal(before, m.entity_embedding.detach())


@pytest.mark.parametrize("model", MODELS)
def test_triplet_classification_matches_reference(g, model):
    """evaluation.find_threshold / accuracy (mkb/evaluation/classif.py) on the positives kernel."""
    m, _, _, _, _ = _setup(g, model)
    X = [tuple(int(v) for v in r) for r in g["clf/X"]]
    y = [int(v) for v in g["clf/y"]]
    thr = evaluation.find_threshold(model=m, X=X, y=y, batch_size=8, device=DEV)
    ref = float(g[f"{model}/threshold"])
    assert abs(thr - ref) <= 1e-4 * max(abs(ref), 1.0)
    acc = evaluation.accuracy(model=m, X=X, y=y, threshold=ref, batch_size=8, device=DEV)
    assert abs(acc - float(g[f"{model}/accuracy"])) <= 1.0 / len(X) + 1e-9


# --------------------------------------------------------------------------------------------------
# distillation.TopKSampling: score all shared candidates with the teacher, keep the k best, append random ones
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ("RotatE", "TransE", "ComplEx"))
def test_top_k_sampling_matches_reference(model):
    from mkb_b200 import distillation

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    teacher = getattr(models, model)(hidden_dim=8, entities=ent_t, relations=rel_t, gamma=6)
    teacher._set_params(torch.from_numpy(d[f"{model}/ent"].copy()), torch.from_numpy(d[f"{model}/rel"].copy()))
    teacher = teacher.to(DEV)
    sample = torch.from_numpy(d["sample"])
    for tag, (ke, kr, ne, nr) in {"a": (4, 2, 2, 1), "b": (7, 3, 0, 0)}.items():
        smp = distillation.TopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                        student_relations=rel_s, batch_size_entity=ke, batch_size_relation=kr,
                                        n_random_entities=ne, n_random_relations=nr, seed=42)
        assert (smp.batch_size_entity, smp.batch_size_relation, smp.supervised) == (ke + ne, kr + nr, False)
        for call in range(2):
            got = smp.get(sample=sample, teacher=teacher, max_ids_per_call=100 if call else 1 << 24)  # chunked too
            for k, t in zip(("ht", "rt", "tt", "hs", "rs", "ts"), got):
                np.testing.assert_array_equal(t.cpu().numpy(), d[f"{model}/{tag}/{call}/{k}"], err_msg=f"{tag}/{call}/{k}")
        assert teacher.training


@pytest.mark.parametrize("model", ("RotatE", "TransE", "ComplEx"))
@pytest.mark.parametrize("kind", ("uniform", "topk"))
def test_distillation_loss_and_student_gradients_match_reference(model, kind):
    """Distillation.distill (distillation/distillation.py:440-677): 3-D samples through the score kernel, three
    KL divergences; loss and the student's gradients against the reference's autograd, two consecutive calls
    (the samplers' RNG streams continue)."""
    from mkb_b200 import distillation

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    teacher = getattr(models, model)(hidden_dim=8, entities=ent_t, relations=rel_t, gamma=6)
    teacher._set_params(torch.from_numpy(d[f"{model}/ent"].copy()), torch.from_numpy(d[f"{model}/rel"].copy()))
    student = getattr(models, model)(hidden_dim=8, entities=ent_s, relations=rel_s, gamma=6)
    student._set_params(torch.from_numpy(d[f"{model}/s_ent"].copy()), torch.from_numpy(d[f"{model}/s_rel"].copy()))
    teacher, student = teacher.to(DEV), student.to(DEV)
    if kind == "uniform":
        smp = distillation.UniformSampling(batch_size_entity=5, batch_size_relation=3, seed=42)
    else:
        smp = distillation.TopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                        student_relations=rel_s, batch_size_entity=4, batch_size_relation=2,
                                        n_random_entities=2, n_random_relations=1, seed=42)
    proc = distillation.Distillation(teacher_entities=ent_t, student_entities=ent_s, teacher_relations=rel_t,
                                     student_relations=rel_s, sampling=smp, device=DEV)
    sample = torch.from_numpy(d["sample2"])
    av = [proc.available(*map(int, row)) for row in sample]
    np.testing.assert_array_equal(np.array([[a["head"], a["relation"], a["tail"]] for a in av]), d[f"{model}/{kind}/avail"])
    for call in range(2):
        student.zero_grad()
        loss = proc.distill(teacher=teacher, student=student, sample=sample)
        loss.backward()
        ref = float(d[f"{model}/{kind}/{call}/loss"])
        assert abs(loss.item() - ref) <= 1e-4 * abs(ref), (call, loss.item(), ref)
        for got, key in ((student.entity_embedding.grad, "g_ent"), (student.relation_embedding.grad, "g_rel")):
            want = d[f"{model}/{kind}/{call}/{key}"]
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max(), (call, key)
        assert teacher.entity_embedding.grad is None


@pytest.mark.parametrize("model", ("RotatE", "ComplEx"))
def test_fast_top_k_sampling_matches_reference(model):
    """FastTopKSampling (top_k_sampling.py:10-318): pre-computed over the teacher's training set with the batched
    kernel path, looked up by searchsorted; same rows and the same RNG stream as the reference's dict tables."""
    from mkb_b200 import datasets, distillation

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    teacher = getattr(models, model)(hidden_dim=8, entities=ent_t, relations=rel_t, gamma=6)
    teacher._set_params(torch.from_numpy(d[f"{model}/ent"].copy()), torch.from_numpy(d[f"{model}/rel"].copy()))
    teacher = teacher.to(DEV)
    train = [tuple(int(x) for x in row) for row in d["fast/train"]]
    ds = datasets.Dataset(train=train, entities=ent_t, relations=rel_t, batch_size=7, shuffle=False, seed=42)
    smp = distillation.FastTopKSampling(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s,
                                        student_relations=rel_s, batch_size_entity=4, batch_size_relation=2,
                                        n_random_entities=2, n_random_relations=1, seed=42, teacher=teacher,
                                        dataset_teacher=ds)
    assert (smp.batch_size_entity, smp.batch_size_relation, smp.supervised) == (6, 3, False)
    q = torch.from_numpy(d["fast/query"])
    for call in range(2):
        got = smp.get(sample=q)
        for k, t in zip(("ht", "rt", "tt", "hs", "rs", "ts"), got):
            np.testing.assert_array_equal(t.cpu().numpy(), d[f"fast/{model}/{call}/{k}"], err_msg=f"{call}/{k}")
    with pytest.raises(KeyError):  # a triple whose (r, t) never occurred in the teacher's training set
        unseen = next((h, r, t) for h in range(60) for r in range(7) for t in range(60)
                      if not any(r == b and t == c for _, b, c in train))
        smp.get(sample=torch.tensor([unseen]))


def _brute_force_transe_neighbours(d, sample, ke, kr):
    """What faiss.IndexFlatL2 specifies for TopKSamplingTransE (top_k_sampling.py:756-790): exact squared-L2 nearest
    rows among the shared entities / relations to t - r, t - h, h + r; returns teacher ids."""
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    E, R = d["TransE/ent"].astype(np.float64), d["TransE/rel"].astype(np.float64)
    se = np.array([i for e, i in ent_t.items() if e in ent_s])
    sr = np.array([i for r, i in rel_t.items() if r in rel_s])
    h, r, t = sample[:, 0], sample[:, 1], sample[:, 2]

    def near(q, rows, ids, k):
        d2 = ((q[:, None, :] - rows[None, :, :]) ** 2).sum(-1)
        return ids[np.argsort(d2, axis=1, kind="stable")[:, :k]]

    return near(E[t] - R[r], E[se], se, ke), near(E[t] - E[h], R[sr], sr, kr), near(E[h] + R[r], E[se], se, ke)


def test_top_k_sampling_transe_is_an_exact_l2_search():
    """TopKSamplingTransE without faiss: the nearest shared rows in exact L2 distance, student ids through the label
    maps, the reference's RNG stream for the random extras; and FastTopKSampling picks it for a TransE teacher."""
    from mkb_b200 import datasets, distillation

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    teacher = models.TransE(hidden_dim=8, entities=ent_t, relations=rel_t, gamma=6)
    teacher._set_params(torch.from_numpy(d["TransE/ent"].copy()), torch.from_numpy(d["TransE/rel"].copy()))
    teacher = teacher.to(DEV)
    kw = dict(teacher_entities=ent_t, teacher_relations=rel_t, student_entities=ent_s, student_relations=rel_s,
              batch_size_entity=4, batch_size_relation=2, seed=42, teacher=teacher)
    smp = distillation.TopKSamplingTransE(n_random_entities=2, n_random_relations=1, **kw)
    sample = d["sample"]
    ht, rt, tt, hs, rs, ts = (x.cpu().numpy() for x in smp.get(sample=torch.from_numpy(sample), teacher=teacher))
    want_h, want_r, want_t = _brute_force_transe_neighbours(d, sample, 4, 2)
    np.testing.assert_array_equal(ht[:, :4], want_h)
    np.testing.assert_array_equal(rt[:, :2], want_r)
    np.testing.assert_array_equal(tt[:, :4], want_t)
    to_s = {i: ent_s[e] for e, i in ent_t.items() if e in ent_s}
    assert all(to_s[a] == b for a, b in zip(ht.ravel(), hs.ravel())) and ht.shape == (len(sample), 6)
    rng = np.random.RandomState(42)  # _randomize_distribution: entities first, then relations
    extra_e = rng.choice(list(to_s.keys()), size=2, replace=False)
    np.testing.assert_array_equal(ht[:, 4:], np.tile(extra_e, (len(sample), 1)))
    np.testing.assert_array_equal(tt[:, 4:], np.tile(extra_e, (len(sample), 1)))
    train = [tuple(int(x) for x in row) for row in d["fast/train"]]
    ds = datasets.Dataset(train=train, entities=ent_t, relations=rel_t, batch_size=7, shuffle=False, seed=42)
    fast = distillation.FastTopKSampling(n_random_entities=0, n_random_relations=0, dataset_teacher=ds, **kw)
    q = np.array(train[2:20:3])
    got = [x.cpu().numpy() for x in fast.get(sample=torch.from_numpy(q))]
    want_h, want_r, want_t = _brute_force_transe_neighbours(d, q, 4, 2)
    np.testing.assert_array_equal(got[0], want_h)
    np.testing.assert_array_equal(got[1], want_r)
    np.testing.assert_array_equal(got[2], want_t)


def test_kdmkb_model_steps_track_reference():
    """KdmkbModel.forward (kdmkb_model.py:286-360): two KBs with partly shared labels, each distilling from the
    other; the rolling losses of both models over four steps and the trained tables against the reference run
    (same shared-pool negatives, same FastTopKSampling RNG streams, torch.optim.Adam)."""
    import collections

    from mkb_b200 import datasets, distillation

    d = load_golden("distill_rows.npz")
    ent_t = {str(e): i for i, e in enumerate(d["labels_t"])}
    ent_s = {str(e): i for i, e in enumerate(d["labels_s"])}
    rel_t = {str(r): i for i, r in enumerate(d["rl_t"])}
    rel_s = {str(r): i for i, r in enumerate(d["rl_s"])}
    spec = {"a": ("RotatE", ent_t, rel_t, d["fast/train"]), "b": ("ComplEx", ent_s, rel_s, d["kd/train_s"])}
    ms, dss = collections.OrderedDict(), collections.OrderedDict()
    for key, (name, ents_, rels_, tr) in spec.items():
        tr = [tuple(int(x) for x in row) for row in tr]
        m = getattr(models, name)(hidden_dim=8, entities=ents_, relations=rels_, gamma=6)
        m._set_params(torch.from_numpy(d[f"kd/{key}/ent0"].copy()), torch.from_numpy(d[f"kd/{key}/rel0"].copy()))
        ms[key] = m.to(DEV)
        dss[key] = datasets.Dataset(train=tr, valid=tr[:6], test=tr[6:12], entities=ents_, relations=rels_, batch_size=6,
                                    shuffle=False, seed=42)

    def per(v):
        return {"a": v, "b": v}

    kd = distillation.KdmkbModel(models=ms, datasets=dss, lr=per(0.01), alpha_kl=per(0.4), alpha_adv=per(0.5),
                                 negative_sampling_size=per(5), batch_size_entity=per(4), batch_size_relation=per(2),
                                 n_random_entities=per(2), n_random_relations=per(1), update_distillation_every=1000,
                                 device=DEV, seed=42, warm_step=0, pool="reference")
    got = []
    for _ in range(4):
        m = kd.forward(dss, ms, weight_kl={"a": 0.4, "b": 0.4})
        got.append([m["a"].get(), m["b"].get()])
    np.testing.assert_allclose(np.array(got), d["kd/rolling_loss"], rtol=2e-4)
    for key in ms:  # Adam turns a near-zero gradient into a full +-lr step: compare as a fraction, like the trainers
        for name, init, ref in (("entity_embedding", "ent0", "ent1"), ("relation_embedding", "rel0", "rel1")):
            upd_ref = d[f"kd/{key}/{ref}"] - d[f"kd/{key}/{init}"]
            upd = getattr(ms[key], name).detach().cpu().numpy() - d[f"kd/{key}/{init}"]
            assert np.abs(upd_ref).max() > 0
            assert (np.abs(upd - upd_ref) > 0.05 * np.abs(upd_ref).max()).mean() < 0.01, (key, name)


# --------------------------------------------------------------------------------------------------
# The reference's OWN known answers for this path: the numbers printed in its doctests
# --------------------------------------------------------------------------------------------------
def _label_map(d, prefix):
    return {str(k): int(v) for k, v in zip(d[f"{prefix}/labels"], d[f"{prefix}/ids"])}


def test_top_k_sampling_reproduces_the_reference_doctest():
    """mkb/distillation/top_k_sampling.py:352-413 (CountriesS1 teacher, CountriesS2 student, RotatE dim 4 under
    torch.manual_seed(42)): the six tensors printed there, literally."""
    from mkb_b200 import distillation

    d = load_golden("distill_doctests.npz")
    ent_t, ent_s = _label_map(d, "topk/ent_t"), _label_map(d, "topk/ent_s")
    rel_t, rel_s = _label_map(d, "topk/rel_t"), _label_map(d, "topk/rel_s")
    teacher = models.RotatE(entities=ent_t, relations=rel_t, gamma=3, hidden_dim=4)
    teacher._set_params(torch.from_numpy(d["topk/ent"].copy()), torch.from_numpy(d["topk/rel"].copy()))
    teacher = teacher.to(DEV)
    smp = distillation.TopKSampling(teacher_relations=rel_t, teacher_entities=ent_t, student_entities=ent_s,
                                    student_relations=rel_s, batch_size_entity=4, batch_size_relation=1,
                                    n_random_entities=1, n_random_relations=0, seed=42)
    sample = torch.tensor([[0, 0, 266], [1, 1, 56]])  # :375-377
    ht, rt, tt, hs, rs, ts = (t.cpu().tolist() for t in smp.get(sample=sample, teacher=teacher))
    assert ht == [[197, 50, 75, 176, 30], [10, 240, 251, 3, 30]]  # :389-391
    assert rt == [[0], [1]]  # :393-395
    assert tt == [[269, 210, 270, 261, 30], [120, 160, 212, 244, 30]]  # :397-399
    assert hs == [[186, 47, 70, 166, 28], [10, 229, 240, 3, 28]]  # :401-403
    assert rs == [[0], [1]]  # :405-407
    assert ts == [[269, 198, 270, 256, 28], [111, 149, 201, 234, 28]]  # :409-413


def test_distillation_reproduces_the_reference_doctest():
    """mkb/distillation/distillation.py:476-501 (Umls, RotatE dim 3 teacher and student under torch.manual_seed(42),
    UniformSampling(3, 3, seed=42)): ``tensor(1.3066)``."""
    from mkb_b200 import distillation

    d = load_golden("distill_doctests.npz")
    ents, rels = _label_map(d, "umls/ent"), _label_map(d, "umls/rel")
    teacher = models.RotatE(hidden_dim=3, entities=ents, relations=rels, gamma=6)
    student = models.RotatE(hidden_dim=3, entities=ents, relations=rels, gamma=6)
    teacher._set_params(torch.from_numpy(d["umls/t_ent"].copy()), torch.from_numpy(d["umls/t_rel"].copy()))
    student._set_params(torch.from_numpy(d["umls/s_ent"].copy()), torch.from_numpy(d["umls/s_rel"].copy()))
    teacher, student = teacher.to(DEV), student.to(DEV)
    proc = distillation.Distillation(teacher_entities=ents, student_entities=ents, teacher_relations=rels,
                                     student_relations=rels, device=DEV,
                                     sampling=distillation.UniformSampling(batch_size_entity=3, batch_size_relation=3,
                                                                           seed=42))
    loss = proc.distill(teacher=teacher, student=student, sample=torch.from_numpy(d["umls/sample"]))
    assert round(loss.item(), 4) == 1.3066  # distillation.py:500-501
    loss.backward()
    assert student.entity_embedding.grad.abs().sum().item() > 0


def test_utils_top_k_reproduces_the_reference_doctest():
    """mkb/utils/top_k.py:24-57 (CountriesS1, RotatE dim 4 under torch.manual_seed(42) — the same tables as the
    TopKSampling doctest): the label lists printed there, through the positives kernel + the exact top-k kernel."""
    d = load_golden("distill_doctests.npz")
    ents, rels = _label_map(d, "topk/ent_t"), _label_map(d, "topk/rel_t")
    model = models.RotatE(entities=ents, relations=rels, gamma=3, hidden_dim=4)
    model._set_params(torch.from_numpy(d["topk/ent"].copy()), torch.from_numpy(d["topk/rel"].copy()))
    model = model.to(DEV)
    top_k = utils.TopK(entities=ents, relations=rels)
    assert top_k.top_heads(k=4, model=model, relation="neighbor", tail="western_africa") == [
        "mauritius", "são_tomé_and_príncipe", "guinea-bissau", "saint_kitts_and_nevis"]  # top_k.py:35-41
    assert top_k.top_relations(k=4, model=model, head="azerbaijan", tail="western_africa") == [
        "locatedin", "neighbor"]  # :43-49
    assert top_k.top_tails(k=4, model=model, head="western_africa", relation="neighbor") == [
        "afghanistan", "barbados", "taiwan", "new_caledonia"]  # :51-57


@pytest.mark.parametrize("pool", ("independent", "reference"))
def test_pipeline_adopts_a_stock_torch_adam(pool):
    """The reference quick-start's own objects — a stock torch.optim.Adam over the model's parameters — are taken
    over by the device-resident step: same trained tables and loss as the generic three-call route with the same
    optimizer stepping through autograd; optimizer.state stays current (step count, moments)."""
    from test_gpu_parity import _toy_pipeline

    ref, p0 = _toy_pipeline("generic", pool)
    m, p = _toy_pipeline("adopted", pool)  # asserts adoption and the optimizer state inside
    torch.testing.assert_close(m.entity_embedding, ref.entity_embedding, rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(m.relation_embedding, ref.relation_embedding, rtol=2e-3, atol=2e-4)
    assert abs(p.metric_loss.get() - p0.metric_loss.get()) < 1e-4
    assert abs(p.test_scores["MR"] - p0.test_scores["MR"]) <= 1.0


def test_fast_top_k_sampling_reproduces_the_reference_doctest():
    """mkb/distillation/top_k_sampling.py:24-81: the same six tensors from tables pre-computed over CountriesS1's
    1 111 training triples (batched kernel path here, a per-triple Python loop in the reference)."""
    from mkb_b200 import datasets, distillation

    d = load_golden("distill_doctests.npz")
    ent_t, ent_s = _label_map(d, "topk/ent_t"), _label_map(d, "topk/ent_s")
    rel_t, rel_s = _label_map(d, "topk/rel_t"), _label_map(d, "topk/rel_s")
    teacher = models.RotatE(entities=ent_t, relations=rel_t, gamma=3, hidden_dim=4)
    teacher._set_params(torch.from_numpy(d["topk/ent"].copy()), torch.from_numpy(d["topk/rel"].copy()))
    teacher = teacher.to(DEV)
    train = [tuple(int(x) for x in row) for row in d["topk/train"]]
    ds = datasets.Dataset(train=train, entities=ent_t, relations=rel_t, batch_size=2, shuffle=False, seed=42)
    smp = distillation.FastTopKSampling(teacher_relations=rel_t, teacher_entities=ent_t, student_entities=ent_s,
                                        student_relations=rel_s, batch_size_entity=4, batch_size_relation=1,
                                        n_random_entities=1, n_random_relations=0, seed=42, teacher=teacher,
                                        dataset_teacher=ds)
    ht, rt, tt, hs, rs, ts = (t.cpu().tolist() for t in smp.get(sample=torch.tensor([[0, 0, 266], [1, 1, 56]])))
    assert ht == [[197, 50, 75, 176, 30], [10, 240, 251, 3, 30]]  # :53-55
    assert rt == [[0], [1]] and rs == [[0], [1]]  # :57-59, :69-71
    assert tt == [[269, 210, 270, 261, 30], [120, 160, 212, 244, 30]]  # :61-63
    assert hs == [[186, 47, 70, 166, 28], [10, 229, 240, 3, 28]]  # :65-67
    assert ts == [[269, 198, 270, 256, 28], [111, 149, 201, 234, 28]]  # :73-79
