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

DEV = "cuda"
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
    np.testing.assert_allclose(got, ref, atol=2e-3 + 0.02 * np.abs(ref))


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
