"""CPU: pin the oracle (oracle/kge_oracle.py) against vectors produced by executing the reference
(tests/golden/make_golden.py) and against the reference's own doctest known answers."""
import numpy as np
import pytest

from conftest import MODELS, MODES, score_tol
from oracle import kge_oracle as ko


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("D", (8, 5))
@pytest.mark.parametrize("mode", MODES)
def test_scores_loss_grads_match_reference(step_cases, model, D, mode):
    g = step_cases
    k = f"{model}_D{D}_{mode}"
    ent, rel = g[f"{k}/ent"], g[f"{k}/rel"]
    sample, neg, w = g[f"{k}/sample"], g[f"{k}/neg"], g[f"{k}/weight"]
    gamma = float(g[f"{k}/gamma"])
    loss, pos, ngs, ge, gr = ko.train_step(model, ent, rel, sample, neg, mode, w, gamma=gamma)
    # fp64 reference (model.double()): the oracle must agree to rounding
    np.testing.assert_allclose(pos, g[f"{k}/f64/pos"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ngs, g[f"{k}/f64/neg_score"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(loss, g[f"{k}/f64/loss"], rtol=1e-12)
    np.testing.assert_allclose(ge, g[f"{k}/f64/grad_ent"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(gr, g[f"{k}/f64/grad_rel"], rtol=1e-10, atol=1e-14)
    # fp32 reference: within the north-star tolerance
    assert np.all(np.abs(pos - g[f"{k}/f32/pos"]) <= score_tol(pos))
    assert np.all(np.abs(ngs - g[f"{k}/f32/neg_score"]) <= score_tol(ngs))
    assert abs(loss - float(g[f"{k}/f32/loss"])) <= 1e-5 * abs(loss)
    assert np.all(np.abs(ge - g[f"{k}/f32/grad_ent"]) <= 1e-4 * np.abs(ge).max())
    assert np.all(np.abs(gr - g[f"{k}/f32/grad_rel"]) <= 1e-4 * np.abs(gr).max())
    # 3-D sample path
    s3 = ko.score(model, ent, rel, g[f"{k}/sample3d"], gamma=gamma)
    assert s3.shape == g[f"{k}/f32/score3d"].shape
    assert np.all(np.abs(s3 - g[f"{k}/f32/score3d"]) <= score_tol(s3))


def test_fp32_mode_tracks_reference_fp32(step_cases):
    g = step_cases
    for model in MODELS:
        k = f"{model}_D8_tail-batch"
        s = ko.score(model, g[f"{k}/ent"], g[f"{k}/rel"], g[f"{k}/sample"], g[f"{k}/neg"], "tail-batch",
                     gamma=float(g[f"{k}/gamma"]), dtype=np.float32)
        assert s.dtype == np.float32
        np.testing.assert_allclose(s, g[f"{k}/f32/neg_score"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
def test_torch_port_is_bit_exact(step_cases, model, mode):
    """oracle/torch_port.py issues the reference's ATen operator sequence: on the same inputs its fp32
    scores, loss and autograd gradients equal the reference's bit for bit (same torch build)."""
    import torch

    from oracle import torch_port as tp

    g = step_cases
    for D in (8, 5):
        k = f"{model}_D{D}_{mode}"
        ent = torch.from_numpy(g[f"{k}/ent"]).requires_grad_()
        rel = torch.from_numpy(g[f"{k}/rel"]).requires_grad_()
        s, n, w = (torch.from_numpy(g[f"{k}/{x}"]) for x in ("sample", "neg", "weight"))
        gamma = float(g[f"{k}/gamma"])
        rng_ = ko.embedding_range(gamma, D)
        pos = tp.forward(model, ent, rel, s, gamma=gamma, embedding_range=rng_)
        ngs = tp.forward(model, ent, rel, s, n, mode, gamma=gamma, embedding_range=rng_)
        loss = tp.adversarial(pos, ngs, w, 0.5)
        loss.backward()
        np.testing.assert_array_equal(pos.detach().numpy(), g[f"{k}/f32/pos"])
        np.testing.assert_array_equal(ngs.detach().numpy(), g[f"{k}/f32/neg_score"])
        np.testing.assert_array_equal(loss.detach().numpy(), g[f"{k}/f32/loss"])
        np.testing.assert_array_equal(ent.grad.numpy(), g[f"{k}/f32/grad_ent"])
        np.testing.assert_array_equal(rel.grad.numpy(), g[f"{k}/f32/grad_rel"])


def test_torch_port_sampler_matches_reference(sampler_cases):
    from oracle import torch_port as tp
    import torch

    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    th, tt = tp.true_sets(triples)
    rng = np.random.RandomState(42)
    for step in range(6):
        out = tp.generate_negatives(rng, torch.from_numpy(g[f"gen{step}/sample"]), str(g[f"gen{step}/mode"]),
                                    th, tt, int(g["N"]), 16)
        np.testing.assert_array_equal(out.numpy(), g[f"gen{step}/neg"])


def test_negative_sampling_doctest(doctest_pins):
    """mkb/sampling/negative_sampling.py:101-126: sampler indices and RotatE scores, both modes."""
    g = doctest_pins
    ent, rel, sample = g["ns/ent"], g["ns/rel"], g["ns/sample"]
    train = [tuple(int(x) for x in r) for r in g["ns/train"]]
    hc = ko.build_filter_csr(train, 4, "head")
    tc = ko.build_filter_csr(train, 4, "tail")
    n_tail = ko.sample_negatives_reference_pool(g["ns/pool0"], sample, "tail-batch", 4, hc, tc, 5)
    n_head = ko.sample_negatives_reference_pool(g["ns/pool1"], sample, "head-batch", 4, hc, tc, 5)
    np.testing.assert_array_equal(n_tail, g["ns/neg_tail"])
    np.testing.assert_array_equal(n_head, g["ns/neg_head"])
    np.testing.assert_array_equal(n_tail, [[2, 3, 0, 2, 2], [3, 0, 3, 0, 0]])  # the doctest's literal
    np.testing.assert_array_equal(n_head, [[2, 2, 2, 2, 2], [2, 2, 2, 2, 3]])
    st = ko.score("RotatE", ent, rel, sample, n_tail, "tail-batch", gamma=3.0)
    sh = ko.score("RotatE", ent, rel, sample, n_head, "head-batch", gamma=3.0)
    np.testing.assert_allclose(st, g["ns/doc_score_tail"], atol=5.1e-5)
    np.testing.assert_allclose(sh, g["ns/doc_score_head"], atol=5.1e-5)
    np.testing.assert_allclose(st, g["ns/score_tail"], rtol=1e-5)
    np.testing.assert_allclose(sh, g["ns/score_head"], rtol=1e-5)


def test_transe_predict_doctest(doctest_pins):
    """mkb/utils/predict.py:89-95."""
    g = doctest_pins
    s = ko.score("TransE", g["predict/ent"], g["predict/rel"], g["predict/sample"], gamma=6.0).reshape(-1)
    np.testing.assert_allclose(s, g["predict/doc_score"], atol=5.1e-5)
    np.testing.assert_allclose(s, g["predict/score"], rtol=1e-5)


def test_weights_and_true_sets(sampler_cases):
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    N = int(g["N"])
    np.testing.assert_allclose(ko.subsampling_weights(triples), g["weights"], rtol=1e-7)
    for side, name in (("head", "true_head"), ("tail", "true_tail")):
        keys, offs, mem = ko.build_filter_csr(triples, N, side)
        ref_keys = g[f"{name}_keys"]  # (r,t) for head; (h,r) for tail
        code = ref_keys[:, 0] * N + ref_keys[:, 1] if side == "head" else ref_keys[:, 1] * N + ref_keys[:, 0]
        order = np.argsort(code, kind="stable")
        np.testing.assert_array_equal(keys, code[order])
        sizes = g[f"{name}_sizes"]
        np.testing.assert_array_equal(np.diff(offs), sizes[order])
        starts = np.concatenate([[0], np.cumsum(sizes)])
        ref_mem = np.concatenate([g[f"{name}_members"][starts[i]:starts[i + 1]] for i in order])
        np.testing.assert_array_equal(mem, ref_mem)


def test_reference_pool_sampler(sampler_cases):
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    N = int(g["N"])
    hc = ko.build_filter_csr(triples, N, "head")
    tc = ko.build_filter_csr(triples, N, "tail")
    for step in range(6):
        mode = str(g[f"gen{step}/mode"])
        out = ko.sample_negatives_reference_pool(g[f"gen{step}/pool"], g[f"gen{step}/sample"], mode, N, hc, tc, 16)
        np.testing.assert_array_equal(out, g[f"gen{step}/neg"])


def test_independent_sampler_invariants(sampler_cases):
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    N = int(g["N"])
    hc = ko.build_filter_csr(triples, N, "head")
    tc = ko.build_filter_csr(triples, N, "tail")
    sample = g["gen0/sample"]
    for mode, csr in (("head-batch", hc), ("tail-batch", tc)):
        out, status = ko.sample_negatives_independent(1234, 3, sample, mode, N, hc, tc, 16)
        assert status == 0 and out.shape == (8, 16)
        assert out.min() >= 0 and out.max() < N
        for i, (h, r, t) in enumerate(sample):
            seg = ko._segment(*csr, int(r) * N + int(t if mode == "head-batch" else h))
            assert not np.isin(out[i], seg).any()
        again, _ = ko.sample_negatives_independent(1234, 3, sample, mode, N, hc, tc, 16)
        np.testing.assert_array_equal(out, again)
        other, _ = ko.sample_negatives_independent(1234, 4, sample, mode, N, hc, tc, 16)
        assert (other != out).any()


def test_philox_known_answer():
    """Random123 kat_vectors: philox4x32-10 with zero and all-ones / pi inputs."""
    z = ko.philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    np.testing.assert_array_equal(z, np.array([0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], np.uint32))
    f = ko.philox4x32_10(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    np.testing.assert_array_equal(f, np.array([0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD], np.uint32))
    p = ko.philox4x32_10(np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32),
                         np.array([0xA4093822, 0x299F31D0], np.uint32))
    np.testing.assert_array_equal(p, np.array([0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1], np.uint32))


@pytest.mark.parametrize("mode", MODES)
def test_filter_lists(eval_cases, mode):
    g = eval_cases
    true = {tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]}
    for i, tr in enumerate(g["test"]):
        cand, bias = ko.test_candidates(tuple(int(x) for x in tr), mode, 50, true)
        np.testing.assert_array_equal(cand, g[f"{mode}/cand"][i])
        np.testing.assert_array_equal(bias, g[f"{mode}/bias"][i])


@pytest.mark.parametrize("model", MODELS)
def test_ranks_and_metrics(eval_cases, model):
    g = eval_cases
    N = 50
    allt = [tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]]
    hc = ko.build_filter_csr(allt, N, "head")
    tc = ko.build_filter_csr(allt, N, "tail")
    gamma = float(g[f"{model}/gamma"])
    ranks = []
    for mode in ("head-batch", "tail-batch"):
        r, contested = ko.rank_all(model, g[f"{model}/ent"], g[f"{model}/rel"], g["test"], mode, hc, tc,
                                   gamma=gamma, tie_margin=1e-5)
        ref = g[f"{model}/{mode}/ranks"]
        # exact wherever the fp64 oracle sees no near-tie; inside the contested window otherwise
        assert np.all(np.abs(r - ref) <= contested)
        assert (contested == 0).mean() > 0.9
        ranks.append(r if contested.sum() == 0 else ref)
    m = ko.rank_metrics(np.concatenate(ranks))
    np.testing.assert_allclose([m[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")],
                               g[f"{model}/metrics"], atol=1e-4)


def test_eval_doctest_replay(eval_doctest):
    """mkb/evaluation/evaluation.py:101-116: replay the doctest's 10 training steps with the oracle's
    fp64 forward/backward + Adam(lr=0.5) — grads accumulate because the doctest never zeroes them —
    then rank and reproduce {'MRR': 0.5417, 'MR': 2.25, 'HITS@1': 0.25, 'HITS@3': 1.0, 'HITS@10': 1.0}."""
    g = eval_doctest
    ent, rel = g["ent0"].astype(np.float64), g["rel0"].astype(np.float64)
    me, ve = np.zeros_like(ent), np.zeros_like(ent)
    mr, vr = np.zeros_like(rel), np.zeros_like(rel)
    acc_e, acc_r = np.zeros_like(ent), np.zeros_like(rel)
    for s in range(int(g["n_steps"])):
        loss, _, _, ge, gr = ko.train_step("RotatE", ent, rel, g[f"step{s}/sample"], g[f"step{s}/neg"],
                                           str(g[f"step{s}/mode"]), g[f"step{s}/weight"], gamma=1.0)
        assert abs(loss - float(g[f"step{s}/loss"])) < 2e-4 * max(1.0, abs(loss))
        acc_e += ge
        acc_r += gr
        ent, me, ve = ko.adam_step(ent, acc_e, me, ve, s + 1, lr=0.5)
        rel, mr, vr = ko.adam_step(rel, acc_r, mr, vr, s + 1, lr=0.5)
    np.testing.assert_allclose(ent, g["ent_final"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(rel, g["rel_final"], rtol=2e-3, atol=2e-3)
    allt = [tuple(int(x) for x in r) for r in g["train"]] + [tuple(int(x) for x in r) for r in g["test"]]
    hc = ko.build_filter_csr(allt, 4, "head")
    tc = ko.build_filter_csr(allt, 4, "tail")
    ranks = np.concatenate([
        ko.rank_all("RotatE", g["ent_final"], g["rel_final"], g["test"], m, hc, tc, gamma=1.0)[0]
        for m in ("head-batch", "tail-batch")])
    m = ko.rank_metrics(ranks)
    assert [m[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")] == list(g["doc_metrics"])


@pytest.mark.parametrize("D", (8, 5))
@pytest.mark.parametrize("mode", MODES)
def test_protate_oracle_matches_reference(D, mode):
    """pRotatE (mkb/models/protate.py:74-93): scores, loss, table gradients and the
    gradient of the trainable modulus against the reference executed in fp64 and fp32."""
    from conftest import load_golden

    g = load_golden("next_rows.npz")
    k = f"pRotatE_D{D}_{mode}"
    ent, rel, mod = g[f"{k}/ent"], g[f"{k}/rel"], float(g[f"{k}/modulus"][0, 0])
    sample, neg, w = g[f"{k}/sample"], g[f"{k}/neg"], g[f"{k}/weight"]
    assert mod == ko.default_modulus(9.0, D)
    loss, pos, ngs, ge, gr, gm = ko.train_step("pRotatE", ent, rel, sample, neg, mode, w, gamma=9.0, modulus=mod)
    np.testing.assert_allclose(pos, g[f"{k}/f64/pos"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ngs, g[f"{k}/f64/neg_score"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(loss, g[f"{k}/f64/loss"], rtol=1e-12)
    np.testing.assert_allclose(ge, g[f"{k}/f64/grad_ent"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(gr, g[f"{k}/f64/grad_rel"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(gm, float(g[f"{k}/f64/grad_modulus"][0, 0]), rtol=1e-10)
    assert np.all(np.abs(pos - g[f"{k}/f32/pos"]) <= score_tol(pos))
    assert np.all(np.abs(ngs - g[f"{k}/f32/neg_score"]) <= score_tol(ngs))
    assert np.all(np.abs(ge - g[f"{k}/f32/grad_ent"]) <= 1e-4 * np.abs(ge).max())
    assert abs(gm - float(g[f"{k}/f32/grad_modulus"][0, 0])) <= 1e-4 * abs(gm)
    s3 = ko.score("pRotatE", ent, rel, np.stack([sample[:4], sample[2:6]]), gamma=9.0, modulus=mod)
    assert np.all(np.abs(s3 - g[f"{k}/f32/score3d"]) <= score_tol(s3))
