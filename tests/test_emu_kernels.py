"""The CUDA kernels' LOGIC on a CPU emulation of the CUDA execution model (tests/emu): the sources of
mkb_b200/csrc are compiled for the host — fibers for threads, rendezvous for __syncthreads and warp
shuffles, host memory for global memory — and called through the same C ABI with numpy buffers.

This container has no GPU; these tests are how a kernel change is checked against the oracle before it
ever reaches a B200 (indexing, tiling, reductions, loss algebra, atomic scatter, sharded addressing,
Philox streams, rank counting).  They are NOT a CPU path of the product (the package refuses to run
without the real CUDA library) and say nothing about performance.  rank_tc.cu (tcgen05/TMA) cannot be
emulated and is covered by the GPU tests only."""
import ctypes as C

import numpy as np
import pytest

from conftest import MODELS, MODES, score_tol
from emu import harness as H
from mkb_b200 import _native as N
from oracle import kge_oracle as ko


def _problem(model, Nn, R, D, B, K, seed, gamma=9.0):
    rng = np.random.RandomState(seed)
    ent, rel = ko.init_tables(model, Nn, R, D, gamma, seed=seed)
    ent *= 2.5
    sample = np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1).astype(np.int64)
    neg = rng.randint(Nn, size=(B, K)).astype(np.int64)
    w = rng.uniform(0.1, 0.5, size=B).astype(np.float32)
    return ent, rel, sample, neg, w


def _close(a, ref, rel=1e-4):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    bad = np.abs(a - ref) > score_tol(ref, rel)
    assert not bad.any(), f"{bad.sum()} / {bad.size} outside tol; max abs err {np.abs(a - ref).max():.3e}"


def _grad_close(a, ref, rel=1e-4):
    err = np.abs(np.asarray(a, np.float64) - ref).max()
    assert err <= rel * max(np.abs(ref).max(), 1e-30), f"grad max err {err:.3e} vs scale {np.abs(ref).max():.3e}"


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("D,B,K", [(16, 5, 19), (5, 3, 7), (132, 2, 40)])
def test_fused_step_vs_oracle(model, mode, D, B, K):
    Nn, R, gamma = 60, 4, 9.0
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=D + B)
    loss, pos, ngs, ge, gr = ko.train_step(model, ent, rel, sample, neg, mode, w, gamma=gamma)
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    _close(f["pos"], pos)
    _close(f["neg"], ngs)
    assert abs(f["stats"][3] - loss) <= 1e-5 * abs(loss)
    _close(H.score(model, ent, rel, gamma, sample), pos)
    _close(H.score(model, ent, rel, gamma, sample, neg, mode), ngs)
    g_ent, g_rel = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f)
    _grad_close(g_ent, ge)
    _grad_close(g_rel, gr)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("G", (1, 2, 3, 16))
def test_sharded_kernels_equal_unsharded(model, mode, G):
    """K7: the forward through the shard table is BIT-identical to the unsharded kernel, the backward
    lands every row gradient in the owner's shard (ragged shards: 43 % G != 0)."""
    Nn, R, D, B, K, gamma = 43, 4, 16, 6, 21, 9.0
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=3 * G)
    f0 = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    ge0, gr0 = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f0)
    shards = H.split_rows(ent, G)
    grads = [np.zeros_like(s) for s in shards]
    st = H.shards_struct(shards, grads)
    f1 = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode, shards=st)
    for k in f0:
        assert np.array_equal(f0[k], f1[k]), k
    assert np.array_equal(H.score(model, ent, rel, gamma, sample, shards=st), f0["pos"])
    assert np.array_equal(H.score(model, ent, rel, gamma, sample, neg, mode, shards=st), f0["neg"])
    _, gr1 = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f1, shards=st)
    ge1 = H.merge_rows(grads, Nn)
    _grad_close(ge1, ge0.astype(np.float64), rel=1e-6)
    _grad_close(gr1, gr0.astype(np.float64), rel=1e-6)
    for s, g in enumerate(grads):  # padding rows of ragged shards stay untouched
        assert not g[(Nn - s + G - 1) // G:].any()
    # scalar-RED debugging switch gives the same gradient
    grads2 = [np.zeros_like(s) for s in shards]
    st2 = H.shards_struct(shards, grads2, scalar_red=True)
    H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f1, shards=st2)
    _grad_close(H.merge_rows(grads2, Nn), ge1.astype(np.float64), rel=1e-6)


def test_unfused_loss_and_score_backward_vs_oracle():
    """kge_adv_loss_fwd/bwd + kge_score_bwd (the generic three-call route) == the oracle."""
    l = H.lib()
    rng = np.random.RandomState(5)
    B, K = 9, 70
    pos = rng.normal(0, 3, size=(B, 1)).astype(np.float32)
    neg = rng.normal(0, 3, size=(B, K)).astype(np.float32)
    w = rng.uniform(0.1, 0.5, B).astype(np.float32)
    stats = np.zeros(4, np.float32)
    ws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, np.uint8)
    H.ok(l.kge_adv_loss_fwd(H.P(pos), H.P(neg), H.P(w), B, K, 0.7, H.P(stats), H.P(ws), None))
    ref = ko.adversarial_loss(pos, neg, w, 0.7)
    assert abs(stats[3] - ref) <= 1e-6 * abs(ref)
    gp, gn = np.empty(B, np.float32), np.empty((B, K), np.float32)
    H.ok(l.kge_adv_loss_bwd(H.P(pos), H.P(neg), H.P(w), B, K, 0.7, H.P(stats), None, H.P(gp), H.P(gn), None))
    rp, rn = ko.adversarial_loss_grads(pos, neg, w, 0.7)
    _grad_close(gp, rp.reshape(-1), 1e-5)
    _grad_close(gn, rn, 1e-5)
    for model in MODELS:
        for mode in MODES:
            ent, rel, sample, ngs, _ = _problem(model, 40, 3, 12, B, 17, seed=9)
            gsc = rng.normal(size=(B, 17)).astype(np.float32)
            ge, gr = np.zeros_like(ent), np.zeros_like(rel)
            tb = H.tables(model, ent, rel, 9.0)
            H.ok(l.kge_score_bwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(ngs), 17, H.P(gsc), H.P(ge),
                                 H.P(gr), None))
            re_, rr_ = ko.score_grads(model, ent, rel, sample, ngs, mode, gsc, gamma=9.0)
            _grad_close(ge, re_)
            _grad_close(gr, rr_)
            gpos = rng.normal(size=(B, 1)).astype(np.float32)
            ge, gr = np.zeros_like(ent), np.zeros_like(rel)
            H.ok(l.kge_score_bwd(C.byref(tb), 0, H.P(sample), B, None, 0, H.P(gpos), H.P(ge), H.P(gr), None))
            re_, rr_ = ko.score_grads(model, ent, rel, sample, None, None, gpos, gamma=9.0)
            _grad_close(ge, re_)
            _grad_close(gr, rr_)


@pytest.mark.parametrize("model", ("RotatE", "DistMult"))
def test_chunked_and_multi_record_backward(model):
    """kge_fused_bwd_chunk over column chunks == kge_fused_bwd; n_records > 1 == per-record launches."""
    l = H.lib()
    Nn, R, D, B, K, gamma, mode = 50, 3, 64, 5, 12, 9.0, "head-batch"
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=1)
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    ge, gr = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f)
    nc, rc = H.NC[model], H.RC[model]
    tb = H.tables(model, ent, rel, gamma)
    for col, wd in ((0, 32), (32, 16), (48, 16)):
        gec, grc = np.zeros((Nn, nc * wd), np.float32), np.zeros((R, rc * wd), np.float32)
        H.ok(l.kge_fused_bwd_chunk(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(f["cpos"]),
                                   H.P(f["cneg"]), H.P(f["stats"]), None, col, wd, 1, 0, H.P(gec), H.P(grc), None))
        _grad_close(gec, ge.reshape(Nn, nc, D)[:, :, col:col + wd].reshape(Nn, nc * wd).astype(np.float64), 1e-6)
        _grad_close(grc, gr.reshape(R, rc, D)[:, :, col:col + wd].reshape(R, rc * wd).astype(np.float64), 1e-6)
    # packed records: [sample | neg | cpos | cneg | stats] per record
    G = 3
    o_neg, o_cp = B * 24, B * 24 + B * K * 8
    o_cn = o_cp + B * 4
    o_st = (o_cn + B * K * 4 + 15) // 16 * 16
    rec = o_st + 16
    buf = np.zeros(G * rec, np.uint8)
    parts = []
    for r in range(G):
        _, _, s_r, n_r, w_r = _problem(model, Nn, R, D, B, K, seed=20 + r)
        f_r = H.fused_fwd(model, ent, rel, gamma, s_r, n_r, w_r, mode)
        base = buf[r * rec:(r + 1) * rec]
        base[:o_neg] = s_r.view(np.uint8).reshape(-1)
        base[o_neg:o_cp] = n_r.view(np.uint8).reshape(-1)
        base[o_cp:o_cn] = f_r["cpos"].view(np.uint8)
        base[o_cn:o_cn + B * K * 4] = f_r["cneg"].view(np.uint8).reshape(-1)
        base[o_st:o_st + 16] = f_r["stats"].view(np.uint8)
        parts.append((s_r, n_r, f_r))
    col, wd = 16, 32
    g1, r1 = np.zeros((Nn, nc * wd), np.float32), np.zeros((R, rc * wd), np.float32)
    p0 = buf.ctypes.data
    H.ok(l.kge_fused_bwd_chunk(C.byref(tb), H.mode_id(mode), p0, B, p0 + o_neg, K, p0 + o_cp, p0 + o_cn, p0 + o_st, None,
                               col, wd, G, rec, H.P(g1), H.P(r1), None))
    total = np.sum([f_r["stats"] for *_, f_r in parts], axis=0).astype(np.float32)
    g2, r2 = np.zeros_like(g1), np.zeros_like(r1)
    for s_r, n_r, f_r in parts:
        H.ok(l.kge_fused_bwd_chunk(C.byref(tb), H.mode_id(mode), H.P(s_r), B, H.P(n_r), K, H.P(f_r["cpos"]),
                                   H.P(f_r["cneg"]), H.P(total), None, col, wd, 1, 0, H.P(g2), H.P(r2), None))
    assert np.abs(g2).max() > 0
    _grad_close(g1, g2.astype(np.float64), 1e-6)
    _grad_close(r1, r2.astype(np.float64), 1e-6)
    # n_records < -1: the same launch with ONE already-global stats buffer (the all-reduced sums)
    g3, r3 = np.zeros_like(g1), np.zeros_like(r1)
    H.ok(l.kge_fused_bwd_chunk(C.byref(tb), H.mode_id(mode), p0, B, p0 + o_neg, K, p0 + o_cp, p0 + o_cn, H.P(total), None,
                               col, wd, -G, rec, H.P(g3), H.P(r3), None))
    _grad_close(g3, g2.astype(np.float64), 1e-6)
    _grad_close(r3, r2.astype(np.float64), 1e-6)


def test_adam_variants_vs_oracle():
    l = H.lib()
    rng = np.random.RandomState(0)
    rows, comps, D = 7, 2, 24
    p = rng.normal(size=(rows, comps * D)).astype(np.float32)
    g = rng.normal(size=p.shape).astype(np.float32)
    m, v = np.abs(rng.normal(size=p.shape)).astype(np.float32) * 0.1, np.abs(rng.normal(size=p.shape)).astype(np.float32) * 0.01
    rp, rm, rv = ko.adam_step(p.astype(np.float64), g.astype(np.float64), m.astype(np.float64), v.astype(np.float64), 3, lr=1e-3)
    p1, g1, m1, v1 = p.copy(), g.copy(), m.copy(), v.copy()
    H.ok(l.kge_adam_step(H.P(p1), H.P(g1), H.P(m1), H.P(v1), p1.size, 3, 1e-3, 0.9, 0.999, 1e-8, 1, None))
    np.testing.assert_allclose(p1, rp, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(m1, rm, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(v1, rv, rtol=2e-6, atol=1e-7)
    assert not g1.any()  # zero_grad folded in
    # column chunks (table layout for param/moments, dense chunk gradient)
    p2, m2, v2 = p.copy(), m.copy(), v.copy()
    for col, wd in ((0, 8), (8, 16)):
        gc = np.ascontiguousarray(g.reshape(rows, comps, D)[:, :, col:col + wd].reshape(rows, comps * wd))
        H.ok(l.kge_adam_step_chunk(H.P(p2), H.P(gc), H.P(m2), H.P(v2), rows, comps, wd, col, comps * D, D, 3, 1e-3, 0.9,
                                   0.999, 1e-8, 1, None))
        assert not gc.any()
    assert np.array_equal(p2, p1) and np.array_equal(m2, m1) and np.array_equal(v2, v1)
    # slice + broadcast into 3 replicas (dense slice moments)
    reps = [p.copy() for _ in range(3)]
    arr = (C.c_void_p * 3)(*[r.ctypes.data for r in reps])
    for col, wd in ((0, 8), (8, 16)):
        sl = lambda x: np.ascontiguousarray(x.reshape(rows, comps, D)[:, :, col:col + wd].reshape(rows, comps * wd))
        gs, ms, vs = sl(g), sl(m), sl(v)
        H.ok(l.kge_adam_slice_bcast(arr, 3, 1, H.P(gs), H.P(ms), H.P(vs), rows, comps, wd, col, comps * D, D, 3, 1e-3,
                                    0.9, 0.999, 1e-8, 1, None))
        assert np.array_equal(ms, sl(m1)) and np.array_equal(vs, sl(v1))
    for r in reps:
        assert np.array_equal(r, p1)


def test_samplers_bit_exact(sampler_cases):
    """kge_sample_negatives == the oracle's Philox specification, kge_filter_pool == the reference's
    own draws (golden vectors), incl. the status word."""
    l = H.lib()
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    Nn = int(g["N"])
    hc, tc = ko.build_filter_csr(triples, Nn, "head"), ko.build_filter_csr(triples, Nn, "tail")
    for call, mode in enumerate(("head-batch", "tail-batch", "tail-batch")):
        sample = np.ascontiguousarray(g[f"gen{call}/sample"], dtype=np.int64)
        fs = H.csr_struct(hc if mode == "head-batch" else tc)
        for sort_rows in (1, 0):
            out = np.full((sample.shape[0], 16), -1, np.int64)
            status = np.zeros(1, np.int32)
            H.ok(l.kge_sample_negatives(C.byref(fs), H.mode_id(mode), H.P(sample), sample.shape[0], 16, Nn, 42, call,
                                        sort_rows, H.P(out), H.P(status), None))
            ref, st = ko.sample_negatives_independent(42, call, sample, mode, Nn, hc, tc, 16, sort_rows=bool(sort_rows))
            assert status[0] == st == 0
            np.testing.assert_array_equal(out, ref)
    rng = np.random.RandomState(42)
    for step in range(6):
        sample = np.ascontiguousarray(g[f"gen{step}/sample"], dtype=np.int64)
        mode = str(g[f"gen{step}/mode"])
        size = g[f"gen{step}/neg"].shape[1]
        pool = rng.randint(Nn, size=2 * size).astype(np.int64)
        fs = H.csr_struct(hc if mode == "head-batch" else tc)
        out = np.full((sample.shape[0], size), -1, np.int64)
        status = np.zeros(1, np.int32)
        H.ok(l.kge_filter_pool(C.byref(fs), H.mode_id(mode), H.P(sample), sample.shape[0], size, Nn, H.P(pool),
                               pool.shape[0], H.P(out), H.P(status), None))
        assert status[0] == 0
        np.testing.assert_array_equal(out, g[f"gen{step}/neg"])
    # a key that is not in the graph sets bit 0 (the reference raises KeyError)
    missing = None
    keys = set(int(k) for k in hc[0])
    for r in range(int(g["R"])):
        for t in range(Nn):
            if r * Nn + t not in keys:
                missing = np.array([[0, r, t]], np.int64)
                break
        if missing is not None:
            break
    if missing is not None:
        fs = H.csr_struct(hc)
        out, status = np.zeros((1, 4), np.int64), np.zeros(1, np.int32)
        H.ok(l.kge_sample_negatives(C.byref(fs), 1, H.P(missing), 1, 4, Nn, 1, 0, 1, H.P(out), H.P(status), None))
        assert status[0] & 1


@pytest.mark.parametrize("model", MODELS)
def test_rank_tile_kernel_vs_reference(eval_cases, model):
    """kge_rank_all (fp32 tile kernel; the tcgen05 variant is GPU-only) == the reference's ranks."""
    l = H.lib()
    g = eval_cases
    Nn = 50
    allt = [tuple(int(x) for x in r) for part in ("train", "valid", "test") for r in g[part]]
    gamma = float(g[f"{model}/gamma"])
    ent = np.ascontiguousarray(g[f"{model}/ent"], np.float32)
    rel = np.ascontiguousarray(g[f"{model}/rel"], np.float32)
    hc, tc = ko.build_filter_csr(allt, Nn, "head"), ko.build_filter_csr(allt, Nn, "tail")
    test = np.ascontiguousarray(g["test"], np.int64)
    tb = H.tables(model, ent, rel, gamma)
    for mode in MODES:
        fs = H.csr_struct(hc if mode == "head-batch" else tc)
        Q = test.shape[0]
        ranks = np.zeros(Q, np.int64)
        scores = np.full((Q, Nn), np.nan, np.float32)
        ws = np.zeros(l.kge_rank_workspace_bytes(C.byref(tb), Q) + 64, np.uint8)
        H.ok(l.kge_rank_all(C.byref(tb), H.mode_id(mode), H.P(test), Q, C.byref(fs), H.P(ranks), H.P(scores), H.P(ws), None))
        ref = g[f"{model}/{mode}/ranks"]
        _, contested = ko.rank_all(model, ent, rel, test, mode, hc, tc, gamma=gamma, tie_margin=1e-5)
        assert np.all(np.abs(ranks - ref) <= contested), (ranks, ref)
        assert (ranks == ref).mean() >= 0.95
        _close(scores, g[f"{model}/{mode}/scores"])


@pytest.fixture(scope="module")
def next_rows():
    from conftest import load_golden

    return load_golden("next_rows.npz")


def test_kl_divergence_vs_reference(next_rows):
    """kge_kl_div_fwd/bwd == the reference's KlDivergence (value and autograd gradient), T = 1 and 3."""
    l = H.lib()
    g = next_rows
    for T in (1, 3):
        s = np.ascontiguousarray(g[f"kl_T{T}/f32/student"], np.float32)
        t = np.ascontiguousarray(g[f"kl_T{T}/f32/teacher"], np.float32)
        B, K = s.shape
        loss = np.zeros(1, np.float32)
        ws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, np.uint8)
        H.ok(l.kge_kl_div_fwd(H.P(s), H.P(t), B, K, float(T), H.P(loss), H.P(ws), None))
        ref = float(g[f"kl_T{T}/f32/loss"])
        assert abs(loss[0] - ref) <= 1e-5 * abs(ref)
        gs, gt = np.empty_like(s), np.empty_like(t)
        H.ok(l.kge_kl_div_bwd(H.P(s), H.P(t), B, K, float(T), None, H.P(gs), H.P(gt), None))
        _grad_close(gs, g[f"kl_T{T}/f32/grad"].astype(np.float64), 1e-4)
        # teacher gradient and fp64 value: closed forms
        s64, t64 = s.astype(np.float64) / T, t.astype(np.float64) / T
        lp = s64 - s64.max(1, keepdims=True)
        lp -= np.log(np.exp(lp).sum(1, keepdims=True))
        lq = t64 - t64.max(1, keepdims=True)
        lq -= np.log(np.exp(lq).sum(1, keepdims=True))
        q = np.exp(lq)
        kl = (q * (lq - lp)).sum(1, keepdims=True)
        assert abs(loss[0] - kl.sum() / (B * K)) <= 1e-5 * abs(ref)
        _grad_close(gt, q * ((lq - lp) - kl) / (B * K * T), 1e-4)
        _grad_close(gs, (np.exp(lp) - q) / (B * K * T), 1e-4)


@pytest.mark.parametrize("cols,k", [(1, 1), (7, 7), (50, 1), (300, 5), (1000, 64), (4099, 1000), (2500, 1024)])
def test_topk_rows_matches_stable_argsort(cols, k):
    """kge_topk_rows: descending scores, ties by ascending column, incl. heavy ties, +-0, infinities."""
    l = H.lib()
    rng = np.random.RandomState(cols + k)
    rows = 4
    x = rng.normal(size=(rows, cols + 3)).astype(np.float32)  # row stride > cols
    x[1] = np.round(x[1])  # many exact ties
    x[2, ::3] = 0.0
    x[2, 1::3] = -0.0
    if cols > 3:
        x[3, 0], x[3, 1], x[3, 2] = np.inf, -np.inf, np.inf
    idx = np.full((rows, k), -1, np.int64)
    val = np.full((rows, k), np.nan, np.float32)
    H.ok(l.kge_topk_rows(H.P(x), rows, cols, cols + 3, k, H.P(idx), H.P(val), None))
    for r in range(rows):
        ref = np.argsort(-x[r, :cols].astype(np.float64), kind="stable")[:k]
        np.testing.assert_array_equal(idx[r], ref)
        np.testing.assert_array_equal(val[r], x[r, ref])
    assert l.kge_topk_rows(H.P(x), rows, cols, cols + 3, cols + 1, H.P(idx), None, None) == -2
    assert l.kge_topk_rows(H.P(x), rows, 5000, 5000, 1025, H.P(idx), None, None) == -6


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("D,B,K", [(16, 5, 19), (5, 3, 7), (132, 2, 40)])
def test_protate_fused_step_vs_oracle(mode, D, B, K):
    """pRotatE through the same kernel template: scores, loss, table gradients, modulus gradient
    (fused and unfused routes), incl. the row-sharded variants."""
    l = H.lib()
    Nn, R, gamma, mod = 60, 4, 9.0, 0.37
    ent, rel, sample, neg, w = _problem("pRotatE", Nn, R, D, B, K, seed=D + B)
    ent *= 4.0  # phases well outside (-pi/2, pi/2)
    loss, pos, ngs, ge, gr, gm = ko.train_step("pRotatE", ent, rel, sample, neg, mode, w, gamma=gamma, modulus=mod)
    f = H.fused_fwd("pRotatE", ent, rel, gamma, sample, neg, w, mode, modulus=mod)
    _close(f["pos"], pos)
    _close(f["neg"], ngs)
    assert abs(f["stats"][3] - loss) <= 1e-5 * abs(loss)
    _close(H.score("pRotatE", ent, rel, gamma, sample, modulus=mod), pos)
    _close(H.score("pRotatE", ent, rel, gamma, sample, neg, mode, modulus=mod), ngs)
    g_ent, g_rel = H.fused_bwd("pRotatE", ent, rel, gamma, sample, neg, mode, f, modulus=mod)
    _grad_close(g_ent, ge)
    _grad_close(g_rel, gr)
    assert abs(H.modulus_grad(f, gamma, mod) - gm) <= 1e-4 * abs(gm)
    # unfused backward with arbitrary upstream gradients
    rng = np.random.RandomState(1)
    gsc = rng.normal(size=(B, K)).astype(np.float32)
    ge2, gr2 = np.zeros_like(ent), np.zeros_like(rel)
    tb = H.tables("pRotatE", ent, rel, gamma, modulus=mod)
    H.ok(l.kge_score_bwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(gsc), H.P(ge2), H.P(gr2), None))
    re_, rr_, rm_ = ko.score_grads("pRotatE", ent, rel, sample, neg, mode, gsc, gamma=gamma, modulus=mod)
    _grad_close(ge2, re_)
    _grad_close(gr2, rr_)
    out = np.zeros(1, np.float32)
    H.ok(l.kge_modulus_grad(H.P(f["neg"]), H.P(gsc), gsc.size, None, None, gamma, H.P(tb._mod), H.P(out), None))
    assert abs(out[0] - rm_) <= 1e-4 * abs(rm_)
    if D % 4 == 0:  # sharded instantiations
        shards = H.split_rows(ent, 3)
        grads = [np.zeros_like(s) for s in shards]
        st = H.shards_struct(shards, grads)
        f1 = H.fused_fwd("pRotatE", ent, rel, gamma, sample, neg, w, mode, shards=st, modulus=mod)
        for k in f:
            assert np.array_equal(f[k], f1[k]), k
        _, gr1 = H.fused_bwd("pRotatE", ent, rel, gamma, sample, neg, mode, f1, shards=st, modulus=mod)
        _grad_close(H.merge_rows(grads, Nn), g_ent.astype(np.float64), rel=1e-6)
    # a model enum without its modulus pointer is rejected
    bad = H.tables("pRotatE", ent, rel, gamma)
    assert l.kge_score_fwd(C.byref(bad), 0, H.P(sample), B, None, 0, H.P(np.zeros((B, 1), np.float32)), None) == -1


def test_protate_ranks_vs_oracle():
    l = H.lib()
    rng = np.random.RandomState(4)
    Nn, R, D, gamma, mod = 70, 3, 12, 9.0, 0.41
    ent, rel = ko.init_tables("pRotatE", Nn, R, D, gamma, seed=2)
    ent *= 4.0
    tri = np.unique(np.stack([rng.randint(Nn, size=400), rng.randint(R, size=400), rng.randint(Nn, size=400)], 1), axis=0)
    hc, tc = ko.build_filter_csr(tri, Nn, "head"), ko.build_filter_csr(tri, Nn, "tail")
    queries = np.ascontiguousarray(tri[:70], np.int64)
    tb = H.tables("pRotatE", ent, rel, gamma, modulus=mod)
    for mode in MODES:
        fs = H.csr_struct(hc if mode == "head-batch" else tc)
        Q = queries.shape[0]
        ranks = np.zeros(Q, np.int64)
        ws = np.zeros(l.kge_rank_workspace_bytes(C.byref(tb), Q) + 64, np.uint8)
        H.ok(l.kge_rank_all(C.byref(tb), H.mode_id(mode), H.P(queries), Q, C.byref(fs), H.P(ranks), None, H.P(ws), None))
        ref, contested = ko.rank_all("pRotatE", ent, rel, queries, mode, hc, tc, gamma=gamma, tie_margin=2e-5, modulus=mod)
        assert np.all(np.abs(ranks - ref) <= contested), (ranks, ref)
        assert (ranks == ref).mean() >= 0.95


@pytest.mark.parametrize("model", MODELS + ("pRotatE",))
@pytest.mark.parametrize("G", (1, 3, 16))
def test_sharded_rank_counts_sum_to_unsharded_ranks(model, G):
    """K5 over row shards: 1 + sum of the per-shard counts == kge_rank_all on the whole table."""
    l = H.lib()
    rng = np.random.RandomState(G)
    Nn, R, D, gamma, mod = 75, 3, 8, 9.0, 0.4
    ent, rel = ko.init_tables(model, Nn, R, D, gamma, seed=G)
    ent *= 4.0
    ent[5] = ent[9]  # exact ties, resolved by entity id across shard boundaries
    ent[17] = ent[9]
    tri = np.unique(np.stack([rng.randint(Nn, size=300), rng.randint(R, size=300), rng.randint(Nn, size=300)], 1), axis=0)
    hc, tc = ko.build_filter_csr(tri, Nn, "head"), ko.build_filter_csr(tri, Nn, "tail")
    queries = np.ascontiguousarray(np.concatenate([tri[:40], [[9, 0, 9], [5, 1, 17]]]), np.int64)
    Q = queries.shape[0]
    mk = dict(modulus=mod) if model == "pRotatE" else {}
    tb = H.tables(model, ent, rel, gamma, **mk)
    shards = H.split_rows(ent, G)
    st = H.shards_struct(shards)
    for mode in MODES:
        fs = H.csr_struct(hc if mode == "head-batch" else tc)
        ws = np.zeros(l.kge_rank_workspace_bytes(C.byref(tb), Q) + 64, np.uint8)
        full = np.zeros(Q, np.int64)
        sc_full = np.full((Q, Nn), np.nan, np.float32)
        H.ok(l.kge_rank_all(C.byref(tb), H.mode_id(mode), H.P(queries), Q, C.byref(fs), H.P(full), H.P(sc_full), H.P(ws), None))
        tbs = H.tables(model, ent, rel, gamma, **mk)
        tbs.entity = None
        total = np.ones(Q, np.int64)
        sc = np.full((Q, Nn), np.nan, np.float32)
        for s in range(G):
            counts = np.full(Q, -7, np.int64)
            H.ok(l.kge_rank_counts_sharded(C.byref(tbs), C.byref(st), s, H.mode_id(mode), H.P(queries), Q, C.byref(fs),
                                           H.P(counts), H.P(sc), H.P(ws), None))
            assert counts.min() >= 0
            total += counts
        np.testing.assert_array_equal(total, full)
        np.testing.assert_array_equal(sc, sc_full)
        ref, contested = ko.rank_all(model, ent, rel, queries, mode, hc, tc, gamma=gamma, tie_margin=2e-5,
                                     **({"modulus": mod} if model == "pRotatE" else {}))
        assert np.all(np.abs(total - ref) <= contested)


@pytest.mark.parametrize("model", MODELS + ("pRotatE",))
@pytest.mark.parametrize("mode", MODES)
def test_by_entity_backward_with_fused_adam_equals_scatter_plus_adam(model, mode):
    """kge_bwd_by_entity_adam (per-step CSR by entity, no atomics, Adam in place) == kge_fused_bwd followed
    by kge_adam_step on the entity table; relation gradient identical; result independent of the
    scatter order (deterministic)."""
    l = H.lib()
    Nn, R, D, B, K, gamma, mod = 37, 3, 24, 9, 14, 9.0, 0.4
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=7)
    sample[3, 2] = sample[3, 0]  # a positive whose head is also its tail
    neg[2, :5] = neg[2, 0]       # duplicates inside a row
    mk = dict(modulus=mod) if model == "pRotatE" else {}
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode, **mk)
    ge, gr = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f, **mk)
    rng = np.random.RandomState(3)
    m0 = (np.abs(rng.normal(size=ent.shape)) * 1e-3).astype(np.float32)
    v0 = (np.abs(rng.normal(size=ent.shape)) * 1e-6).astype(np.float32)
    p_ref, m_ref, v_ref, g_tmp = ent.copy(), m0.copy(), v0.copy(), ge.copy()
    H.ok(l.kge_adam_step(H.P(p_ref), H.P(g_tmp), H.P(m_ref), H.P(v_ref), p_ref.size, 4, 1e-3, 0.9, 0.999, 1e-8, 1, None))
    r_ref, rm_ref, rv_ref, g_tmp = rel.copy(), np.zeros_like(rel), np.zeros_like(rel), gr.copy()
    H.ok(l.kge_adam_step(H.P(r_ref), H.P(g_tmp), H.P(rm_ref), H.P(rv_ref), r_ref.size, 4, 1e-3, 0.9, 0.999, 1e-8, 1, None))
    outs = []
    for trial in range(2):
        p1, m1, v1 = ent.copy(), m0.copy(), v0.copy()
        r1, rm1, rv1 = rel.copy(), np.zeros_like(rel), np.zeros_like(rel)
        tb = H.tables(model, p1, r1, gamma, **mk)
        ws = np.full(l.kge_byent_workspace_bytes(C.byref(tb), B, K) + 64, 0xAB, np.uint8)  # garbage: must not matter
        wsp = (ws.ctypes.data + 63) & ~63
        H.ok(l.kge_bwd_by_entity_adam(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(f["cpos"]),
                                      H.P(f["cneg"]), H.P(f["stats"]), None, H.P(p1), H.P(m1), H.P(v1), H.P(r1), H.P(rm1),
                                      H.P(rv1), 4, 1e-3, 0.9, 0.999, 1e-8, wsp, None), "kge_bwd_by_entity_adam")
        outs.append((p1, m1, v1, r1, rm1, rv1))
    p1, m1, v1, r1, rm1, rv1 = outs[0]
    np.testing.assert_allclose(rm1, rm_ref, rtol=1e-4, atol=1e-6 * np.abs(gr).max())
    assert np.abs((r1 - rel) - (r_ref - rel)).max() <= 2e-3 * np.abs(r_ref - rel).max()
    # Adam is applied to a gradient that differs only by summation order
    np.testing.assert_allclose(m1, m_ref, rtol=1e-4, atol=1e-6 * np.abs(ge).max())
    np.testing.assert_allclose(v1, v_ref, rtol=1e-3, atol=1e-12)
    upd_ref, upd = p_ref - ent, p1 - ent
    assert np.abs(upd_ref).max() > 0
    assert np.abs(upd - upd_ref).max() <= 2e-3 * np.abs(upd_ref).max()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    # wrong table pointer / unsupported dims are rejected before any launch
    other = ent.copy()
    tb = H.tables(model, ent, rel, gamma, **mk)
    assert l.kge_bwd_by_entity_adam(C.byref(tb), 0, H.P(sample), B, H.P(neg), K, H.P(f["cpos"]), H.P(f["cneg"]),
                                    H.P(f["stats"]), None, H.P(other), H.P(m0), H.P(v0), H.P(rel), H.P(rm1), H.P(rv1), 1,
                                    1e-3, 0.9, 0.999, 1e-8, wsp, None) == -6


@pytest.mark.parametrize("Nn,B,K", [(3, 30, 250), (5, 10, 40)])
def test_by_entity_backward_huge_buckets(Nn, B, K):
    """Buckets larger than the in-kernel sort capacity (2048 entries per entity) take the unsorted path;
    buckets of 33..2048 entries the shared-memory bitonic sort (the warp-shuffle sort covers <= 32)."""
    l = H.lib()
    model, mode, R, D, gamma = "TransE", "tail-batch", 2, 4, 6.0
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=2)
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    ge, gr = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f)
    p_ref, m_ref, v_ref, g_tmp = ent.copy(), np.zeros_like(ent), np.zeros_like(ent), ge.copy()
    H.ok(l.kge_adam_step(H.P(p_ref), H.P(g_tmp), H.P(m_ref), H.P(v_ref), p_ref.size, 1, 1e-3, 0.9, 0.999, 1e-8, 1, None))
    p1, m1, v1 = ent.copy(), np.zeros_like(ent), np.zeros_like(ent)
    r1, rm1, rv1 = rel.copy(), np.zeros_like(rel), np.zeros_like(rel)
    tb = H.tables(model, p1, r1, gamma)
    ws = np.zeros(l.kge_byent_workspace_bytes(C.byref(tb), B, K) + 64, np.uint8)
    wsp = (ws.ctypes.data + 63) & ~63
    H.ok(l.kge_bwd_by_entity_adam(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(f["cpos"]), H.P(f["cneg"]),
                                  H.P(f["stats"]), None, H.P(p1), H.P(m1), H.P(v1), H.P(r1), H.P(rm1), H.P(rv1), 1, 1e-3,
                                  0.9, 0.999, 1e-8, wsp, None))
    np.testing.assert_allclose(m1, m_ref, rtol=1e-4, atol=1e-6 * np.abs(ge).max())
    assert np.abs((p1 - ent) - (p_ref - ent)).max() <= 2e-3 * np.abs(p_ref - ent).max()


@pytest.mark.parametrize("model", ("DistMult", "ComplEx"))
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("D,K", [(12, 16), (5, 1), (33, 7)])
def test_pooled_dot_step_equals_fused_step_on_the_same_negatives(model, mode, D, K, sampler_cases):
    """kge_pooled_dot_fwd/bwd (S = Q·Pool^T and two backward GEMMs; fp32 tiles here, tcgen05 on the GPU)
    == kge_fused_fwd/bwd fed the negatives kge_filter_pool selects from the same pool."""
    l = H.lib()
    g = sampler_cases
    triples = [tuple(int(x) for x in r) for r in g["triples"]]
    Nn, R, gamma = int(g["N"]), int(g["R"]), 9.0
    ent, rel = ko.init_tables(model, Nn, R, D, gamma, seed=3)
    ent *= 3.0
    sample = np.ascontiguousarray(g["gen0/sample"], np.int64)
    B = sample.shape[0]
    rng = np.random.RandomState(8)
    w = rng.uniform(0.1, 0.5, B).astype(np.float32)
    pool = rng.randint(Nn, size=max(2 * K, 8)).astype(np.int64)
    pool[5] = pool[2]  # a repeated id inside the pool
    P = pool.shape[0]
    csr = ko.build_filter_csr(triples, Nn, "head" if mode == "head-batch" else "tail")
    fs = H.csr_struct(csr)
    neg, pos_idx = np.full((B, K), -1, np.int64), np.full((B, K), -1, np.int32)
    status = np.zeros(1, np.int32)
    H.ok(l.kge_filter_pool_positions(C.byref(fs), H.mode_id(mode), H.P(sample), B, K, Nn, H.P(pool), P, H.P(neg),
                                     H.P(pos_idx), H.P(status), None))
    assert status[0] == 0 and np.array_equal(pool[pos_idx], neg)
    neg2 = np.full((B, K), -1, np.int64)
    H.ok(l.kge_filter_pool(C.byref(fs), H.mode_id(mode), H.P(sample), B, K, Nn, H.P(pool), P, H.P(neg2), H.P(status), None))
    assert np.array_equal(neg, neg2)
    # reference: the fused kernels on those negatives
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    ge, gr = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f)
    # pooled path
    tb = H.tables(model, ent, rel, gamma)
    ws = np.full(l.kge_pooled_workspace_bytes(C.byref(tb), B, K, P) + 64, 0xCD, np.uint8)
    wsp = (ws.ctypes.data + 63) & ~63
    lws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, np.uint8)
    ps, ns = np.full((B, 1), np.nan, np.float32), np.full((B, K), np.nan, np.float32)
    cpos, stats = np.full(B, np.nan, np.float32), np.zeros(4, np.float32)
    H.ok(l.kge_pooled_dot_fwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(pool), P, H.P(pos_idx), K, H.P(w), 0.5,
                              H.P(ps), H.P(ns), H.P(cpos), H.P(stats), wsp, H.P(lws), None), "kge_pooled_dot_fwd")
    _close(ps, f["pos"].astype(np.float64), 1e-5)
    _close(ns, f["neg"].astype(np.float64), 1e-5)
    np.testing.assert_allclose(cpos, f["cpos"], rtol=1e-5)
    np.testing.assert_allclose(stats, f["stats"], rtol=1e-5)
    g_ent, g_rel = np.zeros_like(ent), np.zeros_like(rel)
    H.ok(l.kge_pooled_dot_bwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(pool), P, K, H.P(cpos), H.P(stats), None,
                              H.P(g_ent), H.P(g_rel), wsp, None), "kge_pooled_dot_bwd")
    _grad_close(g_ent, ge.astype(np.float64), 1e-5)
    _grad_close(g_rel, gr.astype(np.float64), 1e-5)
    # and against the oracle
    loss, _, _, ge64, gr64 = ko.train_step(model, ent, rel, sample, neg, mode, w, gamma=gamma)
    assert abs(stats[3] - loss) <= 1e-5 * abs(loss)
    _grad_close(g_ent, ge64)
    _grad_close(g_rel, gr64)
    # distance models are refused
    tb2 = H.tables("RotatE", *ko.init_tables("RotatE", Nn, R, D, gamma, seed=1), gamma)
    assert l.kge_pooled_dot_fwd(C.byref(tb2), 0, H.P(sample), B, H.P(pool), P, H.P(pos_idx), K, H.P(w), 0.5, None, None,
                                H.P(cpos), H.P(stats), wsp, H.P(lws), None) == -6


def _fuzz_cases(n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        out.append((MODELS + ("pRotatE",))[rng.randint(5)] if True else None)
    shapes = [(int(rng.choice([1, 2, 7])), int(rng.choice([1, 2, 31, 33, 64, 65])),
               int(rng.choice([1, 3, 4, 5, 8, 31, 32, 36, 100, 260])), int(rng.choice([1, 2, 5, 40])),
               int(rng.choice([1, 3])), MODES[rng.randint(2)]) for _ in range(n)]
    return [(m,) + s for m, s in zip(out, shapes)]


@pytest.mark.parametrize("model,B,K,D,Nn,R,mode", _fuzz_cases(36, 123))
def test_fused_step_edge_shapes(model, B, K, D, Nn, R, mode):
    """Degenerate and ragged shapes (B = 1, K = 1, D = 1, one entity, partial warps, D just past a vector or
    block boundary): every route agrees with the oracle, and the specialised paths with the generic one."""
    l = H.lib()
    gamma, mod = 6.0, 0.3
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=B * 1000 + K * 10 + D)
    mk = dict(modulus=mod) if model == "pRotatE" else {}
    ok_ = dict(modulus=mod) if model == "pRotatE" else {}
    out = ko.train_step(model, ent, rel, sample, neg, mode, w, gamma=gamma, **ok_)
    loss, pos, ngs, ge, gr = out[:5]
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode, **mk)
    _close(f["pos"], pos)
    _close(f["neg"], ngs)
    assert abs(f["stats"][3] - loss) <= 1e-5 * abs(loss)
    g_ent, g_rel = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f, **mk)
    # with one or two entities the contributions of a row can cancel to ~0: compare at the scale of one term
    floor = max(1e-3, float(np.abs(gr).max()))  # the relation row receives every term with one sign pattern
    assert np.abs(g_ent - ge).max() <= 1e-4 * max(np.abs(ge).max(), floor)
    assert np.abs(g_rel - gr).max() <= 1e-4 * max(np.abs(gr).max(), floor)
    if model == "pRotatE":
        assert abs(H.modulus_grad(f, gamma, mod) - out[5]) <= 1e-4 * abs(out[5]) + 1e-12
    if D % 4 == 0:
        # row shards (more shards than entities included) and the by-entity backward
        G = 3
        shards = H.split_rows(ent, G)
        grads = [np.zeros_like(s) for s in shards]
        st = H.shards_struct(shards, grads)
        f1 = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode, shards=st, **mk)
        for k in f:
            assert np.array_equal(f[k], f1[k]), k
        H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, f1, shards=st, **mk)
        # same terms, but the order of the float atomics is free (it changes with the block schedule)
        assert np.abs(H.merge_rows(grads, Nn) - g_ent).max() <= 1e-5 * max(np.abs(g_ent).max(), floor)
        p_ref, m_ref, v_ref, g_tmp = ent.copy(), np.zeros_like(ent), np.zeros_like(ent), g_ent.copy()
        H.ok(l.kge_adam_step(H.P(p_ref), H.P(g_tmp), H.P(m_ref), H.P(v_ref), p_ref.size, 1, 1e-2, 0.9, 0.999, 1e-8, 1, None))
        p1, m1, v1 = ent.copy(), np.zeros_like(ent), np.zeros_like(ent)
        r1, rm1, rv1 = rel.copy(), np.zeros_like(rel), np.zeros_like(rel)
        tb = H.tables(model, p1, r1, gamma, **mk)
        ws = np.zeros(l.kge_byent_workspace_bytes(C.byref(tb), B, K) + 64, np.uint8)
        wsp = (ws.ctypes.data + 63) & ~63
        H.ok(l.kge_bwd_by_entity_adam(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(f["cpos"]),
                                      H.P(f["cneg"]), H.P(f["stats"]), None, H.P(p1), H.P(m1), H.P(v1), H.P(r1), H.P(rm1),
                                      H.P(rv1), 1, 1e-2, 0.9, 0.999, 1e-8, wsp, None))
        np.testing.assert_allclose(m1, m_ref, rtol=1e-4, atol=1e-4 * max(np.abs(g_ent).max(), floor))
    # top-k of the negative scores and ranks of the whole (tiny) entity set
    kk = min(K, 3)
    idx = np.zeros((B, kk), np.int64)
    H.ok(l.kge_topk_rows(H.P(f["neg"]), B, K, K, kk, H.P(idx), None, None))
    for r_ in range(B):
        np.testing.assert_array_equal(idx[r_], np.argsort(-f["neg"][r_].astype(np.float64), kind="stable")[:kk])
    tb = H.tables(model, ent, rel, gamma, **mk)
    ranks = np.zeros(B, np.int64)
    ws = np.zeros(l.kge_rank_workspace_bytes(C.byref(tb), B) + 64, np.uint8)
    H.ok(l.kge_rank_all(C.byref(tb), H.mode_id(mode), H.P(sample), B, None, H.P(ranks), None, H.P(ws), None))
    ref, contested = ko.rank_all(model, ent, rel, sample, mode, (np.zeros(0, np.int64),) * 3, (np.zeros(0, np.int64),) * 3,
                                 gamma=gamma, tie_margin=2e-5, **ok_)
    assert np.all(np.abs(ranks - ref) <= contested), (ranks, ref)


def test_sampler_status_word_on_exhausted_true_sets():
    """A positive whose true set covers EVERY entity: the independent sampler flags bit 1 (the reference would
    spin forever), the pool filter flags bit 2 and returns zeros; other rows are unaffected."""
    l = H.lib()
    Nn = 3
    triples = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (1, 0, 2)]  # (h=0, r=0) has every tail; (h=1, r=0) only tail 2
    tc = ko.build_filter_csr(triples, Nn, "tail")
    fs = H.csr_struct(tc)
    sample = np.array([[0, 0, 1], [1, 0, 2]], np.int64)
    out, status = np.full((2, 4), -1, np.int64), np.zeros(1, np.int32)
    H.ok(l.kge_sample_negatives(C.byref(fs), 0, H.P(sample), 2, 4, Nn, 7, 0, 1, H.P(out), H.P(status), None))
    assert status[0] & 2
    assert set(out[1].tolist()) <= {0, 1}  # tail 2 is filtered for the second positive
    pool = np.array([2, 1, 2, 0, 2, 2, 1, 2], np.int64)
    out, pos, status = np.full((2, 4), -1, np.int64), np.full((2, 4), -1, np.int32), np.zeros(1, np.int32)
    H.ok(l.kge_filter_pool_positions(C.byref(fs), 0, H.P(sample), 2, 4, Nn, H.P(pool), 8, H.P(out), H.P(pos), H.P(status), None))
    assert status[0] & 4 and not out[0].any() and not pos[0].any()
    np.testing.assert_array_equal(out[1], [1, 0, 1, 1])  # survivors 1, 0, 1 repeat cyclically
    np.testing.assert_array_equal(pos[1], [1, 3, 6, 1])


@pytest.mark.parametrize("model,K", [("RotatE", 13000), ("TransE", 20000)])
def test_fused_forward_in_the_opt_in_shared_memory_window(model, K):
    """K large enough that one CTA's dynamic shared memory (query + K scores) lies between 48 KB and the 200 KB
    cap: the launch needs cudaFuncAttributeMaxDynamicSharedMemorySize first (the emulation refuses it otherwise,
    like the device).  Beyond the cap the entry point reports KGE_E_UNSUPPORTED and callers take the unfused route."""
    l = H.lib()
    D, B, Nn, R, gamma = 8, 2, 50, 3, 6.0
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=K)
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, "tail-batch")
    loss, pos, ngs, _, _ = ko.train_step(model, ent, rel, sample, neg, "tail-batch", w, gamma=gamma)[:5]
    _close(f["pos"], pos)
    _close(f["neg"], ngs)
    assert abs(f["stats"][3] - loss) <= 1e-5 * abs(loss)
    # past the cap
    K2 = 60000
    neg2 = np.zeros((B, K2), np.int64)
    tb = H.tables(model, ent, rel, gamma)
    ws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, dtype=np.uint8)
    out = [np.zeros((B, 1), np.float32), np.zeros((B, K2), np.float32), np.zeros(B, np.float32),
           np.zeros((B, K2), np.float32), np.zeros(4, np.float32)]
    assert l.kge_fused_fwd(C.byref(tb), 0, H.P(sample), B, H.P(neg2), K2, H.P(w), 0.5, *[H.P(o) for o in out], H.P(ws),
                           None) == -6


def test_pooled_adv_kernel_in_the_opt_in_shared_memory_window():
    """pooled_adv_kernel keeps K scores + P dS entries in dynamic shared memory: (K + P) * 4 > 48 KB takes the
    opt-in path; same scores / loss terms as the gather kernels on the negatives the positions select."""
    l = H.lib()
    model, mode, D, B, Nn, R, gamma = "DistMult", "head-batch", 4, 2, 40, 3, 6.0
    K, Pn = 5000, 9000
    rng = np.random.RandomState(3)
    ent, rel, sample, _, w = _problem(model, Nn, R, D, B, 1, seed=17)
    pool = rng.randint(Nn, size=Pn).astype(np.int64)
    pos_idx = rng.randint(Pn, size=(B, K)).astype(np.int32)
    neg = pool[pos_idx]
    f = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    tb = H.tables(model, ent, rel, gamma)
    ws = np.full(l.kge_pooled_workspace_bytes(C.byref(tb), B, K, Pn) + 64, 0xCD, np.uint8)
    wsp = (ws.ctypes.data + 63) & ~63
    lws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, np.uint8)
    ps, ns = np.full((B, 1), np.nan, np.float32), np.full((B, K), np.nan, np.float32)
    cpos, stats = np.full(B, np.nan, np.float32), np.zeros(4, np.float32)
    H.ok(l.kge_pooled_dot_fwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(pool), Pn, H.P(pos_idx), K, H.P(w), 0.5,
                              H.P(ps), H.P(ns), H.P(cpos), H.P(stats), wsp, H.P(lws), None), "kge_pooled_dot_fwd")
    _close(ps, f["pos"].astype(np.float64), 1e-5)
    _close(ns, f["neg"].astype(np.float64), 1e-5)
    np.testing.assert_allclose(stats, f["stats"], rtol=1e-5)


def test_peer_handshake_primitives():
    """kge_peer_copy / kge_peer_signal / kge_peer_wait (csrc/peer.cu) with three "peers" that are local buffers:
    a record pushed by rank 1 lands in every other peer at its offset and nowhere else, flags are raised on every
    peer, a satisfied wait returns with a clean status and an unsatisfied one times out instead of hanging."""
    l = H.lib()
    G, rec = 3, 64
    bufs = [np.zeros(G * rec + 128, np.uint8) for _ in range(G)]
    src = np.arange(rec, dtype=np.uint8) + 1
    arr = (C.c_void_p * G)(*[H.P(b) for b in bufs])
    H.ok(l.kge_peer_copy(H.P(src), arr, G, 1, 1 * rec, rec, None))
    for r in range(G):
        got = bufs[r][rec:2 * rec]
        assert np.array_equal(got, src) == (r != 1)  # the sender's own slot is written by its forward, not the copy
        assert not bufs[r][:rec].any() and not bufs[r][2 * rec:].any()
    flags = (C.c_void_p * G)(*[H.P(b) + G * rec for b in bufs])
    status = np.zeros(1, np.int32)
    for rank in range(G):
        H.ok(l.kge_peer_signal(flags, G, rank, 7, None))
    for r in range(G):
        f = bufs[r][G * rec:G * rec + 64].view(np.uint32)
        assert f[:G].tolist() == [7] * G and not f[G:].any()
        H.ok(l.kge_peer_wait(H.P(bufs[r]) + G * rec, G, 7, int(1e9), H.P(status), None))
        H.ok(l.kge_peer_wait(H.P(bufs[r]) + G * rec, G, 5, int(1e9), H.P(status), None))  # a later waiter: >=
    assert status[0] == 0
    bufs[0][G * rec + 4:G * rec + 8].view(np.uint32)[0] = 6  # rank 1 never reached step 7 on peer 0
    H.ok(l.kge_peer_wait(H.P(bufs[0]) + G * rec, G, 7, int(2e7), H.P(status), None))
    assert status[0] == 1 << 1
    assert l.kge_peer_copy(H.P(src), arr, G, 1, 8, rec, None) == -5
    assert l.kge_peer_wait(None, G, 1, 1, H.P(status), None) == -1


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
def test_column_sharded_step_equals_the_full_step(model, mode):
    """The arithmetic of DeviceTrainer(mode="colshard") with the collectives replaced by a Python sum: G sub-tables of
    hidden-dim columns, K1 partial scores summed (minus (G-1)*gamma for the distance models), the stand-alone loss
    for the per-score gradients, the fused backward per sub-table with unit statistics — against kge_fused_fwd /
    kge_fused_bwd on the full tables."""
    l = H.lib()
    rng = np.random.RandomState(11)
    Nn, R, D, B, K, gamma, G = 40, 4, 48, 7, 9, 9.0, 3
    ent, rel = ko.init_tables(model, Nn, R, D, gamma, seed=2)
    ent, rel = (ent * 3).astype(np.float32), (rel * 3).astype(np.float32)
    nc, rc = H.NC[model], H.RC[model]
    sample = np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1).astype(np.int64)
    neg = np.sort(rng.randint(Nn, size=(B, K)), axis=1).astype(np.int64)
    w = rng.uniform(0.1, 0.5, B).astype(np.float32)
    full = H.fused_fwd(model, ent, rel, gamma, sample, neg, w, mode)
    ge_ref, gr_ref = H.fused_bwd(model, ent, rel, gamma, sample, neg, mode, full)

    emb_range = float(np.float32((np.float32(gamma) + np.float32(2)) / np.float32(D)))  # the GLOBAL range
    slices = [(0, 16), (16, 16), (32, 16)]
    subs = []
    pos, ngs = np.zeros((B, 1), np.float32), np.zeros((B, K), np.float32)
    for c0, wd in slices:
        cut = lambda t, comps: np.ascontiguousarray(t.reshape(t.shape[0], comps, D)[:, :, c0:c0 + wd].reshape(t.shape[0], comps * wd))
        e_loc, r_loc = cut(ent, nc), cut(rel, rc)
        tb = N.KgeTables(H.P(e_loc), H.P(r_loc), Nn, R, wd, N.MODEL_IDS[model], float(gamma), emb_range, None)
        p_loc, n_loc = np.zeros((B, 1), np.float32), np.zeros((B, K), np.float32)
        H.ok(l.kge_score_fwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, None, 0, H.P(p_loc), None))
        H.ok(l.kge_score_fwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(n_loc), None))
        pos += p_loc
        ngs += n_loc
        subs.append((tb, e_loc, r_loc))
    if model in ("TransE", "RotatE"):
        pos -= (G - 1) * gamma
        ngs -= (G - 1) * gamma
    np.testing.assert_allclose(pos, full["pos"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(ngs, full["neg"], rtol=2e-5, atol=2e-5)
    stats = np.zeros(4, np.float32)
    ws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, np.uint8)
    gp, gn = np.zeros(B, np.float32), np.zeros((B, K), np.float32)
    posv = np.ascontiguousarray(pos.reshape(-1))
    H.ok(l.kge_adv_loss_fwd(H.P(posv), H.P(ngs), H.P(w), B, K, 0.5, H.P(stats), H.P(ws), None))
    H.ok(l.kge_adv_loss_bwd(H.P(posv), H.P(ngs), H.P(w), B, K, 0.5, H.P(stats), None, H.P(gp), H.P(gn), None))
    np.testing.assert_allclose(stats[3], full["stats"][3], rtol=1e-5)
    unit = np.array([0, 0, 0.5, 0], np.float32)
    ge, gr = np.zeros_like(ent), np.zeros_like(rel)
    for (c0, wd), (tb, e_loc, r_loc) in zip(slices, subs):
        ge_l, gr_l = np.zeros_like(e_loc), np.zeros_like(r_loc)
        H.ok(l.kge_fused_bwd(C.byref(tb), H.mode_id(mode), H.P(sample), B, H.P(neg), K, H.P(gp), H.P(gn), H.P(unit), None,
                             H.P(ge_l), H.P(gr_l), None))
        ge.reshape(Nn, nc, D)[:, :, c0:c0 + wd] = ge_l.reshape(Nn, nc, wd)
        gr.reshape(R, rc, D)[:, :, c0:c0 + wd] = gr_l.reshape(R, rc, wd)
    np.testing.assert_allclose(ge, ge_ref, rtol=2e-4, atol=2e-6 * np.abs(ge_ref).max())
    np.testing.assert_allclose(gr, gr_ref, rtol=2e-4, atol=2e-6 * np.abs(gr_ref).max())
