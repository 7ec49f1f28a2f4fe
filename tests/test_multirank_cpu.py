"""world_size-2 gloo test (CPU) of the multi-GPU decomposition in mkb_b200/compose/parallel.py:
per-rank partial loss sums + all-reduce + per-rank gradients with the GLOBAL normaliser + gradient
all-reduce reproduce the single-process result on the global batch.  The per-rank arithmetic is the
oracle's (no GPU here); the collectives and the slicing are the product code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kge_oracle as ko

MODEL, N, R, D, B, K, GAMMA = "RotatE", 80, 5, 8, 12, 9, 6.0


def _problem():
    rng = np.random.RandomState(3)
    ent, rel = ko.init_tables(MODEL, N, R, D, GAMMA, seed=5)
    ent *= 3
    sample = np.stack([rng.randint(N, size=2 * B), rng.randint(R, size=2 * B), rng.randint(N, size=2 * B)], 1)
    neg = rng.randint(N, size=(2 * B, K))
    w = rng.uniform(0.1, 0.5, size=2 * B)
    return ent, rel, sample.astype(np.int64), neg.astype(np.int64), w


def _worker(rank, world, port, out):
    from mkb_b200.compose import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ent, rel, sample, neg, w = _problem()
    mine = parallel.rank_slices(np.arange(2 * B), B, world, rank)[0]
    s, n, ww = sample[mine], neg[mine], w[mine]
    pos = ko.score(MODEL, ent, rel, s, gamma=GAMMA)
    ngs = ko.score(MODEL, ent, rel, s, n, "tail-batch", gamma=GAMMA)
    a = ko._softmax(ngs * 0.5, axis=1)
    stats = torch.tensor([(ww * ko._logsigmoid(pos[:, 0])).sum(), (ww * (a * ko._logsigmoid(-ngs)).sum(1)).sum(),
                          ww.sum(), 0.0], dtype=torch.float64)
    parallel.allreduce_loss_sums(stats)
    Wg = stats[2].item()
    gpos = -(ww / (2 * Wg)) * ko._sigmoid(-pos[:, 0])
    gneg = (ww / (2 * Wg))[:, None] * a * ko._sigmoid(ngs)
    ge1, gr1 = ko.score_grads(MODEL, ent, rel, s, None, None, gpos[:, None], gamma=GAMMA)
    ge2, gr2 = ko.score_grads(MODEL, ent, rel, s, n, "tail-batch", gneg, gamma=GAMMA)
    flat = torch.from_numpy(np.concatenate([(ge1 + ge2).ravel(), (gr1 + gr2).ravel()]))
    parallel.allreduce_gradients(flat)
    if rank == 0:
        out["loss"] = parallel.loss_from_sums(stats).item()
        out["flat"] = flat.numpy().copy()
        out["mine"] = mine
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_two_rank_decomposition_matches_single_process():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ent, rel, sample, neg, w = _problem()
    loss, _, _, ge, gr = ko.train_step(MODEL, ent, rel, sample, neg, "tail-batch", w, gamma=GAMMA)
    assert abs(out["loss"] - loss) < 1e-12
    np.testing.assert_allclose(out["flat"], np.concatenate([ge.ravel(), gr.ravel()]), rtol=1e-10, atol=1e-15)
    np.testing.assert_array_equal(out["mine"], np.arange(B))


def test_rank_slices_cover_the_order_once():
    from mkb_b200.compose import parallel

    order = np.random.RandomState(0).permutation(1000)
    for world in (1, 2, 4, 8):
        per_rank = [parallel.rank_slices(order, 64, world, r) for r in range(world)]
        assert len({len(p) for p in per_rank}) == 1
        seen = np.concatenate([np.concatenate(p) for p in per_rank])
        assert sorted(seen.tolist()) == list(range(1000))
        for g in range(len(per_rank[0])):
            sizes = [len(per_rank[r][g]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= 64


def _eval_worker(rank, world, port, out):
    """Evaluation(distributed=True): slices + all-gather restore query order (gloo; the per-slice ranks
    come from the oracle instead of the CUDA kernel)."""
    from mkb_b200 import evaluation

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    Nn, R, D = 30, 3, 4
    ent, rel = ko.init_tables("TransE", Nn, R, D, 6.0, seed=1)
    tri = np.unique(np.stack([rng.randint(Nn, size=120), rng.randint(R, size=120), rng.randint(Nn, size=120)], 1), axis=0)
    hc, tc = ko.build_filter_csr(tri, Nn, "head"), ko.build_filter_csr(tri, Nn, "tail")
    queries = [tuple(map(int, r)) for r in tri[:23]]  # 23 queries over 2 ranks: uneven slices

    class Model:
        entity_embedding = torch.zeros(1)

    ev = evaluation.Evaluation(entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, batch_size=4,
                               true_triples=[tuple(map(int, r)) for r in tri], distributed=True)
    seen = []

    def local(model, q, mode):
        seen.append(q.shape[0])
        r, _ = ko.rank_all("TransE", ent, rel, q.numpy(), mode, hc, tc, gamma=6.0)
        return torch.from_numpy(r)

    ev._ranks_local = local
    got = {mode: ev.ranks(Model(), queries, mode).numpy() for mode in ("head-batch", "tail-batch")}
    metrics = ev.eval(Model(), queries)
    if rank == 0:
        out["ranks"] = got
        out["metrics"] = metrics
        out["seen"] = seen
        out["ref"] = {mode: ko.rank_all("TransE", ent, rel, np.array(queries), mode, hc, tc, gamma=6.0)[0]
                      for mode in ("head-batch", "tail-batch")}
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_evaluation_restores_query_order():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_eval_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for mode in ("head-batch", "tail-batch"):
        np.testing.assert_array_equal(out["ranks"][mode], out["ref"][mode])
    assert all(n in (11, 12) for n in out["seen"])  # each rank only ranked its slice
    both = np.concatenate([out["ref"]["head-batch"], out["ref"]["tail-batch"]])
    assert out["metrics"] == ko.rank_metrics(both)


def test_weight_zero_padding_rows_change_nothing():
    """compose.Pipeline._rank_slice and DeviceTrainer's colshard step pad short per-rank batches with copies of row 0
    at weight 0.  In the reference's loss every term of a positive carries its weight (losses/adversarial.py:22-30),
    so such rows must leave the loss and every gradient untouched — checked on the oracle's closed forms."""
    ent, rel, sample, neg, w = _problem()
    s, n, ww = sample[:B], neg[:B], w[:B]
    ref = ko.train_step(MODEL, ent, rel, s, n, "tail-batch", ww, gamma=GAMMA)
    pad = 5
    s2 = np.concatenate([s, np.repeat(s[:1], pad, axis=0)])
    n2 = np.concatenate([n, np.repeat(n[:1], pad, axis=0)])
    w2 = np.concatenate([ww, np.zeros(pad)])
    got = ko.train_step(MODEL, ent, rel, s2, n2, "tail-batch", w2, gamma=GAMMA)
    assert abs(got[0] - ref[0]) <= 1e-12 * abs(ref[0])  # loss
    np.testing.assert_allclose(got[3], ref[3], rtol=0, atol=1e-15)  # entity gradient
    np.testing.assert_allclose(got[4], ref[4], rtol=0, atol=1e-15)  # relation gradient


def _colshard_worker(rank, world, port, out, model):
    """The exchange pattern of DeviceTrainer(mode="colshard") under a real process group (gloo): rank r holds the
    hidden-dim columns slice r of every row, scores the GLOBAL batch over its columns, ONE all-reduce sums the partial
    scores, every rank forms the per-score gradients redundantly and differentiates its own sub-table.  Per-rank
    arithmetic: the oracle's (the kernels' version of the same algebra is tested on the emulation,
    test_column_sharded_step_equals_the_full_step).  TransE covers the "(G-1) * gamma" correction of the distance
    models, ComplEx the plain sum of the dot-product ones (both have no hidden-dim-dependent constant)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(3)
    ent, rel = ko.init_tables(model, N, R, D, GAMMA, seed=5)
    ent, rel = ent.astype(np.float64) * 3, rel.astype(np.float64)
    sample = np.stack([rng.randint(N, size=2 * B), rng.randint(R, size=2 * B), rng.randint(N, size=2 * B)], 1)
    neg, w = rng.randint(N, size=(2 * B, K)), rng.uniform(0.1, 0.5, size=2 * B)
    nc, rc = ko.entity_dim(model, D) // D, ko.relation_dim(model, D) // D
    wd = D // world
    c0 = rank * wd
    cut = lambda t, comps: np.ascontiguousarray(t.reshape(t.shape[0], comps, D)[:, :, c0:c0 + wd].reshape(t.shape[0], comps * wd))
    e_loc, r_loc = cut(ent, nc), cut(rel, rc)
    part = np.concatenate([ko.score(model, e_loc, r_loc, sample, neg, "tail-batch", gamma=GAMMA),
                           ko.score(model, e_loc, r_loc, sample, gamma=GAMMA)], axis=1)  # [2B, K + 1] partial scores
    t = torch.from_numpy(part)
    dist.all_reduce(t)  # the scheme's one exchange
    full = t.numpy() - ((world - 1) * GAMMA if model == "TransE" else 0.0)
    ngs, pos = full[:, :K], full[:, K]
    gpos, gneg = ko.adversarial_loss_grads(pos, ngs, w)
    ge1, gr1 = ko.score_grads(model, e_loc, r_loc, sample, None, None, gpos[:, None], gamma=GAMMA)[:2]
    ge2, gr2 = ko.score_grads(model, e_loc, r_loc, sample, neg, "tail-batch", gneg, gamma=GAMMA)[:2]
    parts_e = [torch.zeros(N, nc * wd, dtype=torch.float64) for _ in range(world)]
    parts_r = [torch.zeros(R, rc * wd, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts_e, torch.from_numpy(ge1 + ge2))  # only to compare: the scheme itself never gathers
    dist.all_gather(parts_r, torch.from_numpy(gr1 + gr2))
    if rank == 0:
        ge, gr = np.zeros_like(ent), np.zeros_like(rel)
        for rr in range(world):
            ge.reshape(N, nc, D)[:, :, rr * wd:(rr + 1) * wd] = parts_e[rr].numpy().reshape(N, nc, wd)
            gr.reshape(R, rc, D)[:, :, rr * wd:(rr + 1) * wd] = parts_r[rr].numpy().reshape(R, rc, wd)
        ref = ko.train_step(model, ent, rel, sample, neg, "tail-batch", w, gamma=GAMMA)
        out["loss_err"] = abs(ko.adversarial_loss(pos, ngs, w) - ref[0])
        out["ent_err"] = float(np.abs(ge - ref[3]).max() / np.abs(ref[3]).max())
        out["rel_err"] = float(np.abs(gr - ref[4]).max() / np.abs(ref[4]).max())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("model", ("TransE", "ComplEx"))
def test_two_rank_column_sharded_step_matches_single_process(model):
    out = mp.Manager().dict()
    mp.spawn(_colshard_worker, args=(2, _free_port(), out, model), nprocs=2, join=True)
    assert out["loss_err"] < 1e-12 and out["ent_err"] < 1e-12 and out["rel_err"] < 1e-12, dict(out)
