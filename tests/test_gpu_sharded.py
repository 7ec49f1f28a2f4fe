"""K7 — row-sharded entity table (BASELINE config 4).

Single-GPU tests exercise the sharded ADDRESSING with every shard resident on the one GPU ("virtual
shards"): the kernels are the ones a multi-GPU run launches, only the shard base pointers are local
allocations instead of NVLink peer mappings.  The forward must be BIT-identical to the unsharded
kernel (same arithmetic, same order); the backward differs only in the order of the atomic adds.
The 2-GPU test (skipped unless two devices are visible) runs the real thing through
DeviceTrainer(mode="rowshard") and compares with one GPU replaying the global batch."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import MODELS, MODES
from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from mkb_b200 import models, ops, sampling
    from mkb_b200.compose import DeviceTrainer

from conftest import DEV  # "cuda" (or "cpu" under the KGE_TEST_EMU developer shim)


def _problem(model, Nn, R, D, B, K, seed, gamma=9.0):
    rng = np.random.RandomState(seed)
    ent, rel = ko.init_tables(model, Nn, R, D, gamma, seed=seed)
    ent *= 2.5
    sample = np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1).astype(np.int64)
    neg = rng.randint(Nn, size=(B, K)).astype(np.int64)
    w = rng.uniform(0.1, 0.5, size=B).astype(np.float32)
    return ent, rel, sample, neg, w


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _updates_disagree(a, b, init):
    """Fraction of parameters whose total UPDATE differs by more than 5 % of the largest update.
    Two correct runs differ only in the order of the atomic gradient adds; Adam can turn that into a
    visible difference on the rare element whose gradient contributions cancel to ~0, hence a
    fraction, not a max."""
    ua, ub = a - init, b - init
    assert ua.abs().max().item() > 0
    return ((ua - ub).abs() > 0.05 * ua.abs().max()).float().mean().item()


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("G", (1, 2, 3, 4, 16))
def test_sharded_forward_is_bit_identical_and_backward_matches(model, mode, G):
    Nn, R, D, B, K, gamma = 403, 7, 64, 37, 50, 9.0  # 403 % G != 0 for G in (2, 3, 4, 16): ragged last rows
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=11 + G)
    spec = ops.TableSpec(model, D, gamma, (gamma + 2) / D)
    E, Rl, s, n, wt = _t(ent), _t(rel), _t(sample), _t(neg), _t(w)
    f32 = dict(dtype=torch.float32, device=DEV)
    ws = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)

    # unsharded reference run (the kernels the parity tests pin to the oracle)
    cp0, cn0, st0 = torch.empty(B, **f32), torch.empty(B, K, **f32), torch.zeros(4, **f32)
    ps0, ns0 = torch.empty(B, 1, **f32), torch.empty(B, K, **f32)
    ops.fused_forward_raw(spec, E, Rl, s, n, wt, mode, 0.5, cp0, cn0, st0, ws, ps0, ns0)
    ge0, gr0 = torch.zeros_like(E), torch.zeros_like(Rl)
    ops.fused_backward_raw(spec, E, Rl, s, n, mode, cp0, cn0, st0, ge0, gr0)

    shards = ops.split_rows(E, G)
    grads = [torch.zeros_like(t) for t in shards]
    ss = ops.ShardSet.of_tensors(shards, grads)
    cp1, cn1, st1 = torch.empty(B, **f32), torch.empty(B, K, **f32), torch.zeros(4, **f32)
    ps1, ns1 = torch.empty(B, 1, **f32), torch.empty(B, K, **f32)
    ops.fused_forward_sharded_raw(spec, ss, Nn, Rl, s, n, wt, mode, 0.5, cp1, cn1, st1, ws, ps1, ns1)
    for a, b in ((ps0, ps1), (ns0, ns1), (cp0, cp1), (cn0, cn1), (st0, st1)):
        assert torch.equal(a, b)
    # unfused scorer through the shard table
    assert torch.equal(ops.score_sharded(spec, ss, Nn, Rl, s), ps0)
    assert torch.equal(ops.score_sharded(spec, ss, Nn, Rl, s, n, mode), ns0)

    gr1 = torch.zeros_like(Rl)
    ops.fused_backward_sharded_raw(spec, ss, Nn, Rl, s, n, mode, cp1, cn1, st1, gr1)
    ge1 = ops.merge_rows(grads, Nn)
    scale = ge0.abs().max().item()
    assert (ge1 - ge0).abs().max().item() <= 1e-5 * scale
    assert (gr1 - gr0).abs().max().item() <= 1e-5 * gr0.abs().max().item()
    # padding rows of ragged shards are never touched
    for sidx, g in enumerate(grads):
        assert g[ops.shard_rows(Nn, G, sidx):].abs().sum().item() == 0.0
    # and the whole thing still agrees with the fp64 oracle
    _, _, _, ge_ref, gr_ref = ko.train_step(model, ent, rel, sample, neg, mode, w, gamma=gamma)
    assert np.abs(ge1.cpu().numpy() - ge_ref).max() <= 1e-4 * np.abs(ge_ref).max()
    assert np.abs(gr1.cpu().numpy() - gr_ref).max() <= 1e-4 * np.abs(gr_ref).max()


def test_sharded_scalar_red_switch_matches_vector_red():
    model, mode, G = "RotatE", "head-batch", 3
    Nn, R, D, B, K, gamma = 200, 5, 32, 16, 24, 9.0
    ent, rel, sample, neg, w = _problem(model, Nn, R, D, B, K, seed=5)
    spec = ops.TableSpec(model, D, gamma, (gamma + 2) / D)
    E, Rl, s, n, wt = _t(ent), _t(rel), _t(sample), _t(neg), _t(w)
    f32 = dict(dtype=torch.float32, device=DEV)
    ws = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    out = []
    for scalar in (False, True):
        shards = ops.split_rows(E, G)
        grads = [torch.zeros_like(t) for t in shards]
        ss = ops.ShardSet.of_tensors(shards, grads, scalar_red=scalar)
        cp, cn, st = torch.empty(B, **f32), torch.empty(B, K, **f32), torch.zeros(4, **f32)
        ops.fused_forward_sharded_raw(spec, ss, Nn, Rl, s, n, wt, mode, 0.5, cp, cn, st, ws)
        gr = torch.zeros_like(Rl)
        ops.fused_backward_sharded_raw(spec, ss, Nn, Rl, s, n, mode, cp, cn, st, gr)
        out.append(ops.merge_rows(grads, Nn))
    assert (out[0] - out[1]).abs().max().item() <= 1e-5 * out[0].abs().max().item()  # atomic order differs


def test_sharded_argument_validation():
    spec = ops.TableSpec("TransE", 6, 9.0, 11 / 6)  # hidden_dim % 4 != 0 -> unsupported
    E = torch.zeros(10, 6, device=DEV)
    Rl = torch.zeros(3, 6, device=DEV)
    ss = ops.ShardSet.of_tensors(ops.split_rows(E, 2))
    with pytest.raises(ops.N.KgeError, match="not supported"):
        ops.score_sharded(spec, ss, 10, Rl, torch.zeros(2, 3, dtype=torch.int64, device=DEV))
    with pytest.raises(ValueError):
        ops.ShardSet([1] * 17)


@pytest.mark.parametrize("model,G", [("RotatE", 4), ("ComplEx", 3), ("TransE", 2)])
def test_virtual_shard_trainer_tracks_single_gpu_trainer(model, G):
    """DeviceTrainer(virtual_shards=G) == DeviceTrainer() over the same batches and negatives."""
    Nn, R, D, B, K, gamma = 1000, 9, 64, 48, 32, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=8000), rng.randint(R, size=8000), rng.randint(Nn, size=8000)], 1), axis=0)
    w_all = torch.from_numpy(rng.uniform(0.1, 0.5, len(tri)).astype(np.float32)).to(DEV)
    T = torch.from_numpy(tri).to(DEV)
    tables = []
    for vs in (None, G):
        torch.manual_seed(3)
        m = getattr(models, model)(hidden_dim=D, entities={i: i for i in range(Nn)},
                                   relations={i: i for i in range(R)}, gamma=gamma).to(DEV)
        ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=7)
        tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, virtual_shards=vs)
        assert tr.mode == ("rowshard" if vs else "single")
        losses = []
        for step in range(6):
            idx = torch.arange(step * B, (step + 1) * B, device=DEV)
            tr.step(T[idx], w_all[idx], "head-batch" if step % 2 == 0 else "tail-batch")
            losses.append(tr.loss())
        tr.sync_model()
        ns.check_status(DEV)
        tables.append((m.entity_embedding.detach().clone(), m.relation_embedding.detach().clone(), losses))
    (e0, r0, l0), (e1, r1, l1) = tables
    assert np.allclose(l0, l1, rtol=1e-4), (l0, l1)
    torch.manual_seed(3)
    init = getattr(models, model)(hidden_dim=D, entities={i: i for i in range(Nn)},
                                  relations={i: i for i in range(R)}, gamma=gamma).to(DEV)
    assert _updates_disagree(e0, e1, init.entity_embedding.detach()) < 1e-3
    assert _updates_disagree(r0, r1, init.relation_embedding.detach()) < 1e-2


# --------------------------------------------------------------------------------------------------
# two real GPUs
# --------------------------------------------------------------------------------------------------
def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    Nn, R, D, B, K, gamma = 2001, 11, 128, 64, 32, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=20000), rng.randint(R, size=20000), rng.randint(Nn, size=20000)], 1), axis=0)
    w_all = rng.uniform(0.1, 0.5, len(tri)).astype(np.float32)
    torch.manual_seed(1)
    m = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, gamma=gamma).to(dev)
    init = m.entity_embedding.detach().clone()
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=5 + rank)
    tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, distributed=True, mode="rowshard")
    assert tr.mode == "rowshard"
    used = []
    for step in range(4):
        idx = np.arange(step * B * world, (step + 1) * B * world).reshape(world, B)[rank]
        s = torch.from_numpy(tri[idx]).to(dev)
        w = torch.from_numpy(w_all[idx]).to(dev)
        mode = "head-batch" if step % 2 == 0 else "tail-batch"
        tr.step(s, w, mode)
        used.append((s.clone(), tr.neg[:B].clone(), w.clone(), mode))
    loss = tr.loss()
    tr.sync_model()
    # replay the GLOBAL batches on rank 0 with the single-GPU kernels
    gathered = []
    for s, n, w, mode in used:
        gs = [torch.empty_like(s) for _ in range(world)]
        gn = [torch.empty_like(n) for _ in range(world)]
        gw = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(gs, s)
        dist.all_gather(gn, n)
        dist.all_gather(gw, w)
        gathered.append((torch.cat(gs), torch.cat(gn), torch.cat(gw), mode))
    if rank == 0:
        from mkb_b200 import optim

        torch.manual_seed(1)
        ref = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, gamma=gamma).to(dev)
        opt = optim.DenseAdam(ref.parameters(), lr=1e-3)
        for S, Ng, W, mode in gathered:
            l = ops.fused_adversarial_step(ref.spec, ref.entity_embedding, ref.relation_embedding, S, Ng, W, mode, 0.5)
            opt.zero_grad()
            l.backward()
            opt.step()
        out["loss_err"] = abs(loss - l.item()) / abs(l.item())
        out["frac_bad"] = _updates_disagree(ref.entity_embedding.detach(), m.entity_embedding.detach(), init)
        out["moved"] = (ref.entity_embedding.detach() - init).abs().max().item()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_rowshard_matches_single_gpu_replay():
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["moved"] > 0
    assert out["loss_err"] < 1e-4, dict(out)
    assert out["frac_bad"] < 1e-3, dict(out)


@pytest.mark.parametrize("model,G", [("RotatE", 4), ("DistMult", 3)])
def test_sharded_ranks_equal_replicated_ranks(model, G):
    """Evaluation straight from the row shards (per-shard counts, summed) == kge_rank_all on the
    gathered table, and Evaluation.eval(ranks_fn=...) gives the same metrics."""
    import functools

    from mkb_b200 import evaluation

    Nn, R, D, K, gamma = 333, 5, 32, 8, 9.0
    rng = np.random.RandomState(1)
    tri = np.unique(np.stack([rng.randint(Nn, size=2500), rng.randint(R, size=2500), rng.randint(Nn, size=2500)], 1), axis=0)
    torch.manual_seed(5)
    m = getattr(models, model)(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                               gamma=gamma).to(DEV)
    with torch.no_grad():
        m.entity_embedding.mul_(3.0)
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=7)
    tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=16, virtual_shards=G)
    ev = evaluation.Evaluation(entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, batch_size=2,
                               true_triples=[tuple(map(int, r)) for r in tri])
    q = [tuple(map(int, r)) for r in tri[:150]]
    for mode in MODES:
        got = tr.sharded_ranks(ev, q, mode)
        ref = ev.ranks(m, q, mode)  # model still holds the same (untrained) table
        if model == "DistMult":  # the replicated path ranks on the tensor cores (3xTF32): near-ties may flip
            assert (got == ref).float().mean().item() >= 0.97 and (got - ref).abs().max().item() <= 2
        else:
            assert torch.equal(got, ref)
    if model != "DistMult":
        assert ev.eval(m, q, ranks_fn=functools.partial(tr.sharded_ranks, ev)) == ev.eval(m, q)
