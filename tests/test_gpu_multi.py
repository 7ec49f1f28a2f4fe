"""2-GPU NCCL parity: the batch-parallel step (per-rank fused forward, all-reduce of the loss sums,
per-rank fused backward, gradient all-reduce) equals one GPU running the global batch.
Skipped unless two CUDA devices are visible (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from mkb_b200 import models, ops, sampling
    from mkb_b200.compose import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    Nn, R, D, B, K = 2000, 11, 128, 64, 32
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=20000), rng.randint(R, size=20000), rng.randint(Nn, size=20000)], 1), axis=0)
    torch.manual_seed(1)
    m = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)}, gamma=9.0).to(dev)
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=5 + rank)
    order = np.random.RandomState(2).permutation(len(tri))[: B * world]
    mine = parallel.rank_slices(order, B, world, rank)[0]
    s = torch.from_numpy(tri[mine]).to(dev)
    w = torch.from_numpy(np.random.RandomState(3).uniform(0.1, 0.5, len(tri)).astype(np.float32)[mine]).to(dev)
    neg = ns.generate(s, "head-batch", check=True)
    ent, rel = m.entity_embedding.detach(), m.relation_embedding.detach()
    cp, cn = torch.empty(B, device=dev), torch.empty(B, K, device=dev)
    stats, ws = torch.zeros(4, device=dev), torch.zeros(1 << 16, dtype=torch.uint8, device=dev)
    flat = torch.zeros(ent.numel() + rel.numel(), device=dev)
    ge, gr = flat[: ent.numel()].view_as(ent), flat[ent.numel():].view_as(rel)
    ops.fused_forward_raw(m.spec, ent, rel, s, neg, w, "head-batch", 0.5, cp, cn, stats, ws)
    parallel.allreduce_loss_sums(stats)
    ops.fused_backward_raw(m.spec, ent, rel, s, neg, "head-batch", cp, cn, stats, ge, gr)
    parallel.allreduce_gradients(flat)
    gs = [torch.empty_like(s) for _ in range(world)]
    gn = [torch.empty_like(neg) for _ in range(world)]
    gw = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(gs, s)
    dist.all_gather(gn, neg)
    dist.all_gather(gw, w)
    if rank == 0:
        S, Ng, W = torch.cat(gs), torch.cat(gn), torch.cat(gw)
        loss = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, S, Ng, W, "head-batch", 0.5)
        loss.backward()
        out["loss_err"] = abs(loss.item() - parallel.loss_from_sums(stats).item())
        out["ent_err"] = (m.entity_embedding.grad - ge).abs().max().item() / m.entity_embedding.grad.abs().max().item()
        out["rel_err"] = (m.relation_embedding.grad - gr).abs().max().item() / m.relation_embedding.grad.abs().max().item()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_matches_single_gpu_global_batch():
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["loss_err"] < 1e-6
    assert out["ent_err"] < 1e-5 and out["rel_err"] < 1e-5


def _worker_trainer(rank, world, port, out, mode, packed=False, handshake=None):
    """DeviceTrainer in a multi-GPU mode vs rank 0 replaying the GLOBAL batch on one GPU."""
    import torch.distributed as dist

    from mkb_b200 import models, ops, optim, sampling
    from mkb_b200.compose import DeviceTrainer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    Nn, R, D, B, K = 3000, 11, 256, 64, 32
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=30000), rng.randint(R, size=30000), rng.randint(Nn, size=30000)], 1), axis=0)
    ents, rels = {i: i for i in range(Nn)}, {i: i for i in range(R)}
    torch.manual_seed(1)
    m = models.RotatE(hidden_dim=D, entities=ents, relations=rels, gamma=9.0).to(dev)
    torch.manual_seed(1)
    ref = models.RotatE(hidden_dim=D, entities=ents, relations=rels, gamma=9.0).to(dev)
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=ents, relations=rels, seed=5 + rank)
    tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, distributed=True, mode=mode, packed_records=packed,
                       handshake=handshake)
    opt = optim.DenseAdam([ref.entity_embedding, ref.relation_embedding], lr=1e-3)
    errs = []
    for step in range(4):
        md = "head-batch" if step % 2 == 0 else "tail-batch"
        idx = np.random.RandomState(100 + step).permutation(len(tri))[: B * world].reshape(world, B)[rank]
        s = torch.from_numpy(tri[idx]).to(dev)
        w = torch.full((B,), 0.25, device=dev)
        tr.step(s, w, md)
        gs = [torch.empty_like(s) for _ in range(world)]
        gn = [torch.empty_like(tr.neg[:B]) for _ in range(world)]
        dist.all_gather(gs, s)
        dist.all_gather(gn, tr.neg[:B].contiguous())
        S, Ng = torch.cat(gs), torch.cat(gn)
        loss = ops.fused_adversarial_step(ref.spec, ref.entity_embedding, ref.relation_embedding, S, Ng,
                                          torch.full((B * world,), 0.25, device=dev), md, 0.5)
        loss.backward()
        opt.step()
        opt.zero_grad()
        errs.append(abs(loss.item() - tr.loss()))
    tr.sync_model()  # colshard: gather the column slices back into the model's tables
    torch.cuda.synchronize()
    e = (m.entity_embedding - ref.entity_embedding).abs().max().item()
    r = (m.relation_embedding - ref.relation_embedding).abs().max().item()
    moved = (ref.entity_embedding - models.RotatE(hidden_dim=D, entities=ents, relations=rels, gamma=9.0).to(dev).entity_embedding).abs().max().item()
    res = torch.tensor([e, r, max(errs)], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["ent"], out["rel"], out["loss"] = res.tolist()
        out["mode"], out["note"], out["moved"] = tr.mode, tr.mode_note, moved
        out["handshake"] = tr.handshake
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode,packed,handshake", (("colpar", False, "peer"), ("colpar", False, "nccl"),
                                                   ("colpar", True, "nccl"), ("allreduce", False, None),
                                                   ("colshard", False, None)))
def test_two_gpu_trainer_matches_single_gpu(mode, packed, handshake):
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker_trainer, args=(2, port, out, mode, packed, handshake), nprocs=2, join=True)
    print(dict(out))
    assert out["mode"] == mode, out["note"]
    assert out["handshake"] == handshake
    assert out["loss"] < 1e-5
    # 4 Adam steps of lr 1e-3 move parameters by ~4e-3; replicas must agree with the replay to ~1e-6
    assert out["ent"] < 2e-5 and out["rel"] < 2e-5


def _worker_pipeline(rank, world, port, out, mode):
    """compose.Pipeline.learn under torch.distributed: every batch of the Dataset is the GLOBAL batch, rank r trains on
    its r-th block with its own negative stream; replicas / gathered shards end up identical on all
    ranks and the loss falls."""
    import torch.distributed as dist

    from mkb_b200 import compose, datasets, losses, models, optim, sampling

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    Nn, R, D, B, K = 2000, 7, 128, 130, 16  # 130: not a multiple of world * anything -> padded blocks
    rng = np.random.RandomState(0)
    tri = sorted({(int(rng.randint(Nn)), int(rng.randint(R)), int(rng.randint(Nn))) for _ in range(6000)})
    ents, rels = {i: i for i in range(Nn)}, {i: i for i in range(R)}
    ds = datasets.Dataset(train=tri, entities=ents, relations=rels, batch_size=B, shuffle=True, seed=42)
    torch.manual_seed(3)
    m = models.RotatE(hidden_dim=D, entities=ents, relations=rels, gamma=9.0).to(dev)
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=ents, relations=rels, seed=5)
    opt = optim.DenseAdam([p for p in m.parameters() if p.requires_grad], lr=5e-3)
    pipe = compose.Pipeline(epochs=1, device=dev, trainer_options={"mode": mode})
    pipe.learn(model=m, dataset=ds, sampling=ns, optimizer=opt, loss=losses.Adversarial(0.5))
    first = pipe.metric_loss.get()
    pipe.epochs = 3  # the same Pipeline keeps its DeviceTrainer (moments, shards) across learn() calls
    pipe.learn(model=m, dataset=ds, sampling=ns, optimizer=opt, loss=losses.Adversarial(0.5))
    tr = pipe._trainer
    gathered = [torch.empty_like(m.entity_embedding.data) for _ in range(world)]
    dist.all_gather(gathered, m.entity_embedding.data.contiguous())
    diff = max((g - gathered[0]).abs().max().item() for g in gathered)
    if rank == 0:
        out.update(mode=tr.mode, per_rank_batch=tr.max_batch, first=first, last=pipe.metric_loss.get(), diff=diff,
                   steps=tr.t, expected_steps=4 * 2 * -(-len(tri) // B))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ("colpar", "colshard", "allreduce"))
def test_two_gpu_pipeline_learn(mode):
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker_pipeline, args=(2, port, out, mode), nprocs=2, join=True)
    print(dict(out))
    # 130 triples per global batch -> 65 per rank (colshard rounds its blocks up to a multiple of 4: 16-byte units)
    assert out["mode"] == mode and out["per_rank_batch"] == (68 if mode == "colshard" else 65)
    assert out["steps"] == out["expected_steps"]
    assert out["diff"] == 0.0  # every rank holds the same table after training
    assert out["last"] < out["first"] - 0.02
