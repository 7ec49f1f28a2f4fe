"""Diagnostic: event timeline of one pipelined multi-GPU step (run under torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from mkb_b200 import models, ops, sampling
from mkb_b200.compose import DeviceTrainer, parallel
import bench

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mode_adam = sys.argv[1] if len(sys.argv) > 1 else "side"   # side | after | none
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ds, mname, N, R, T, D, B, K, gamma = bench.CONFIGS["cfg2"]
graph = bench.synth_graph("cfg2")
torch.manual_seed(42)
model = getattr(models, mname)(hidden_dim=D, entities={i: i for i in range(N)}, relations={i: i for i in range(R)}, gamma=gamma).to(dev)
ns = sampling.NegativeSampling(size=K, train_triples=graph, entities=range(N), relations=range(R), seed=42 + rank, device=dev)
tr = DeviceTrainer(model, ns, lr=5e-5, max_batch=B, distributed=world > 1, chunks=chunks)
samples = torch.from_numpy(graph[np.random.RandomState(rank).choice(len(graph), (20, B))]).to(dev)
w = torch.full((B,), 0.3, device=dev)

def step(i, rec=None):
    sample = samples[i]; mode = "head-batch" if i % 2 == 0 else "tail-batch"
    neg = tr.neg[:B]; cp, cn = tr.coef_pos[:B], tr.coef_neg[:B]
    main = torch.cuda.current_stream(dev)
    def mark(name, stream=None):
        if rec is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(stream or main); rec.append((name, e))
    mark("start")
    ops.sample_negatives(tr._csr[mode], sample, mode, K, N, ns.seed, ns._calls, tr.status, neg); ns._calls += 1
    ops.fused_forward_raw(tr.spec, tr.ent, tr.rel, sample, neg, w, mode, 0.5, cp, cn, tr.stats, tr.ws)
    mark("fwd")
    if world > 1: parallel.allreduce_loss_sums(tr.stats)
    mark("ar_stats")
    tr.t += 1
    for c, (col, wd, flat, ge, gr) in enumerate(tr.chunks):
        ops.fused_backward_chunk_raw(tr.spec, tr.ent, tr.rel, sample, neg, mode, cp, cn, tr.stats, col, wd, ge, gr)
        tr._chunk_done[c].record(main); mark(f"bwd{c}")
        with torch.cuda.stream(tr.side):
            tr.side.wait_event(tr._chunk_done[c])
            if world > 1: parallel.allreduce_gradients(flat)
            mark(f"  ar{c}", tr.side)
            if mode_adam == "side":
                ops.adam_step_chunk(tr.ent, ge, tr.m_ent, tr.v_ent, tr.nc, wd, col, D, tr.t, tr.lr)
                ops.adam_step_chunk(tr.rel, gr, tr.m_rel, tr.v_rel, tr.rc, wd, col, D, tr.t, tr.lr)
                mark(f"  adam{c}", tr.side)
    main.wait_stream(tr.side)
    if mode_adam == "after":
        for c, (col, wd, flat, ge, gr) in enumerate(tr.chunks):
            ops.adam_step_chunk(tr.ent, ge, tr.m_ent, tr.v_ent, tr.nc, wd, col, D, tr.t, tr.lr)
            ops.adam_step_chunk(tr.rel, gr, tr.m_rel, tr.v_rel, tr.rc, wd, col, D, tr.t, tr.lr)
    mark("end")

for i in range(6): step(i)
torch.cuda.synchronize()
if world > 1: dist.barrier()
rec = []
step(6, rec); step(7, rec2 := [])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(8, 18): step(i)
b.record(); torch.cuda.synchronize()
if rank == 0:
    t0 = rec[0][1]
    print(f"adam={mode_adam} chunks={chunks} world={world}: avg step {a.elapsed_time(b)/10:.3f} ms")
    print("  " + " | ".join(f"{n.strip()}@{t0.elapsed_time(e):.3f}" for n, e in rec))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
