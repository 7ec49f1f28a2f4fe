"""Time the filtered all-entity ranking kernel (K5) at BASELINE config 5 shapes (development aid)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mkb_b200 import evaluation, models, ops

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="RotatE,TransE,ComplEx,DistMult")
ap.add_argument("--N", type=int, default=40943)
ap.add_argument("--D", type=int, default=1000)
ap.add_argument("--Q", type=int, default=3134)
ap.add_argument("--cfg5", action="store_true",
                help="BASELINE config 5 on the real Wn18rr graph of tests/golden/cfg5_wn18rr.npz: all 3134 test "
                     "triples, both modes, filtered metrics; prints ONE JSON line with the SFU roofline and exits")
args = ap.parse_args()
dev = torch.device("cuda:0")


def cfg5_line():
    """Wn18rr RotatE dim=1000 full-entity evaluation (filtered MRR / Hits@10) — evaluation/evaluation.py:185-279."""
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    import cfg5_tables

    g = np.load(os.path.join(root, "tests", "golden", "cfg5_wn18rr.npz"))
    N, R, D, gamma = int(g["n_entity"]), int(g["n_relation"]), 1000, 9.0
    train = g["train"].astype(np.int64)
    test = g["test"].astype(np.int64)
    true = np.concatenate([train, g["valid"].astype(np.int64), test])
    ent, rel = cfg5_tables.make_tables("RotatE", train, N, R, D, gamma)
    m = models.RotatE(hidden_dim=D, entities={i: i for i in range(N)}, relations={i: i for i in range(R)}, gamma=gamma)
    m._set_params(torch.from_numpy(ent), torch.from_numpy(rel))
    m = m.to(dev).eval()
    ev = evaluation.Evaluation(entities={i: i for i in range(N)}, relations={i: i for i in range(R)}, batch_size=64,
                               true_triples=[tuple(r) for r in true.tolist()], device=dev)
    q = torch.from_numpy(test).to(dev)
    times = {}
    for mode in ("head-batch", "tail-batch"):
        csr = ev._filter("head" if mode == "head-batch" else "tail", dev)
        ops.rank_all(m.spec, m.entity_embedding, m.relation_embedding, q[:64], mode, csr)  # warm-up
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.rank_all(m.spec, m.entity_embedding, m.relation_embedding, q, mode, csr)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        times[mode] = best
    t0 = time.time()
    metrics = ev.eval(m, [tuple(r) for r in test.tolist()])  # the public call: both modes, filtered
    wall = time.time() - t0
    ms = sum(times.values())
    Q = len(test)
    sqrt_per_s = 2 * Q * N * D / (ms * 1e-3)
    peak = 148 * 16 * 1.965e9  # one MUFU op per SM sub-partition per clock x 4 x 148 SMs at the max SM clock
    print(json.dumps({
        "metric": "filtered rankings/s (all-entity evaluation)", "value": 2 * Q / (ms * 1e-3), "unit": "rankings/s",
        "config": {"workload": "Wn18rr (real graph) RotatE dim=1000 full-entity evaluation, filtered, both modes (cfg5)",
                   "queries": Q, "entities": N, "tables": "tests/golden/cfg5_tables.py (seeded, structured)"},
        "ms": times, "ms_total": ms, "evaluation_eval_wall_s": wall, "metrics": metrics, "dtype": "f32",
        "roofline": {"kernel": "rank_tile_kernel<RotatE> (K5)", "bound": "sfu", "achieved": sqrt_per_s / 1e12,
                     "peak": peak / 1e12, "unit": "Tsqrt/s", "frac": sqrt_per_s / peak,
                     "note": "one MUFU.SQRT per (query, entity, dim); peak = 16 MUFU/clk/SM x 148 SMs x 1.965 GHz"},
    }))


if args.cfg5:
    cfg5_line()
    sys.exit(0)
N, R, D, Q = args.N, 11, args.D, args.Q
rng = np.random.RandomState(0)
tri = np.unique(np.stack([rng.randint(N, size=93003), rng.randint(R, size=93003), rng.randint(N, size=93003)], 1), axis=0)
test = tri[rng.choice(len(tri), Q, replace=False)]
for name in args.model.split(","):
    torch.manual_seed(0)
    m = getattr(models, name)(hidden_dim=D, entities={i: i for i in range(N)}, relations={i: i for i in range(R)}, gamma=6.0).to(dev)
    ev = evaluation.Evaluation(entities={i: i for i in range(N)}, relations={i: i for i in range(R)}, batch_size=64,
                               true_triples=tri)
    q = torch.from_numpy(test).to(dev)
    for mode in ("head-batch", "tail-batch"):
        csr = ev._filter("head" if mode == "head-batch" else "tail", dev)
        ops.rank_all(m.spec, m.entity_embedding, m.relation_embedding, q[:64], mode, csr)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ranks = ops.rank_all(m.spec, m.entity_embedding, m.relation_embedding, q, mode, csr)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        pairs = Q * N * D
        print(f"{name} {mode}: {ms:.1f} ms for {Q} queries x {N} entities x D={D} -> {Q/ms*1e3:.0f} rankings/s, "
              f"{pairs/ms/1e6:.0f} G pair-dims/s, mean rank {ranks.float().mean().item():.1f}", flush=True)
    t0 = time.time()
    out = ev.eval(m, [tuple(map(int, r)) for r in test[:512]])
    print(f"   Evaluation.eval on 512 triples (1024 rankings): {time.time()-t0:.2f} s wall -> {out}", flush=True)

# K8: exact top-k of score rows against torch.topk / a full argsort (what utils.TopK replaced)
for rows, cols, k in ((1, N, 10), (64, N, 100), (1024, N, 64)):
    x = torch.randn(rows, cols, device=dev)
    for name, fn in (("kge_topk_rows", lambda: ops.topk_rows(x, k)), ("torch.topk", lambda: torch.topk(x, k, dim=1)),
                     ("torch.argsort[:k]", lambda: torch.argsort(x, dim=1, descending=True)[:, :k])):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f"top-{k} of {rows} x {cols}: {name} {a.elapsed_time(b) / 10 * 1e3:.0f} us", flush=True)

# distillation.TopKSampling.get: the reference loops over the batch with three model calls + three full argsorts per
# triple (mkb/distillation/top_k_sampling.py:577-604); here the whole batch is scored and reduced per chunk
from mkb_b200 import distillation

ents, rels = {i: i for i in range(N)}, {i: i for i in range(R)}
torch.manual_seed(0)
teacher = models.RotatE(hidden_dim=D, entities=ents, relations=rels, gamma=6.0).to(dev)
smp = distillation.TopKSampling(teacher_entities=ents, teacher_relations=rels, student_entities=ents,
                                student_relations=rels, batch_size_entity=100, batch_size_relation=5,
                                n_random_entities=10, n_random_relations=2, seed=42)
batch = torch.from_numpy(test[:256]).to(dev)
smp.get(batch[:8], teacher)
torch.cuda.synchronize()
t0 = time.time()
out = smp.get(batch, teacher)
torch.cuda.synchronize()
print(f"TopKSampling.get: {1e3 * (time.time() - t0):.1f} ms for {batch.shape[0]} triples x {N} candidate entities "
      f"(2 modes) + {R} relations, k = 100 / 5 -> {tuple(out[0].shape)}, {tuple(out[1].shape)}", flush=True)
