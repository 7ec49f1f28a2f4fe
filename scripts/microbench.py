"""Quick kernel timings at BASELINE config shapes (development aid; bench.py is the contract)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mkb_b200 import models, ops, optim

CFG = {
    "cfg1": ("TransE", 40943, 11, 200, 256, 64, 6.0),
    "cfg2": ("RotatE", 14541, 237, 1000, 1024, 256, 9.0),
    "cfg3": ("ComplEx", 14541, 237, 1000, 1024, 256, 9.0),
    "cfg4": ("RotatE", 123182, 37, 500, 1024, 256, 24.0),
    "cfg2t": ("TransE", 14541, 237, 1000, 1024, 256, 9.0),
    "cfg2d": ("DistMult", 14541, 237, 1000, 1024, 256, 9.0),
}


def timeit(fn, iters=20, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg2,cfg3,cfg1,cfg4,cfg2t")
    ap.add_argument("--pooled", action="store_true")
    ap.add_argument("--band", type=int, default=0, help="restrict negative ids to [0, band)")
    ap.add_argument("--sorted", action="store_true", help="sort each row of negatives by id")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.cfg.split(","):
        mname, N, R, D, B, K, gamma = CFG[name]
        torch.manual_seed(42)
        m = getattr(models, mname)(hidden_dim=D, entities={i: i for i in range(N)}, relations={i: i for i in range(R)},
                                   gamma=gamma).to(dev)
        g = torch.Generator().manual_seed(43)
        s = torch.stack([torch.randint(N, (B,), generator=g), torch.randint(R, (B,), generator=g),
                         torch.randint(N, (B,), generator=g)], 1).to(dev)
        if args.pooled:
            pool = torch.randint(N, (2 * K,), generator=g)
            n = pool[:K].repeat(B, 1).to(dev)
        else:
            n = torch.randint(args.band or N, (B, K), generator=g)
            if args.sorted:
                n = n.sort(dim=1).values
            n = n.to(dev)
        w = (torch.rand(B, generator=g) * 0.4 + 0.1).to(dev)
        ent, rel = m.entity_embedding.detach(), m.relation_embedding.detach()
        coef_pos = torch.empty(B, device=dev); coef_neg = torch.empty(B, K, device=dev)
        stats = torch.empty(4, device=dev); ws = torch.zeros(1 << 16, dtype=torch.uint8, device=dev)
        g_ent = torch.zeros_like(ent); g_rel = torch.zeros_like(rel)
        row_e = ent.shape[1] * 4; row_r = rel.shape[1] * 4
        fwd_bytes = B * (2 * row_e + row_r) + B * K * row_e + 8 * (3 * B + B * K) + 4 * B + 4 * B * (1 + K) + 4
        bwd_bytes = fwd_bytes + B * K * row_e + 2 * B * row_e + B * row_r
        for mode in ("tail-batch", "head-batch"):
            f = lambda: ops.fused_forward_raw(m.spec, ent, rel, s, n, w, mode, 0.5, coef_pos, coef_neg, stats, ws)
            b = lambda: ops.fused_backward_raw(m.spec, ent, rel, s, n, mode, coef_pos, coef_neg, stats, g_ent, g_rel)
            tf, tfm = timeit(f)
            tb, tbm = timeit(b)
            tfc, _ = timeit(f, flush=flush)
            tbc, _ = timeit(b, flush=flush)
            print(f"{name} {mname} {mode} pooled={args.pooled}: fwd {tf*1e3:.1f} us (min {tfm*1e3:.1f}, L2-flushed {tfc*1e3:.1f}) "
                  f"= {fwd_bytes/tf/1e6:.0f} GB/s logical | bwd {tb*1e3:.1f} us (min {tbm*1e3:.1f}, flushed {tbc*1e3:.1f}) "
                  f"= {bwd_bytes/tb/1e6:.0f} GB/s logical | loss {stats[3].item():.5f}", flush=True)
        sc = lambda: ops._score_fwd(m.spec, ent, rel, s, n, "tail-batch")
        ts_, _ = timeit(sc)
        z = lambda: g_ent.zero_()
        tz, _ = timeit(z)
        mm, vv = torch.zeros_like(ent), torch.zeros_like(ent)
        ad = lambda: ops.adam_step(ent, g_ent, mm, vv, 1, 1e-4, zero_grad=True)
        ta, _ = timeit(ad)
        print(f"   unfused score fwd {ts_*1e3:.1f} us | grad memset {tz*1e3:.1f} us | adam+zero {ta*1e3:.1f} us "
              f"({7*ent.numel()*4/ta/1e6:.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
