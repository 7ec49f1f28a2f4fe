import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mkb_b200 import models, ops, sampling
dev = "cuda"
rng = np.random.RandomState(1)
Nn, R, D, B, K = 700, 9, 512, 48, 40
ents, rels = {i: i for i in range(Nn)}, {i: i for i in range(R)}
for name in ("TransE", "ComplEx", "RotatE"):
    torch.manual_seed(3)
    m = getattr(models, name)(hidden_dim=D, entities=ents, relations=rels, gamma=9.0).to(dev)
    ent, rel = m.entity_embedding.detach(), m.relation_embedding.detach()
    s = torch.from_numpy(np.stack([rng.randint(Nn, size=B), rng.randint(R, size=B), rng.randint(Nn, size=B)], 1)).to(dev)
    n = torch.from_numpy(rng.randint(Nn, size=(B, K))).to(dev)
    w = torch.full((B,), 0.25, device=dev)
    cp, cn = torch.empty(B, device=dev), torch.empty(B, K, device=dev)
    stats, ws = torch.zeros(4, device=dev), torch.zeros(1 << 16, dtype=torch.uint8, device=dev)
    for mode in ("tail-batch", "head-batch"):
        ops.fused_forward_raw(m.spec, ent, rel, s, n, w, mode, 0.5, cp, cn, stats, ws)
        ge, gr = torch.zeros_like(ent), torch.zeros_like(rel)
        ops.fused_backward_raw(m.spec, ent, rel, s, n, mode, cp, cn, stats, ge, gr)
        nc, rc = ent.shape[1] // D, rel.shape[1] // D
        for col, wd in ((0, 128), (128, 128), (256, 128), (384, 128)):
            gec = torch.zeros(Nn, nc * wd, device=dev); grc = torch.zeros(R, rc * wd, device=dev)
            ops.fused_backward_chunk_raw(m.spec, ent, rel, s, n, mode, cp, cn, stats, col, wd, gec, grc)
            ref_e = ge.view(Nn, nc, D)[:, :, col:col + wd].reshape(Nn, nc * wd)
            ref_r = gr.view(R, rc, D)[:, :, col:col + wd].reshape(R, rc * wd)
            print(name, mode, col, "ent err %.3e (scale %.3e) rel err %.3e (scale %.3e)" % (
                (gec - ref_e).abs().max().item(), ref_e.abs().max().item(), (grc - ref_r).abs().max().item(), ref_r.abs().max().item()))
        # adam chunk vs full adam
        p1, p2 = ent.clone(), ent.clone()
        m1, v1, m2, v2 = (torch.zeros_like(ent) for _ in range(4))
        g1 = ge.clone()
        ops.adam_step(p1, g1, m1, v1, 1, 1e-3, zero_grad=True)
        for col, wd in ((0, 128), (128, 128), (256, 128), (384, 128)):
            gec = ge.view(Nn, nc, D)[:, :, col:col + wd].reshape(Nn, nc * wd).contiguous()
            ops.adam_step_chunk(p2, gec, m2, v2, nc, wd, col, D, 1, 1e-3)
        print(name, mode, "adam chunk vs full: param err %.3e m err %.3e" % ((p1 - p2).abs().max().item(), (m1 - m2).abs().max().item()))
