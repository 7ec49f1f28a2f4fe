"""Full-shape oracle spot checks for the BASELINE configs that had none in round 1 (VERDICT weak #1):
cfg1 (Wn18rr TransE D=200 B=256 K=64), cfg3 (FB15k-237 ComplEx D=1000 B=1024 K=256, the gather path) and
cfg4 (Yago3-10 RotatE D=500 B=1024 K=256, N = 123 182), both modes.  Same recipe as
test_gpu_parity.py::test_full_size_rows_vs_oracle (cfg2): the whole batch runs through the fused kernels at
the real shape; the fp64 oracle re-scores a random subset of positives, re-derives the loss from all scores,
and checks the gradients of a sub-batch."""
import numpy as np
import pytest
import torch

from conftest import DEV, score_tol
from mkb_b200 import models, ops
from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

SHAPES = {  # name: (model, N, R, D, B, K, gamma)
    "cfg1": ("TransE", 40943, 11, 200, 256, 64, 6.0),
    "cfg3": ("ComplEx", 14541, 237, 1000, 1024, 256, 9.0),
    "cfg4": ("RotatE", 123182, 37, 500, 1024, 256, 24.0),
}


def _close(a, ref, rel=1e-4):
    a = a.detach().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, np.float64)
    bad = np.abs(a - ref) > score_tol(ref, rel)
    assert not bad.any(), f"{bad.sum()} / {bad.size} outside tol; max abs err {np.abs(a - ref).max():.3e}"


def _grad_close(a, ref, rel=1e-4):
    a = a.detach().cpu().numpy().astype(np.float64)
    err = np.abs(a - ref).max()
    assert err <= rel * max(np.abs(ref).max(), 1e-30), f"grad max err {err:.3e} vs scale {np.abs(ref).max():.3e}"


@pytest.mark.parametrize("cfg", sorted(SHAPES))
@pytest.mark.parametrize("mode", ("tail-batch", "head-batch"))
def test_full_shape_rows_vs_oracle(cfg, mode):
    name, Nn, R, D, B, K, gamma = SHAPES[cfg]
    torch.manual_seed(42)
    ents, rels = {i: i for i in range(Nn)}, {i: i for i in range(R)}
    m = getattr(models, name)(hidden_dim=D, entities=ents, relations=rels, gamma=gamma).to(DEV)
    with torch.no_grad():  # widen the init so that softmax weights differ and scores are not all ~gamma
        m.entity_embedding.mul_(3.0)
        m.relation_embedding.mul_(3.0)
    g = torch.Generator().manual_seed(43)
    s = torch.stack([torch.randint(Nn, (B,), generator=g), torch.randint(R, (B,), generator=g),
                     torch.randint(Nn, (B,), generator=g)], 1)
    n = torch.randint(Nn, (B, K), generator=g)
    side = 2 if mode == "tail-batch" else 0
    n[:, 0] = s[:, side]  # candidate == the positive's own entity must reproduce the positive score
    w = torch.rand(B, generator=g) * 0.4 + 0.1
    sd, nd, wd = s.to(DEV), n.to(DEV), w.to(DEV)
    loss, ps, ns = ops.fused_adversarial_step(m.spec, m.entity_embedding, m.relation_embedding, sd, nd, wd, mode, 0.5,
                                              return_scores=True)
    loss.backward()
    torch.testing.assert_close(ns[:, 0], ps[:, 0], rtol=1e-5, atol=1e-5)
    # unfused route (models.*.forward) == fused
    with torch.no_grad():
        torch.testing.assert_close(m(sd, nd, mode), ns, rtol=1e-6, atol=1e-6)
    rows = np.random.RandomState(3).choice(B, 12, replace=False)
    ent, rel = m.entity_embedding.detach().cpu().numpy(), m.relation_embedding.detach().cpu().numpy()
    _close(ns[rows], ko.score(name, ent, rel, s.numpy()[rows], n.numpy()[rows], mode, gamma=gamma))
    _close(ps[rows], ko.score(name, ent, rel, s.numpy()[rows], gamma=gamma))
    ref_loss = ko.adversarial_loss(ps.cpu().numpy(), ns.cpu().numpy(), w.numpy(), 0.5)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    # gradients of a sub-batch against the oracle's closed forms (fp64)
    sub = rows[:4]
    m2 = getattr(models, name)(hidden_dim=D, entities=ents, relations=rels, gamma=gamma).to(DEV)
    m2._set_params(m.entity_embedding.detach(), m.relation_embedding.detach())
    l2 = ops.fused_adversarial_step(m2.spec, m2.entity_embedding, m2.relation_embedding, sd[sub], nd[sub], wd[sub],
                                    mode, 0.5)
    l2.backward()
    _, _, _, ge, gr = ko.train_step(name, ent, rel, s.numpy()[sub], n.numpy()[sub], mode, w.numpy()[sub], gamma=gamma)
    _grad_close(m2.entity_embedding.grad, ge)
    _grad_close(m2.relation_embedding.grad, gr)
    # the full-batch dense gradient has the rows it must have: every sampled id and both sides of every positive
    touched = torch.zeros(Nn, dtype=torch.bool)
    touched[n.flatten()] = True
    touched[s[:, 0]] = True
    touched[s[:, 2]] = True
    nz = (m.entity_embedding.grad.abs().sum(1) > 0).cpu()
    assert not (nz & ~touched).any()
