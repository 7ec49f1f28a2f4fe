"""The atomics-free by-entity backward with Adam fused in (csrc/byent.cu, DeviceTrainer(backward="by_entity"))
against the default scatter path: same training trajectory up to summation order, and bit-reproducible."""
import numpy as np
import pytest
import torch

from conftest import DEV

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from mkb_b200 import models, sampling
    from mkb_b200.compose import DeviceTrainer


def _run(model, backward, steps=6, D=64):
    Nn, R, B, K, gamma = 800, 7, 48, 32, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=8000), rng.randint(R, size=8000), rng.randint(Nn, size=8000)], 1), axis=0)
    w_all = torch.from_numpy(rng.uniform(0.1, 0.5, len(tri)).astype(np.float32)).to(DEV)
    T = torch.from_numpy(tri).to(DEV)
    torch.manual_seed(3)
    m = getattr(models, model)(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                               gamma=gamma).to(DEV)
    init = m.entity_embedding.detach().clone()
    ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=7)
    tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, backward=backward)
    assert tr.backward == backward
    losses = []
    for step in range(steps):
        idx = torch.arange(step * B, (step + 1) * B, device=DEV)
        tr.step(T[idx], w_all[idx], "head-batch" if step % 2 == 0 else "tail-batch")
        losses.append(tr.loss())
    ns.check_status(DEV)
    return m, init, losses, tr


@pytest.mark.parametrize("model", ("RotatE", "ComplEx", "TransE", "DistMult", "pRotatE"))
def test_by_entity_trainer_tracks_scatter_trainer(model):
    m0, init, l0, _ = _run(model, "scatter")
    m1, _, l1, tr = _run(model, "by_entity")
    assert np.allclose(l0, l1, rtol=1e-4), (l0, l1)
    u0, u1 = m0.entity_embedding.detach() - init, m1.entity_embedding.detach() - init
    assert u0.abs().max().item() > 0
    assert ((u0 - u1).abs() > 0.05 * u0.abs().max()).float().mean().item() < 1e-3
    r0, r1 = m0.relation_embedding.detach(), m1.relation_embedding.detach()
    assert ((r0 - r1).abs() > 3e-4).float().mean().item() < 1e-2
    assert torch.count_nonzero(tr.g_ent).item() == 0  # the dense entity gradient is never touched
    if model == "pRotatE":
        assert abs(m0.modulus.item() - m1.modulus.item()) <= 1e-5


def test_by_entity_training_is_bit_reproducible():
    a, _, la, _ = _run("RotatE", "by_entity", steps=4)
    b, _, lb, _ = _run("RotatE", "by_entity", steps=4)
    assert torch.equal(a.entity_embedding, b.entity_embedding)
    assert torch.equal(a.relation_embedding, b.relation_embedding)
    assert la == lb


def test_by_entity_full_size_step_matches_scatter():
    """Config-2 shapes (N=14541, D=1000, B=1024, K=256): one by-entity step == one scatter + Adam step."""
    from mkb_b200 import ops

    Nn, R, D, B, K, gamma = 14541, 237, 1000, 1024, 256, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=60000), rng.randint(R, size=60000), rng.randint(Nn, size=60000)], 1), axis=0)
    s = torch.from_numpy(tri[:B]).to(DEV)
    w = torch.from_numpy(rng.uniform(0.1, 0.5, B).astype(np.float32)).to(DEV)
    out = []
    for backward in ("scatter", "by_entity"):
        torch.manual_seed(1)
        m = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                          gamma=gamma).to(DEV)
        init = m.entity_embedding.detach().clone()
        ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=3)
        tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, backward=backward)
        tr.step(s, w, "tail-batch")
        tr.step(s, w, "head-batch")
        out.append((m.entity_embedding.detach() - init, m.relation_embedding.detach().clone(), tr.loss()))
    (u0, r0, l0), (u1, r1, l1) = out
    assert abs(l0 - l1) <= 1e-4 * abs(l0)
    assert u0.abs().max().item() > 0
    assert ((u0 - u1).abs() > 0.05 * u0.abs().max()).float().mean().item() < 1e-3
    assert ((r0 - r1).abs() > 3e-4).float().mean().item() < 1e-2


@pytest.mark.parametrize("model", ("ComplEx", "DistMult"))
def test_pooled_gemm_trainer_tracks_gather_trainer(model):
    """pool='reference' + pooled_gemm=True (S = Q·Pool^T on the tensor cores, csrc/pooled.cu) follows the
    gather kernels on the same pools: same losses, same updates up to summation order / 3xTF32 rounding."""
    Nn, R, D, B, K, gamma = 800, 7, 64, 48, 32, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=8000), rng.randint(R, size=8000), rng.randint(Nn, size=8000)], 1), axis=0)
    w_all = torch.from_numpy(rng.uniform(0.1, 0.5, len(tri)).astype(np.float32)).to(DEV)
    T = torch.from_numpy(tri).to(DEV)
    out = []
    for pooled in (False, True):
        torch.manual_seed(3)
        m = getattr(models, model)(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                                   gamma=gamma).to(DEV)
        with torch.no_grad():
            m.entity_embedding.mul_(3.0)
        init = m.entity_embedding.detach().clone()
        ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=7,
                                       pool="reference")
        tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, pooled_gemm=pooled)
        assert tr.pooled_gemm == pooled
        losses = []
        for step in range(5):
            idx = torch.arange(step * B, (step + 1) * B, device=DEV)
            tr.step(T[idx], w_all[idx], "head-batch" if step % 2 == 0 else "tail-batch")
            losses.append(tr.loss())
        ns.check_status(DEV)
        out.append((m.entity_embedding.detach() - init, m.relation_embedding.detach().clone(), losses))
    (u0, r0, l0), (u1, r1, l1) = out
    assert np.allclose(l0, l1, rtol=1e-4), (l0, l1)
    assert u0.abs().max().item() > 0
    assert ((u0 - u1).abs() > 0.05 * u0.abs().max()).float().mean().item() < 1e-3
    assert ((r0 - r1).abs() > 3e-4).float().mean().item() < 1e-2
    # a distance model silently keeps the gather kernels
    m = models.RotatE(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                      gamma=gamma).to(DEV)
    assert not DeviceTrainer(m, ns, max_batch=B, pooled_gemm=True).pooled_gemm


def test_pooled_gemm_full_size_step_matches_gather():
    """Config-3 shapes (ComplEx D=1000, B=1024, K=256, pool of 512): the three-GEMM step == the gather step."""
    Nn, R, D, B, K, gamma = 14541, 237, 1000, 1024, 256, 9.0
    rng = np.random.RandomState(0)
    tri = np.unique(np.stack([rng.randint(Nn, size=60000), rng.randint(R, size=60000), rng.randint(Nn, size=60000)], 1), axis=0)
    s = torch.from_numpy(tri[:B]).to(DEV)
    w = torch.from_numpy(rng.uniform(0.1, 0.5, B).astype(np.float32)).to(DEV)
    out = []
    for pooled in (False, True):
        torch.manual_seed(1)
        m = models.ComplEx(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                           gamma=gamma).to(DEV)
        with torch.no_grad():
            m.entity_embedding.mul_(20.0)  # scores of order 1 so the softmax weights are not uniform
            m.relation_embedding.mul_(20.0)
        init = m.entity_embedding.detach().clone()
        ns = sampling.NegativeSampling(size=K, train_triples=tri, entities=range(Nn), relations=range(R), seed=3,
                                       pool="reference")
        tr = DeviceTrainer(m, ns, lr=1e-3, max_batch=B, pooled_gemm=pooled)
        tr.step(s, w, "tail-batch")
        l_first = tr.loss()
        tr.step(s, w, "head-batch")
        out.append((m.entity_embedding.detach() - init, l_first, tr.loss()))
    (u0, a0, b0), (u1, a1, b1) = out
    assert abs(a0 - a1) <= 5e-5 * abs(a0) and abs(b0 - b1) <= 2e-4 * abs(b0)
    assert u0.abs().max().item() > 0
    assert ((u0 - u1).abs() > 0.05 * u0.abs().max()).float().mean().item() < 1e-3


def test_kl_divergence_full_size_properties():
    """[1024, 14541] score matrices: KL >= 0, == 0 for identical inputs, invariant to per-row shifts, and both
    gradients sum to zero along every row (softmax is shift-invariant)."""
    from mkb_b200 import losses

    gen = torch.Generator(device=DEV).manual_seed(0)
    s = (3 * torch.randn(1024, 14541, device=DEV, generator=gen)).requires_grad_()
    t = (3 * torch.randn(1024, 14541, device=DEV, generator=gen)).requires_grad_()
    kl = losses.KlDivergence()
    loss = kl(s, t, T=2)
    loss.backward()
    assert loss.item() > 0
    assert kl(s.detach(), s.detach(), T=2).item() == pytest.approx(0.0, abs=1e-9)
    shift = torch.randn(1024, 1, device=DEV, generator=gen)
    assert kl(s.detach() + shift, t.detach() - shift, T=2).item() == pytest.approx(loss.item(), rel=1e-4)
    assert s.grad.sum(1).abs().max().item() <= 1e-3 * s.grad.abs().sum(1).max().item()
    assert t.grad.sum(1).abs().max().item() <= 1e-3 * t.grad.abs().sum(1).max().item()
    ref = torch.mean(torch.nn.functional.kl_div(torch.log_softmax(s.detach() / 2, 1), torch.softmax(t.detach() / 2, 1),
                                                reduction="none"))
    assert loss.item() == pytest.approx(ref.item(), rel=1e-4)
