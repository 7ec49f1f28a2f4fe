"""Regression tests for the round-1 advisor findings: unaligned relation-gradient offset in the
flat gradient buffer, ``seed=None`` samplers, optimizer hyper-parameters edited between ``learn()`` calls."""
import numpy as np
import pytest
import torch

from conftest import DEV
from mkb_b200 import compose, datasets, losses, models, optim, sampling

pytestmark = pytest.mark.gpu


def _graph(Nn=121, R=3, T=700, seed=4):
    rng = np.random.RandomState(seed)
    tri = sorted({(int(rng.randint(Nn)), int(rng.randint(R)), int(rng.randint(Nn))) for _ in range(T)})
    return tri, {i: i for i in range(Nn)}, {i: i for i in range(R)}


def _learn(route, model_name, D, seed=42, lr=0.01, epochs=1, opt=None, m=None, pipe=None, ns=None):
    tri, ents, rels = _graph()
    ds = datasets.Dataset(train=tri, entities=ents, relations=rels, batch_size=32, shuffle=False, seed=42)
    if m is None:
        torch.manual_seed(7)
        m = getattr(models, model_name)(hidden_dim=D, entities=ents, relations=rels, gamma=6).to(DEV)
        ns = sampling.NegativeSampling(size=8, train_triples=tri, entities=ents, relations=rels, seed=seed)
        params = [p for p in m.parameters() if p.requires_grad]
        opt = optim.DenseAdam(params, lr=lr) if route == "device" else torch.optim.Adam(params, lr=lr)
        pipe = compose.Pipeline(epochs=epochs, device=DEV, fused=(route != "generic"))
    pipe.learn(model=m, dataset=ds, sampling=ns, optimizer=opt, loss=losses.Adversarial(0.5))
    return m, opt, pipe, ns


@pytest.mark.parametrize("model_name,D", [("TransE", 50), ("DistMult", 5), ("RotatE", 3)])
def test_device_route_with_unaligned_table_sizes(model_name, D):
    """(N * entity_dim) % 4 != 0: the relation gradient must still start on a 16-byte boundary (ADVICE high)."""
    m_dev, _, pipe, _ = _learn("device", model_name, D)
    assert getattr(pipe, "_trainer", None) is not None
    assert m_dev.entity_embedding.numel() % 4 != 0
    assert pipe._trainer.g_rel.data_ptr() % 16 == 0
    m_ref, _, _, _ = _learn("generic", model_name, D)
    torch.testing.assert_close(m_dev.entity_embedding, m_ref.entity_embedding, rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(m_dev.relation_embedding, m_ref.relation_embedding, rtol=2e-3, atol=2e-4)


def test_seed_none_sampler_runs_on_both_routes():
    """NegativeSampling(seed=None) is legal in the reference (RandomState(None)); KdmkbModel's default."""
    for route in ("fused", "device"):
        m, _, pipe, ns = _learn(route, "TransE", 8, seed=None)
        assert isinstance(ns.seed, int) and np.isfinite(pipe.metric_loss.get())
    a = sampling.NegativeSampling(size=4, train_triples=_graph()[0], entities=range(121), relations=range(3), seed=None)
    b = sampling.NegativeSampling(size=4, train_triples=_graph()[0], entities=range(121), relations=range(3), seed=None)
    assert a.seed != b.seed  # fresh entropy per sampler, like RandomState(None)


def test_learning_rate_edits_between_learn_calls_take_effect():
    """The cached DeviceTrainer re-reads lr / betas / eps from the optimizer (LR schedulers, manual edits)."""
    m, opt, pipe, ns = _learn("device", "TransE", 8)
    before = m.entity_embedding.detach().clone()
    opt.param_groups[0]["lr"] = 0.0
    _learn("device", "TransE", 8, opt=opt, m=m, pipe=pipe, ns=ns)
    torch.testing.assert_close(m.entity_embedding, before, rtol=0, atol=0)
    opt.param_groups[0]["lr"] = 0.01
    _learn("device", "TransE", 8, opt=opt, m=m, pipe=pipe, ns=ns)
    assert (m.entity_embedding - before).abs().max().item() > 1e-4
