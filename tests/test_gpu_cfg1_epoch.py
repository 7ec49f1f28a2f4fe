"""BASELINE config 1 as worded — "Wn18rr TransE dim=200 batch=256 neg=64 Adversarial loss on CPU (reference
compose.Pipeline, 1 epoch)" — replayed on the CUDA path against the UNMODIFIED reference's own run of that epoch
(tests/golden/cfg1_epoch.npz, `make_golden.py cfg1`: mkb.datasets.Wn18rr -> models.TransE -> NegativeSampling ->
Adversarial -> torch.optim.Adam under mkb.compose.Pipeline(epochs=1).learn, 680 steps).

Same torch seed => same table init and the same shuffled loader order (both golden-tested elsewhere); the sampler
runs with pool="reference" (the reference's shared RandomState pool, bit-exact negatives).  Every route of
compose.Pipeline must land on the reference's trajectory: the per-step losses (generic route), the rolling loss, the
trained relation table and 512 sampled rows of the trained entity table."""
import numpy as np
import pytest
import torch

from conftest import DEV, load_golden
from mkb_b200 import compose, datasets, losses, models, optim, sampling

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def graph():
    w = load_golden("cfg5_wn18rr.npz")  # carries the real Wn18rr triples
    as_list = lambda a: [tuple(r) for r in a.astype(np.int64).tolist()]
    return as_list(w["train"]), as_list(w["valid"]), as_list(w["test"]), int(w["n_entity"]), int(w["n_relation"])


@pytest.mark.parametrize("route", ("generic", "fused", "device", "adopted"))
def test_cfg1_epoch_lands_on_the_reference_trajectory(graph, route):
    g = load_golden("cfg1_epoch.npz")
    train, valid, test, N, R = graph
    D, B, K, seed = int(g["hidden_dim"]), int(g["batch"]), int(g["neg"]), int(g["seed"])
    ents, rels = {i: i for i in range(N)}, {i: i for i in range(R)}
    torch.manual_seed(seed)
    ds = datasets.Dataset(train=train, valid=valid, test=test, entities=ents, relations=rels, batch_size=B,
                          shuffle=True, seed=seed, pin_memory=(route == "device"))
    model = models.TransE(hidden_dim=D, entities=ents, relations=rels, gamma=float(g["gamma"]))
    assert abs(model.entity_embedding.detach().double().sum().item() - float(g["ent0_checksum"])) < 1e-6, "init differs"
    model = model.to(DEV)
    sampler = sampling.NegativeSampling(size=K, train_triples=train, entities=ents, relations=rels, seed=seed,
                                        pool="reference")
    params = [p for p in model.parameters() if p.requires_grad]
    opt = optim.DenseAdam(params, lr=float(g["lr"])) if route == "device" else torch.optim.Adam(params, lr=float(g["lr"]))
    seen = []

    class RecLoss(losses.Adversarial):  # the generic route calls the loss object once per step
        def __call__(self, positive_score, negative_score, weight):
            err = super().__call__(positive_score, negative_score, weight)
            seen.append(err.detach())
            return err

    # generic: three model / loss calls + autograd; fused: one fused autograd step; both step a stock torch.optim.Adam.
    # device: the device-resident step with optim.DenseAdam; adopted: the same step taking over a torch.optim.Adam
    pipe = compose.Pipeline(epochs=1, device=DEV, fused=(route != "generic"), adopt_torch_adam=(route == "adopted"))
    pipe.learn(model=model, dataset=ds, sampling=sampler, optimizer=opt,
               loss=RecLoss(0.5) if route == "generic" else losses.Adversarial(0.5))
    ref_losses = g["losses"]
    if route == "generic":
        got = torch.stack(seen).cpu().numpy().astype(np.float64)
        assert got.shape == ref_losses.shape == (680,)
        np.testing.assert_allclose(got, ref_losses, rtol=2e-5)
    if route in ("device", "adopted"):
        assert getattr(pipe, "_trainer", None) is not None and pipe._trainer.t == 680
    else:
        assert getattr(pipe, "_trainer", None) is None
    assert abs(pipe.metric_loss.get() - float(g["rolling_loss"])) < 2e-5
    ent = model.entity_embedding.detach().cpu().numpy()
    rel = model.relation_embedding.detach().cpu().numpy()
    for got_t, ref_t, what in ((rel, g["rel_final"], "relation table"), (ent[g["rows"]], g["ent_rows_final"], "entity rows")):
        err = np.abs(got_t - ref_t)
        # Adam turns a flipped sign of a ~0 L1 residual into a +-lr step: isolated elements may sit a few lr apart
        assert err.max() <= 5e-4, (what, err.max())
        assert (err <= 1e-5).mean() >= 0.999, (what, (err <= 1e-5).mean())
    assert abs(np.abs(ent).astype(np.float64).sum() - float(g["ent_abs_checksum"])) <= 1e-4 * float(g["ent_abs_checksum"])
