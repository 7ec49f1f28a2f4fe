"""K2-TMA (csrc/score_tma.cuh): the fused forward with candidate rows staged through cp.async.bulk must be
BIT-IDENTICAL to the LDG kernel (same per-lane arithmetic in the same order) — scores, coefficients, loss sums —
at the BASELINE shapes and at ragged ones (K not a multiple of 8 or 32, more / fewer positives than resident CTAs,
every ring depth).  The backward variant is compared with the scatter kernel within atomic-order noise."""
import os

import numpy as np
import pytest
import torch

from conftest import DEV, EMU
from mkb_b200 import models, ops

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(EMU, reason="TMA / mbarrier PTX is not emulated")]


class _env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _problem(name, Nn, R, D, B, K, seed=0):
    torch.manual_seed(seed)
    m = getattr(models, name)(hidden_dim=D, entities={i: i for i in range(Nn)}, relations={i: i for i in range(R)},
                              gamma=9.0).to(DEV)
    with torch.no_grad():
        m.entity_embedding.mul_(3.0)
        m.relation_embedding.mul_(3.0)
    g = torch.Generator().manual_seed(seed + 1)
    s = torch.stack([torch.randint(Nn, (B,), generator=g), torch.randint(R, (B,), generator=g),
                     torch.randint(Nn, (B,), generator=g)], 1).to(DEV)
    n = torch.randint(Nn, (B, K), generator=g).sort(dim=1).values.to(DEV)
    w = (torch.rand(B, generator=g) * 0.4 + 0.1).to(DEV)
    return m, s, n, w


def _forward(m, s, n, w, mode):
    B, K = n.shape
    cp, cn = torch.empty(B, device=DEV), torch.empty(B, K, device=DEV)
    ps, ns = torch.empty(B, 1, device=DEV), torch.empty(B, K, device=DEV)
    stats = torch.zeros(4, device=DEV)
    ws = torch.zeros(max(ops.N.load().kge_loss_workspace_bytes(B), 64), dtype=torch.uint8, device=DEV)
    ops.fused_forward_raw(m.spec, m.entity_embedding.data, m.relation_embedding.data, s, n, w, mode, 0.5, cp, cn,
                          stats, ws, ps, ns)
    torch.cuda.synchronize()
    assert int(ws[4:8].view(torch.int32).item()) == 0, "a TMA wait timed out"
    return cp, cn, ps, ns, stats


SHAPES = [  # model, N, R, D, B, K
    ("RotatE", 14541, 237, 1000, 1024, 256),   # config 2
    ("ComplEx", 14541, 237, 1000, 300, 256),   # config 3's rows, B not a multiple of the grid
    ("TransE", 40943, 11, 200, 256, 64),       # config 1
    ("RotatE", 5000, 37, 500, 100, 250),       # config 4's dim; K % 8 != 0
    ("DistMult", 3000, 5, 1000, 7, 50),        # fewer positives than SMs, ragged K
    ("RotatE", 3000, 5, 64, 500, 33),          # tiny rows (D < 128: idle lanes), K = 33
    ("TransE", 3000, 5, 1024, 40, 8),          # the largest supported dim, one row per warp
]


@pytest.mark.parametrize("name,Nn,R,D,B,K", SHAPES)
@pytest.mark.parametrize("mode", ("tail-batch", "head-batch"))
def test_tma_forward_is_bit_identical(name, Nn, R, D, B, K, mode):
    m, s, n, w = _problem(name, Nn, R, D, B, K)
    with _env(KGE_FWD_TMA=0):
        ref = _forward(m, s, n, w, mode)
    for minb, stages in ((1, 0), (1, 1), (1, 2), (2, 0)):
        with _env(KGE_FWD_TMA=1, KGE_TMA_MINB=minb, KGE_TMA_STAGES=stages):
            got = _forward(m, s, n, w, mode)
        for a, b, what in zip(got, ref, ("coef_pos", "coef_neg", "pos_score", "neg_score", "stats")):
            assert torch.equal(a, b), (what, minb, stages, (a - b).abs().max().item())


def _backward(m, s, n, mode, cp, cn, stats):
    ge, gr = torch.zeros_like(m.entity_embedding.data), torch.zeros_like(m.relation_embedding.data)
    ops.fused_backward_raw(m.spec, m.entity_embedding.data, m.relation_embedding.data, s, n, mode, cp, cn, stats, ge, gr)
    torch.cuda.synchronize()
    return ge, gr


@pytest.mark.parametrize("name,Nn,R,D,B,K", SHAPES)
@pytest.mark.parametrize("mode", ("tail-batch", "head-batch"))
def test_tma_backward_matches_scatter_kernel(name, Nn, R, D, B, K, mode):
    """K3-TMA (bulk loads in, cp.reduce.async.bulk adds out) against K3 (LDG + RED.128): same gradients up to
    the order in which floating-point adds land (both are atomic scatters)."""
    m, s, n, w = _problem(name, Nn, R, D, B, K, seed=3)
    with _env(KGE_FWD_TMA=0):
        cp, cn, _, _, stats = _forward(m, s, n, w, mode)
    with _env(KGE_BWD_TMA=0, KGE_BWD_BULK=0):
        ge_ref, gr_ref = _backward(m, s, n, mode, cp, cn, stats)
    for variant in ({"KGE_BWD_TMA": 1, "KGE_TMA_STAGES": 0}, {"KGE_BWD_TMA": 1, "KGE_TMA_STAGES": 1},
                    {"KGE_BWD_BULK": 1, "KGE_BWD_TMA": 0}):  # K3-TMA (two ring depths), K3b (LDG in, bulk reduce out)
        stages = variant
        with _env(**variant):
            ge, gr = _backward(m, s, n, mode, cp, cn, stats)
        assert ops.N.load().kge_tma_fail_flag() == 0, "a TMA wait timed out"
        for got, ref, what in ((ge, ge_ref, "entity"), (gr, gr_ref, "relation")):
            scale = ref.abs().max().item()
            err = (got - ref).abs().max().item()
            assert err <= 2e-5 * scale, (what, stages, err, scale)
        # rows nobody touched stay exactly zero (a bulk reduction must not spill over its row)
        touched = torch.zeros(Nn, dtype=torch.bool, device=DEV)
        touched[n.flatten()] = True
        touched[s[:, 0]] = True
        touched[s[:, 2]] = True
        assert not ge[~touched].any()
