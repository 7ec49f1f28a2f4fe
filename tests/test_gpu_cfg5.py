"""BASELINE config 5 AT SIZE: filtered ranking on the real Wn18rr graph (N = 40 943 entities, 11 relations,
filter = train + valid + test = 93 003 triples) at D = 1000, against the REFERENCE's own ranks.

tests/golden/cfg5_wn18rr.npz (made by `python tests/golden/make_golden.py cfg5` where /root/reference exists)
holds the graph, 64 test queries per model (32 for RotatE: the reference needs 17 s per ranking there) and, for
both modes, the ranks the unmodified mkb.evaluation.Evaluation.compute_score produced (evaluation.py:217-279
over datasets/base.py:196-241's candidate / filter_bias lists), the fp64 oracle's ranks, and per query the number
of unfiltered candidates within 1e-6 / 1e-5 / 1e-4 (relative) of the positive's score.  Tables are rebuilt from a
seed on both sides (tests/golden/cfg5_tables.py; checksums pinned).

Bar: ranks identical to the reference's, except where the fp64 oracle sees candidates within 1e-5 relative of the
positive (then within that count); at least 90 % of the ranks must be identical outright, and the aggregated
metrics must match the reference's Evaluation.eval output."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import DEV, GOLDEN, MODELS, load_golden
from mkb_b200 import evaluation, models

sys.path.insert(0, GOLDEN)
import cfg5_tables  # noqa: E402

pytestmark = pytest.mark.gpu
MARGIN_COL = 1  # contested[:, 1] = within 1e-5 relative


@pytest.fixture(scope="module")
def cfg5():
    g = load_golden("cfg5_wn18rr.npz")
    N, R = int(g["n_entity"]), int(g["n_relation"])
    train = g["train"].astype(np.int64)
    true = np.concatenate([train, g["valid"].astype(np.int64), g["test"].astype(np.int64)])
    ev = evaluation.Evaluation(entities={i: i for i in range(N)}, relations={i: i for i in range(R)}, batch_size=8,
                               true_triples=[tuple(r) for r in true.tolist()], device=DEV)
    return g, train, N, R, ev


@pytest.mark.parametrize("name", MODELS)
def test_cfg5_ranks_match_reference_at_size(cfg5, name):
    g, train, N, R, ev = cfg5
    D, gamma = int(g["hidden_dim"]), float(g[f"{name}/gamma"])
    ent, rel = cfg5_tables.make_tables(name, train, N, R, D, gamma)
    assert abs(ent.astype(np.float64).sum() - float(g[f"{name}/ent_checksum"])) < 1e-9, "tables differ from the fixture's"
    assert abs(rel.astype(np.float64).sum() - float(g[f"{name}/rel_checksum"])) < 1e-9
    m = getattr(models, name)(hidden_dim=D, entities={i: i for i in range(N)}, relations={i: i for i in range(R)},
                              gamma=gamma)
    m._set_params(torch.from_numpy(ent), torch.from_numpy(rel))
    m = m.to(DEV).eval()
    queries = [tuple(r) for r in g[f"{name}/queries"].tolist()]
    all_ranks, identical, total = [], 0, 0
    for mi, mode in enumerate(("head-batch", "tail-batch")):
        got = ev.ranks(m, queries, mode).cpu().numpy()
        ref = g[f"{name}/ref_ranks"][mi]
        r64 = g[f"{name}/{mode}/rank64"]
        contested = g[f"{name}/{mode}/contested"][:, MARGIN_COL]
        # within the near-tie band of the exact (fp64) ranking, and of the reference's own fp32 ranking
        assert np.all(np.abs(got - r64) <= contested), (name, mode, got[np.abs(got - r64) > contested], r64)
        assert np.all(np.abs(got - ref) <= contested + np.abs(ref - r64)), (name, mode)
        identical += int((got == ref).sum())
        total += len(ref)
        all_ranks.append(got)
    frac = identical / total
    print(f"cfg5 {name}: {identical}/{total} ranks identical to the reference's ({100 * frac:.1f} %)")
    assert frac >= 0.90
    # Evaluation.eval (head-batch stream then tail-batch stream, evaluation.py:185-199) vs the reference's metrics
    out = ev.eval(m, queries)
    ref_m = g[f"{name}/ref_metrics"]
    r = np.concatenate(all_ranks).astype(np.float64)
    exp = np.array([np.mean(1 / r), np.mean(r), np.mean(r <= 1), np.mean(r <= 3), np.mean(r <= 10)])
    np.testing.assert_allclose([out[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")], np.round(exp, 4), atol=1e-4)
    if frac == 1.0:
        np.testing.assert_allclose([out[k] for k in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")], ref_m, atol=1e-4)
    else:  # a flipped near-tie moves MR by <= 1/len and MRR by less
        np.testing.assert_allclose(out["MR"], ref_m[1], atol=1.0)
        np.testing.assert_allclose(out["MRR"], ref_m[0], atol=2e-2)
