"""CPU pinning of the config-5-at-size fixture (tests/golden/cfg5_wn18rr.npz): the graph is the real Wn18rr, the
seeded tables reproduce bit-identically, the fp64 oracle reproduces its stored ranks on a sample of queries, and
the reference's fp32 ranks sit within the oracle's near-tie band everywhere."""
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import kge_oracle as ko

sys.path.insert(0, GOLDEN)
import cfg5_tables  # noqa: E402


@pytest.fixture(scope="module")
def g():
    return load_golden("cfg5_wn18rr.npz")


def test_graph_is_wn18rr(g):
    # mkb/datasets/wn18rr.py:44-49
    assert (int(g["n_entity"]), int(g["n_relation"])) == (40943, 11)
    assert (len(g["train"]), len(g["valid"]), len(g["test"])) == (86835, 3034, 3134)
    assert int(g["train"][:, [0, 2]].max()) < 40943 and int(g["train"][:, 1].max()) == 10


def test_reference_ranks_within_oracle_band(g):
    for name in ("TransE", "DistMult", "ComplEx", "RotatE"):
        for mi, mode in enumerate(("head-batch", "tail-batch")):
            ref, r64 = g[f"{name}/ref_ranks"][mi], g[f"{name}/{mode}/rank64"]
            band = g[f"{name}/{mode}/contested"][:, 1]
            assert np.all(np.abs(ref - r64) <= band), (name, mode)
            assert (ref == r64).mean() >= 0.9
            assert np.all(np.diff(g[f"{name}/{mode}/contested"], axis=1) >= 0)  # wider margin, more candidates


def test_tables_and_oracle_ranks_reproduce(g):
    name = "TransE"
    N, R, D, gamma = int(g["n_entity"]), int(g["n_relation"]), int(g["hidden_dim"]), float(g[f"{name}/gamma"])
    train = g["train"].astype(np.int64)
    ent, rel = cfg5_tables.make_tables(name, train, N, R, D, gamma)
    assert abs(ent.astype(np.float64).sum() - float(g[f"{name}/ent_checksum"])) < 1e-9
    assert abs(rel.astype(np.float64).sum() - float(g[f"{name}/rel_checksum"])) < 1e-9
    true = np.concatenate([train, g["valid"].astype(np.int64), g["test"].astype(np.int64)])
    hc, tc = ko.build_filter_csr(true, N, "head"), ko.build_filter_csr(true, N, "tail")
    q = g[f"{name}/queries"][:3]
    for mode in ("head-batch", "tail-batch"):
        r, c = ko.rank_all(name, ent, rel, q, mode, hc, tc, gamma=gamma, rel_margins=tuple(g["margins"]))
        np.testing.assert_array_equal(r, g[f"{name}/{mode}/rank64"][:3])
        np.testing.assert_array_equal(c, g[f"{name}/{mode}/contested"][:3])
