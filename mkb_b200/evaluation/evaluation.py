"""Link-prediction evaluation; mirror of mkb/evaluation/evaluation.py (Evaluation.eval /
eval_relations / compute_score) with mkb/datasets/base.py's TestDataset folded into the kernel.

The reference builds, per query and in Python, an N-long candidate list and filter-bias vector
(base.py:196-241), scores it, argsorts and looks the positive up (evaluation.py:237-263).  Here the
true triples are a device CSR and ``kge_rank_all`` returns the filtered rank of every query directly.
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from .. import ops
from ..utils.filters import build_filter_csr

__all__ = ["Evaluation"]

_METRICS = ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")


class _Mean:
    """river.stats.Mean stand-in (evaluation.py:187-189): incremental arithmetic mean."""

    def __init__(self):
        self.n, self.m = 0, 0.0

    def update(self, x):
        self.n += 1
        self.m += (x - self.m) / self.n
        return self

    def get(self):
        return self.m


class Evaluation:
    """``Evaluation(entities, relations, batch_size, true_triples=[], device='cpu', num_workers=1)``
    (evaluation.py:137-146).  ``device`` is where the reference would move each batch; the kernels
    run on the model's CUDA device."""

    def __init__(self, entities, relations, batch_size, true_triples=[], device="cpu", num_workers=1):
        self.entities = entities
        self.relations = relations
        self.true_triples = true_triples
        self.batch_size = batch_size
        self.device = device
        self.num_workers = num_workers
        self._csr = {}

    def _filter(self, side, device):
        key = (side, str(device))
        if key not in self._csr:
            if len(self.true_triples) == 0:
                self._csr[key] = None
            else:
                k, o, m = build_filter_csr(self.true_triples, len(self.entities), side)
                self._csr[key] = ops.FilterCSR(torch.from_numpy(k).to(device), torch.from_numpy(o).to(device),
                                               torch.from_numpy(m).to(device))
        return self._csr[key]

    def ranks(self, model, dataset, mode):
        """int64 ranks (on the model's device) of the true head (head-batch) / tail (tail-batch)."""
        dev = model.entity_embedding.device
        queries = torch.as_tensor(np.asarray(dataset, dtype=np.int64).reshape(-1, 3)).to(dev)
        csr = self._filter("head" if mode == "head-batch" else "tail", dev)
        chunk = max(int(self.batch_size), 1) * 64  # rank tiles are 64 queries; keep launches large
        out = [ops.rank_all(model.spec, model.entity_embedding, model.relation_embedding, queries[lo:lo + chunk],
                            mode, csr) for lo in range(0, queries.shape[0], chunk)]
        return torch.cat(out) if out else torch.zeros(0, dtype=torch.int64, device=dev)

    @staticmethod
    def _update(metrics, ranks):
        for ranking in ranks.tolist():  # evaluation.py:266-274, same update order
            metrics["MRR"].update(1.0 / ranking)
            metrics["MR"].update(ranking)
            metrics["HITS@1"].update(1.0 if ranking <= 1 else 0.0)
            metrics["HITS@3"].update(1.0 if ranking <= 3 else 0.0)
            metrics["HITS@10"].update(1.0 if ranking <= 10 else 0.0)
        return metrics

    def eval(self, model, dataset):
        """MRR, MR, HITS@1/3/10 over head-batch then tail-batch queries, rounded to 4 dp
        (evaluation.py:185-199)."""
        metrics = collections.OrderedDict({m: _Mean() for m in _METRICS})
        with torch.no_grad():
            for mode in ("head-batch", "tail-batch"):
                metrics = self._update(metrics, self.ranks(model, dataset, mode))
        return {name: round(metric.get(), 4) for name, metric in metrics.items()}

    def eval_relations(self, model, dataset):
        """Relation prediction (evaluation.py:201-215, TestDatasetRelation base.py:254-305): score the
        triple under every relation, bias -1 on relations that form another true triple (those slots
        are replaced by the true relation), rank the true relation."""
        metrics = collections.OrderedDict({m: _Mean() for m in _METRICS})
        dev = model.entity_embedding.device
        n_rel = len(self.relations)
        true = set(map(tuple, self.true_triples)) if len(self.true_triples) else set()
        triples = np.asarray(dataset, dtype=np.int64).reshape(-1, 3)
        with torch.no_grad():
            for lo in range(0, len(triples), max(int(self.batch_size), 1)):
                part = triples[lo:lo + max(int(self.batch_size), 1)]
                cand = np.repeat(part[:, None, :], n_rel, axis=1)  # [b, R, 3]
                bias = np.zeros((len(part), n_rel), dtype=np.float32)
                for i, (h, r, t) in enumerate(part):
                    for rr in range(n_rel):
                        if (int(h), rr, int(t)) in true:
                            cand[i, rr, 1] = r
                            bias[i, rr] = -1.0
                        else:
                            cand[i, rr, 1] = rr
                    bias[i, r] = 0.0
                score = model(torch.from_numpy(cand).to(dev)) + torch.from_numpy(bias).to(dev)
                order = torch.argsort(score, dim=1, descending=True, stable=True)
                pos = torch.from_numpy(part[:, 1]).to(dev)
                first = (order == pos[:, None]).int().argmax(dim=1) + 1
                metrics = self._update(metrics, first)
        return {f"{name}_relations": round(metric.get(), 4) for name, metric in metrics.items()}
