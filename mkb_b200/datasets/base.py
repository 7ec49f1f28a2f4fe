"""Evaluation-side datasets; mirrors of mkb/datasets/base.py (TestDataset, TestDatasetRelation).

The reference builds every item with a Python loop over all entities and a set lookup per candidate
(base.py:196-241, 34 ms per query at Wn18rr).  The fast path of this package never materialises these
lists — ``kge_rank_all`` filters inside the kernel from a CSR — but the classes are part of the
reference's public surface (``Evaluation.get_entity_stream`` / ``compute_score`` consume them), so they
exist here with the same item layout, built with vectorised numpy from the same CSR.
"""
from __future__ import annotations

import numpy as np
import torch

from ..utils.filters import build_filter_csr, triples_to_array

__all__ = ["TestDataset", "TestDatasetRelation", "BatchStream"]


class TestDataset:
    """``TestDataset(triples, true_triples, entities, relations, mode)`` (base.py:185-194).

    ``ds[i] -> (sample int64[3], negative_sample int64[N], filter_bias float32[N], mode)``: candidate j is
    entity j with bias 0, except candidates that form another true triple, which are replaced by the
    positive's id and carry bias -1e5 (base.py:203-233)."""

    __test__ = False  # not a pytest class

    def __init__(self, triples, true_triples, entities, relations, mode):
        if mode not in ("head-batch", "tail-batch", "relation-batch"):
            raise ValueError(f"unknown mode {mode!r}")
        self.len = len(triples)
        self.triples = triples
        self.true_triples = set(map(tuple, true_triples))
        self.n_entity = len(entities)
        self.n_relation = len(relations)
        self.mode = mode
        self._arr = triples_to_array(list(triples))
        if mode != "relation-batch":
            self._keys, self._offsets, self._members = build_filter_csr(
                list(true_triples), self.n_entity, "head" if mode == "head-batch" else "tail")

    def __len__(self):
        return self.len

    def _true_set(self, fixed, relation):
        if self._keys.shape[0] == 0:
            return self._members[:0]
        code = relation * self.n_entity + fixed
        k = int(np.searchsorted(self._keys, code))
        if k >= self._keys.shape[0] or self._keys[k] != code:
            return self._members[:0]
        return self._members[self._offsets[k]:self._offsets[k + 1]]

    def __getitem__(self, idx):
        head, relation, tail = (int(x) for x in self._arr[idx])
        positive, fixed = (head, tail) if self.mode == "head-batch" else (tail, head)
        cand = np.arange(self.n_entity, dtype=np.int64)
        bias = np.zeros(self.n_entity, dtype=np.float32)
        members = self._true_set(fixed, relation)
        cand[members] = positive
        bias[members] = -1e5
        if 0 <= positive < self.n_entity:
            bias[positive] = 0.0  # "actual target" branch (base.py:209-210, :224-225)
        return (torch.tensor((head, relation, tail), dtype=torch.int64), torch.from_numpy(cand),
                torch.from_numpy(bias), self.mode)

    @staticmethod
    def collate_fn(data):
        return {
            "sample": torch.stack([d[0] for d in data], dim=0),
            "negative_sample": torch.stack([d[1] for d in data], dim=0),
            "filter_bias": torch.stack([d[2] for d in data], dim=0),
            "mode": data[0][3],
        }


class TestDatasetRelation(TestDataset):
    """``TestDatasetRelation(triples, true_triples, entities, relations)`` (base.py:254-305): the triple
    under every relation; relations that form another true triple are replaced by the true relation and
    carry bias -1 (int64, as the reference's LongTensor)."""

    __test__ = False

    def __init__(self, triples, true_triples, entities, relations):
        super().__init__(triples=triples, true_triples=true_triples, entities=entities, relations=relations,
                         mode="relation-batch")
        true = triples_to_array(list(true_triples))
        # sorted codes of the true triples: membership of (h, r', t) for all r' is one searchsorted
        self._codes = np.unique((true[:, 0] * self.n_relation + true[:, 1]) * self.n_entity + true[:, 2]) \
            if true.shape[0] else np.zeros(0, dtype=np.int64)

    def is_true(self, head, tail):
        """bool[R]: (head, r, tail) is a true triple, for every relation r."""
        codes = (head * self.n_relation + np.arange(self.n_relation, dtype=np.int64)) * self.n_entity + tail
        pos = np.searchsorted(self._codes, codes)
        pos[pos >= self._codes.shape[0]] = 0
        return self._codes[pos] == codes if self._codes.shape[0] else np.zeros(self.n_relation, dtype=bool)

    def __getitem__(self, idx):
        head, relation, tail = (int(x) for x in self._arr[idx])
        true = self.is_true(head, tail)
        rel = np.where(true, relation, np.arange(self.n_relation, dtype=np.int64))
        bias = np.where(true, -1, 0).astype(np.int64)
        if 0 <= relation < self.n_relation:
            bias[relation] = 0
        cand = np.stack([np.full(self.n_relation, head, dtype=np.int64), rel,
                         np.full(self.n_relation, tail, dtype=np.int64)], axis=-1)
        return (torch.tensor((head, relation, tail), dtype=torch.int64), torch.from_numpy(cand),
                torch.from_numpy(bias), self.mode)


class BatchStream:
    """What ``torch.utils.data.DataLoader(dataset, batch_size, collate_fn=...)`` yields in the reference
    (evaluation.py:147-163), without worker processes: collated batches in order."""

    def __init__(self, dataset, batch_size):
        self.dataset = dataset
        self.batch_size = max(int(batch_size), 1)

    def __len__(self):
        return -(-len(self.dataset) // self.batch_size)

    def __iter__(self):
        n = len(self.dataset)
        for lo in range(0, n, self.batch_size):
            yield self.dataset.collate_fn([self.dataset[i] for i in range(lo, min(lo + self.batch_size, n))])
