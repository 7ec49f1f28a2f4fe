"""Training-batch feed; mirror of mkb/datasets/dataset.py (Dataset) + mkb/datasets/base.py
(TrainDataset) for the non-classification path.

The reference runs two torch DataLoaders (head-batch, tail-batch) over per-triple Python objects.
Here the triples and their sub-sampling weights are two tensors built once (optionally resident on
the GPU with ``device='cuda'``), and a batch is one index_select — the per-epoch permutations consume
torch's global RNG in the same pattern as DataLoader(shuffle=True) so a seeded run visits the same
batches as the reference.
"""
from __future__ import annotations

import copy
import csv
import json
import os

import numpy as np
import torch

__all__ = ["Dataset", "from_directory"]


def subsampling_weights(triples: np.ndarray, start: int = 3) -> torch.Tensor:
    """``sqrt(1 / (cnt(h,r) + cnt(t,-r-1)))`` with counts starting at ``start``
    (mkb/datasets/base.py:102-121), vectorised."""
    if triples.shape[0] == 0:
        return torch.zeros(0)
    h, r, t = triples[:, 0], triples[:, 1], triples[:, 2]
    span = int(r.max()) + 1
    hr = h * span + r
    tr = t * span + r
    _, inv_hr, cnt_hr = np.unique(hr, return_inverse=True, return_counts=True)
    _, inv_tr, cnt_tr = np.unique(tr, return_inverse=True, return_counts=True)
    total = (cnt_hr[inv_hr] + start) + (cnt_tr[inv_tr] + start)
    return torch.sqrt(1 / torch.from_numpy(total.astype(np.float32)))


class Dataset:
    """``Dataset(train, batch_size, entities=None, relations=None, valid=None, test=None,
    shuffle=True, classification=False, pre_compute=True, num_workers=1, seed=42, ...)``
    (mkb/datasets/dataset.py:94-186).  Iterating yields, alternately, a head-batch and a tail-batch
    dict ``{"sample": int64[B,3], "weight": float32[B], "mode": str}`` (:188-194).

    Extra keywords (not in the reference): ``device`` — keep triples/weights on that device and yield
    batches there (no per-step H2D copy); ``pin_memory`` — host batches are gathered into a small ring of
    page-locked buffers so the consumer's ``.to(device, non_blocking=True)`` is a true async copy.
    """

    def __init__(self, train, batch_size, entities=None, relations=None, valid=None, test=None,
                 shuffle=True, classification=False, pre_compute=True, num_workers=1, seed=42,
                 classification_valid=None, classification_test=None, device=None, pin_memory=False):
        if classification:
            raise NotImplementedError(
                "classification mode feeds ConvE/BCE, which is outside the KGE hot path this package covers")
        self.train, self.valid, self.test = train, valid, test
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.classification = classification
        self.pre_compute = pre_compute
        self.num_workers = num_workers
        self.seed = seed
        self.device = torch.device(device) if device is not None else None
        self.pin_memory = bool(pin_memory) and self.device is None and torch.cuda.is_available()
        self._ring, self._slot = None, 0

        if entities is None:  # dataset.py:116-127
            self.entities = self.mapping_entities()
            self.train = [(self.entities[h], r, self.entities[t]) for h, r, t in self.train]
            if self.valid is not None:
                self.valid = [(self.entities[h], r, self.entities[t]) for h, r, t in self.valid]
            if self.test is not None:
                self.test = [(self.entities[h], r, self.entities[t]) for h, r, t in self.test]
        else:
            self.entities = entities
        if relations is None:  # dataset.py:129-141
            self.relations = self.mapping_relations()
            self.train = [(h, self.relations[r], t) for h, r, t in self.train]
            if self.valid is not None:
                self.valid = [(h, self.relations[r], t) for h, r, t in self.valid]
            if self.test is not None:
                self.test = [(h, self.relations[r], t) for h, r, t in self.test]
        else:
            self.relations = relations

        self.n_entity = len(self.entities)
        self.n_relation = len(self.relations)

        arr = np.asarray(self.train, dtype=np.int64).reshape(-1, 3)
        self._triples = torch.from_numpy(arr)
        self._weights = subsampling_weights(arr)
        if self.device is not None:
            self._triples = self._triples.to(self.device)
            self._weights = self._weights.to(self.device)
        self.len = int(2 * len(self.train) / self.batch_size)  # dataset.py:170-172
        self.step = 0
        self.fetch_head = self.fetch("head-batch")
        self.fetch_tail = self.fetch("tail-batch")
        self.classification_valid = classification_valid
        self.classification_test = classification_test
        if self.seed:
            torch.manual_seed(self.seed)  # dataset.py:185-186 (side effect kept on purpose)

    # -- id mappings (dataset.py, mapping_entities / mapping_relations) -----------------------
    def mapping_entities(self):
        """All heads first, then all tails, first-seen order over train+valid+test (dataset.py:322-331)."""
        tt = self.true_triples
        return {e: i for i, e in enumerate(dict.fromkeys([h for h, _, _ in tt] + [t for _, _, t in tt]))}

    def mapping_relations(self):
        return {r: i for i, r in enumerate(dict.fromkeys([r for _, r, _ in self.true_triples]))}

    # -- iteration -----------------------------------------------------------------------------
    # One epoch's visiting order for a loader, drawn the way torch's DataLoader(shuffle=True) does it so
    # a seeded run visits the same batches as the reference: creating the loader iterator draws a base
    # seed from the global RNG; the RandomSampler draws its own seed and runs randperm on a private
    # generator.  With worker processes (the reference's default num_workers=1) the sampler is advanced
    # while the iterator is being built (index prefetch), so the two draws are back to back per loader;
    # in the single-process case the sampler seed is drawn lazily at the first next(), i.e. after BOTH
    # loader iterators of zip(head, tail) exist.
    @staticmethod
    def _draw_seed():
        return int(torch.empty((), dtype=torch.int64).random_().item())

    def _perm(self, seed):
        n = self._triples.shape[0]
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g)

    def _order(self):
        n = self._triples.shape[0]
        if not self.shuffle:
            return torch.arange(n)
        self._draw_seed()  # the iterator's base seed
        return self._perm(self._draw_seed())

    def _epoch_orders(self):
        """(head order, tail order) for one pass of __iter__."""
        n = self._triples.shape[0]
        if not self.shuffle:
            return torch.arange(n), torch.arange(n)
        if self.num_workers and self.num_workers > 0:
            return self._order(), self._order()
        self._draw_seed()
        self._draw_seed()
        return self._perm(self._draw_seed()), self._perm(self._draw_seed())

    def _batches(self, mode, order):
        order = order.to(self._triples.device)
        if self.pin_memory and self._ring is None:
            self._ring = [(torch.empty((self.batch_size, 3), dtype=torch.int64).pin_memory(),
                           torch.empty(self.batch_size, dtype=torch.float32).pin_memory()) for _ in range(8)]
        for lo in range(0, order.shape[0], self.batch_size):
            idx = order[lo:lo + self.batch_size]
            if self.pin_memory:
                sbuf, wbuf = self._ring[self._slot]
                self._slot = (self._slot + 1) % len(self._ring)
                n = idx.shape[0]
                torch.index_select(self._triples, 0, idx, out=sbuf[:n])
                torch.index_select(self._weights, 0, idx, out=wbuf[:n])
                yield {"sample": sbuf[:n], "weight": wbuf[:n], "mode": mode}
            else:
                yield {"sample": self._triples.index_select(0, idx), "weight": self._weights.index_select(0, idx),
                       "mode": mode}

    def __iter__(self):
        head_order, tail_order = self._epoch_orders()
        head = self._batches("head-batch", head_order)
        tail = self._batches("tail-batch", tail_order)
        for hb, tb in zip(head, tail):
            yield hb
            yield tb

    def fetch(self, mode):
        while True:
            yield from self._batches(mode, self._order())

    def __next__(self):  # dataset.py:196-205
        self.step += 1
        return next(self.fetch_head) if self.step % 2 == 0 else next(self.fetch_tail)

    def __len__(self):
        return self.len

    @property
    def true_triples(self):
        out = copy.deepcopy(self.train)
        if self.valid is not None:
            out += self.valid
        if self.test is not None:
            out += self.test
        return out

    @property
    def train_triples(self):
        return self.train

    @property
    def name(self):
        return self.__class__.__name__

    @property
    def _repr_content(self):
        return {
            "Batch size": f"{self.batch_size}",
            "Entities": f"{self.n_entity}",
            "Relations": f"{self.n_relation}",
            "Shuffle": f"{self.shuffle}",
            "Train triples": f"{len(self.train) if self.train else 0}",
            "Validation triples": f"{len(self.valid) if self.valid else 0}",
            "Test triples": f"{len(self.test) if self.test else 0}",
        }

    def __repr__(self):
        l_len = max(map(len, self._repr_content.keys()))
        r_len = max(map(len, self._repr_content.values()))
        return f"{self.name} dataset\n" + "\n".join(
            k.rjust(l_len) + "  " + v.ljust(r_len) for k, v in self._repr_content.items())


def _read_csv(path):
    with open(path, newline="") as f:
        return [(int(h), int(r), int(t)) for h, r, t in csv.reader(f)]


def from_directory(path, batch_size, shuffle=True, seed=42, device=None):
    """Load a dataset stored the way mkb bundles its own (mkb/datasets/wn18rr.py:62-82):
    ``train.csv / valid.csv / test.csv`` with integer ``h,r,t`` rows, ``entities.json``,
    ``relations.json``.  E.g. ``from_directory('<site-packages>/mkb/datasets/fb15k237', 1024)``."""
    with open(os.path.join(path, "entities.json")) as f:
        entities = json.load(f)
    with open(os.path.join(path, "relations.json")) as f:
        relations = json.load(f)
    parts = {}
    for name in ("train", "valid", "test"):
        p = os.path.join(path, f"{name}.csv")
        parts[name] = _read_csv(p) if os.path.exists(p) else None
    return Dataset(train=parts["train"], valid=parts["valid"], test=parts["test"], entities=entities,
                   relations=relations, batch_size=batch_size, shuffle=shuffle, seed=seed, device=device)
