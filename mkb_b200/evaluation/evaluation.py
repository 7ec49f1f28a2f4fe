"""Link-prediction evaluation; mirror of mkb/evaluation/evaluation.py (Evaluation) with
mkb/datasets/base.py's TestDataset folded into the kernel.

The reference builds, per query and in Python, an N-long candidate list and filter-bias vector
(base.py:196-241), scores it, argsorts and looks the positive up (evaluation.py:237-263).  Here the
true triples are a device CSR and ``kge_rank_all`` returns the filtered rank of every query directly:
``eval`` and ``detail_eval`` run on it.  The reference's stream-based entry points
(``get_entity_stream`` / ``get_relation_stream`` / ``compute_score`` / ``compute_detailled_score``) are
kept for callers that drive them directly; they score through the same CUDA kernels and replace the
argsort + lookup by a comparison count (the position of the positive in a stable descending sort).
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from .. import ops
from ..datasets.base import BatchStream, TestDataset, TestDatasetRelation
from ..utils import Bar
from ..utils.filters import build_filter_csr, triples_to_array

__all__ = ["Evaluation"]

_METRICS = ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10")
_TYPES = ("1_1", "1_M", "M_1", "M_M")


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


def _new_metrics():
    return collections.OrderedDict({m: _Mean() for m in _METRICS})


def _update(metrics, ranks):
    for ranking in ranks:  # evaluation.py:266-274, same update order
        metrics["MRR"].update(1.0 / ranking)
        metrics["MR"].update(ranking)
        metrics["HITS@1"].update(1.0 if ranking <= 1 else 0.0)
        metrics["HITS@3"].update(1.0 if ranking <= 3 else 0.0)
        metrics["HITS@10"].update(1.0 if ranking <= 10 else 0.0)
    return metrics


def _position_of(score, positive_arg):
    """1-based position of column ``positive_arg[i]`` in a stable descending sort of ``score[i]``:
    ``1 + #{j: s_j > s_p} + #{j < p: s_j == s_p}`` (what evaluation.py:245-263 reads off the argsort)."""
    idx = positive_arg.view(-1, 1)
    sp = score.gather(1, idx)
    cols = torch.arange(score.shape[1], device=score.device).view(1, -1)
    return 1 + (score > sp).sum(1) + ((score == sp) & (cols < idx)).sum(1)


class Evaluation:
    """``Evaluation(entities, relations, batch_size, true_triples=[], device='cpu', num_workers=1)``
    (evaluation.py:137-146).  ``device`` is where the reference would move each batch; the kernels
    run on the model's CUDA device."""

    def __init__(self, entities, relations, batch_size, true_triples=[], device="cpu", num_workers=1,
                 distributed=False):
        """``distributed=True`` (not in the reference): under an initialised ``torch.distributed`` group with
        replicated tables, every rank ranks a contiguous slice of the queries and the ranks are all-gathered
        back in query order, so every rank reports the same metrics as a single-GPU run."""
        self.distributed = distributed
        self.entities = entities
        self.relations = relations
        self.true_triples = true_triples
        self.batch_size = batch_size
        self.device = device
        self.num_workers = num_workers
        self._csr = {}

    # ------------------------------------------------------------------------------------------
    # fast path: ranks straight from the kernel
    # ------------------------------------------------------------------------------------------
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
        if self.distributed and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            return self._ranks_distributed(model, queries, mode)
        return self._ranks_local(model, queries, mode)

    def _ranks_local(self, model, queries, mode):
        dev = queries.device
        csr = self._filter("head" if mode == "head-batch" else "tail", dev)
        chunk = max(int(self.batch_size), 1) * 64  # rank tiles are 64 queries; keep launches large
        out = [ops.rank_all(model.spec, model.entity_embedding, model.relation_embedding, queries[lo:lo + chunk],
                            mode, csr, modulus=getattr(model, "kernel_modulus", None))
               for lo in range(0, queries.shape[0], chunk)]
        return torch.cat(out) if out else torch.zeros(0, dtype=torch.int64, device=dev)

    def _ranks_distributed(self, model, queries, mode):
        """Queries are independent units: rank r takes the r-th contiguous slice (sizes differ by at most
        one), one all-gather of the padded int64 ranks restores query order on every rank."""
        dist = torch.distributed
        world, rank = dist.get_world_size(), dist.get_rank()
        bounds = np.linspace(0, queries.shape[0], world + 1).round().astype(np.int64)
        width = int((bounds[1:] - bounds[:-1]).max()) if queries.shape[0] else 0
        mine = torch.zeros(width, dtype=torch.int64, device=queries.device)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        if hi > lo:
            mine[: hi - lo] = self._ranks_local(model, queries[lo:hi], mode)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        if width:
            dist.all_gather(parts, mine)
        return torch.cat([parts[r][: int(bounds[r + 1] - bounds[r])] for r in range(world)]) if width else mine

    @staticmethod
    def _update(metrics, ranks):
        return _update(metrics, ranks.tolist())

    def eval(self, model, dataset, ranks_fn=None):
        """MRR, MR, HITS@1/3/10 over head-batch then tail-batch queries, rounded to 4 dp
        (evaluation.py:185-199).  ``ranks_fn(dataset, mode)`` (not in the reference) substitutes another
        source of ranks, e.g. ``functools.partial(trainer.sharded_ranks, evaluation)`` for a row-sharded
        table that is never gathered."""
        metrics = _new_metrics()
        with torch.no_grad():
            for mode in ("head-batch", "tail-batch"):
                ranks = self.ranks(model, dataset, mode) if ranks_fn is None else ranks_fn(dataset, mode)
                metrics = self._update(metrics, ranks)
        return {name: round(metric.get(), 4) for name, metric in metrics.items()}

    def eval_relations(self, model, dataset):
        """Relation prediction (evaluation.py:201-215, TestDatasetRelation base.py:254-305): score the
        triple under every relation (one positives-only kernel launch per batch over the 3-D sample),
        bias -1 on relations that form another true triple (those slots carry the true relation),
        rank of the true relation."""
        metrics = _new_metrics()
        dev = model.entity_embedding.device
        ds = TestDatasetRelation(triples=dataset, true_triples=self.true_triples, entities=self.entities,
                                 relations=self.relations)
        with torch.no_grad():
            for data in BatchStream(ds, max(int(self.batch_size), 1) * 64):
                score = model(data["negative_sample"].to(dev)) + data["filter_bias"].to(dev)
                metrics = self._update(metrics, _position_of(score, data["sample"][:, 1].to(dev)))
        return {f"{name}_relations": round(metric.get(), 4) for name, metric in metrics.items()}

    # ------------------------------------------------------------------------------------------
    # relation categories (Bordes et al. 2013) and the per-category table
    # ------------------------------------------------------------------------------------------
    def types_relations(self, model=None, dataset=None, threshold=1.5):
        """{relation label: '1_1' | '1_M' | 'M_1' | 'M_M'} from the mean number of heads per
        (tail, relation) and of tails per (head, relation) over ``true_triples`` (rows counted as they
        come, duplicates included), each '1' when <= ``threshold`` (evaluation.py:335-383).

        Like the reference, the i-th relation PRESENT in ``true_triples`` (ascending id) is reported
        under the label of relation id i — identical to the relation's own label whenever every relation
        occurs, which is the only case the reference handles correctly."""
        arr = triples_to_array(list(self.true_triples))
        label = {i: name for name, i in self.relations.items()}
        out = {}
        for row, r in enumerate(np.unique(arr[:, 1])):
            sel = arr[arr[:, 1] == r]
            heads_per_tail = np.unique(sel[:, 2], return_counts=True)[1].mean()
            tails_per_head = np.unique(sel[:, 0], return_counts=True)[1].mean()
            out[label[row]] = ("1" if heads_per_tail <= threshold else "M") + "_" + \
                              ("1" if tails_per_head <= threshold else "M")
        return out

    def detail_eval(self, model, dataset, threshold=1.5):
        """Metrics per relation category and side as the reference's DataFrame (evaluation.py:385-464):
        index relation in (1_1, 1_M, M_1, M_M), columns (head|tail) x metric + (metadata, frequency)."""
        import pandas as pd

        by_id = {self.relations[name]: t for name, t in
                 self.types_relations(model=model, dataset=dataset, threshold=threshold).items()}
        metrics = collections.OrderedDict(
            (mode, collections.OrderedDict((t, _new_metrics()) for t in _TYPES))
            for mode in ("head-batch", "tail-batch"))
        rel_of = np.asarray(dataset, dtype=np.int64).reshape(-1, 3)[:, 1].tolist()
        with torch.no_grad():
            for mode in ("head-batch", "tail-batch"):
                for r, ranking in zip(rel_of, self.ranks(model, dataset, mode).tolist()):
                    _update(metrics[mode][by_id[r]], [ranking])
        return self._detail_frame(pd, metrics, by_id)

    @staticmethod
    def _detail_frame(pd, metrics, by_id):
        sides = {}
        for mode, side in (("head-batch", "head"), ("tail-batch", "tail")):
            rows = [{m: round(metrics[mode][t][m].get(), 4) for m in _METRICS} for t in _TYPES]
            frame = pd.DataFrame(rows)
            frame.columns = pd.MultiIndex.from_product([[side], frame.columns])
            sides[side] = frame
        results = pd.concat([sides["head"], sides["tail"]], axis="columns")
        results = results.set_index(pd.Series(list(_TYPES)))
        results.index.name = "relation"
        freq = collections.OrderedDict((t, 0) for t in _TYPES)
        for t in by_id.values():
            freq[t] += 1
        frequency = pd.DataFrame.from_dict({t: c / len(by_id) for t, c in freq.items()}, orient="index",
                                           columns=["frequency"])
        frequency.columns = pd.MultiIndex.from_product([["metadata"], frequency.columns])
        return pd.concat([results, frequency], axis="columns")

    # ------------------------------------------------------------------------------------------
    # the reference's stream-based entry points
    # ------------------------------------------------------------------------------------------
    def _get_test_loader(self, triples, mode):
        return BatchStream(TestDataset(triples=triples, true_triples=self.true_triples, entities=self.entities,
                                       relations=self.relations, mode=mode), self.batch_size)

    def get_entity_stream(self, dataset):
        """[head-batch loader, tail-batch loader] (evaluation.py:165-169)."""
        return [self._get_test_loader(dataset, "head-batch"), self._get_test_loader(dataset, "tail-batch")]

    def get_relation_stream(self, dataset):
        """Relation-prediction loader (evaluation.py:171-183)."""
        return BatchStream(TestDatasetRelation(triples=dataset, true_triples=self.true_triples,
                                               entities=self.entities, relations=self.relations), self.batch_size)

    @staticmethod
    def _stream_ranks(model, data, device):
        sample = data["sample"].to(device)
        negative_sample = data["negative_sample"].to(device)
        filter_bias = data["filter_bias"].to(device)
        mode = data["mode"]
        if mode in ("head-batch", "tail-batch"):
            score = model(sample=sample, negative_sample=negative_sample, mode=mode)
            positive_arg = sample[:, 0] if mode == "head-batch" else sample[:, 2]
        elif mode == "relation-batch":
            score = model(negative_sample)
            positive_arg = sample[:, 1]
        else:
            raise ValueError(f"unknown mode {mode!r}")
        return sample, _position_of(score + filter_bias, positive_arg.to(score.device)).tolist()

    @classmethod
    def compute_score(cls, model, test_set, metrics, device):
        """Consume a stream of ``{"sample", "negative_sample", "filter_bias", "mode"}`` batches and update
        ``metrics`` (objects with ``.update``) exactly like evaluation.py:217-279."""
        training = model.training
        model = model.eval()
        bar = Bar(dataset=test_set, update_every=1)
        bar.set_description("Evaluation")
        with torch.no_grad():
            for data in bar:
                _, ranks = cls._stream_ranks(model, data, model.entity_embedding.device)
                _update(metrics, ranks)
        if training:
            model.train()
        return metrics

    @classmethod
    def compute_detailled_score(cls, model, test_set, metrics, types_relations, device):
        """Same, metrics keyed by ``[mode][relation type]`` (evaluation.py:281-333)."""
        training = model.training
        model = model.eval()
        bar = Bar(dataset=test_set, update_every=1)
        bar.set_description("Evaluation")
        with torch.no_grad():
            for data in bar:
                sample, ranks = cls._stream_ranks(model, data, model.entity_embedding.device)
                for r, ranking in zip(sample[:, 1].tolist(), ranks):
                    _update(metrics[data["mode"]][types_relations[r]], [ranking])
        if training:
            model.train()
        return metrics
