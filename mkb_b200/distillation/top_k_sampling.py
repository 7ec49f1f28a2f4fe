"""``TopKSampling`` — mirror of mkb/distillation/top_k_sampling.py:321-677.

For every training triple the reference scores ALL shared entities as head candidates, all shared
relations, and all shared entities as tail candidates with the teacher, argsorts each score row and
keeps the first ``batch_size_entity`` / ``batch_size_relation`` positions (``_get_rank_entities`` :672-677,
``_get_rank_relations`` :665-669) — a Python loop over the batch with three model calls and three full
sorts per triple (:577-604).  Here the whole batch is three kernel calls per candidate chunk:
``kge_score_fwd`` over ``[B, n_candidates]`` (head-batch, tail-batch) and over the 3-D ``[B, n_relations, 3]``
sample, each followed by ``kge_topk_rows`` (exact radix select, ties by candidate order = a stable
descending argsort).  The random entities / relations appended afterwards consume the seeded
``numpy.RandomState`` exactly like ``_randomize_distribution`` (:877-960).
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from .. import ops

__all__ = ["FastTopKSampling", "TopKSampling", "TopKSamplingTransE"]


class TopKSampling:
    """Same constructor as the reference (:486-499).  ``get(sample, teacher)`` returns the six int64 tensors
    ``(head_teacher, relation_teacher, tail_teacher, head_student, relation_student, tail_student)`` of shape
    ``[B, batch_size_entity]`` / ``[B, batch_size_relation]`` on the teacher's device."""

    def __init__(self, teacher_entities, teacher_relations, student_entities, student_relations, batch_size_entity,
                 batch_size_relation, n_random_entities, n_random_relations, device="cpu", seed=None, **kwargs):
        self.batch_size_entity_top_k = batch_size_entity
        self.batch_size_relation_top_k = batch_size_relation
        self.n_random_entities = n_random_entities
        self.n_random_relations = n_random_relations
        self.device = device
        self._rng = np.random.RandomState(seed)
        self.mapping_entities = collections.OrderedDict(
            {i: student_entities[e] for e, i in teacher_entities.items() if e in student_entities})
        self.mapping_relations = collections.OrderedDict(
            {i: student_relations[r] for r, i in teacher_relations.items() if r in student_relations})
        i64 = torch.int64
        self.entities_teacher_selection = torch.tensor(list(self.mapping_entities.keys()), dtype=i64)
        self.entities_teacher_prediction = self.entities_teacher_selection.view(1, -1)
        self.entities_student = torch.tensor(list(self.mapping_entities.values()), dtype=i64)
        self.relations_teacher = torch.tensor(list(self.mapping_relations.keys()), dtype=i64)
        self.relations_student = torch.tensor(list(self.mapping_relations.values()), dtype=i64)
        self._dev_cache = None

    @property
    def supervised(self):
        """Do not include the ground truth (:551-554)."""
        return False

    @property
    def batch_size_entity(self):
        return self.batch_size_entity_top_k + self.n_random_entities

    @property
    def batch_size_relation(self):
        return self.batch_size_relation_top_k + self.n_random_relations

    def _on(self, dev):
        if self._dev_cache is None or self._dev_cache[0] != dev:
            self._dev_cache = (dev, self.entities_teacher_selection.to(dev), self.entities_student.to(dev),
                               self.relations_teacher.to(dev), self.relations_student.to(dev))
        return self._dev_cache[1:]

    def get(self, sample, teacher, max_ids_per_call=1 << 24, **kwargs):
        dev = teacher.entity_embedding.device
        ent_t, ent_s, rel_t, rel_s = self._on(dev)
        sample = sample.to(dev)
        B, n_ent, n_rel = sample.shape[0], ent_t.shape[0], rel_t.shape[0]
        k_e = min(int(self.batch_size_entity_top_k), n_ent)  # argsort(...)[:, :k] never returns more than exist
        k_r = min(int(self.batch_size_relation_top_k), n_rel)
        rank_h = torch.empty((B, k_e), dtype=torch.int64, device=dev)
        rank_t = torch.empty((B, k_e), dtype=torch.int64, device=dev)
        rank_r = torch.empty((B, k_r), dtype=torch.int64, device=dev)
        training = teacher.training
        teacher.eval()
        with torch.no_grad():
            step = max(1, int(max_ids_per_call) // max(n_ent, 1))  # bound the [b, n_candidates] id matrix
            for lo in range(0, B, step):
                s = sample[lo:lo + step]
                b = s.shape[0]
                cand = ent_t.view(1, -1).expand(b, n_ent).contiguous()
                if k_e:
                    rank_h[lo:lo + b] = ops.topk_rows(teacher(s, cand, "head-batch"), k_e)
                    rank_t[lo:lo + b] = ops.topk_rows(teacher(s, cand, "tail-batch"), k_e)
                if k_r:
                    rels = torch.stack([s[:, 0:1].expand(b, n_rel), rel_t.view(1, -1).expand(b, n_rel),
                                        s[:, 2:3].expand(b, n_rel)], dim=2).contiguous()  # [b, n_rel, 3]
                    rank_r[lo:lo + b] = ops.topk_rows(teacher(rels), k_r)
        if training:
            teacher.train()
        head_teacher, head_student = ent_t[rank_h], ent_s[rank_h]
        tail_teacher, tail_student = ent_t[rank_t], ent_s[rank_t]
        # the reference indexes relations_student for BOTH relation outputs (:604-607); kept
        relation_teacher, relation_student = rel_s[rank_r], rel_s[rank_r]

        if self.n_random_entities > 0:  # _randomize_distribution :894-925, same RNG consumption
            rnd_t = self._rng.choice(list(self.mapping_entities.keys()), size=self.n_random_entities, replace=False)
            rnd_s = torch.tensor([[self.mapping_entities[i] for i in rnd_t]], dtype=torch.int64, device=dev).expand(B, -1)
            rnd_t = torch.tensor(np.asarray(rnd_t)[None, :], dtype=torch.int64, device=dev).expand(B, -1)
            head_teacher, head_student = torch.cat([head_teacher, rnd_t], 1), torch.cat([head_student, rnd_s], 1)
            tail_teacher, tail_student = torch.cat([tail_teacher, rnd_t], 1), torch.cat([tail_student, rnd_s], 1)
        if self.n_random_relations > 0:  # :927-949
            rnd_t = self._rng.choice(list(self.mapping_relations.keys()), size=self.n_random_relations, replace=False)
            rnd_s = torch.tensor([[self.mapping_relations[i] for i in rnd_t]], dtype=torch.int64, device=dev).expand(B, -1)
            rnd_t = torch.tensor(np.asarray(rnd_t)[None, :], dtype=torch.int64, device=dev).expand(B, -1)
            relation_teacher = torch.cat([relation_teacher, rnd_t], 1)
            relation_student = torch.cat([relation_student, rnd_s], 1)
        return (head_teacher, relation_teacher, tail_teacher, head_student, relation_student, tail_student)


class FastTopKSampling:
    """``TopKSampling`` evaluated ONCE for every training triple of the teacher's dataset, then looked up
    (mkb/distillation/top_k_sampling.py:10-318).  The reference fills six Python dicts keyed by the strings
    ``"r_t"`` / ``"h_t"`` / ``"h_r"`` in a loop over every triple of every head-batch; here the pre-computation
    is the batched kernel path of ``TopKSampling.get`` and the tables are three sorted int64 key vectors with
    their ``[n_keys, k]`` rows on the teacher's device, so ``get`` is three ``searchsorted`` calls.

    A TransE teacher switches the pre-computation to the nearest-neighbour sampler ``TopKSamplingTransE`` like
    the reference (:168-171)."""

    def __init__(self, teacher_entities, teacher_relations, student_entities, student_relations, batch_size_entity,
                 batch_size_relation, n_random_entities, n_random_relations, dataset_teacher, teacher, device="cpu",
                 seed=None, **kwargs):
        base_method = TopKSamplingTransE if teacher.name == "TransE" else TopKSampling  # :168-171
        base = base_method(teacher_entities=teacher_entities, teacher_relations=teacher_relations,
                           student_entities=student_entities, student_relations=student_relations,
                           batch_size_entity=batch_size_entity, batch_size_relation=batch_size_relation,
                           n_random_entities=0, n_random_relations=0, device=device, seed=seed, teacher=teacher)
        self.mapping_entities, self.mapping_relations = base.mapping_entities, base.mapping_relations
        self.batch_size_entity_top_k = batch_size_entity
        self.batch_size_relation_top_k = batch_size_relation
        self.n_random_entities = n_random_entities
        self.n_random_relations = n_random_relations
        self._rng = np.random.RandomState(seed)
        self.device = device
        dev = teacher.entity_embedding.device
        self._span = int(max(teacher.n_entity, teacher.n_relation)) + 1
        samples = [data["sample"] for data in dataset_teacher if data["mode"] == "head-batch"]  # :218-220
        samples = torch.cat(samples).to(dev) if samples else torch.zeros((0, 3), dtype=torch.int64, device=dev)
        out = base.get(samples, teacher)
        h, r, t = samples[:, 0], samples[:, 1], samples[:, 2]
        # the head candidates of a triple depend on (r, t) only, the relations on (h, t), the tails on (h, r):
        # one row per distinct key (the reference's dict keeps the last occurrence; all occurrences are equal)
        self._tables = {}
        for name, key, teacher_rows, student_rows in (("head", r * self._span + t, out[0], out[3]),
                                                      ("relation", h * self._span + t, out[1], out[4]),
                                                      ("tail", h * self._span + r, out[2], out[5])):
            keys, inverse = torch.unique(key, return_inverse=True)  # keys come back sorted
            order = torch.argsort(inverse, stable=True)
            grouped = inverse[order]
            starts = torch.ones_like(grouped, dtype=torch.bool)
            starts[1:] = grouped[1:] != grouped[:-1]
            first = order[starts]  # one representative triple per key, in key order
            self._tables[name] = (keys, teacher_rows[first], student_rows[first])

    @property
    def supervised(self):
        return False

    @property
    def batch_size_entity(self):
        return self.batch_size_entity_top_k + self.n_random_entities

    @property
    def batch_size_relation(self):
        return self.batch_size_relation_top_k + self.n_random_relations

    def _lookup(self, name, key):
        keys, teacher_rows, student_rows = self._tables[name]
        pos = torch.searchsorted(keys, key).clamp_(max=max(keys.shape[0] - 1, 0))
        if keys.shape[0] == 0 or not bool((keys[pos] == key).all()):
            bad = key[(keys[pos] != key)][0].item() if keys.shape[0] else key[0].item()
            raise KeyError(f"{bad // self._span}_{bad % self._span}")  # the reference's dict lookup fails the same way
        return teacher_rows[pos], student_rows[pos]

    def get(self, sample, **kwargs):
        dev = self._tables["head"][0].device
        sample = sample.to(dev)
        B = sample.shape[0]
        h, r, t = sample[:, 0], sample[:, 1], sample[:, 2]
        head_teacher, head_student = self._lookup("head", r * self._span + t)
        relation_teacher, relation_student = self._lookup("relation", h * self._span + t)
        tail_teacher, tail_student = self._lookup("tail", h * self._span + r)
        if self.n_random_entities > 0:  # _randomize_distribution :894-925, same RNG consumption
            rnd_t = self._rng.choice(list(self.mapping_entities.keys()), size=self.n_random_entities, replace=False)
            rnd_s = torch.tensor([[self.mapping_entities[i] for i in rnd_t]], dtype=torch.int64, device=dev).expand(B, -1)
            rnd_t = torch.tensor(np.asarray(rnd_t)[None, :], dtype=torch.int64, device=dev).expand(B, -1)
            head_teacher, head_student = torch.cat([head_teacher, rnd_t], 1), torch.cat([head_student, rnd_s], 1)
            tail_teacher, tail_student = torch.cat([tail_teacher, rnd_t], 1), torch.cat([tail_student, rnd_s], 1)
        if self.n_random_relations > 0:  # :927-949
            rnd_t = self._rng.choice(list(self.mapping_relations.keys()), size=self.n_random_relations, replace=False)
            rnd_s = torch.tensor([[self.mapping_relations[i] for i in rnd_t]], dtype=torch.int64, device=dev).expand(B, -1)
            rnd_t = torch.tensor(np.asarray(rnd_t)[None, :], dtype=torch.int64, device=dev).expand(B, -1)
            relation_teacher = torch.cat([relation_teacher, rnd_t], 1)
            relation_student = torch.cat([relation_student, rnd_s], 1)
        return (head_teacher, relation_teacher, tail_teacher, head_student, relation_student, tail_student)


class TopKSamplingTransE:
    """Top-k candidates of a TransE teacher by NEAREST NEIGHBOUR in embedding space instead of by score
    (mkb/distillation/top_k_sampling.py:680-875): the best heads of ``(h, r, t)`` are the shared entities closest to
    ``t - r``, the best relations those closest to ``t - h``, the best tails those closest to ``h + r``
    (``TransE._top_k``), in exact L2 distance over a snapshot of the teacher's rows taken at construction.

    The reference builds that exact search with ``faiss.IndexFlatL2`` (brute force); faiss is not a dependency
    here: the squared distances of a whole batch against the snapshot are one ``torch.cdist`` on the teacher's
    device and the selection is ``kge_topk_rows`` (exact, ties by candidate order).  Same constructor, same
    outputs, same RNG stream for the random extras.  Parity with faiss itself is unpinned (faiss is absent from
    the build container): the tests pin it to a numpy brute-force search, which is what IndexFlatL2 specifies."""

    def __init__(self, teacher_entities, teacher_relations, student_entities, student_relations, teacher,
                 batch_size_entity, batch_size_relation, n_random_entities, n_random_relations, seed=None, **kwargs):
        self.batch_size_entity_top_k = batch_size_entity
        self.batch_size_relation_top_k = batch_size_relation
        self.n_random_entities = n_random_entities
        self.n_random_relations = n_random_relations
        self._rng = np.random.RandomState(seed)
        self.mapping_entities = collections.OrderedDict(
            {i: student_entities[e] for e, i in teacher_entities.items() if e in student_entities})
        self.mapping_relations = collections.OrderedDict(
            {i: student_relations[e] for e, i in teacher_relations.items() if e in student_relations})
        dev = teacher.entity_embedding.device
        i64 = dict(dtype=torch.int64, device=dev)
        self._ent_t = torch.tensor(list(self.mapping_entities.keys()), **i64)
        self._ent_s = torch.tensor(list(self.mapping_entities.values()), **i64)
        self._rel_t = torch.tensor(list(self.mapping_relations.keys()), **i64)
        self._rel_s = torch.tensor(list(self.mapping_relations.values()), **i64)
        with torch.no_grad():  # the "trees": a snapshot, like the rows added to the faiss index (:760-770)
            self._ent_rows = teacher.entity_embedding.detach()[self._ent_t].clone()
            self._rel_rows = teacher.relation_embedding.detach()[self._rel_t].clone()

    @property
    def supervised(self):
        return False

    @property
    def batch_size_entity(self):
        return self.batch_size_entity_top_k + self.n_random_entities

    @property
    def batch_size_relation(self):
        return self.batch_size_relation_top_k + self.n_random_relations

    @staticmethod
    def _nearest(queries, rows, k):
        """Positions of the ``k`` rows nearest to each query (ascending squared L2 distance)."""
        d2 = torch.cdist(queries, rows, p=2.0, compute_mode="donot_use_mm_for_euclid_dist").square_()
        return ops.topk_rows(-d2, k)

    def get(self, sample, teacher, **kwargs):
        dev = self._ent_rows.device
        sample = sample.to(dev)
        B = sample.shape[0]
        with torch.no_grad():
            q_head, q_relation, q_tail = teacher._top_k(sample)
            k_e = min(int(self.batch_size_entity_top_k), self._ent_rows.shape[0])
            k_r = min(int(self.batch_size_relation_top_k), self._rel_rows.shape[0])
            top_h = self._nearest(q_head.reshape(B, -1), self._ent_rows, k_e)
            top_r = self._nearest(q_relation.reshape(B, -1), self._rel_rows, k_r)
            top_t = self._nearest(q_tail.reshape(B, -1), self._ent_rows, k_e)
        head_teacher, head_student = self._ent_t[top_h], self._ent_s[top_h]
        relation_teacher, relation_student = self._rel_t[top_r], self._rel_s[top_r]
        tail_teacher, tail_student = self._ent_t[top_t], self._ent_s[top_t]
        if self.n_random_entities > 0:  # _randomize_distribution :894-925, same RNG consumption
            rnd_t = self._rng.choice(list(self.mapping_entities.keys()), size=self.n_random_entities, replace=False)
            rnd_s = torch.tensor([[self.mapping_entities[i] for i in rnd_t]], dtype=torch.int64, device=dev).expand(B, -1)
            rnd_t = torch.tensor(np.asarray(rnd_t)[None, :], dtype=torch.int64, device=dev).expand(B, -1)
            head_teacher, head_student = torch.cat([head_teacher, rnd_t], 1), torch.cat([head_student, rnd_s], 1)
            tail_teacher, tail_student = torch.cat([tail_teacher, rnd_t], 1), torch.cat([tail_student, rnd_s], 1)
        if self.n_random_relations > 0:  # :927-949
            rnd_t = self._rng.choice(list(self.mapping_relations.keys()), size=self.n_random_relations, replace=False)
            rnd_s = torch.tensor([[self.mapping_relations[i] for i in rnd_t]], dtype=torch.int64, device=dev).expand(B, -1)
            rnd_t = torch.tensor(np.asarray(rnd_t)[None, :], dtype=torch.int64, device=dev).expand(B, -1)
            relation_teacher = torch.cat([relation_teacher, rnd_t], 1)
            relation_student = torch.cat([relation_student, rnd_s], 1)
        return (head_teacher, relation_teacher, tail_teacher, head_student, relation_student, tail_student)
