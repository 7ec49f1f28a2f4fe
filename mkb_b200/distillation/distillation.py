"""``Distillation`` — mirror of mkb/distillation/distillation.py:217-677.

``distill(teacher, student, sample)`` asks the sampler for a candidate set per positive (heads, relations,
tails, as teacher ids and as student ids), scores the resulting triples with both models and sums three
KL divergences (teacher no-grad, student differentiable).  The reference assembles the 3-D samples triple by
triple in Python (:576-622: a loop with ``.item()``, deep copies and ``torch.stack``); here availability and
the id mappings are two lookup tensors on the device and the six ``[n, k, 3]`` samples are built with a
handful of indexing ops.  The arithmetic is this package's kernels: ``model(sample[n,k,3])`` is K1's 3-D
path (``kge_score_fwd`` / ``kge_score_bwd``), the loss ``kge_kl_div_fwd`` / ``kge_kl_div_bwd``.
"""
from __future__ import annotations

import collections

import torch

from ..losses import KlDivergence

__all__ = ["Distillation"]


class Distillation:
    def __init__(self, teacher_entities, student_entities, teacher_relations, student_relations, sampling,
                 device="cpu"):
        self.teacher_entities = teacher_entities
        self.student_entities = student_entities
        self.teacher_relations = teacher_relations
        self.student_relations = student_relations
        self.sampling = sampling
        self.device = device
        self.mapping_entities = collections.OrderedDict(
            {i: student_entities[e] for e, i in teacher_entities.items() if e in student_entities})
        self.mapping_relations = collections.OrderedDict(
            {i: student_relations[e] for e, i in teacher_relations.items() if e in student_relations})
        # teacher id -> student id, -1 where the label is not shared
        self._ent_map = torch.full((max(teacher_entities.values(), default=-1) + 1,), -1, dtype=torch.int64)
        for t, s in self.mapping_entities.items():
            self._ent_map[t] = s
        self._rel_map = torch.full((max(teacher_relations.values(), default=-1) + 1,), -1, dtype=torch.int64)
        for t, s in self.mapping_relations.items():
            self._rel_map[t] = s
        self._maps_dev = None

    def available(self, head, relation, tail):
        """Which of the three distillations a triple (teacher ids) takes part in (:250-288)."""
        h, r, t = head in self.mapping_entities, relation in self.mapping_relations, tail in self.mapping_entities
        if self.sampling.supervised:
            every = h and r and t
            return {"head": every, "relation": every, "tail": every}
        return {"head": r and t, "relation": h and t, "tail": h and r}

    def _maps(self, dev):
        if self._maps_dev is None or self._maps_dev[0] != dev:
            self._maps_dev = (dev, self._ent_map.to(dev), self._rel_map.to(dev))
        return self._maps_dev[1:]

    @staticmethod
    def _triples(h, r, t, k):
        """[n, k, 3] sample from three operands that are either [n] (fixed part) or [n, k] (the candidates)."""
        parts = [x.view(-1, 1).expand(-1, k) if x.dim() == 1 else x for x in (h, r, t)]
        return torch.stack(parts, dim=2).contiguous()

    def distill(self, teacher, student, sample):
        dev = student.entity_embedding.device
        sample = sample.to(dev)
        dists = self.sampling.get(sample=sample, mapping_entities=self.mapping_entities,
                                  mapping_relations=self.mapping_relations, positive_sample_size=sample.shape[0],
                                  teacher=teacher)
        ht, rt, tt, hs, rs, ts = (d.to(device=dev, dtype=torch.int64) for d in dists)
        ent_map, rel_map = self._maps(dev)
        h, r, t = sample[:, 0], sample[:, 1], sample[:, 2]
        sh, sr, st = ent_map[h], rel_map[r], ent_map[t]  # student ids of the positives (-1: not shared)
        ok_h, ok_r, ok_t = sh >= 0, sr >= 0, st >= 0
        sup = bool(self.sampling.supervised)
        if sup:
            use = {"head": ok_h & ok_r & ok_t}
            use["relation"] = use["tail"] = use["head"]
        else:
            use = {"head": ok_r & ok_t, "relation": ok_h & ok_t, "tail": ok_h & ok_r}

        def with_truth(dist, truth):  # supervised samplers carry the ground truth in the last slot (:308-309)
            if not sup:
                return dist
            dist = dist.clone()
            dist[:, -1] = truth
            return dist

        kl = KlDivergence()
        loss = 0
        groups = (
            ("head", ht, hs, lambda m, d, k: self._triples(with_truth(d[m], h[m]), r[m], t[m], k),
             lambda m, d, k: self._triples(with_truth(d[m], sh[m]), sr[m], st[m], k)),
            ("relation", rt, rs, lambda m, d, k: self._triples(h[m], with_truth(d[m], r[m]), t[m], k),
             lambda m, d, k: self._triples(sh[m], with_truth(d[m], sr[m]), st[m], k)),
            ("tail", tt, ts, lambda m, d, k: self._triples(h[m], r[m], with_truth(d[m], t[m]), k),
             lambda m, d, k: self._triples(sh[m], sr[m], with_truth(d[m], st[m]), k)),
        )
        for name, d_teacher, d_student, build_t, build_s in groups:
            m = use[name]
            if not bool(m.any()):
                continue
            k = d_teacher.shape[1]
            with torch.no_grad():
                scores_teacher = teacher(build_t(m, d_teacher, k))
            loss = loss + kl(teacher_score=scores_teacher, student_score=student(build_s(m, d_student, k)))
        return loss
