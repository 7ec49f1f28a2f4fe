"""Top heads / relations / tails of a query; mirror of mkb/utils/top_k.py (TopK).

The reference scores one triple per candidate with ``model(sample[N,3])`` and argsorts all N scores to
keep ``k`` (top_k.py:226-234).  Here the scores come from the positives kernel on the same ``[N,3]``
sample (same formula, same rounding as the reference's call) and the selection is ``kge_topk_rows``:
an exact radix-select top-k, ties by candidate order — no sort of the N scores.
"""
import torch

from .. import ops

__all__ = ["TopK"]


class TopK:
    """``TopK(entities, relations, device='cpu')``; ``top_heads(k, model, relation, tail)``,
    ``top_relations(k, model, head, tail)``, ``top_tails(k, model, head, relation)`` -> list of labels,
    best first.  Ids or labels are accepted for the fixed parts, as in the reference."""

    def __init__(self, entities, relations, device="cpu"):
        self.mapping_entities = entities
        self.mapping_relations = relations
        self.reverse_mapping_entities = {value: key for key, value in entities.items()}
        self.reverse_mapping_relations = {value: key for key, value in relations.items()}
        self.entities = torch.tensor([e for _, e in entities.items()], dtype=torch.int64)
        self.relations = torch.tensor([r for _, r in relations.items()], dtype=torch.int64)
        self.device = device

    def _entity(self, x):
        return self.mapping_entities[x] if isinstance(x, str) else x

    def _relation(self, x):
        return self.mapping_relations[x] if isinstance(x, str) else x

    @staticmethod
    def _get_rank(model, sample, k, device=None):
        """Positions (into ``sample``) of the ``k`` best-scoring triples, best first."""
        dev = model.entity_embedding.device
        training = model.training
        model.eval()
        with torch.no_grad():
            scores = model(sample.to(dev)).view(1, -1)
            rank = ops.topk_rows(scores, min(int(k), scores.shape[1])).flatten().cpu()
        if training:
            model.train()
        return rank

    def top_heads(self, k, model, relation, tail):
        n = self.entities.shape[0]
        sample = torch.stack([self.entities, torch.full((n,), self._relation(relation), dtype=torch.int64),
                              torch.full((n,), self._entity(tail), dtype=torch.int64)], dim=1)
        rank = self._get_rank(model=model, sample=sample, k=k, device=self.device)
        return [self.reverse_mapping_entities[e.item()] for e in self.entities[rank]]

    def top_relations(self, k, model, head, tail):
        n = self.relations.shape[0]
        sample = torch.stack([torch.full((n,), self._entity(head), dtype=torch.int64), self.relations,
                              torch.full((n,), self._entity(tail), dtype=torch.int64)], dim=1)
        rank = self._get_rank(model=model, sample=sample, k=k, device=self.device)
        return [self.reverse_mapping_relations[r.item()] for r in self.relations[rank]]

    def top_tails(self, k, model, head, relation):
        n = self.entities.shape[0]
        sample = torch.stack([torch.full((n,), self._entity(head), dtype=torch.int64),
                              torch.full((n,), self._relation(relation), dtype=torch.int64), self.entities], dim=1)
        rank = self._get_rank(model=model, sample=sample, k=k, device=self.device)
        return [self.reverse_mapping_entities[e.item()] for e in self.entities[rank]]
