"""TransE / DistMult / ComplEx / RotatE: constructors as in mkb/models/{transe,distmult,complex,rotate}.py."""
from math import pi

import torch
import torch.nn as nn

from .base import BaseModel

__all__ = ["TransE", "DistMult", "ComplEx", "RotatE", "pRotatE"]


class TransE(BaseModel):
    """score = gamma - || h + r - t ||_1   (mkb/models/transe.py:55-76)."""

    def _top_k(self, sample):
        """The points whose nearest entity / relation rows are the best heads, relations and tails of each
        triple: ``t - r``, ``t - h``, ``h + r`` (transe.py:78-84)."""
        head, relation, tail, _ = self.batch(sample=sample)
        return tail - relation, tail - head, head + relation


class DistMult(BaseModel):
    """score = sum_d h * r * t   (mkb/models/distmult.py:53-75)."""


class ComplEx(BaseModel):
    """score = Re<h, r, conj(t)>; rows are [re | im]   (mkb/models/complex.py:55-85)."""

    _entity_mult = 2
    _relation_mult = 2


class RotatE(BaseModel):
    """score = gamma - sum_d | h∘r - t | with r = exp(i * relation / (range / pi))
    (mkb/models/rotate.py:60-99).  ``modulus`` is carried only because the reference has it
    (rotate.py:66-67): trainable, never used, its grad stays None."""

    _entity_mult = 2

    def __init__(self, hidden_dim, entities, relations, gamma):
        super().__init__(hidden_dim=hidden_dim, entities=entities, relations=relations, gamma=gamma)
        self.pi = pi
        self.modulus = nn.Parameter(torch.Tensor([[0.5 * self.embedding_range.item()]]))


class pRotatE(BaseModel):
    """score = gamma - modulus * sum_d | sin((h + r - t) / (range / pi)) | — RotatE on phases only, with a
    TRAINABLE scalar ``modulus`` initialised to 0.5 * range (mkb/models/protate.py:60-93).  Runs on the
    same kernel template as the other models; the modulus is read on the device."""

    def __init__(self, hidden_dim, entities, relations, gamma):
        super().__init__(hidden_dim=hidden_dim, entities=entities, relations=relations, gamma=gamma)
        self.pi = pi
        self.modulus = nn.Parameter(torch.Tensor([[0.5 * self.embedding_range.item()]]))
