"""Host-side mirror of mkb/models/base.py (Base, BaseModel): same constructor, attributes and call
contract; the gathers + scoring run in the CUDA library instead of ATen."""
from __future__ import annotations

import pickle

import torch
import torch.nn as nn

from .. import ops

__all__ = ["BaseModel"]


class BaseModel(nn.Module):
    """Knowledge-graph-embedding model: two fp32 tables + constants (mkb/models/base.py:66-100).

    Parameters mirror the reference: ``entities`` / ``relations`` map label -> id; ``gamma`` sets
    the margin and the init range ``(gamma + 2) / hidden_dim``.
    """

    _entity_mult = 1
    _relation_mult = 1

    def __init__(self, hidden_dim, entities, relations, gamma):
        super().__init__()
        self.entities = {i: e for e, i in entities.items()}
        self.relations = {i: r for r, i in relations.items()}
        self.n_entity = len(entities)
        self.n_relation = len(relations)
        self.hidden_dim = hidden_dim
        self.entity_dim = hidden_dim * self._entity_mult
        self.relation_dim = hidden_dim * self._relation_mult
        self.epsilon = 2
        self.gamma = nn.Parameter(torch.Tensor([gamma]), requires_grad=False)
        self.embedding_range = nn.Parameter(
            torch.Tensor([(self.gamma.item() + self.epsilon) / self.hidden_dim]), requires_grad=False)
        # same creation order and init calls as the reference => same values under a given torch seed
        self.entity_embedding = nn.Parameter(torch.zeros(self.n_entity, self.entity_dim))
        nn.init.uniform_(self.entity_embedding, a=-self.embedding_range.item(), b=self.embedding_range.item())
        self.relation_embedding = nn.Parameter(torch.zeros(self.n_relation, self.relation_dim))
        nn.init.uniform_(self.relation_embedding, a=-self.embedding_range.item(), b=self.embedding_range.item())
        self._spec = None

    # -- reference API ---------------------------------------------------------------------
    @property
    def name(self):
        return self.__class__.__name__

    @property
    def embeddings(self):
        """{'entities': {label: tensor}, 'relations': {label: tensor}} (base.py:102-116)."""
        ent = {self.entities[i]: self.entity_embedding[i].detach() for i in range(self.n_entity)}
        rel = {self.relations[i]: self.relation_embedding[i].detach() for i in range(self.n_relation)}
        return {"entities": ent, "relations": rel}

    def save(self, path):
        with open(path, "wb") as handle:
            pickle.dump(self.cpu().eval(), handle, protocol=pickle.HIGHEST_PROTOCOL)

    def _set_params(self, entities_embeddings, relations_embeddings, **kwargs):
        self.entity_embedding.data.copy_(entities_embeddings)
        self.relation_embedding.data.copy_(relations_embeddings)
        for parameter, weights in kwargs.items():
            self._parameters[parameter].data.copy_(weights)
        return self

    def distill(self, sample, negative_sample=None, mode=None):
        return self(sample=sample, negative_sample=negative_sample, mode=mode)

    @property
    def _repr_content(self):
        return {
            "Entities embeddings dim": f"{self.entity_dim}",
            "Relations embeddings dim": f"{self.relation_dim}",
            "Gamma": f"{self.gamma.item()}",
            "Number of entities": f"{self.n_entity}",
            "Number of relations": f"{self.n_relation}",
        }

    def __repr__(self):
        l_len = max(map(len, self._repr_content.keys()))
        r_len = max(map(len, self._repr_content.values()))
        return f"{self.name} model\n" + "\n".join(
            k.rjust(l_len) + "  " + v.ljust(r_len) for k, v in self._repr_content.items())

    @staticmethod
    def format_sample(sample, negative_sample=None):
        """Shape bookkeeping of mkb/models/base.py:132-151."""
        if sample.dim() == 2:
            if negative_sample is None:
                return sample, (sample.size(0), 1)
            return sample, tuple(negative_sample.shape)
        if sample.dim() == 3:
            return sample.reshape(sample.size(0) * sample.size(1), 3), (sample.size(0), sample.size(1))
        raise ValueError("sample must be 2-D [B,3] or 3-D [n,b,3]")

    def batch(self, sample, negative_sample=None, mode=None):
        """``(head, relation, tail, shape)`` embedding rows of a sample, shaped ``[B,1,dim]`` / ``[B,K,dim]`` as
        mkb/models/base.py:153-207 returns them.  An accessor for callers that want the rows themselves
        (``TransE._top_k``, nearest-neighbour samplers): ``forward`` never materialises these tensors — the
        gathers happen inside the scoring kernels."""
        flat, shape = self.format_sample(sample, negative_sample)
        ent, rel = self.entity_embedding, self.relation_embedding
        head, relation, tail = ent[flat[:, 0]].unsqueeze(1), rel[flat[:, 1]].unsqueeze(1), ent[flat[:, 2]].unsqueeze(1)
        if mode == "head-batch":
            head = ent[negative_sample.reshape(-1)].view(negative_sample.size(0), negative_sample.size(1), -1)
        elif mode == "tail-batch":
            tail = ent[negative_sample.reshape(-1)].view(negative_sample.size(0), negative_sample.size(1), -1)
        return head, relation, tail, shape

    # -- kernels ---------------------------------------------------------------------------
    @property
    def spec(self):
        """Kernel-side constants (cached; gamma / embedding_range are frozen Parameters)."""
        if self._spec is None:
            self._spec = ops.TableSpec(self.name, self.hidden_dim, self.gamma.item(), self.embedding_range.item())
        return self._spec

    @property
    def kernel_modulus(self):
        """The modulus the kernels read: pRotatE's trainable scalar, None for every other model (RotatE
        carries an unused one, rotate.py:66-67)."""
        return self.modulus if self.name == "pRotatE" else None

    def forward(self, sample, negative_sample=None, mode=None):
        """``model(sample)``, ``model(sample, negative_sample, mode)``, ``model(sample[n,b,3])``
        (mkb/models/base.py:153-207 + transe.py:65-76 / distmult.py:63-75 / complex.py:65-85 /
        rotate.py:69-99) through kge_score_fwd; differentiable through kge_score_bwd."""
        flat, shape = self.format_sample(sample, negative_sample)
        if sample.dim() == 3:
            negative_sample, mode = None, None
        out = ops.score(self.spec, self.entity_embedding, self.relation_embedding, flat, negative_sample, mode,
                        self.kernel_modulus)
        return out.view(shape)
