"""CPU port of the reference's per-batch training step in eager PyTorch — TEST / BASELINE
INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/kge_oracle.py's header for who may import this).

Why a second oracle: the reference *is* eager PyTorch on the CPU, so its cost (and its fp32 rounding)
is that of the ATen operator sequence it issues, multi-threaded by ATen's intra-op pool.  This module
restates that operator sequence function by function so that
  * ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs can time "the reference's CPU path" on
    the GPU box, where /root/reference does not exist, and
  * tests can check that the sequence reproduces the reference's fp32 outputs bit for bit
    (tests/test_oracle_golden.py::test_torch_port_is_bit_exact).

Each function cites the reference lines whose operator sequence it follows.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def gather_rows(ent, rel, sample, neg, mode):
    """index_select gathers of mkb/models/base.py:166-207 -> head, relation, tail as [B,1|K,dim]."""
    rel_rows = torch.index_select(rel, 0, sample[:, 1]).unsqueeze(1)
    if mode == "head-batch":
        b, k = neg.shape
        head = torch.index_select(ent, 0, neg.view(-1)).view(b, k, -1)
        tail = torch.index_select(ent, 0, sample[:, 2]).unsqueeze(1)
    elif mode == "tail-batch":
        b, k = neg.shape
        head = torch.index_select(ent, 0, sample[:, 0]).unsqueeze(1)
        tail = torch.index_select(ent, 0, neg.view(-1)).view(b, k, -1)
    else:
        head = torch.index_select(ent, 0, sample[:, 0]).unsqueeze(1)
        tail = torch.index_select(ent, 0, sample[:, 2]).unsqueeze(1)
    return head, rel_rows, tail


def forward(model, ent, rel, sample, neg=None, mode=None, *, gamma, embedding_range):
    """The four forward()s: transe.py:65-76, distmult.py:63-75, complex.py:65-85, rotate.py:69-99."""
    if neg is None:
        mode = None
    shape = (sample.size(0), 1) if neg is None else tuple(neg.shape)
    head, relation, tail = gather_rows(ent, rel, sample, neg, mode)
    if model == "TransE":
        s = head + (relation - tail) if mode == "head-batch" else (head + relation) - tail
        out = gamma - torch.norm(s, p=1, dim=2)
    elif model == "DistMult":
        s = head * (relation * tail) if mode == "head-batch" else (head * relation) * tail
        out = s.sum(dim=2)
    elif model == "ComplEx":
        re_h, im_h = torch.chunk(head, 2, dim=2)
        re_r, im_r = torch.chunk(relation, 2, dim=2)
        re_t, im_t = torch.chunk(tail, 2, dim=2)
        if mode == "head-batch":
            re_s = re_r * re_t + im_r * im_t
            im_s = re_r * im_t - im_r * re_t
            s = re_h * re_s + im_h * im_s
        else:
            re_s = re_h * re_r - im_h * im_r
            im_s = re_h * im_r + im_h * re_r
            s = re_s * re_t + im_s * im_t
        out = s.sum(dim=2)
    elif model == "RotatE":
        re_h, im_h = torch.chunk(head, 2, dim=2)
        re_t, im_t = torch.chunk(tail, 2, dim=2)
        phase = relation / (embedding_range / math.pi)
        re_r, im_r = torch.cos(phase), torch.sin(phase)
        if mode == "head-batch":
            re_s = re_r * re_t + im_r * im_t
            im_s = re_r * im_t - im_r * re_t
            re_s = re_s - re_h
            im_s = im_s - im_h
        else:
            re_s = re_h * re_r - im_h * im_r
            im_s = re_h * im_r + im_h * re_r
            re_s = re_s - re_t
            im_s = im_s - im_t
        s = torch.stack([re_s, im_s], dim=0).norm(dim=0)  # rotate.py:95-96 (the slow CPU op, App. C.2)
        out = gamma - s.sum(dim=2)
    else:
        raise ValueError(model)
    return out.view(shape)


def adversarial(pos, neg, weight, alpha=0.5):
    """mkb/losses/adversarial.py:21-30."""
    p = F.logsigmoid(pos).squeeze(dim=1)
    n = (F.softmax(neg * alpha, dim=1).detach() * F.logsigmoid(-neg)).sum(dim=1)
    return (-(weight * p).sum() / weight.sum() - (weight * n).sum() / weight.sum()) / 2


def true_sets(triples):
    """positive_triples (mkb/sampling/negative_sampling.py:7-28) as dicts of numpy arrays."""
    th, tt = {}, {}
    for h, r, t in triples:
        tt.setdefault((h, r), set()).add(t)
        th.setdefault((r, t), set()).add(h)
    return ({k: np.array(list(v)) for k, v in th.items()}, {k: np.array(list(v)) for k, v in tt.items()})


def generate_negatives(rng, sample, mode, true_head, true_tail, n_entity, size):
    """NegativeSampling.generate (negative_sampling.py:158-201): one shared pool per call, Python
    loop over the positives, np.in1d filter, first `size` survivors (re-filtering until enough)."""
    pool = rng.randint(n_entity, size=size * 2)
    rows = []
    for h, r, t in sample.tolist():
        rec = true_head[(r, t)] if mode == "head-batch" else true_tail[(h, r)]
        got, parts = 0, []
        while got < size:
            keep = pool[np.isin(pool, rec, assume_unique=True, invert=True)]
            if keep.size == 0:
                raise RuntimeError("empty filtered pool")
            got += keep.size
            parts.append(keep)
        rows.append(torch.from_numpy(np.concatenate(parts)[:size].astype(np.int64)))
    return torch.stack(rows, dim=0)


class CpuTrainer:
    """The loop body of mkb/compose/pipeline.py:206-242 on CPU tensors with a dense torch Adam
    (README.md:123-126).  ``step`` returns the loss as a Python float (``error.item()``, :242)."""

    def __init__(self, model, n_entity, n_relation, hidden_dim, gamma, lr=5e-5, seed=42, alpha=0.5, device="cpu"):
        """``device="cuda"`` runs the SAME eager operator sequence on a GPU (what the reference does with
        ``device='cuda'``): bench.py reports it next to the CPU baseline, SURVEY §8(d)."""
        from .kge_oracle import embedding_range, entity_dim, relation_dim

        self.model, self.gamma, self.alpha = model, float(gamma), alpha
        self.n_entity = n_entity
        self.embedding_range = embedding_range(gamma, hidden_dim)
        g = torch.Generator().manual_seed(seed)
        r = self.embedding_range
        self.ent = torch.nn.Parameter(
            ((torch.rand(n_entity, entity_dim(model, hidden_dim), generator=g) * 2 - 1) * r).to(device))
        self.rel = torch.nn.Parameter(
            ((torch.rand(n_relation, relation_dim(model, hidden_dim), generator=g) * 2 - 1) * r).to(device))
        self.opt = torch.optim.Adam([self.ent, self.rel], lr=lr)

    def step(self, sample, weight, mode, negative_sample):
        pos = forward(self.model, self.ent, self.rel, sample, gamma=self.gamma, embedding_range=self.embedding_range)
        neg = forward(self.model, self.ent, self.rel, sample, negative_sample, mode, gamma=self.gamma,
                      embedding_range=self.embedding_range)
        err = adversarial(pos, neg, weight, self.alpha)
        err.backward()
        self.opt.step()
        self.opt.zero_grad()
        return err.item()
