"""Negative sampling on the device; mirror of mkb/sampling/negative_sampling.py."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..utils.filters import build_filter_csr

__all__ = ["NegativeSampling", "positive_triples"]


def positive_triples(triples):
    """Reference-shaped dictionaries (negative_sampling.py:7-28): ``true_head[(r,t)]`` and
    ``true_tail[(h,r)]`` -> np.array of unique entities.  Kept for API compatibility; the kernels
    use the CSR form."""
    true_head, true_tail = {}, {}
    for h, r, t in triples:
        true_tail.setdefault((h, r), set()).add(t)
        true_head.setdefault((r, t), set()).add(h)
    return ({k: np.array(sorted(v)) for k, v in true_head.items()},
            {k: np.array(sorted(v)) for k, v in true_tail.items()})


class NegativeSampling:
    """``NegativeSampling(size, train_triples, entities, relations, seed=42).generate(sample, mode)``
    -> ``LongTensor[B, size]`` of corrupted heads (head-batch) or tails (tail-batch), none of which
    forms a training triple with the positive (negative_sampling.py:133-201).

    pool="independent" (default): every output slot draws independently from its own Philox4x32-10
        stream on the GPU — no host work, no H2D copy on the step.  Rows come back sorted by id
        (``sort_rows=True``): same multiset, order chosen for the scoring kernels' cache behaviour.
    pool="reference": the reference's exact semantics — ONE pool of ``2*size`` candidates per call
        drawn from ``np.random.RandomState(seed).randint`` (:151,:166), every positive takes the first
        ``size`` survivors of that shared pool (cyclically repeated when fewer survive, :176-195).
        The 2*size ids are the only bytes copied to the device.
    """

    def __init__(self, size, train_triples, entities, relations, seed=42, pool="independent", device=None,
                 sort_rows=True):
        if pool not in ("independent", "reference"):
            raise ValueError("pool must be 'independent' or 'reference'")
        self.size = size
        self.n_entity = len(entities)
        self.n_relation = len(relations)
        # seed=None is legal in the reference (np.random.RandomState(None), negative_sampling.py:151; KdmkbModel's
        # default): resolve it once to a concrete value so the device Philox key and the RandomState agree
        if seed is None:
            seed = int(np.random.SeedSequence().entropy) & (2 ** 32 - 1)
        self.seed = int(seed)
        self.pool = pool
        self.sort_rows = sort_rows  # independent pool only: rows ascending by id (L2-friendly gathers)
        self._rng = np.random.RandomState(self.seed)
        self._calls = 0
        self._host_csr = {side: build_filter_csr(train_triples, self.n_entity, side) for side in ("head", "tail")}
        self._dev_csr = {}
        self._status = {}
        if device is not None:
            self._csr("head", torch.device(device))
            self._csr("tail", torch.device(device))

    # reference attribute names, built lazily (they are only needed by callers poking at them)
    @property
    def true_head(self):
        k, o, m = self._host_csr["head"]
        return {(int(c // self.n_entity), int(c % self.n_entity)): m[o[i]:o[i + 1]] for i, c in enumerate(k)}

    @property
    def true_tail(self):
        k, o, m = self._host_csr["tail"]
        return {(int(c % self.n_entity), int(c // self.n_entity)): m[o[i]:o[i + 1]] for i, c in enumerate(k)}

    def _csr(self, side, device):
        key = (side, device)
        if key not in self._dev_csr:
            k, o, m = self._host_csr[side]
            self._dev_csr[key] = ops.FilterCSR(torch.from_numpy(k).to(device), torch.from_numpy(o).to(device),
                                               torch.from_numpy(m).to(device))
            self._status[device] = torch.zeros(1, dtype=torch.int32, device=device)
        return self._dev_csr[key]

    def check_status(self, device=None):
        """Synchronising check of the device-side status word: raises KeyError when a positive's key
        was absent from the training triples (the reference's dict lookup raises at :180/:187) and
        RuntimeError when a true set left no admissible candidate (the reference spins forever)."""
        for dev, st in self._status.items():
            if device is not None and dev != torch.device(device):
                continue
            v = int(st.item())
            st.zero_()
            if v & 1:
                raise KeyError("a positive (head, relation) / (relation, tail) key is not in train_triples")
            if v & 6:
                raise RuntimeError("no admissible negative: the true set covers every candidate")

    def generate(self, sample, mode, check=False):
        if mode not in ("head-batch", "tail-batch"):
            raise ValueError(f"unknown mode {mode!r}")
        src_device = sample.device
        if not sample.is_cuda:
            if not torch.cuda.is_available():
                raise ops.N.KgeError("mkb_b200.sampling needs a CUDA device (no CPU fallback)")
            sample = sample.cuda()
        device = sample.device
        sample = sample.to(torch.int64).contiguous()
        csr = self._csr("head" if mode == "head-batch" else "tail", device)
        status = self._status[device]
        if self.pool == "reference":
            pool = torch.from_numpy(self._rng.randint(self.n_entity, size=self.size * 2).astype(np.int64))
            out, _ = ops.filter_pool(csr, sample, mode, self.size, self.n_entity, pool.to(device), status)
        else:
            out, _ = ops.sample_negatives(csr, sample, mode, self.size, self.n_entity, self.seed, self._calls, status,
                                          sort_rows=self.sort_rows)
        self._calls += 1
        if check:
            self.check_status(device)
        return out if src_device == device else out.to(src_device)
