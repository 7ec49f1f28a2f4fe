"""Multi-GPU plumbing of the training step (one process per GPU, torch.distributed).

The path shards by positives: rank r scores its own slice of the global batch against its replica of
the tables.  Two exchanges make the result identical to one process running the global batch:

  1. the three loss sums (S_p = Σ w logσ(p), S_n = Σ w Σ a logσ(-n), W = Σ w) are summed over ranks
     BEFORE the backward, because every per-score gradient carries the factor 1/(2W) with W the
     global sum of weights (mkb/losses/adversarial.py:28-30 applied to the global batch);
  2. the dense gradients are summed over ranks before the (replicated) optimizer step.

Backend-agnostic on purpose (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["rank_slices", "allreduce_loss_sums", "allreduce_gradients", "loss_from_sums"]


def rank_slices(order, batch_size, world_size, rank):
    """Slice a global visiting order into this rank's batches: global batch g is
    ``order[g*world*B : (g+1)*world*B]``, rank r takes its r-th block of B.  A trailing partial
    global batch is split as evenly as possible (ranks may get one positive more or fewer)."""
    order = np.asarray(order)
    out = []
    step = batch_size * world_size
    for lo in range(0, len(order), step):
        chunk = order[lo:lo + step]
        parts = np.array_split(chunk, world_size)
        out.append(parts[rank])
    return out


def allreduce_loss_sums(stats, group=None):
    """In-place SUM all-reduce of stats[0:3] = (S_p, S_n, W)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats[:3], op=dist.ReduceOp.SUM, group=group)
    return stats


def allreduce_gradients(flat_grad, group=None):
    """In-place SUM all-reduce of the flat [entity | relation] gradient buffer."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def loss_from_sums(stats):
    """L = -(S_p + S_n) / (2 W) on whatever device `stats` lives."""
    return -(stats[0] + stats[1]) / (2 * stats[2])
