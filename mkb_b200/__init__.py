"""mkb_b200 — the mkb (raphaelsty/mkb) training/evaluation hot path on B200 (sm_100a) kernels.

Same public surface as the reference for the path it covers::

    from mkb_b200 import datasets, models, sampling, losses, evaluation, compose

Everything numeric runs in ``libkge_b200.so`` (hand-written CUDA behind the C ABI declared in
``include/kge_b200.h``); there is no CPU fallback.
"""
from . import compose, datasets, distillation, evaluation, losses, models, optim, ops, sampling, utils  # noqa: F401

__version__ = "0.1.0"
__all__ = ["compose", "datasets", "distillation", "evaluation", "losses", "models", "optim", "ops", "sampling", "utils"]
