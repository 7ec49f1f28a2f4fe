"""The consumers of the score / top-k kernels in mkb's distillation add-on.

Only the sampler that IS a "score every candidate -> keep the k best" reduce lives here
(``TopKSampling``, mkb/distillation/top_k_sampling.py:321-677); the distillation training loop
(``Distillation``, ``KdmkbModel``) and the faiss-based samplers are outside the hot path.
"""
from .top_k_sampling import TopKSampling

__all__ = ["TopKSampling"]
