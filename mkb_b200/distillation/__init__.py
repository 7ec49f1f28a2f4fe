"""mkb's distillation add-on on this package's kernels: the loss of a distillation
step (``Distillation.distill``: 3-D samples through the score kernel, KL-divergence kernels) and the two
samplers that need no nearest-neighbour index (``UniformSampling``; ``TopKSampling`` and its pre-computed form ``FastTopKSampling``, a "score every candidate
-> keep the k best" reduce on the score + exact top-k kernels).  Mirrors mkb/distillation/{distillation,
uniform_sampling,top_k_sampling}.py; ``KdmkbModel`` drives several KBs through those pieces; ``TopKSamplingTransE`` is the
nearest-neighbour sampler of TransE teachers without its faiss dependency (exact L2 search on the device).
"""
from .distillation import Distillation
from .kdmkb_model import KdmkbModel
from .top_k_sampling import FastTopKSampling, TopKSampling, TopKSamplingTransE
from .uniform_sampling import UniformSampling

__all__ = ["Distillation", "KdmkbModel", "FastTopKSampling", "TopKSampling", "TopKSamplingTransE", "UniformSampling"]
