"""Losses that run as CUDA kernels behind the C ABI: the self-adversarial negative-sampling loss
(kge_adv_loss_*, fused into kge_fused_fwd on the training path) and the KL-divergence distillation loss
(kge_kl_div_*)."""
from .adversarial import Adversarial
from .kl_divergence import KlDivergence

__all__ = ["Adversarial", "KlDivergence"]
