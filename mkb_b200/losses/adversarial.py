"""Self-adversarial negative-sampling loss (Sun et al. 2019), mirror of mkb/losses/adversarial.py."""
from .. import ops

__all__ = ["Adversarial"]


class Adversarial:
    """``Adversarial(alpha)(positive_score[B,1], negative_score[B,K], weight[B]) -> 0-d tensor``.

    A plain callable like the reference's (which subclasses nn.Module without initialising it,
    adversarial.py:18-21).  One CUDA kernel forward, one backward."""

    def __init__(self, alpha=0.5):
        self.alpha = alpha

    def __call__(self, positive_score, negative_score, weight):
        return ops.adversarial_loss(positive_score, negative_score, weight, self.alpha)
