"""KL-divergence distillation loss, mirror of mkb/losses/kl_divergence.py."""
from .. import ops

__all__ = ["KlDivergence"]


class KlDivergence:
    """``KlDivergence()(student_score[n,k], teacher_score[n,k], T=1) -> 0-d tensor``:
    ``mean(kl_div(log_softmax(student / T, dim=1), softmax(teacher / T, dim=1), reduction='none'))``
    (kl_divergence.py:22-29).  A plain callable like the reference's; one CUDA kernel forward, one
    backward (gradients w.r.t. the student and, when it requires grad, the teacher)."""

    def __init__(self):
        pass

    def __call__(self, student_score, teacher_score, T=1):
        return ops.kl_divergence(student_score, teacher_score, T)
