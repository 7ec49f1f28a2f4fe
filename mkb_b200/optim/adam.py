"""Dense Adam on the kge_adam_step kernel: torch.optim.Adam's update rule (no weight decay, no
amsgrad) for fp32 CUDA parameters, one launch per tensor, ``zero_grad`` optionally folded in."""
import torch

from .. import ops

__all__ = ["DenseAdam"]


class DenseAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, fused_zero_grad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.fused_zero_grad = fused_zero_grad

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                ops.adam_step(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], st["step"], group["lr"], b1, b2,
                              group["eps"], zero_grad=self.fused_zero_grad)
        return loss

    def zero_grad(self, set_to_none=True):
        if self.fused_zero_grad:
            return  # gradients were cleared by the step kernel and stay allocated
        super().zero_grad(set_to_none=set_to_none)
