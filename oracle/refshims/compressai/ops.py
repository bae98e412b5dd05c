import torch
from torch import nn


class _LowerBoundFn(torch.autograd.Function):
    """max(x, bound); the gradient passes where x >= bound or where it pushes x up (the
    behaviour of compressai's LowerBound, needed only by tools/train_reference_ckpt.py)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (grad_output < 0)) * grad_output, None


class LowerBound(nn.Module):
    """max(x, bound) with a registered buffer `bound`."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)]))

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            return _LowerBoundFn.apply(x, self.bound)
        return torch.max(x, self.bound)
