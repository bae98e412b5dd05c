import torch
from torch import nn


class LowerBound(nn.Module):
    """max(x, bound) with a registered buffer `bound` (forward value only; the
    straight-through gradient of the original is irrelevant for eval)."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)]))

    def forward(self, x):
        return torch.max(x, self.bound)
