import torch
from torch import nn

from .ops import LowerBound


class EntropyBottleneck(nn.Module):  # import-only on the eval_model path
    def __init__(self, channels=0, *args, **kwargs):
        super().__init__()

    def loss(self):
        return 0.0


class GaussianConditional(nn.Module):
    """Only what GaussianConditionalLossless(GMM) inherits on the hot path:
    buffers, lower_bound_scale, likelihood_lower_bound, _standardized_cumulative."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9,
                 likelihood_bound=1e-9, **kwargs):
        super().__init__()
        self.tail_mass = float(tail_mass)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.register_buffer("scale_table", torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self.lower_bound_scale = LowerBound(scale_bound)

    @staticmethod
    def _standardized_cumulative(inputs):
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)
