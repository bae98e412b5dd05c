from torch import nn


class GDN1(nn.Module):  # import-only in the reference (activfun == "GDN1" is not a shipped config)
    def __init__(self, in_channels):
        super().__init__()
        raise NotImplementedError("GDN1 is not used by the shipped configs")
