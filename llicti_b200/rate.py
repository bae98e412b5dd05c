"""Rate bookkeeping of the eval loop: bits per pixel per (scale, band, channel) stream
(mirrors graphs/losses/rate_dist.py:125-135 and the 'te' table of loggers/rate.py)."""
import logging

import numpy as np


class CompressionRLossList:
    """len(stream) * 8 / numel * 3 for every stream of a bytestream_list (bits per pixel;
    like the reference, container overhead inside a stream counts, nothing else does)."""

    def __init__(self):
        self.rate1list = []

    def forward(self, numel_x, bytestream_list):
        self.rate1list = [[len(s) * 8 / numel_x * 3 for s in row] for row in bytestream_list]
        return self.rate1list

    __call__ = forward


class RateLogger:
    """Accumulates per-image rate tables and prints their mean (header row + one row per
    scale, nine columns each = 3 bands x (Y, Co, Cg))."""

    def __init__(self):
        self.rate = []
        self.current_iteration = 0
        self.current_epoch = 0
        self.logger = logging.getLogger("Rate Loss")

    def __call__(self, rate):
        self.current_iteration += 1
        self.rate.append(rate)

    def mean(self):
        self.current_epoch += 1
        m = np.array(self.rate, dtype=np.float64).mean(axis=0)
        self.rate = []
        return m

    def state_dict(self):
        return {"rate": self.rate, "it": self.current_iteration, "ep": self.current_epoch}

    def load_state_dict(self, info):
        self.rate, self.current_iteration, self.current_epoch = info["rate"], info["it"], info["ep"]

    def display(self, lr=0.0, typ="te"):
        rate = self.mean()
        assert rate.shape[1] == 9, "expected 3 bands x 3 colour channels per scale"
        total = float(rate.sum())
        self.logger.info("  {} rate: {:.3f} bpp = {:.3f} bpsp".format(typ, total, total / 3))
        self.logger.info("    hdr : {:.3f}".format(float(rate[0].sum())))
        for i in range(1, rate.shape[0]):
            scl = rate.shape[0] - 1 - i
            cells = "  ".join("b{}: {:.3f}+{:.3f}+{:.3f}".format(b, *rate[i][3 * b:3 * b + 3]) for b in range(3))
            self.logger.info("    scl{}: {:.3f}  ({})".format(scl, float(rate[i].sum()), cells))
        return total, 0.0
