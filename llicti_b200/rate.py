"""Rate bookkeeping of the eval loop: bits per pixel per (scale, band, channel) stream
(mirrors graphs/losses/rate_dist.py:125-135 and the 'te' table of loggers/rate.py)."""
import logging
from datetime import datetime

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


class TrainRLossList:
    """Rate estimate from the self-informations of forward() (graphs/losses/rate_dist.py:97-104): per scale the nine
    sums over (batch, rows, columns) / numel * 3 bits per pixel, and their total.  When the self-informations carry an
    autograd edge (the training step) the total is a tensor to call `.backward()` on, as in the reference; otherwise a
    float summed in double precision (validate())."""

    def __init__(self):
        self.rate1 = 0.0
        self.rate1list = []

    def forward(self, numel_x, sinfoslist):
        if any(getattr(s, "requires_grad", False) for s in sinfoslist):
            import torch
            self.rate1, self.rate1list = 0, []
            for s in sinfoslist:
                per = torch.sum(s, dim=(0, 2, 3)) / numel_x * 3
                self.rate1list.append(per.tolist())
                self.rate1 = self.rate1 + torch.sum(per)
            return self.rate1, self.rate1list
        self.rate1list = [[float(v) / numel_x * 3 for v in s.double().sum(dim=(0, 2, 3)).tolist()] for s in sinfoslist]
        self.rate1 = float(sum(sum(r) for r in self.rate1list))
        return self.rate1, self.rate1list

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
        self.logger.info(self.format_table(self.current_epoch, rate, lr, typ))
        return float(rate.sum()), 0.0

    # first line / continuation-line prefixes per table type, as the reference prints them
    _HEAD = {"tr": "  Train Epoch: {:3d}  Rates: scl", "te": "   Test Epoch: {:3d}  Rates: hdr ",
             "va": "  Valid Epoch: {:3d}  Rates: scl", "it": "Train Itera: {:3d}  Rates: scl"}
    _CONT = {"tr": " " * 35 + "scl", "te": " " * 35 + "scl", "va": " " * 35 + "scl", "it": " " * 33 + "scl"}

    @classmethod
    def format_table(cls, epoch, rate, lr=0.0, typ="te", now=None):
        """The reference's table text (loggers/rate.py:120-168, text_log_list): one line per row of `rate`
        ([rows][9] bits per pixel, 3 bands x (Y, Co, Cg)); for typ 'te' row 0 is the header row ("hdr ->", "hd=")
        and row s > 0 is printed as scale s-1."""
        rate = np.asarray(rate, dtype=np.float64)
        assert rate.ndim == 2 and rate.shape[1] == 9, "expected 3 bands x 3 colour channels per row"
        te = typ == "te"
        lines, total = [], 0.0
        for s in range(rate.shape[0]):
            label = "-> " if te and s == 0 else "{:d}-> ".format(s - 1 if te else s)
            cells, row_sum = "", 0.0
            for b in range(3):
                y, co, cg = rate[s][3 * b:3 * b + 3]
                band_sum = y + co + cg
                cells += "{:.2f}+{:.2f}+{:.2f}(b{:d}={:.3f}) ".format(y, co, cg, b, band_sum)
                row_sum += band_sum
            tail = "(hd={:.3f}) ".format(row_sum) if te and s == 0 else "(s{:d}={:.3f}) ".format(s - 1 if te else s, row_sum)
            total += row_sum
            prefix = cls._HEAD[typ].format(epoch) if s == 0 else cls._CONT[typ]
            lines.append(prefix + label + cells + tail)
        stamp = (now or datetime.now()).strftime("%H:%M:%S")
        end = "(({:.3f})) ".format(total) + ("  (lr: {:.6f}) ({})".format(lr, stamp) if typ in ("tr", "it") else " ({})".format(stamp))
        return "\n".join(lines) + end
