"""Synthetic inputs of the bench and the command-line tools: photographic-like images and deterministic stand-in
weights in the reference's state_dict naming (the shipped checkpoint is absent).

These are data generators, not part of the codec.  `oracle/llicti_oracle.py` carries the same two functions for
the tests (the oracle imports nothing from this package); `tests/test_abi_host.py` checks that both produce the
same bytes, so the golden fixtures, the parity tests and the bench see the same inputs.
"""
from typing import Dict

import numpy as np

# layer-0 branches of the three band models (LLICTI_nets.py:650-675)
_BAND_BRANCHES = {0: ["layer0_00_11"], 1: ["layer0_00_01", "layer0_11_01"], 2: ["layer0_00_10", "layer0_11_10", "layer0_01_10"]}


def synthetic_state_dict(chs: int = 88, num_mixtures: int = 5, evens: int = 4, odds: int = 3, seed: int = 1337) -> Dict[str, np.ndarray]:
    """Deterministic stand-in weights in the reference's state_dict naming
    (SURVEY.md section 8b).  The shipped checkpoint is absent, so tests and the bench use
    these.  Random weights alone would put every mean near zero and cost ~16 bit per
    symbol, so a hand-wired "interpolation path" is laid over the random weights: six
    hidden units of the mean sub-network carry +/- the average of the nearest known
    neighbours of each colour channel through both ReLU layers, so the predicted means
    interpolate the image, spreads are a few grey levels and symbol costs are realistic."""
    rng = np.random.default_rng(seed)
    g = int(chs)
    Ch = 4 * g
    M = int(num_mixtures)
    sd: Dict[str, np.ndarray] = {}
    pre = "entropymodel.entmdls_scale_band.0."
    Ev, Od = int(evens), int(odds)
    shapes = {"layer0_00_11": (Ev, Ev), "layer0_00_01": (Od, Ev), "layer0_11_01": (Ev, Od),
              "layer0_00_10": (Ev, Od), "layer0_11_10": (Od, Ev), "layer0_01_10": (Ev, Ev)}
    # nearest-neighbour taps (dy, dx, weight) inside each layer-0 kernel window
    taps = {"layer0_00_11": [(1, 1, .25), (1, 2, .25), (2, 1, .25), (2, 2, .25)],
            "layer0_00_01": [(1, 1, .25), (1, 2, .25)], "layer0_11_01": [(1, 1, .25), (2, 1, .25)],
            "layer0_00_10": [(1, 1, .25), (2, 1, .25)], "layer0_11_10": [(1, 1, .25), (1, 2, .25)],
            "layer0_01_10": []}
    for b, branches in _BAND_BRANCHES.items():
        for name in branches:
            kh, kw = shapes[name]
            fan = 3 * kh * kw * len(branches)
            w = rng.standard_normal((Ch, 3, kh, kw)).astype(np.float32) * np.float32(1.5 / np.sqrt(fan))
            bias = (rng.standard_normal(Ch) * 0.05).astype(np.float32)
            for c in range(3):
                for sgn, u in ((1.0, g + 2 * c), (-1.0, g + 2 * c + 1)):
                    w[u] = 0
                    bias[u] = 0
                    for dy, dx, tw in taps[name]:
                        w[u, c, dy, dx] = sgn * tw
            sd[f"{pre}{b}.{name}.weight"] = w
            sd[f"{pre}{b}.{name}.bias"] = bias
        w1 = (rng.standard_normal((Ch, g, 1, 1)) * (1.0 / np.sqrt(g))).astype(np.float32)
        b1 = (rng.standard_normal(Ch) * 0.05).astype(np.float32)
        for u in range(6):
            w1[g + u] = 0
            w1[g + u, u] = 1.0
            b1[g + u] = 0
        sd[f"{pre}{b}.layers1toL.0.weight"] = w1
        sd[f"{pre}{b}.layers1toL.0.bias"] = b1
        w2 = (rng.standard_normal((12 * M, g, 1, 1)) * (0.25 / np.sqrt(g))).astype(np.float32)
        b2 = np.zeros(12 * M, dtype=np.float32)
        # spread head: 1.5..10 grey levels; mean head: interpolation path + small offsets;
        # weight head: positive; coupling head: small
        b2[0:3 * M] = (np.abs(rng.standard_normal(3 * M)) * 3.0 + 1.5) / 255.0
        w2[0:3 * M] *= np.float32(0.02)
        b2[3 * M:6 * M] = rng.standard_normal(3 * M) * (2.0 / 255.0)
        w2[3 * M:6 * M] *= np.float32(0.05)
        for c in range(3):
            for m in range(M):
                w2[3 * M + c * M + m, 2 * c] = 1.0
                w2[3 * M + c * M + m, 2 * c + 1] = -1.0
        b2[6 * M:9 * M] = np.abs(rng.standard_normal(3 * M)) * 0.5 + 0.2
        b2[9 * M:12 * M] = rng.standard_normal(3 * M) * 0.02
        w2[9 * M:12 * M] *= np.float32(0.1)
        sd[f"{pre}{b}.layers1toL.2.weight"] = w2
        sd[f"{pre}{b}.layers1toL.2.bias"] = b2.astype(np.float32)
    return sd


def synthetic_image(H: int, W: int, index: int = 0, noise: float = 2.0) -> np.ndarray:
    """uint8 [3,H,W]: shared luminance field of random 2-D cosines with 1/f amplitudes,
    small chroma fields, Gaussian noise and a few hard edges.  Seeded by index."""
    rng = np.random.default_rng(1337 + index)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)

    def field_(ncomp, amp):
        f = np.zeros((H, W), dtype=np.float32)
        for _ in range(ncomp):
            fx, fy = rng.uniform(0.2, 24.0, size=2)
            ph = rng.uniform(0, 2 * np.pi)
            a = amp / np.sqrt(fx * fx + fy * fy)
            f += np.float32(a) * np.cos(np.float32(2 * np.pi) * (np.float32(fx) * xx / W + np.float32(fy) * yy / H)
                                        + np.float32(ph))
        return f

    lum = 128 + field_(32, 90.0)
    img = np.stack([lum + field_(8, 25.0), lum + field_(8, 15.0), lum + field_(8, 25.0)])
    for _ in range(3):
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        x1, y1 = int(rng.integers(x0, W + 1)), int(rng.integers(y0, H + 1))
        img[:, y0:y1, x0:x1] += rng.uniform(-40, 40, size=(3, 1, 1)).astype(np.float32)
    img += rng.standard_normal(img.shape).astype(np.float32) * np.float32(noise)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_batch_torch(n: int, H: int, W: int, first_index: int, device, noise: float = 2.0):
    """n DISTINCT synthetic images (image k is seeded by first_index + k) as a uint8 [n,3,H,W] tensor on `device`,
    built with torch ops from the same recipe as synthetic_image (random 2-D cosines with 1/f amplitudes, chroma
    fields, three rectangles, Gaussian noise).  Large batches (4K images, 50k-image sweeps) take milliseconds per image
    on a GPU instead of seconds per image on the host; the pixels are not bit-identical to synthetic_image's (other
    cosine / noise generators), which only matters to tests that pin inputs -- those use synthetic_image."""
    import torch
    dev = torch.device(device)
    yy = (torch.arange(H, device=dev, dtype=torch.float32) / H)[:, None]
    xx = (torch.arange(W, device=dev, dtype=torch.float32) / W)[None, :]
    out = torch.empty((n, 3, H, W), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    two_pi = float(2 * np.pi)
    for k in range(n):
        rng = np.random.default_rng(1337 + first_index + k)

        def field_(ncomp, amp):
            f = torch.zeros((H, W), dtype=torch.float32, device=dev)
            for _ in range(ncomp):
                fx, fy = rng.uniform(0.2, 24.0, size=2)
                ph = rng.uniform(0, 2 * np.pi)
                a = amp / np.sqrt(fx * fx + fy * fy)
                # cos(u + v) = cos u cos v - sin u sin v: two 1-D tables and one outer product per component
                u = two_pi * float(fx) * xx + float(ph)
                v = two_pi * float(fy) * yy
                f.add_(float(a) * (torch.cos(v) * torch.cos(u) - torch.sin(v) * torch.sin(u)))
            return f

        lum = 128 + field_(32, 90.0)
        img = torch.stack([lum + field_(8, 25.0), lum + field_(8, 15.0), lum + field_(8, 25.0)])
        for _ in range(3):
            x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
            x1, y1 = int(rng.integers(x0, W + 1)), int(rng.integers(y0, H + 1))
            img[:, y0:y1, x0:x1] += torch.as_tensor(rng.uniform(-40, 40, size=(3, 1, 1)).astype(np.float32), device=dev)
        gen.manual_seed(1337 + first_index + k)
        img += torch.randn(img.shape, generator=gen, device=dev, dtype=torch.float32) * float(noise)
        out[k] = torch.clamp(torch.round(img), 0, 255).to(torch.uint8)
    return out
