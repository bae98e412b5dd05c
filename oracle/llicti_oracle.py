"""oracle/llicti_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy for the integer/byte work, torch CPU fp32 ops for the
two floating-point stages) of the reference's compress/decompress hot path.
Every function cites the reference lines (relative to /root/reference) it
follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import this module; the product path
(llicti_b200/) never does and fails loudly without its CUDA library.

Pinning (see tests/golden/make_golden.py and DESIGN.md):
  * stages implemented inside /root/reference (colour transform, pyramid, header,
    CNN, GMM CDF tables, symbol mapping, stream order) are pinned bit-exactly
    against the unmodified reference files executed in the build container;
  * the arithmetic coder is third-party (torchac==0.9.3, not vendored, not
    installable): oracle/torchac_port.c restates its algorithm -> byte-level
    parity with the real torchac binary is "parity unpinned".

The oracle deliberately keeps the reference's cost structure (dense CDF tables
of shape H x W x Lp, one sequential coder call per stream) so that timing it is
an honest CPU baseline of the reference's algorithm.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------
# C coder (oracle/torchac_port.c)
# --------------------------------------------------------------------------
_LIB = None


def build_coder(force: bool = False) -> str:
    """Compile oracle/torchac_port.c -> oracle/_build/libtorchac_port.so (gcc)."""
    so = os.path.join(_HERE, "_build", "libtorchac_port.so")
    src = os.path.join(_HERE, "torchac_port.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
    return so


def _coder():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_coder())
        lib.oracle_ac_encode_table.restype = ctypes.c_size_t
        lib.oracle_ac_encode_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_size_t]
        lib.oracle_ac_encode_bounds.restype = ctypes.c_size_t
        lib.oracle_ac_encode_bounds.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.oracle_ac_decode_table.restype = None
        lib.oracle_ac_decode_table.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_size_t, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def ac_encode_table(cdf: np.ndarray, sym: np.ndarray) -> bytes:
    """torchac.encode_int16_normalized_cdf restated (call site LLICTI_nets.py:406-407).
    cdf int16 [n, Lp], sym int16 [n]."""
    cdf = np.ascontiguousarray(cdf, dtype=np.int16)
    sym = np.ascontiguousarray(sym, dtype=np.int16)
    n, Lp = cdf.shape
    assert sym.shape == (n,)
    cap = 2 * n + n // 64 + 16
    out = np.empty(cap, dtype=np.uint8)
    ln = _coder().oracle_ac_encode_table(cdf.ctypes.data, sym.ctypes.data, n, Lp, out.ctypes.data, cap)
    assert ln <= cap
    return out[:ln].tobytes()


def ac_encode_bounds(bounds: np.ndarray) -> bytes:
    """Same coder fed with packed (c_low | (c_high-1) << 16) per symbol."""
    bounds = np.ascontiguousarray(bounds, dtype=np.uint32)
    n = bounds.shape[0]
    cap = 2 * n + n // 64 + 16
    out = np.empty(cap, dtype=np.uint8)
    ln = _coder().oracle_ac_encode_bounds(bounds.ctypes.data, n, out.ctypes.data, cap)
    assert ln <= cap
    return out[:ln].tobytes()


def ac_decode_table(cdf: np.ndarray, stream: bytes) -> np.ndarray:
    """torchac.decode_int16_normalized_cdf restated (call site LLICTI_nets.py:492-493)."""
    cdf = np.ascontiguousarray(cdf, dtype=np.int16)
    n, Lp = cdf.shape
    buf = np.frombuffer(stream, dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    _coder().oracle_ac_decode_table(cdf.ctypes.data, n, Lp, buf.ctypes.data if len(buf) else None, len(buf),
                                    out.ctypes.data)
    return out


def ac_encode_table_py(cdf: np.ndarray, sym: np.ndarray) -> bytes:
    """Pure-Python twin of torchac_port.c's encoder (small cases only); used to
    cross-check the C file."""
    n, Lp = cdf.shape
    cdf = cdf.view(np.uint16)
    low, high, pending = 0, 0xFFFFFFFF, 0
    bits: List[int] = []

    def put(b):
        nonlocal pending
        bits.append(b)
        bits.extend([1 - b] * pending)
        pending = 0

    for i in range(n):
        s = int(sym[i])
        span = high - low + 1
        c_low = int(cdf[i, s])
        c_high = 0x10000 if s == Lp - 2 else int(cdf[i, s + 1])
        high = (low - 1 + ((span * c_high) >> 16)) & 0xFFFFFFFF
        low = (low + ((span * c_low) >> 16)) & 0xFFFFFFFF
        while True:
            if high < 0x80000000:
                put(0)
            elif low >= 0x80000000:
                put(1)
            elif low >= 0x40000000 and high < 0xC0000000:
                pending += 1
                low = (low << 1) & 0x7FFFFFFF
                high = ((high << 1) | 0x80000001) & 0xFFFFFFFF
                continue
            else:
                break
            low = (low << 1) & 0xFFFFFFFF
            high = ((high << 1) | 1) & 0xFFFFFFFF
    pending += 1
    put(0 if low < 0x40000000 else 1)
    while len(bits) % 8:
        bits.append(0)
    return np.packbits(np.array(bits, dtype=np.uint8)).tobytes()


# --------------------------------------------------------------------------
# Configuration / weights
# --------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """The subset of configs/llicti_*.json the eval_model path reads
    (LLICTI_nets.py:19-26, 260-278, 590-603)."""
    dwtlevels: Sequence[int] = (0, 1, 2, 3, 4)
    chs: int = 88            # config.chs[0]; hidden width per sub-network
    num_mixtures: int = 5
    evens: int = 4
    odds: int = 3

    @staticmethod
    def from_dict(cfg) -> "OracleConfig":
        assert cfg["clrchs"] == 3 and cfg["clr_joint_mode"] == 2 and cfg["ycocg"] and not cfg["mwsa_joint"]
        assert cfg["conv_layers"] == 3 and not cfg["combine_layers1toL"] and not cfg["subtract_mean"]
        assert cfg["distribution"] == "normal" and cfg["activfun"] == "ReLU" and cfg["lif_prec_bits"] == 8
        assert all(cfg["useprevlevNN"][1:]), "one shared model set over scales"
        lv = list(cfg["dwtlevels"])
        assert lv == list(range(len(lv)))
        return OracleConfig(dwtlevels=tuple(lv), chs=int(cfg["chs"][0]), num_mixtures=int(cfg["num_mixtures"]),
                            evens=int(cfg["Evens"][0]), odds=int(cfg["Odds"][0]))


# layer-0 branches per band: (state_dict name, source phase index, (padl, padr, padt, padb))
# LLICTI_nets.py:650-675 (with Ev=4, Od=3): kernel shapes follow from the weights themselves.
def band_branches(cfg: OracleConfig):
    Ev, Od = cfg.evens, cfg.odds
    return {
        0: [("layer0_00_11", 0, (Ev // 2 - 1, Ev // 2, Ev // 2 - 1, Ev // 2))],
        1: [("layer0_00_01", 0, (Ev // 2 - 1, Ev // 2, Od // 2, Od // 2)),
            ("layer0_11_01", 1, (Od // 2, Od // 2, Ev // 2, Ev // 2 - 1))],
        2: [("layer0_00_10", 0, (Od // 2, Od // 2, Ev // 2 - 1, Ev // 2)),
            ("layer0_11_10", 1, (Ev // 2, Ev // 2 - 1, Od // 2, Od // 2)),
            ("layer0_01_10", 2, (Ev // 2, Ev // 2 - 1, Ev // 2 - 1, Ev // 2))],
    }


def synthetic_state_dict(cfg: OracleConfig, seed: int = 1337) -> Dict[str, np.ndarray]:
    """Deterministic stand-in weights in the reference's state_dict naming
    (SURVEY.md section 8b).  The shipped checkpoint is absent, so tests and the bench use
    these.  Random weights alone would put every mean near zero and cost ~16 bit per
    symbol, so a hand-wired "interpolation path" is laid over the random weights: six
    hidden units of the mean sub-network carry +/- the average of the nearest known
    neighbours of each colour channel through both ReLU layers, so the predicted means
    interpolate the image, spreads are a few grey levels and symbol costs are realistic."""
    rng = np.random.default_rng(seed)
    g = cfg.chs
    Ch = 4 * g
    M = cfg.num_mixtures
    sd: Dict[str, np.ndarray] = {}
    pre = "entropymodel.entmdls_scale_band.0."
    Ev, Od = cfg.evens, cfg.odds
    shapes = {"layer0_00_11": (Ev, Ev), "layer0_00_01": (Od, Ev), "layer0_11_01": (Ev, Od),
              "layer0_00_10": (Ev, Od), "layer0_11_10": (Od, Ev), "layer0_01_10": (Ev, Ev)}
    # nearest-neighbour taps (dy, dx, weight) inside each layer-0 kernel window
    taps = {"layer0_00_11": [(1, 1, .25), (1, 2, .25), (2, 1, .25), (2, 2, .25)],
            "layer0_00_01": [(1, 1, .25), (1, 2, .25)], "layer0_11_01": [(1, 1, .25), (2, 1, .25)],
            "layer0_00_10": [(1, 1, .25), (2, 1, .25)], "layer0_11_10": [(1, 1, .25), (1, 2, .25)],
            "layer0_01_10": []}
    for b, branches in band_branches(cfg).items():
        for name, _, _ in branches:
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


def jittered_state_dict(cfg: OracleConfig, seed: int = 1337, jitter: float = 2e-3) -> Dict[str, np.ndarray]:
    """synthetic_state_dict with every weight moved by N(0, jitter^2).  The hand-wired interpolation path (taps of exactly
    +-0.25 over integer / 255 samples, zero weights elsewhere) puts pre-activations EXACTLY on the ReLU's kink wherever
    four neighbours cancel; there the gradient depends on rounding noise (the reference's float lifting leaves +-1e-9
    where integers would cancel).  Gradient parity is tested with these generic weights, as trained ones are."""
    rng = np.random.default_rng(seed + 977)
    sd = synthetic_state_dict(cfg, seed)
    return {k: (sd[k] + jitter * rng.standard_normal(sd[k].shape)).astype(np.float32) for k in sorted(sd)}


# --------------------------------------------------------------------------
# a2: integer YCoCg-R  (LLICTI_nets.py:62-74, 77-88, 571-582)
# --------------------------------------------------------------------------
def rgb_to_ycocg_r(rgb: np.ndarray) -> np.ndarray:
    """uint8 [3,H,W] -> int16 [3,H,W]; floor division like torch's `//` on int16."""
    r, g, b = (rgb[i].astype(np.int16) for i in range(3))
    co = r - b
    t = b + np.floor_divide(co, 2)
    cg = g - t
    y = t + np.floor_divide(cg, 2)
    return np.stack([y, co, cg]).astype(np.int16)


def ycocg_r_to_rgb(ycc: np.ndarray) -> np.ndarray:
    """int16 [3,H,W] (Y in 0..255) -> int16 RGB [3,H,W] (LLICTI_nets.py:77-88)."""
    y, co, cg = (ycc[i].astype(np.int16) for i in range(3))
    t = y - np.floor_divide(cg, 2)
    g = cg + t
    b = t - np.floor_divide(co, 2)
    r = b + co
    return np.stack([r, g, b]).astype(np.int16)


# --------------------------------------------------------------------------
# a5: scale pyramid with replicate padding of the short phases (LLICTI_nets.py:218-245)
# --------------------------------------------------------------------------
def pyramid_split(x: np.ndarray, levels: Sequence[int]):
    """x int16 [3,H,W] -> (list over levels of int16 [12,Hs,Ws] in phase order
    x00,x11,x01,x10; pad flags [[padH,padW],...]; padHW_int)."""
    out, flags, pad_int = [], [], 0
    for lev in range(0, max(levels) + 1):
        if lev not in levels:
            continue
        st = 2 ** (lev + 1)
        of = st // 2
        x00 = x[:, 0::st, 0::st]
        x01 = x[:, 0::st, of::st]
        x10 = x[:, of::st, 0::st]
        x11 = x[:, of::st, of::st]
        padH = x00.shape[1] > x11.shape[1]
        padW = x00.shape[2] > x11.shape[2]
        flags.append([bool(padH), bool(padW)])
        pad_int = 4 * pad_int + 2 * int(padH) + int(padW)
        Hs, Ws = x00.shape[1], x00.shape[2]

        def rep(p):
            ph, pw = Hs - p.shape[1], Ws - p.shape[2]
            return np.pad(p, ((0, 0), (0, ph), (0, pw)), mode="edge") if (ph or pw) else p

        out.append(np.concatenate([x00, rep(x11), rep(x01), rep(x10)], axis=0).astype(np.int16))
    return out, flags, pad_int


def pad_flags_from_int(pad_int: int, num_scales: int):
    """LLICTI_nets.py:533-542."""
    flags = []
    for _ in range(num_scales):
        padW = pad_int % 2 == 1
        pad_int //= 2
        padH = pad_int % 2 == 1
        pad_int //= 2
        flags.append([padH, padW])
    flags.reverse()
    return flags


def crop_shape(band: int, Hs: int, Ws: int, padH: bool, padW: bool) -> Tuple[int, int]:
    """Coded region of a band (LLICTI_nets.py:396-397): band 0 (x11) drops the
    replicated row and column, band 1 (x01) the column, band 2 (x10) the row."""
    ch = Hs - int(padH) if band in (0, 2) else Hs
    cw = Ws - int(padW) if band in (0, 1) else Ws
    return ch, cw


# --------------------------------------------------------------------------
# a9: interpolator CNN (LLICTI_nets.py:721-753, 695-712, 822-825)
# --------------------------------------------------------------------------
class OracleNet:
    def __init__(self, cfg: OracleConfig, state_dict: Dict[str, np.ndarray]):
        self.cfg = cfg
        self.sd = {k: torch.as_tensor(np.asarray(v)).float() for k, v in state_dict.items()
                   if k.startswith("entropymodel.entmdls_scale_band.0.") and "conditional_prob_model" not in k}
        self.branches = band_branches(cfg)

    def params(self, band: int, planes_int: np.ndarray) -> np.ndarray:
        """planes_int int16 [12,Hs,Ws] (only phases 0..band are read) -> fp32 [12*M,Hs,Ws].
        The network sees value/255 in fp32 (LLICTI_nets.py:143-144)."""
        y = torch.from_numpy(planes_int[None].astype(np.int16)) / 255  # int16 / 255 -> fp32 true division
        return self.params_float(band, y)[0].numpy()

    def params_float(self, band: int, y: torch.Tensor) -> torch.Tensor:
        """y fp32 [B, >= 3 (band + 1), Hs, Ws] -> fp32 [B, 12 M, Hs, Ws] (get_params, LLICTI_nets.py:822-825)."""
        pre = f"entropymodel.entmdls_scale_band.0.{band}."
        acc = None
        for name, phase, pad in self.branches[band]:
            inp = F.pad(y[:, 3 * phase:3 * phase + 3], pad=pad, mode="replicate")
            o = F.conv2d(inp, self.sd[pre + name + ".weight"], self.sd[pre + name + ".bias"])
            acc = o if acc is None else acc + o
        h = torch.relu(acc)
        h = torch.relu(F.conv2d(h, self.sd[pre + "layers1toL.0.weight"], self.sd[pre + "layers1toL.0.bias"], groups=4))
        return F.conv2d(h, self.sd[pre + "layers1toL.2.weight"], self.sd[pre + "layers1toL.2.bias"], groups=4)


# --------------------------------------------------------------------------
# a10-a13: coupling, GMM CDF on the sampling grid, integer table
# --------------------------------------------------------------------------
SCALE_BOUND = 0.11 / 255.0       # entropy_layer_nets.py:149
WEIGHT_BOUND = 1e-6              # entropy_layer_nets.py:158


def sampling_points(min_val: int, max_val: int) -> torch.Tensor:
    """LLICTI_nets.py:941-942."""
    pts = torch.linspace(min_val - 0.5, max_val + 0.5, steps=max_val - min_val + 1 + 1) / 255
    pts[0], pts[-1] = (min_val - 0.5 - 20) / 255, (max_val + 0.5 + 20) / 255
    return pts


def gmm_cdf_table(sigma: np.ndarray, mu: np.ndarray, w: np.ndarray, min_val: int, max_val: int,
                  sum_order: str = "torch") -> np.ndarray:
    """sigma, mu, w fp32 [M,H,W] -> int16 [H,W,Lp] (entropy_layer_nets.py:185-204 then
    LLICTI_nets.py:955-983).  sum_order="torch" lets torch.sum pick the order (what the
    reference does on this host); "seq" forces ((((t0+t1)+t2)+t3)+t4); "ilp4" forces
    (((t0+t4)+t1)+t2)+t3 (the order of ATen's 4-accumulator reduction)."""
    s = torch.from_numpy(np.ascontiguousarray(sigma))[None]
    m = torch.from_numpy(np.ascontiguousarray(mu))[None]
    ww = torch.from_numpy(np.ascontiguousarray(w))[None]
    pts = sampling_points(min_val, max_val)
    B, X, H, W = m.shape
    P = pts.shape[0]

    def rsum(t, dim):
        if sum_order == "torch":
            return torch.sum(t, dim=dim, keepdim=True)
        parts = torch.unbind(t, dim=dim)
        if sum_order == "seq":
            acc = parts[0]
            for q in parts[1:]:
                acc = acc + q
        elif sum_order == "ilp4":
            assert len(parts) == 5
            acc = (((parts[0] + parts[4]) + parts[1]) + parts[2]) + parts[3]
        else:
            raise ValueError(sum_order)
        return acc.unsqueeze(dim)

    s = torch.max(s, torch.tensor([SCALE_BOUND]))
    ww = torch.max(ww.permute(0, 2, 3, 1).reshape(B, H, W, 1, X), torch.tensor([WEIGHT_BOUND]))
    ww = ww / (1e-9 + rsum(ww, 4))
    z = (pts - m.unsqueeze(4)) / s.unsqueeze(4)
    cm = 0.5 * torch.erfc(float(-(2 ** -0.5)) * z)                       # B X H W P
    cdf = rsum(ww.unsqueeze(5) * cm.permute(0, 2, 3, 1, 4).reshape(B, H, W, 1, X, P), 4)
    cdf = cdf.reshape(B, H, W, P)[0]
    # integer table: round(cdf * (2^16 - (Lp-1))) -> int16 (wraps) + arange(Lp)
    scale = torch.tensor(2, dtype=torch.float32).pow_(16) - (P - 1)
    q = cdf.mul(scale).round().to(torch.int16)
    q.add_(torch.arange(P, dtype=torch.int16))
    return q.numpy()


def couple_means(params: np.ndarray, clr: int, y_band: np.ndarray, M: int) -> np.ndarray:
    """LLICTI_nets.py:381-392: mu_Co += a*y_Y ; mu_Cg += b*y_Y + d*y_Co with y = int/255 fp32.
    y_band: int16 [3,Hs,Ws] centred integer values of the band being coded."""
    mu = torch.from_numpy(params[(3 + clr) * M:(3 + clr + 1) * M].copy())
    yf = torch.from_numpy(y_band.astype(np.int16)) / 255
    a = torch.from_numpy(params[9 * M:10 * M])
    b = torch.from_numpy(params[10 * M:11 * M])
    d = torch.from_numpy(params[11 * M:12 * M])
    if clr == 1:
        mu += a * yf[0:1]
    elif clr == 2:
        mu += b * yf[0:1] + d * yf[1:2]
    return mu.numpy()


# --------------------------------------------------------------------------
# Interleaved-substream container (new in this repo; not in the reference)
# --------------------------------------------------------------------------
def scale_sub_len(sub_len: int, scl: int) -> int:
    """Target symbols per substream at scale `scl`: sub_len at the finest scale, half of it at the coarser ones.  (A
    substream is one serial decoder chain: the coarser scales have a quarter and less of the symbols, and with chains
    of the same length they would have too few chains to fill a GPU.  Container version 2.)"""
    return sub_len if scl == 0 else max(sub_len // 2, 1)


def num_substreams(n: int, sub_len: int) -> int:
    """Number of interleaved substreams of a stream with n symbols: about one per
    sub_len symbols, a multiple of 32 once there are more than 32 (one warp each)."""
    s = max(1, -(-n // sub_len))
    if s > 32:
        s = -(-s // 32) * 32
    return min(s, 65535)


def pack_substreams(parts: List[bytes]) -> bytes:
    """[u16 S][u16 len_0..len_{S-1}][payload_0 | payload_1 | ...], little endian."""
    S = len(parts)
    head = np.array([S] + [len(p) for p in parts], dtype="<u2")
    assert all(len(p) < 65536 for p in parts)
    return head.tobytes() + b"".join(parts)


def unpack_substreams(blob: bytes) -> List[bytes]:
    S = int(np.frombuffer(blob[:2], dtype="<u2")[0])
    lens = np.frombuffer(blob[2:2 + 2 * S], dtype="<u2").astype(np.int64)
    off = 2 + 2 * S
    parts = []
    for ln in lens:
        parts.append(blob[off:off + int(ln)])
        off += int(ln)
    assert off == len(blob)
    return parts


# --------------------------------------------------------------------------
# a1, a7, a15: compress / decompress drivers
# --------------------------------------------------------------------------
@dataclass
class StageDump:
    """Intermediate results kept for stage-boundary parity tests."""
    ycocg: np.ndarray = None
    minmax: List[int] = None
    planes: List[np.ndarray] = None
    pad_flags: List[List[bool]] = None
    pad_int: int = 0
    params: Dict[Tuple[int, int], np.ndarray] = field(default_factory=dict)
    tables: Dict[Tuple[int, int, int], np.ndarray] = field(default_factory=dict)
    symbols: Dict[Tuple[int, int, int], np.ndarray] = field(default_factory=dict)


class OracleCodec:
    def __init__(self, cfg: OracleConfig, state_dict, sub_len: int = 0, sum_order: str = "torch"):
        """sub_len == 0: torchac-compatible single stream per (scale, band, channel) -- the
        reference's bitstream.  sub_len > 0: interleaved-substream container."""
        self.cfg = cfg
        self.net = OracleNet(cfg, state_dict)
        self.sub_len = int(sub_len)
        self.sum_order = sum_order
        self.S = len(cfg.dwtlevels)
        self.M = cfg.num_mixtures

    # -- helpers -------------------------------------------------------------
    def _ranges(self, minmax, clr):
        lo = -127 if clr == 0 else int(minmax[clr])
        hi = 128 if clr == 0 else int(minmax[3 + clr])
        return lo, hi

    def _shift(self, minmax, clr):
        return 127 if clr == 0 else -int(minmax[clr])       # LLICTI_nets.py:544-547

    def _table(self, params, clr, y_band, minmax):
        M = self.M
        sigma = params[clr * M:(clr + 1) * M]
        w = params[(6 + clr) * M:(7 + clr) * M]
        mu = couple_means(params, clr, y_band, M)
        lo, hi = self._ranges(minmax, clr)
        return gmm_cdf_table(sigma, mu, w, lo, hi, self.sum_order)

    def _encode_stream(self, table: np.ndarray, sym: np.ndarray, scl: int) -> bytes:
        n = sym.shape[0]
        if self.sub_len <= 0:
            return ac_encode_table(table, sym)
        S = num_substreams(n, scale_sub_len(self.sub_len, scl))
        return pack_substreams([ac_encode_table(table[j::S], sym[j::S]) for j in range(S)])

    def _decode_stream(self, table: np.ndarray, blob: bytes, scl: int) -> np.ndarray:
        n = table.shape[0]
        if self.sub_len <= 0:
            return ac_decode_table(table, blob)
        parts = unpack_substreams(blob)
        S = len(parts)
        assert S == num_substreams(n, scale_sub_len(self.sub_len, scl))
        out = np.empty(n, dtype=np.int16)
        for j in range(S):
            out[j::S] = ac_decode_table(table[j::S], parts[j])
        return out

    # -- compress (LLICTI_nets.py:125-159, 344-413) ------------------------------
    def compress(self, rgb: np.ndarray, dump: StageDump = None):
        """rgb uint8 [3,H,W] -> bytestream_list (list[1+S] of list[9] of bytes)."""
        assert rgb.dtype == np.uint8 and rgb.ndim == 3 and rgb.shape[0] == 3
        S, M = self.S, self.M
        st_last = 2 ** (max(self.cfg.dwtlevels) + 1)
        x00_last_rgb = np.ascontiguousarray(rgb[:, 0::st_last, 0::st_last])           # :132, :248-252
        ycc = rgb_to_ycocg_r(rgb)
        minmax = [0, int(ycc[1].min()), int(ycc[2].min()), 255, int(ycc[1].max()), int(ycc[2].max())]  # :137-139
        cen = ycc.copy()
        cen[0] -= 127                                                                 # :143
        planes, flags, pad_int = pyramid_split(cen, self.cfg.dwtlevels)
        h_last, w_last = planes[-1].shape[1:]
        header = [np.array([S, h_last, w_last], dtype=np.uint8).tobytes(),              # :347
                  np.array(minmax, dtype=np.int16).tobytes(),                         # :348
                  np.array([pad_int], dtype=np.int16).tobytes(),                      # :349
                  x00_last_rgb.tobytes(),                                             # :350
                  self._mode_tag(), b"", b"", b"", b""]
        out = [header]
        if dump is not None:
            dump.ycocg, dump.minmax, dump.planes, dump.pad_flags, dump.pad_int = ycc, minmax, planes, flags, pad_int
        for scl in range(S - 1, -1, -1):                                              # :364
            pl = planes[scl]
            Hs, Ws = pl.shape[1:]
            padH, padW = flags[scl]
            row = []
            for b in range(3):
                params = self.net.params(b, pl)
                y_band = pl[3 * (b + 1):3 * (b + 2)]
                ch, cw = crop_shape(b, Hs, Ws, padH, padW)
                if dump is not None:
                    dump.params[(scl, b)] = params.copy()
                for clr in range(3):
                    table = self._table(params, clr, y_band, minmax)
                    sym = (y_band[clr].astype(np.int32) + self._shift(minmax, clr)).astype(np.int16)   # :549-557
                    t = np.ascontiguousarray(table[:ch, :cw]).reshape(ch * cw, -1)
                    s = np.ascontiguousarray(sym[:ch, :cw]).reshape(-1)
                    if dump is not None:
                        dump.tables[(scl, b, clr)] = t
                        dump.symbols[(scl, b, clr)] = s
                    row.append(self._encode_stream(t, s, scl))
            out.append(row)
        return out

    def _mode_tag(self) -> bytes:
        """Header slot 4 is b'' in the reference (LLICTI_nets.py:351-354); this repo uses it
        to flag the substream container: b'' = torchac-compatible, else [2, sub_len as u32 LE] (2 = container
        version: substreams of sub_len symbols at scale 0, sub_len / 2 at the coarser scales)."""
        if self.sub_len <= 0:
            return b""
        return bytes([2]) + int(self.sub_len).to_bytes(4, "little")

    # -- decompress (LLICTI_nets.py:161-179, 415-509) -----------------------------
    def decompress(self, bsl, dump: StageDump = None) -> np.ndarray:
        """bytestream_list -> uint8 RGB [3,H,W].  `dump` receives the decoder's own network outputs, tables and
        symbols (diagnose_round_trip compares them with the encoder's)."""
        S, M = self.S, self.M
        hdr = bsl[0]
        ns, h_last, w_last = (int(v) for v in np.frombuffer(hdr[0], dtype=np.uint8))
        assert ns == S                                                                # :424
        minmax = [int(v) for v in np.frombuffer(hdr[1], dtype=np.int16)]
        pad_int = int(np.frombuffer(hdr[2], dtype=np.int16)[0])
        flags = pad_flags_from_int(pad_int, S)
        tag = hdr[4]
        sub_len = int.from_bytes(tag[1:5], "little") if len(tag) else 0
        assert sub_len == self.sub_len
        x00 = rgb_to_ycocg_r(np.frombuffer(hdr[3], dtype=np.uint8).reshape(3, h_last, w_last))  # :429-430
        x00[0] -= 127                                                                 # :444
        cur = None
        for scl in range(S - 1, -1, -1):
            if scl == S - 1:
                Hs, Ws = h_last, w_last
                pl = np.zeros((12, Hs, Ws), dtype=np.int16)
                pl[0:3] = x00
            else:                                                                     # :446-454
                pH, pW = flags[scl + 1]
                full = _interleave(cur)[:, :cur.shape[1] * 2 - int(pH), :cur.shape[2] * 2 - int(pW)]
                Hs, Ws = full.shape[1:]
                pl = np.zeros((12, Hs, Ws), dtype=np.int16)
                pl[0:3] = full
            padH, padW = flags[scl]
            row = bsl[len(bsl) - 1 - scl]
            for b in range(3):
                params = self.net.params(b, pl)
                ch, cw = crop_shape(b, Hs, Ws, padH, padW)
                if dump is not None:
                    dump.params[(scl, b)] = params.copy()
                for clr in range(3):
                    y_band = pl[3 * (b + 1):3 * (b + 2)]
                    table = self._table(params, clr, y_band, minmax)
                    t = np.ascontiguousarray(table[:ch, :cw]).reshape(ch * cw, -1)
                    sym = self._decode_stream(t, row[3 * b + clr], scl).reshape(ch, cw)
                    if dump is not None:
                        dump.tables[(scl, b, clr)] = t
                        dump.symbols[(scl, b, clr)] = sym.reshape(-1).copy()
                    val = (sym.astype(np.int32) - self._shift(minmax, clr)).astype(np.int16)
                    val = np.pad(val, ((0, Hs - ch), (0, Ws - cw)), mode="edge")       # :512-530
                    pl[3 * (b + 1) + clr] = val
            cur = pl
        pH, pW = flags[0]
        full = _interleave(cur)[:, :cur.shape[1] * 2 - int(pH), :cur.shape[2] * 2 - int(pW)]  # :501-509
        full = full.copy()
        full[0] += 127                                                                # :174
        rgb = ycocg_r_to_rgb(full)
        return rgb.astype(np.uint8)


# --------------------------------------------------------------------------
# Rate estimation: LLICTI.forward (LLICTI_nets.py:101-123) -> LLICTIEntropyLayer.forward (:318-342)
# -> LLICTIEntropyModel4.forward / get_self_infos (:802-811, :827-935) -> GaussianConditionalLosslessGMM.forward
# (entropy_layer_nets.py:160-183) with _likelihood_fk (:121-139).  The training / validation path: no coding,
# -log2 of the probability mass of every sample.
# --------------------------------------------------------------------------
LIKELIHOOD_BOUND = 1e-9          # compressai GaussianConditional(likelihood_bound=1e-9), applied at entropy_layer_nets.py:181-182


def float_ycocg_r(x: torch.Tensor) -> torch.Tensor:
    """LLICTI_nets.py:40-49 with RNDFACTOR = 255 (lif_prec_bits = 8): the lifting steps in fp32 on values k/255,
    rounding half to even.  NOT the integer transform compress() uses (:62-74 floors): training sees this one."""
    R, G, B = x[:, 0:1], x[:, 1:2], x[:, 2:3]
    Co = R - B
    t = B + torch.round(Co * 255 / 2) / 255
    Cg = G - t
    Y = t + torch.round(Cg * 255 / 2) / 255
    return torch.cat((Y, Co, Cg), dim=1)


class _LowerBoundFn(torch.autograd.Function):
    """compressai.ops.LowerBound as the reference's entropy layer uses it (entropy_layer_nets.py:9, 135, 158, 176):
    max(x, bound) whose gradient passes where x >= bound or where the step would raise x."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (grad_output < 0)) * grad_output, None


def _lower_bound(x: torch.Tensor, bound: float) -> torch.Tensor:
    b = torch.tensor([bound])
    if torch.is_grad_enabled() and x.requires_grad:
        return _LowerBoundFn.apply(x, b)
    return torch.max(x, b)


def self_informations_torch(cfg: OracleConfig, net: "OracleNet", x: torch.Tensor) -> List[torch.Tensor]:
    """LLICTI.forward (LLICTI_nets.py:101-123, 318-342, 827-935) on x float32 [B,3,H,W] = uint8 / 255, H and W multiples of
    2^S: per scale fp32 [B,9,Hs,Ws], -log2 p of band b, channel clr at index 3 b + clr.  Differentiable with respect to the
    tensors of `net.sd` when they require grad (the training step differentiates exactly this graph)."""
    S, M = len(cfg.dwtlevels), cfg.num_mixtures
    assert x.shape[2] % 2 ** S == 0 and x.shape[3] % 2 ** S == 0
    x = float_ycocg_r(x)
    x[:, 0] = x[:, 0] - (2 ** 7 - 1) / (2 ** 8 - 1)                                  # :110, mean_y_ycocg :26
    out = []
    half = float(0.5 / 255.0)
    for lev in cfg.dwtlevels:                                                        # lazyDWT(pad=False), :218-241
        st, of = 2 ** (lev + 1), 2 ** lev
        y = torch.cat((x[:, :, 0::st, 0::st], x[:, :, of::st, of::st], x[:, :, 0::st, of::st], x[:, :, of::st, 0::st]), dim=1)
        B_, _, H, W = y.shape
        infos = []
        for b in range(3):
            params = net.params_float(b, y[:, 0:3 * (b + 1)])
            yt = y[:, 3 * (b + 1):3 * (b + 2)]
            stdev, mean, wts = params[:, 0:3 * M], params[:, 3 * M:6 * M].clone(), params[:, 6 * M:9 * M]
            a, bb, d = params[:, 9 * M:10 * M], params[:, 10 * M:11 * M], params[:, 11 * M:12 * M]
            mean[:, M:2 * M] = mean[:, M:2 * M] + a * yt[:, 0:1]                     # :858-860
            mean[:, 2 * M:3 * M] = mean[:, 2 * M:3 * M] + bb * yt[:, 0:1] + d * yt[:, 1:2]
            # GaussianConditionalLosslessGMM.forward: channels-last, inputs repeated per mixture
            inp = yt.permute(0, 2, 3, 1).repeat_interleave(M, dim=3)
            sc = _lower_bound(stdev.permute(0, 2, 3, 1), SCALE_BOUND)
            values = torch.abs(inp - mean.permute(0, 2, 3, 1))
            const = float(-(2 ** -0.5))
            upper = 0.5 * torch.erfc(const * ((half - values) / sc))
            lower = 0.5 * torch.erfc(const * ((-half - values) / sc))
            lik_m = (upper - lower).view(B_, H, W, 3, M)
            w = _lower_bound(wts.permute(0, 2, 3, 1).view(B_, H, W, 3, M), WEIGHT_BOUND)
            w = w / torch.sum(w, dim=4, keepdim=True)
            lik = torch.sum(w * lik_m, dim=4).permute(0, 3, 1, 2)
            lik = _lower_bound(lik, LIKELIHOOD_BOUND)
            infos.append(-torch.log2(lik))
        out.append(torch.cat(infos, dim=1))
    return out


def forward_self_informations(cfg: OracleConfig, net: "OracleNet", rgb: np.ndarray) -> List[np.ndarray]:
    """rgb uint8 [3,H,W], H and W multiples of 2^S (the un-padded lazyDWT needs equal phase sizes; the reference
    trains on such patches and pads validation images, agents/llicti_agent.py:105-116) -> per scale fp32
    [9,Hs,Ws]: -log2 p of band b, channel clr at index 3 b + clr."""
    x = torch.from_numpy(rgb.astype(np.float32) / np.float32(255.0))[None]          # ToTensor()
    with torch.no_grad():
        return [t[0].numpy() for t in self_informations_torch(cfg, net, x)]


def train_loss_and_grads(cfg: OracleConfig, state_dict, rgb: np.ndarray, grad_acc_iters: int = 1):
    """One backward pass of the reference's training step (agents/llicti_agent.py:52-61) on a batch rgb uint8 [B,3,H,W]:
    self-informations -> TrainRLossList (graphs/losses/rate_dist.py:97-104: sum over scales of sum(sinfo) / numel(x) * 3) ->
    (loss / grad_acc_iters).backward().  Returns (loss, {state_dict key: gradient as fp32 array})."""
    net = OracleNet(cfg, state_dict)
    for t in net.sd.values():
        t.requires_grad_(True)
    x = torch.from_numpy(rgb.astype(np.float32) / np.float32(255.0))
    sinfos = self_informations_torch(cfg, net, x)
    loss = 0
    for t in sinfos:
        loss = loss + torch.sum(torch.sum(t, dim=(0, 2, 3)) / x.numel() * 3)
    (loss / grad_acc_iters).backward()
    return float(loss.item()), {k: t.grad.numpy().copy() for k, t in net.sd.items()}


def diagnose_round_trip(codec: "OracleCodec", rgb: np.ndarray) -> str:
    """Compress and decompress `rgb` once more with stage dumps on both sides and say where the decoder first
    departs from the encoder, in coding order: the network outputs of a (scale, band), the integer table of a
    (scale, band, channel), or the decoded symbols.  The two sides run the same functions on the same inputs, so a
    difference in `params` or `tables` means a floating-point library call (conv2d, erfc, sum) did not reproduce
    itself inside one process."""
    enc, dec = StageDump(), StageDump()
    bsl = codec.compress(rgb, enc)
    rec = codec.decompress(bsl, dec)
    S = codec.S
    for scl in range(S - 1, -1, -1):
        for b in range(3):
            pe, pd = enc.params[(scl, b)], dec.params[(scl, b)]
            if not np.array_equal(pe, pd):
                bad = np.argwhere(pe != pd)
                return (f"network outputs differ at scale {scl} band {b}: {len(bad)} of {pe.size} values, first at "
                        f"(channel, row, col) = {tuple(int(v) for v in bad[0])} (max abs difference {float(np.abs(pe - pd).max()):.3e}); "
                        f"torch threads {torch.get_num_threads()}")
            for clr in range(3):
                te, td = enc.tables[(scl, b, clr)], dec.tables[(scl, b, clr)]
                if not np.array_equal(te, td):
                    bad = np.argwhere(te != td)
                    return (f"integer CDF tables differ at scale {scl} band {b} channel {clr} although the network outputs "
                            f"agree: {len(bad)} of {te.size} entries, first at (position, entry) = {tuple(int(v) for v in bad[0])}; "
                            f"torch threads {torch.get_num_threads()}")
                if not np.array_equal(enc.symbols[(scl, b, clr)], dec.symbols[(scl, b, clr)]):
                    return f"tables agree but the coder returned other symbols at scale {scl} band {b} channel {clr}"
    return "this repetition round-tripped" if np.array_equal(rec, rgb) else "stages agree but the pixels differ (inverse transform)"


def _interleave(pl: np.ndarray) -> np.ndarray:
    """Inverse lazy DWT (LLICTI_nets.py:446-452): phases x00,x11,x01,x10 -> [3,2Hs,2Ws]."""
    _, Hs, Ws = pl.shape
    full = np.zeros((3, 2 * Hs, 2 * Ws), dtype=np.int16)
    full[:, 0::2, 0::2] = pl[0:3]
    full[:, 0::2, 1::2] = pl[6:9]
    full[:, 1::2, 0::2] = pl[9:12]
    full[:, 1::2, 1::2] = pl[3:6]
    return full


# --------------------------------------------------------------------------
# Synthetic photographic-like test images
# --------------------------------------------------------------------------
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
