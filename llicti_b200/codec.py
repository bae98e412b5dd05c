"""Host-side handle on one libllicti_b200 context (one per process / GPU).

`Codec` owns the C context, turns a reference-format state_dict into the C weight struct,
assembles / parses the reference's `bytestream_list` around the batch entry points, and
exposes the stage-level calls used by the parity tests.  torch is used for device memory and
streams only; every computation happens inside the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import container

PREFIX = "entropymodel.entmdls_scale_band.0."
# branch order of llicti_weights.l0_* (include/llicti.h)
L0_NAMES = [(0, "layer0_00_11"), (1, "layer0_00_01"), (1, "layer0_11_01"),
            (2, "layer0_00_10"), (2, "layer0_11_10"), (2, "layer0_01_10")]
L0_SHAPES = {"layer0_00_11": (4, 4), "layer0_00_01": (3, 4), "layer0_11_01": (4, 3),
             "layer0_00_10": (4, 3), "layer0_11_10": (3, 4), "layer0_01_10": (4, 4)}


@dataclass
class CodecConfig:
    """What the eval_model path reads from configs/llicti_*.json plus this repo's knobs."""
    num_scales: int = 5
    chs: int = 88
    num_mixtures: int = 5
    sub_len: int = 0                 # 0 = torchac-compatible streams; >0 = interleaved substreams
    numerics: int = L.NUM_TORCH_CUDA
    cnn_impl: int = L.CNN_TCGEN05     # CNN_FP32 = CUDA-core exactness reference
    device: int = 0
    decode_impl: int = 0             # 0 = CDF windows + serial chains, 1 = legacy one-warp-per-chain

    @staticmethod
    def from_json_dict(cfg, **over) -> "CodecConfig":
        """Validate the model hyper-parameters exactly as far as the CUDA path supports them
        (the shipped llicti_A / llicti_B settings; everything else is a dead branch in the
        reference)."""
        def need(cond, what):
            if not cond:
                raise ValueError(f"unsupported configuration for the B200 path: {what}")
        need(cfg["clrchs"] == 3 and cfg["clr_joint_mode"] == 2, "clrchs=3, clr_joint_mode=2 required")
        need(bool(cfg["ycocg"]) and not cfg["mwsa_joint"], "ycocg=true, mwsa_joint=false required")
        need(cfg["conv_layers"] == 3 and not cfg["combine_layers1toL"], "conv_layers=3, combine_layers1toL=false")
        need(not cfg["subtract_mean"] and cfg["activfun"] == "ReLU", "subtract_mean=false, activfun=ReLU")
        need(cfg["distribution"] == "normal" and cfg["num_mixtures"] == 5, "5-component normal mixture")
        need(cfg["lif_prec_bits"] == 8 and cfg["ent_mdl_num"] == 4, "lif_prec_bits=8, ent_mdl_num=4")
        lv = list(cfg["dwtlevels"])
        need(lv == list(range(len(lv))) and 1 <= len(lv) <= L.MAX_SCALES, "dwtlevels must be 0..S-1")
        need(all(cfg["useprevlevNN"][1:len(lv)]), "one shared model set over scales (useprevlevNN)")
        need(all(e == 4 for e in cfg["Evens"][:len(lv)]) and all(o == 3 for o in cfg["Odds"][:len(lv)]), "Evens=4, Odds=3")
        need(int(cfg["chs"][0]) in (88, 60), "chs[0] must be 88 or 60")
        return CodecConfig(num_scales=len(lv), chs=int(cfg["chs"][0]), **over)


def _as_f32(v) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(v, dtype=np.float32)


class Codec:
    def __init__(self, cfg: CodecConfig, state_dict: Dict[str, "np.ndarray | torch.Tensor"]):
        self.lib = L.load()
        if not torch.cuda.is_available() or self.lib.llicti_device_count() <= 0:
            raise RuntimeError("llicti_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device("cuda", cfg.device)
        g = cfg.chs
        keep = []            # keep the numpy buffers alive during llicti_create
        w = L.Weights()
        for i, (band, name) in enumerate(L0_NAMES):
            kh, kw = L0_SHAPES[name]
            a = _as_f32(state_dict[f"{PREFIX}{band}.{name}.weight"])
            b = _as_f32(state_dict[f"{PREFIX}{band}.{name}.bias"])
            if a.shape != (4 * g, 3, kh, kw) or b.shape != (4 * g,):
                raise ValueError(f"{name}: unexpected shape {a.shape}")
            keep += [a, b]
            w.l0_w[i], w.l0_b[i] = a.ctypes.data, b.ctypes.data
        for band in range(3):
            a1 = _as_f32(state_dict[f"{PREFIX}{band}.layers1toL.0.weight"]).reshape(4 * g, g)
            b1 = _as_f32(state_dict[f"{PREFIX}{band}.layers1toL.0.bias"])
            a2 = _as_f32(state_dict[f"{PREFIX}{band}.layers1toL.2.weight"]).reshape(12 * cfg.num_mixtures, g)
            b2 = _as_f32(state_dict[f"{PREFIX}{band}.layers1toL.2.bias"])
            keep += [a1, b1, a2, b2]
            w.l1_w[band], w.l1_b[band] = a1.ctypes.data, b1.ctypes.data
            w.l2_w[band], w.l2_b[band] = a2.ctypes.data, b2.ctypes.data
        self._ccfg = L.Config(cfg.num_scales, cfg.chs, cfg.num_mixtures, cfg.sub_len, cfg.numerics, cfg.cnn_impl,
                              cfg.device, cfg.decode_impl)
        self._ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.llicti_create(C.byref(self._ccfg), C.byref(w), C.byref(self._ctx)))
        self._reserved = (0, 0, 0)
        crc = 0
        for a in keep:                     # fixed order: l0 (w, b) x 6 branches, then per band l1 w, b, l2 w, b
            crc = zlib.crc32(a.data, crc)
        self.weights_crc = crc & 0xFFFFFFFF
        # what the CDFs depend on besides the weights: the CNN's arithmetic (0 fp32, 1 tcgen05 / bf16 operands,
        # 2 tcgen05 / fp16 operands) and the numerics profile of the CDF stage
        self.cnn_operands = int(self.lib.llicti_cnn_operands(self._ctx))
        self.fingerprint = container.fingerprint(self.cnn_operands, cfg.numerics, self.weights_crc)
        del keep

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.llicti_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- geometry / workspace -------------------------------------------------------------
    def geometry(self, H: int, W: int) -> L.Geom:
        g = L.Geom()
        L.check(self.lib.llicti_geometry(C.byref(self._ccfg), H, W, C.byref(g)))
        return g

    def reserve(self, n: int, H: int, W: int):
        cur_n, cur_h, cur_w = self._reserved
        if (H, W) != (cur_h, cur_w) or n > cur_n:
            L.check(self.lib.llicti_reserve(self._ctx, n, H, W))
            self._reserved = (n, H, W)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def launches(self) -> int:
        return int(self.lib.llicti_launch_count(self._ctx))

    def check_status(self):
        """Raise if a kernel of an asynchronous *_dev call flagged an error (waits for the stream)."""
        L.check(self.lib.llicti_status(self._ctx, self._stream()))

    def profile(self, enable: bool):
        L.check(self.lib.llicti_profile(self._ctx, int(enable)))

    def profile_read(self):
        """{class: (milliseconds, launch groups)} since profiling was enabled / last read."""
        ms = (C.c_double * len(L.KERNEL_CLASSES))()
        cnt = (C.c_int64 * len(L.KERNEL_CLASSES))()
        L.check(self.lib.llicti_profile_read(self._ctx, ms, cnt))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(L.KERNEL_CLASSES)}

    def selftest_fdiv(self, n_pairs: int, seed: int = 1) -> int:
        """Mismatches between the CDF stage's hoisted division and div.rn.f32 on n_pairs random operand pairs."""
        bad = C.c_uint64(0)
        L.check(self.lib.llicti_selftest_fdiv(self._ctx, n_pairs, seed, C.byref(bad)))
        return int(bad.value)

    def decode_stats(self, reset: bool = True):
        """Decode-side diagnostics: window misses and pipeline waits since the last reset."""
        out = (C.c_uint64 * 8)()
        L.check(self.lib.llicti_decode_stats(self._ctx, out, int(reset)))
        return {"slow_path_symbols": int(out[0]), "consumer_polls": int(out[1]), "chunks_redone": int(out[3]),
                "consumer_wait_cycles": int(out[4]), "consumer_cycles": int(out[5]), "consumer_runs": int(out[6]),
                "consumer_redo_cycles": int(out[7]), "consumer_longest_run_cycles": int(out[2])}

    # -- full path, host buffers (the timed end-to-end call) -------------------------------
    def encode_host(self, rgb: np.ndarray, out: Optional[np.ndarray] = None):
        """rgb uint8 [n,3,H,W] host array (pinned or pageable).  Returns (blob view, stream_off
        uint64 [n*9S+1], minmax int16 [n,6])."""
        assert rgb.dtype == np.uint8 and rgb.ndim == 4 and rgb.shape[1] == 3 and rgb.flags.c_contiguous
        n, _, H, W = rgb.shape
        self.reserve(n, H, W)
        g = self.geometry(H, W)
        ns = 9 * self.cfg.num_scales
        cap = n * int(g.max_stream_bytes)
        if out is None or out.nbytes < cap:
            out = np.empty(cap, dtype=np.uint8)
        off = np.empty(n * ns + 1, dtype=np.uint64)
        mm = np.empty((n, 6), dtype=np.int16)
        L.check(self.lib.llicti_encode_host(self._ctx, rgb.ctypes.data, n, H, W, out.ctypes.data, out.nbytes,
                                            off.ctypes.data, mm.ctypes.data, self._stream()))
        return out[:int(off[-1])], off, mm

    def decode_host(self, blob: np.ndarray, off: np.ndarray, mm: np.ndarray, x00: np.ndarray, n: int, H: int, W: int,
                    out: Optional[np.ndarray] = None) -> np.ndarray:
        self.reserve(n, H, W)
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        mm = np.ascontiguousarray(mm, dtype=np.int16)
        x00 = np.ascontiguousarray(x00, dtype=np.uint8)
        if out is None:
            out = np.empty((n, 3, H, W), dtype=np.uint8)
        L.check(self.lib.llicti_decode_host(self._ctx, blob.ctypes.data, off.ctypes.data, mm.ctypes.data,
                                            x00.ctypes.data, n, H, W, out.ctypes.data, self._stream()))
        return out

    # -- full path, device buffers (asynchronous) --------------------------------------------
    def encode_dev(self, rgb: torch.Tensor, out=None):
        """rgb uint8 [n,3,H,W] CUDA tensor.  Returns CUDA tensors (blob, stream_off, minmax);
        `out` = a previous return value to write into (no allocation on the timed path)."""
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        n, _, H, W = rgb.shape
        self.reserve(n, H, W)
        g = self.geometry(H, W)
        ns = 9 * self.cfg.num_scales
        if out is not None:
            blob, off, mm = out
            assert blob.numel() >= n * int(g.max_stream_bytes) and off.numel() == n * ns + 1 and mm.shape == (n, 6)
        else:
            blob = torch.empty(n * int(g.max_stream_bytes), dtype=torch.uint8, device=self.device)
            off = torch.empty(n * ns + 1, dtype=torch.int64, device=self.device)
            mm = torch.empty((n, 6), dtype=torch.int16, device=self.device)
        L.check(self.lib.llicti_encode_dev(self._ctx, rgb.data_ptr(), n, H, W, blob.data_ptr(), blob.numel(),
                                           off.data_ptr(), mm.data_ptr(), self._stream()))
        return blob, off, mm

    def decode_dev(self, blob: torch.Tensor, off: torch.Tensor, mm: torch.Tensor, x00: torch.Tensor, n: int, H: int,
                   W: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.reserve(n, H, W)
        if out is None:
            out = torch.empty((n, 3, H, W), dtype=torch.uint8, device=self.device)
        L.check(self.lib.llicti_decode_dev(self._ctx, blob.data_ptr(), off.data_ptr(), mm.data_ptr(), x00.data_ptr(),
                                           n, H, W, out.data_ptr(), self._stream()))
        return out

    # -- batches of mixed sizes (llicti_encode_batch_host / llicti_decode_batch_host) ------------------
    def compress_mixed(self, images: Sequence[np.ndarray]):
        """images: uint8 [3,H_i,W_i] arrays of any sizes -> list of bytestream_lists, in order (one C call: the
        library groups the descriptors by size)."""
        n = len(images)
        ns = 9 * self.cfg.num_scales
        items = (L.EncodeItem * n)()
        keep = []
        for i, img in enumerate(images):
            img = np.ascontiguousarray(img, dtype=np.uint8)
            assert img.ndim == 3 and img.shape[0] == 3
            g = self.geometry(img.shape[1], img.shape[2])
            out = np.empty(int(g.max_stream_bytes), dtype=np.uint8)
            off = np.empty(ns + 1, dtype=np.uint64)
            mm = np.empty(6, dtype=np.int16)
            keep.append((img, out, off, mm, g))
            items[i] = L.EncodeItem(img.ctypes.data, img.shape[1], img.shape[2], out.ctypes.data, out.nbytes, off.ctypes.data, mm.ctypes.data)
        L.check(self.lib.llicti_encode_batch_host(self._ctx, items, n, self._stream()))
        self._reserved = (0, 0, 0)          # the library re-reserved per size group
        S = self.cfg.num_scales
        res = []
        for img, out, off, mm, g in keep:
            res.append(container.assemble(S, self.cfg.sub_len, g.Hs[S - 1], g.Ws[S - 1], g.pad_int, img[None], out[:int(off[-1])], off,
                                          mm[None], fp=self.fingerprint, checksum=True)[0])
        return res

    def decompress_mixed(self, bsls: Sequence) -> List[np.ndarray]:
        """list of bytestream_lists of any image sizes -> list of uint8 [3,H_i,W_i] arrays, in order."""
        n = len(bsls)
        items = (L.DecodeItem * n)()
        keep = []
        for i, bsl in enumerate(bsls):
            blob, off, mm, x00, _, H, W = self.from_bytestream_lists([bsl])
            blob, off, mm, x00 = (np.ascontiguousarray(a) for a in (blob, off, mm, x00))
            out = np.empty((3, H, W), dtype=np.uint8)
            keep.append((blob, off, mm, x00, out))
            items[i] = L.DecodeItem(blob.ctypes.data, off.ctypes.data, mm.ctypes.data, x00.ctypes.data, H, W, out.ctypes.data)
        L.check(self.lib.llicti_decode_batch_host(self._ctx, items, n, self._stream()))
        self._reserved = (0, 0, 0)
        for i, bsl in enumerate(bsls):
            want = container.image_checksum(bsl[0])
            if want is not None and (zlib.crc32(keep[i][4].data) & 0xFFFFFFFF) != want:
                raise ValueError(f"image {i}: the decoded pixels fail the stream's checksum")
        return [k[4] for k in keep]

    # -- rate estimation (LLICTI.forward) ----------------------------------------------------------
    def forward_dev(self, rgb: torch.Tensor):
        """rgb uint8 [n,3,H,W] CUDA tensor, H and W multiples of 2^S -> list over scales of float32 [n,9,Hs,Ws]
        self-informations in bits (index 3*band + clr), as LLICTI.forward returns them."""
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        n, _, H, W = rgb.shape
        self.reserve(n, H, W)
        g = self.geometry(H, W)
        S = self.cfg.num_scales
        fpl = [torch.empty((n, 12, g.Hs[s], g.Ws[s]), dtype=torch.float32, device=self.device) for s in range(S)]
        out = [torch.empty((n, 9, g.Hs[s], g.Ws[s]), dtype=torch.float32, device=self.device) for s in range(S)]
        fp = (C.c_void_p * S)(*[t.data_ptr() for t in fpl])
        op = (C.c_void_p * S)(*[t.data_ptr() for t in out])
        L.check(self.lib.llicti_forward_dev(self._ctx, rgb.data_ptr(), n, H, W, fp, op, self._stream()))
        return out

    # -- training step (agents/llicti_agent.py:48-83): weights in from the optimizer, gradients out ------------
    def _weights_struct(self, tensors: Dict[str, torch.Tensor]):
        """llicti_weights holding the DEVICE pointers of fp32 CUDA tensors under the reference's state_dict names."""
        g, w, keep = self.cfg.chs, L.Weights(), []

        def ptr(name, shape):
            t = tensors[name]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == self.device):
                raise ValueError(f"{name}: expected a contiguous float32 tensor on {self.device}")
            if tuple(t.shape) != shape:
                raise ValueError(f"{name}: unexpected shape {tuple(t.shape)}, expected {shape}")
            keep.append(t)
            return t.data_ptr()

        for i, (band, name) in enumerate(L0_NAMES):
            kh, kw = L0_SHAPES[name]
            w.l0_w[i] = ptr(f"{PREFIX}{band}.{name}.weight", (4 * g, 3, kh, kw))
            w.l0_b[i] = ptr(f"{PREFIX}{band}.{name}.bias", (4 * g,))
        for band in range(3):
            w.l1_w[band] = ptr(f"{PREFIX}{band}.layers1toL.0.weight", (4 * g, g, 1, 1))
            w.l1_b[band] = ptr(f"{PREFIX}{band}.layers1toL.0.bias", (4 * g,))
            w.l2_w[band] = ptr(f"{PREFIX}{band}.layers1toL.2.weight", (12 * self.cfg.num_mixtures, g, 1, 1))
            w.l2_b[band] = ptr(f"{PREFIX}{band}.layers1toL.2.bias", (12 * self.cfg.num_mixtures,))
        return w, keep

    def set_weights_dev(self, tensors: Dict[str, torch.Tensor]):
        """Replace the context's weights by these CUDA tensors (the parameters after an optimizer step).  fp32-CNN
        contexts only; the stream fingerprint of this codec no longer describes its weights afterwards."""
        w, keep = self._weights_struct(tensors)
        L.check(self.lib.llicti_set_weights_dev(self._ctx, C.byref(w), self._stream()))
        self.fingerprint = None
        del keep

    def train_forward_dev(self, rgb: torch.Tensor):
        """forward_dev for a training step: returns (self-informations per scale, kept) where `kept` = (fp32 planes,
        network outputs of every band) is what backward_dev needs to skip the colour split and the CNN."""
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        n, _, H, W = rgb.shape
        self.reserve(n, H, W)
        g = self.geometry(H, W)
        S = self.cfg.num_scales
        fpl = [torch.empty((n, 12, g.Hs[s], g.Ws[s]), dtype=torch.float32, device=self.device) for s in range(S)]
        out = [torch.empty((n, 9, g.Hs[s], g.Ws[s]), dtype=torch.float32, device=self.device) for s in range(S)]
        keep = torch.empty((180 * n * int(g.positions),), dtype=torch.float32, device=self.device)
        fp = (C.c_void_p * S)(*[t.data_ptr() for t in fpl])
        op = (C.c_void_p * S)(*[t.data_ptr() for t in out])
        L.check(self.lib.llicti_train_forward_dev(self._ctx, rgb.data_ptr(), n, H, W, fp, op, keep.data_ptr(), self._stream()))
        return out, (fpl, keep)

    def backward_dev(self, rgb: torch.Tensor, gsinfo: Sequence[torch.Tensor], names: Sequence[str], kept=None) -> Dict[str, torch.Tensor]:
        """Gradients of a loss with respect to every weight, given dL / d self-information per scale (float32
        [n,9,Hs,Ws], the shapes forward_dev returns).  `names`: the state_dict keys wanted (all 24 weight tensors);
        `kept`: the second result of train_forward_dev for this batch (consumed), or None to recompute."""
        assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        n, _, H, W = rgb.shape
        self.reserve(n, H, W)
        g = self.geometry(H, W)
        S, G, M = self.cfg.num_scales, self.cfg.chs, self.cfg.num_mixtures
        gs = []
        for s in range(S):
            t = gsinfo[s].to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != (n, 9, g.Hs[s], g.Ws[s]):
                raise ValueError(f"gradient of scale {s}: shape {tuple(t.shape)}, expected {(n, 9, g.Hs[s], g.Ws[s])}")
            gs.append(t)
        if kept is not None:
            fpl, keep_params = kept
            assert keep_params.numel() == 180 * n * int(g.positions) and len(fpl) == S
        else:
            fpl = [torch.empty((n, 12, g.Hs[s], g.Ws[s]), dtype=torch.float32, device=self.device) for s in range(S)]
            keep_params = None
        grads = {}
        for band, name in L0_NAMES:
            kh, kw = L0_SHAPES[name]
            grads[f"{PREFIX}{band}.{name}.weight"] = torch.empty((4 * G, 3, kh, kw), dtype=torch.float32, device=self.device)
            grads[f"{PREFIX}{band}.{name}.bias"] = torch.empty((4 * G,), dtype=torch.float32, device=self.device)
        for band in range(3):
            grads[f"{PREFIX}{band}.layers1toL.0.weight"] = torch.empty((4 * G, G, 1, 1), dtype=torch.float32, device=self.device)
            grads[f"{PREFIX}{band}.layers1toL.0.bias"] = torch.empty((4 * G,), dtype=torch.float32, device=self.device)
            grads[f"{PREFIX}{band}.layers1toL.2.weight"] = torch.empty((12 * M, G, 1, 1), dtype=torch.float32, device=self.device)
            grads[f"{PREFIX}{band}.layers1toL.2.bias"] = torch.empty((12 * M,), dtype=torch.float32, device=self.device)
        w, keep = self._weights_struct(grads)
        fp = (C.c_void_p * S)(*[t.data_ptr() for t in fpl])
        gp = (C.c_void_p * S)(*[t.data_ptr() for t in gs])
        L.check(self.lib.llicti_backward_dev(self._ctx, rgb.data_ptr(), n, H, W, fp, gp,
                                             keep_params.data_ptr() if keep_params is not None else None, C.byref(w), self._stream()))
        del keep
        return {k: grads[k] for k in names}

    # -- bytestream_list assembly (LLICTI_nets.py:346-354, 409-411) ----------------------------
    def to_bytestream_lists(self, rgb: np.ndarray, blob: np.ndarray, off: np.ndarray, mm: np.ndarray):
        S = self.cfg.num_scales
        g = self.geometry(rgb.shape[2], rgb.shape[3])
        return container.assemble(S, self.cfg.sub_len, g.Hs[S - 1], g.Ws[S - 1], g.pad_int, rgb, blob, off, mm,
                                  fp=self.fingerprint, checksum=True)

    def from_bytestream_lists(self, bsls: Sequence):
        """Parse headers; returns (blob, off, minmax, x00, n, H, W)."""
        return container.parse(self.cfg.num_scales, self.cfg.sub_len, bsls, fp=self.fingerprint)

    def compress_images(self, rgb: np.ndarray):
        """uint8 [n,3,H,W] -> list of n bytestream_lists."""
        blob, off, mm = self.encode_host(rgb)
        return self.to_bytestream_lists(rgb, blob, off, mm)

    def decompress_images(self, bsls: Sequence) -> np.ndarray:
        blob, off, mm, x00, n, H, W = self.from_bytestream_lists(bsls)
        rgb = self.decode_host(blob, off, mm, x00, n, H, W)
        for i, bsl in enumerate(bsls):     # streams that carry the checksum of their image are verified
            want = container.image_checksum(bsl[0])
            if want is not None and (zlib.crc32(rgb[i].data) & 0xFFFFFFFF) != want:
                raise ValueError(f"image {i}: the decoded pixels fail the stream's checksum (corrupt stream, or coded with "
                                 "other weights / another CNN implementation than this codec's)")
        return rgb

    # -- stage-level calls (parity tests) --------------------------------------------------------
    def color_split(self, rgb: torch.Tensor):
        n, _, H, W = rgb.shape
        g = self.geometry(H, W)
        S = self.cfg.num_scales
        planes = [torch.empty((n, 12, g.Hs[s], g.Ws[s]), dtype=torch.int16, device=self.device) for s in range(S)]
        mm = torch.empty((n, 4), dtype=torch.int32, device=self.device)
        ptrs = (C.c_void_p * S)(*[p.data_ptr() for p in planes])
        L.check(self.lib.llicti_color_split(self._ctx, rgb.data_ptr(), n, H, W, ptrs, mm.data_ptr(), self._stream()))
        return planes, mm

    def merge_color(self, planes0: torch.Tensor, H: int, W: int) -> torch.Tensor:
        n = planes0.shape[0]
        out = torch.empty((n, 3, H, W), dtype=torch.uint8, device=self.device)
        L.check(self.lib.llicti_merge_color(self._ctx, planes0.data_ptr(), n, H, W, out.data_ptr(), self._stream()))
        return out

    def cnn_params(self, band: int, planes: torch.Tensor) -> torch.Tensor:
        assert planes.dtype == torch.int16 and planes.is_cuda and planes.is_contiguous() and planes.shape[1] == 12
        n, _, Hs, Ws = planes.shape
        out = torch.empty((n, 12 * self.cfg.num_mixtures, Hs, Ws), dtype=torch.float32, device=self.device)
        L.check(self.lib.llicti_cnn_params(self._ctx, band, planes.data_ptr(), n, Hs, Ws, out.data_ptr(), self._stream()))
        return out

    def cdf_table(self, params: torch.Tensor, yband: torch.Tensor, clr: int, min_val: int, max_val: int) -> torch.Tensor:
        """params float [60,P], yband int16 [3,P] -> int16 [P, Lp]."""
        P = params.shape[1]
        Lp = max_val - min_val + 2
        out = torch.empty((P, Lp), dtype=torch.int16, device=self.device)
        L.check(self.lib.llicti_cdf_table(self._ctx, params.data_ptr(), yband.data_ptr(), clr, min_val, max_val, P,
                                          out.data_ptr(), self._stream()))
        return out

    def cdf_bounds(self, params: torch.Tensor, yband: torch.Tensor, clr: int, min_val: int, max_val: int) -> torch.Tensor:
        P = params.shape[1]
        out = torch.empty(P, dtype=torch.int32, device=self.device)
        L.check(self.lib.llicti_cdf_bounds(self._ctx, params.data_ptr(), yband.data_ptr(), clr, min_val, max_val, P,
                                           out.data_ptr(), self._stream()))
        return out

    def ac_encode_bounds(self, bounds: torch.Tensor, S: int = 1) -> List[bytes]:
        """bounds int32/uint32 [n_sym] on device -> list of S substream payloads."""
        n_sym = bounds.numel()
        per = -(-n_sym // S)
        slot = (2 * per + per // 64 + 16 + 15) // 16 * 16
        out = torch.empty(S * slot, dtype=torch.uint8, device=self.device)
        lens = torch.empty(S, dtype=torch.int32, device=self.device)
        L.check(self.lib.llicti_ac_encode_bounds(self._ctx, bounds.data_ptr(), n_sym, S, out.data_ptr(), slot,
                                                 lens.data_ptr(), self._stream()))
        out_h, lens_h = out.cpu().numpy(), lens.cpu().numpy()
        return [out_h[j * slot:j * slot + int(lens_h[j])].tobytes() for j in range(S)]

    def ac_decode_table(self, table: torch.Tensor, parts: List[bytes]) -> torch.Tensor:
        """table int16 [n_sym, Lp] on device, parts = S substream payloads -> int16 [n_sym]."""
        n_sym, Lp = table.shape
        S = len(parts)
        offs = np.zeros(S + 1, dtype=np.uint32)
        offs[1:] = np.cumsum([len(p) for p in parts])
        blob = np.frombuffer(b"".join(parts) + b"\0", dtype=np.uint8)
        d_blob = torch.from_numpy(blob.copy()).to(self.device)
        d_offs = torch.from_numpy(offs.astype(np.int32)).to(self.device)
        sym = torch.empty(n_sym, dtype=torch.int16, device=self.device)
        L.check(self.lib.llicti_ac_decode_table(self._ctx, table.data_ptr(), n_sym, Lp, S, d_blob.data_ptr(),
                                                d_offs.data_ptr(), sym.data_ptr(), self._stream()))
        return sym
