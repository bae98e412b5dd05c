"""Host-side mirror of the reference's model interface, `graphs.models.LLICTI_nets.LLICTI`
(reference graphs/models/LLICTI_nets.py:91-179), for the compress / decompress path.

Same class name, constructor argument, method names (including the reference's spelling
`decompres`), return shapes and state_dict key names, so `LLICTIAgent.eval_model`
(agents/llicti_agent.py:123-164) runs against it unchanged -- but every computation is done by
libllicti_b200.so on the GPU.  The nn.Module tree below only holds the parameters (so that the
reference's checkpoints load by name); it has no PyTorch forward path and no CPU fallback.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch
from torch import nn

from . import _lib as L
from .codec import Codec, CodecConfig, L0_SHAPES


def _cfg_get(config, key, default=None):
    if isinstance(config, dict):
        return config.get(key, default)
    return getattr(config, key, default)


def _cfg_dict(config) -> dict:
    return dict(config) if isinstance(config, dict) else {k: getattr(config, k) for k in vars(config)}


class _ProbModelBuffers(nn.Module):
    """Buffer names compressai's GaussianConditional registers under
    `...conditional_prob_model.` in the reference's checkpoints; kept so
    such checkpoints load without unexpected-key errors.  The CUDA path uses the two bounds as
    compile-time constants (0.11/255 and 1e-6, entropy_layer_nets.py:149,158)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("scale_bound", torch.tensor([0.11 / 255.0]))


class LLICTIEntropyModel4(nn.Module):
    """Parameter holder + stage-level entry points of one band's interpolator
    (reference LLICTI_nets.py:585-952)."""

    BRANCHES = {0: ["layer0_00_11"], 1: ["layer0_00_01", "layer0_11_01"],
                2: ["layer0_00_10", "layer0_11_10", "layer0_01_10"]}

    def __init__(self, band: int, chs: int, num_mixtures: int, owner):
        super().__init__()
        self.band = band
        self.num_mixtures = num_mixtures
        Ch = 4 * chs
        for name in self.BRANCHES[band]:
            setattr(self, name, nn.Conv2d(3, Ch, kernel_size=L0_SHAPES[name]))
        self.layers1toL = nn.Sequential(nn.Conv2d(Ch, Ch, 1, groups=4), nn.ReLU(inplace=True),
                                        nn.Conv2d(Ch, 12 * num_mixtures, 1, groups=4))
        self.conditional_prob_model = _ProbModelBuffers()
        self._owner = [owner]   # list: keep the parent out of the module tree

    def get_params(self, y_condition: torch.Tensor) -> torch.Tensor:
        """y_condition float [B, 3*(band+1), Hs, Ws] holding integer/255 values ->
        float [B, 60, Hs, Ws]  (reference :822-825)."""
        codec = self._owner[0]._codec()
        B, C_, Hs, Ws = y_condition.shape
        assert C_ == 3 * (self.band + 1)
        planes = torch.zeros((B, 12, Hs, Ws), dtype=torch.int16, device=codec.device)
        planes[:, :C_] = torch.round(y_condition.to(codec.device) * 255).to(torch.int16)
        return codec.cnn_params(self.band, planes)

    def get_cdfs(self, stdevs, means, weights, clrch=0, int_cdf=False, minVal=0, maxVal=255):
        """Integer CDF table int16 [B,1,H,W,Lp] from (already coupled) GMM parameters
        (reference :938-952 with int_cdf=True).  The float table (int_cdf=False) is never
        materialised by this implementation."""
        if not int_cdf:
            raise NotImplementedError("only the integer CDF (int_cdf=True) exists on the B200 path")
        codec = self._owner[0]._codec()
        B, M, H, W = means.shape
        P = B * H * W
        params = torch.zeros((12 * M, P), dtype=torch.float32, device=codec.device)

        def flat(t):
            return t.to(codec.device).permute(1, 0, 2, 3).reshape(M, P)

        params[0:M], params[3 * M:4 * M], params[6 * M:7 * M] = flat(stdevs), flat(means), flat(weights)
        yband = torch.zeros((3, P), dtype=torch.int16, device=codec.device)
        table = codec.cdf_table(params, yband, 0, int(minVal), int(maxVal))
        return table.reshape(B, 1, H, W, -1)


class LLICTIEntropyLayer(nn.Module):
    def __init__(self, chs: int, num_mixtures: int, owner):
        super().__init__()
        bands = nn.ModuleList([LLICTIEntropyModel4(b, chs, num_mixtures, owner) for b in range(3)])
        self.entmdls_scale_band = nn.ModuleList([bands])


class _SelfInformations(torch.autograd.Function):
    """forward() with an autograd edge: the reference's training step calls `self.model(x)`, sums the result into a
    loss and calls `.backward()` (agents/llicti_agent.py:56-61).  Both directions run in libllicti_b200
    (`llicti_train_forward_dev`, `llicti_backward_dev`); the fp32 planes and the network outputs (240 B per position and
    band) are kept between them, no hidden activation is."""

    @staticmethod
    def forward(ctx, model, rgb, names, *params):
        codec = model._train_codec()
        codec.set_weights_dev(dict(zip(names, [p.detach() for p in params])))
        outs, kept = codec.train_forward_dev(rgb)
        ctx.model, ctx.rgb, ctx.names, ctx.kept = model, rgb, names, kept
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gsinfo):
        codec = ctx.model._train_codec()
        shapes = None
        if any(g is None for g in gsinfo):                           # a scale the loss did not touch
            n, _, H, W = ctx.rgb.shape
            geo = codec.geometry(H, W)
            shapes = [(n, 9, geo.Hs[s], geo.Ws[s]) for s in range(len(gsinfo))]
        gs = [g if g is not None else torch.zeros(shapes[s], dtype=torch.float32, device=codec.device) for s, g in enumerate(gsinfo)]
        kept, ctx.kept = ctx.kept, None                                # consumed: a second backward recomputes
        grads = codec.backward_dev(ctx.rgb, gs, ctx.names, kept=kept)
        return (None, None, None) + tuple(grads[k] for k in ctx.names)


class LLICTI(nn.Module):
    """Drop-in for the reference's LLICTI on the eval_model path."""

    def __init__(self, config, sub_len: int = None, numerics: int = None, cnn_impl: int = None):
        super().__init__()
        cfgd = _cfg_dict(config)
        over = {}
        over["sub_len"] = int(cfgd.get("b200_sub_len", 0) if sub_len is None else sub_len)
        over["numerics"] = int(cfgd.get("b200_numerics", L.NUM_TORCH_CUDA) if numerics is None else numerics)
        over["cnn_impl"] = int(cfgd.get("b200_cnn_impl", L.CNN_TCGEN05) if cnn_impl is None else cnn_impl)
        self.codec_config = CodecConfig.from_json_dict(cfgd, **over)
        self.list_scales = list(cfgd["dwtlevels"])
        self.num_scales = len(self.list_scales)
        self.ycocg = True
        self.entropymodel = LLICTIEntropyLayer(self.codec_config.chs, self.codec_config.num_mixtures, self)
        self.__dict__["_codec_obj"] = None
        self.__dict__["_train_codec_obj"] = None

    # -- parameter plumbing -------------------------------------------------------------------
    def _invalidate(self):
        c = self.__dict__.get("_codec_obj")
        if c is not None:
            c.close()
        self.__dict__["_codec_obj"] = None
        t = self.__dict__.get("_train_codec_obj")
        if t is not None:
            t.close()
        self.__dict__["_train_codec_obj"] = None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in state_dict.items()
              if not ("conditional_prob_model" in k and k not in own)}    # compressai's extra buffers
        for k in own:
            if "conditional_prob_model" in k and k not in sd:
                sd[k] = self.state_dict()[k]
        res = super().load_state_dict(sd, strict=strict, **kw)
        self._invalidate()
        return res

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._invalidate()
        return r

    def _codec(self) -> Codec:
        c = self.__dict__.get("_codec_obj")
        if c is None:
            if not torch.cuda.is_available():
                raise RuntimeError("llicti_b200.LLICTI needs a CUDA device; there is no CPU fallback")
            p = next(self.parameters())
            dev = p.device.index if p.is_cuda else torch.cuda.current_device()
            self.codec_config.device = int(dev or 0)
            c = Codec(self.codec_config, self.state_dict())
            self.__dict__["_codec_obj"] = c
        return c

    def _train_codec(self) -> Codec:
        """The training context: fp32 CNN (the reference trains in fp32), weights pushed from the parameters before every
        forward pass (`llicti_set_weights_dev`)."""
        c = self.__dict__.get("_train_codec_obj")
        if c is None:
            if not torch.cuda.is_available():
                raise RuntimeError("llicti_b200.LLICTI needs a CUDA device; there is no CPU fallback")
            p = next(self.parameters())
            if not p.is_cuda:
                raise RuntimeError("training needs the model on a CUDA device: model.to('cuda')")
            cfg = CodecConfig(**{**self.codec_config.__dict__, "cnn_impl": L.CNN_FP32, "device": int(p.device.index or 0)})
            c = Codec(cfg, self.state_dict())
            self.__dict__["_train_codec_obj"] = c
        return c

    # -- reference interface ----------------------------------------------------------------------
    def forward(self, x):
        """x float32 [B,3,H,W] in [0,1] (uint8/255), H and W multiples of 2^num_scales -> list[num_scales] of
        float32 [B,9,Hs,Ws] self-informations (reference :101-123): the rate-estimation path of validate() and the
        forward half of the training step.  In training mode (`model.train()`, the reference's agent sets it before every
        epoch), with gradients enabled and parameters that require them, the result carries an autograd edge whose backward is `llicti_backward_dev` (fp32 CNN), so the reference's
        `loss.backward(); optimizer.step()` works on it unchanged; otherwise it is inference only."""
        assert x.dim() == 4 and x.shape[1] == 3, "expected [B,3,H,W]"
        named = [(k, p) for k, p in self.named_parameters() if p.requires_grad]
        if self.training and torch.is_grad_enabled() and named:
            dev = named[0][1].device
            rgb = torch.round(x.detach().to(dev) * 255).to(torch.uint8).contiguous()
            c = self.__dict__.get("_codec_obj")
            if c is not None:                      # the coding context holds weights that are about to change
                c.close()
                self.__dict__["_codec_obj"] = None
            names = tuple(k for k, _ in named)
            if len(names) != 24:
                raise RuntimeError("training needs all 24 weight tensors to require gradients")
            return list(_SelfInformations.apply(self, rgb, names, *[p for _, p in named]))
        with torch.no_grad():
            codec = self._codec()
            rgb = torch.round(x.to(codec.device) * 255).to(torch.uint8).contiguous()
            return codec.forward_dev(rgb)

    @torch.no_grad()
    def compress(self, x: torch.Tensor):
        """x float32 [1,3,H,W] in [0,1] (uint8/255, what the dataloader yields) ->
        (bytestream_list, x_ycocg) as the reference (:125-159)."""
        assert x.dim() == 4 and x.shape[1] == 3, "expected [B,3,H,W]"
        if x.shape[0] != 1:
            raise ValueError("compress() mirrors the reference's batch-1 call; use compress_batch() for batches")
        codec = self._codec()
        rgb = torch.round(x.to(codec.device) * 255).to(torch.uint8).contiguous()
        blob, off, mm = codec.encode_dev(rgb)
        codec.check_status()
        off_h = off.cpu().numpy().astype(np.uint64)
        blob_h = blob[:int(off_h[-1])].cpu().numpy()
        bsl = codec.to_bytestream_lists(rgb.cpu().numpy(), blob_h, off_h, mm.cpu().numpy())[0]
        return bsl, self._ycocg_float(rgb)

    @torch.no_grad()
    def decompres(self, bytestream_list, devc=None, xorg=None):
        """bytestream_list -> float32 [1,3,H,W] RGB with values k/255 on `devc` (:161-179)."""
        codec = self._codec()
        rgb = codec.decompress_images([bytestream_list])
        out = torch.from_numpy(rgb).to(devc if devc is not None else codec.device).to(torch.float32) / 255
        return out

    decompress = decompres

    @torch.no_grad()
    def compress_batch(self, rgb_uint8) -> List[list]:
        """uint8 [n,3,H,W] (numpy or tensor) -> list of n bytestream_lists."""
        if isinstance(rgb_uint8, torch.Tensor):
            rgb_uint8 = rgb_uint8.cpu().numpy()
        return self._codec().compress_images(np.ascontiguousarray(rgb_uint8))

    @torch.no_grad()
    def decompress_batch(self, bytestream_lists) -> np.ndarray:
        return self._codec().decompress_images(bytestream_lists)

    @staticmethod
    def _ycocg_float(rgb_u8: torch.Tensor) -> torch.Tensor:
        """Second return value of compress(): YCoCg-R (Y-127) / 255 as float (reference :135-144).
        Bookkeeping output only; the coded planes are produced inside the CUDA library."""
        r, g, b = (rgb_u8[:, i:i + 1].to(torch.int16) for i in range(3))
        co = r - b
        t = b + torch.div(co, 2, rounding_mode="floor")
        cg = g - t
        y = t + torch.div(cg, 2, rounding_mode="floor") - 127
        return torch.cat((y, co, cg), dim=1) / 255
