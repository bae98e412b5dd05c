"""Diagnostics for numerics decisions (run on the GPU box): which summation order / division
form reproduces PyTorch's CUDA ops bit for bit, CNN error levels, and basic timings."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import llicti_oracle as O  # noqa: E402
from llicti_b200 import Codec, CodecConfig, _lib as L  # noqa: E402
from test_gpu_parity import torch_reference_table, stage_c_inputs  # noqa: E402

print("device:", torch.cuda.get_device_name(0), "torch", torch.__version__)
ocfg, sd, d = stage_c_inputs()
M = 5
for numerics, nm in ((L.NUM_TORCH_CUDA, "torch_cuda"), (L.NUM_TORCH_CPU, "torch_cpu")):
    codec = Codec(CodecConfig(numerics=numerics), sd)
    tot = {"cuda": [0, 0], "cpu": [0, 0]}
    for (scale, band), params in d.params.items():
        yb = d.planes[scale][3 * (band + 1):3 * (band + 2)]
        d_params = torch.from_numpy(params.reshape(60, -1).copy()).cuda()
        d_y = torch.from_numpy(yb.reshape(3, -1).copy()).cuda()
        for clr in range(3):
            lo = -127 if clr == 0 else d.minmax[clr]
            hi = 128 if clr == 0 else d.minmax[3 + clr]
            got = codec.cdf_table(d_params, d_y, clr, lo, hi)
            for dev in ("cuda", "cpu"):
                pt = torch.from_numpy(params).to(dev)
                yf = torch.from_numpy(yb.astype(np.int16)).to(dev) / 255
                mu = pt[(3 + clr) * M:(4 + clr) * M].clone()
                if clr == 1:
                    mu += pt[9 * M:10 * M] * yf[0:1]
                elif clr == 2:
                    mu += pt[10 * M:11 * M] * yf[0:1] + pt[11 * M:12 * M] * yf[1:2]
                ref = torch_reference_table(params[clr * M:(clr + 1) * M], mu.cpu().numpy(),
                                            params[(6 + clr) * M:(7 + clr) * M], lo, hi, dev)
                diff = (got.cpu().to(torch.int32) - ref.cpu().to(torch.int32)).ne(0).sum().item()
                tot[dev][0] += diff
                tot[dev][1] += got.numel()
    print(f"profile {nm}: mismatching table entries vs torch-cuda {tot['cuda'][0]}/{tot['cuda'][1]}, "
          f"vs torch-cpu {tot['cpu'][0]}/{tot['cpu'][1]}")
    codec.close()

# isolate the primitives
x = torch.linspace(-6, 6, 100001, device="cuda")
print("x/255 == x*(1/255f) on cuda:", torch.equal(x / 255, x * torch.tensor(1.0 / 255.0, dtype=torch.float32, device="cuda")),
      " == true division:", torch.equal(x / 255, (x.double() / 255).float()))
w = torch.rand(1, 64, 64, 1, 5, device="cuda")
wp = w.permute(0, 4, 1, 2, 3).contiguous().permute(0, 2, 3, 4, 1)
s = torch.sum(wp, dim=4)
seq = (((wp[..., 0] + wp[..., 1]) + wp[..., 2]) + wp[..., 3]) + wp[..., 4]
ilp = (((wp[..., 0] + wp[..., 4]) + wp[..., 1]) + wp[..., 2]) + wp[..., 3]
print("cuda sum(dim of 5): == sequential", torch.equal(s, seq), " == ilp4", torch.equal(s, ilp))

# CNN error levels and a first timing
net = O.OracleNet(ocfg, sd)
codec = Codec(CodecConfig(numerics=L.NUM_TORCH_CPU), sd)
for (scale, band), params in sorted(d.params.items()):
    got = codec.cnn_params(band, torch.from_numpy(d.planes[scale][None]).cuda())[0].cpu().numpy()
    err = np.abs(got - params)
    print(f"cnn scale {scale} band {band}: max abs err {err.max():.3e}, max |ref| {np.abs(params).max():.3e}")
imgs = np.stack([O.synthetic_image(512, 768, i) for i in range(4)])
for sub_len in (0, 2048):
    c = Codec(CodecConfig(sub_len=sub_len), sd)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        blob, off, mm = c.encode_host(imgs)
        torch.cuda.synchronize(); t1 = time.time()
        bsls = c.to_bytestream_lists(imgs, blob, off, mm)
        b2, o2, m2, x00, n, H, W = c.from_bytestream_lists(bsls)
        torch.cuda.synchronize(); t2 = time.time()
        rec = c.decode_host(b2, o2, m2, x00, n, H, W)
        torch.cuda.synchronize(); t3 = time.time()
    print(f"sub_len {sub_len}: 4x768x512 encode {t1 - t0:.4f}s decode {t3 - t2:.4f}s lossless {np.array_equal(rec, imgs)} "
          f"bytes {blob.size} bpsp {blob.size * 8 / imgs.size:.3f}")
    c.close()
