"""File-level front end of the hot path:

    python -m llicti_b200.cli encode --config configs/llicti_A.json [--checkpoint ckpt.pth.tar] in.png out.llicti
    python -m llicti_b200.cli decode --config configs/llicti_A.json [--checkpoint ckpt.pth.tar] in.llicti out.png

The checkpoint is the reference's (dict with 'state_dict', agents/base.py:84-95).  Without one the
model keeps its untrained initialisation, like the reference when model_best.pth.tar is missing
(base.py:78-80), seeded with `--seed` (default 1337, the reference's manual seed) so that encoder
and decoder agree.  `--sub-len N` selects the interleaved-substream container (default:
torchac-compatible streams).
"""
import argparse
import json
import sys
import time

import numpy as np
import torch


def _model(args):
    from . import LLICTI
    cfg = json.load(open(args.config))
    torch.manual_seed(args.seed)
    model = LLICTI(cfg, sub_len=args.sub_len)
    if args.checkpoint:
        ckpt = torch.load(args.checkpoint, map_location="cpu", weights_only=False)
        model.load_state_dict(ckpt["state_dict"])
    return model.to(torch.device("cuda", args.device)).eval()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="llicti_b200.cli")
    ap.add_argument("command", choices=["encode", "decode"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--config", required=True)
    ap.add_argument("--checkpoint")
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--sub-len", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    from PIL import Image
    from . import fileformat
    if not torch.cuda.is_available():
        raise SystemExit("llicti_b200 needs a CUDA device; there is no CPU fallback")
    if args.command == "encode":
        img = np.asarray(Image.open(args.src).convert("RGB"))
        x = (torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).float() / 255)[None].cuda(args.device)
        model = _model(args)
        t0 = time.perf_counter()
        bsl, _ = model.compress(x)
        dt = time.perf_counter() - t0
        H, W = img.shape[:2]
        nbytes = fileformat.write(args.dst, bsl, args.sub_len, H, W)
        print(f"{args.src}: {W}x{H} -> {nbytes} bytes ({8 * nbytes / (H * W):.3f} bpp) in {dt * 1e3:.1f} ms")
    else:
        bsl, sub_len, H, W = fileformat.read(args.src)
        args.sub_len = sub_len
        model = _model(args)
        t0 = time.perf_counter()
        rec = model.decompres(bsl, torch.device("cuda", args.device))
        dt = time.perf_counter() - t0
        out = (rec[0] * 255).round().clamp(0, 255).byte().permute(1, 2, 0).cpu().numpy()
        Image.fromarray(out, "RGB").save(args.dst)
        print(f"{args.src}: {W}x{H} decoded in {dt * 1e3:.1f} ms -> {args.dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
