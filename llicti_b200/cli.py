"""File-level front end of the hot path:

    python -m llicti_b200.cli encode --config configs/llicti_A.json [--checkpoint ckpt.pth.tar] in.png out.llicti
    python -m llicti_b200.cli decode --config configs/llicti_A.json [--checkpoint ckpt.pth.tar] in.llicti out.png
    python -m llicti_b200.cli encode-dir --config ... [--batch 16] [--workers 0] images/ streams/
    python -m llicti_b200.cli decode-dir --config ... [--batch 16] streams/ images_out/

The directory commands (section 8f, rank 3) code every *.png / *.jpg (every *.llicti) of a directory in batches of
equally sized images: the files are decoded by a pool of host threads into pinned staging buffers while the GPU
works on the previous batch (llicti_b200/ingest.py).

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


def _dir_command(args):
    import os
    from . import fileformat, ingest
    from .image_dl import list_images
    os.makedirs(args.dst, exist_ok=True)
    t0 = time.perf_counter()
    if args.command == "encode-dir":
        paths = list_images(args.src)
        model = _model(args)
        bsls = ingest.compress_files(model._codec(), paths, max_batch=args.batch, workers=args.workers)
        px = nbytes = 0
        for path, bsl in zip(paths, bsls):
            from . import container
            H, W = container.stream_size(bsl)
            nbytes += fileformat.write(os.path.join(args.dst, os.path.splitext(os.path.basename(path))[0] + ".llicti"), bsl,
                                       args.sub_len, H, W)
            px += H * W
        dt = time.perf_counter() - t0
        print(f"{len(paths)} images, {px / 1e6:.2f} MP -> {nbytes} bytes ({8 * nbytes / max(px, 1):.3f} bpp) in {dt:.2f} s "
              f"({px / 1e6 / dt:.1f} MP/s file to file)")
        return 0
    files = [os.path.join(args.src, f) for f in sorted(os.listdir(args.src)) if f.endswith(".llicti")]
    if not files:
        raise SystemExit(f"no .llicti files in {args.src}")
    loaded = [fileformat.read(f) for f in files]
    sub_lens = {x[1] for x in loaded}
    if len(sub_lens) != 1:
        raise SystemExit("the files of a directory must share one container mode")
    args.sub_len = sub_lens.pop()
    model = _model(args)
    dst = [os.path.join(args.dst, os.path.splitext(os.path.basename(f))[0] + ".png") for f in files]
    ingest.decompress_to_files(model._codec(), [x[0] for x in loaded], dst, max_batch=args.batch, workers=args.workers)
    px = sum(x[2] * x[3] for x in loaded)
    dt = time.perf_counter() - t0
    print(f"{len(files)} streams, {px / 1e6:.2f} MP decoded in {dt:.2f} s ({px / 1e6 / dt:.1f} MP/s file to file)")
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="llicti_b200.cli")
    ap.add_argument("command", choices=["encode", "decode", "encode-dir", "decode-dir"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--config", required=True)
    ap.add_argument("--checkpoint")
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--sub-len", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16, help="images per batch of the directory commands")
    ap.add_argument("--workers", type=int, default=0, help="host threads decoding / writing images (0 = one per core, at most 32)")
    args = ap.parse_args(argv)
    from PIL import Image
    from . import fileformat
    if not torch.cuda.is_available():
        raise SystemExit("llicti_b200 needs a CUDA device; there is no CPU fallback")
    if args.command in ("encode-dir", "decode-dir"):
        return _dir_command(args)
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
