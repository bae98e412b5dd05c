#!/usr/bin/env python
"""How fast is the file-to-file path (PNG directory -> .llicti directory -> PNG directory) next to the GPU codec?

Writes N synthetic PNGs, runs `llicti_b200.cli encode-dir` / `decode-dir` on them (host-thread PIL decode into pinned staging
buffers, size-grouped batches, one batch of decode ahead of the GPU: llicti_b200/ingest.py) and prints the throughputs beside
the time the GPU codec alone needs for the same pixels -- i.e. how busy the ingest keeps the GPU.

    python tools/ingest_probe.py [--images 48] [--size 2040x1356] [--workers 0] [--batch 16]
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=48)
    ap.add_argument("--size", default="2040x1356")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    W, H = (int(v) for v in args.size.split("x"))
    from PIL import Image
    from llicti_b200 import cli, synth, ingest
    import json
    import torch
    cfg = os.path.join(ROOT, "configs", "llicti_A.json")
    cj = json.load(open(cfg))
    with tempfile.TemporaryDirectory() as tmp:
        src, mid, dst = (os.path.join(tmp, d) for d in ("png", "llicti", "out"))
        for d in (src, mid, dst):
            os.makedirs(d)
        # the bench's stand-in weights as a checkpoint in the reference's format (the CLI's default is the untrained initialisation)
        ckpt = os.path.join(tmp, "model_best.pth.tar")
        sd = synth.synthetic_state_dict(int(cj["chs"][0]), int(cj["num_mixtures"]), int(cj["Evens"][0]), int(cj["Odds"][0]))
        torch.save({"state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, ckpt)
        t0 = time.perf_counter()
        for i in range(args.images):
            Image.fromarray(np.ascontiguousarray(synth.synthetic_image(H, W, 3000 + i).transpose(1, 2, 0)), "RGB").save(
                os.path.join(src, f"img{i:04d}.png"), compress_level=1)
        px = args.images * H * W / 1e6
        print(f"{args.images} PNGs of {W}x{H} ({px:.0f} MP) written in {time.perf_counter() - t0:.1f} s, host cores {os.cpu_count()}")
        paths = sorted(os.path.join(src, f) for f in os.listdir(src))
        t0 = time.perf_counter()
        with __import__("concurrent.futures").futures.ThreadPoolExecutor(args.workers or (os.cpu_count() or 4)) as ex:
            list(ex.map(ingest.load_rgb, paths))
        t_pil = time.perf_counter() - t0
        print(f"host PNG decode alone ({args.workers or os.cpu_count()} threads): {px / t_pil:.0f} MP/s")
        common = ["--config", cfg, "--checkpoint", ckpt, "--sub-len", "2048", "--batch", str(args.batch), "--workers", str(args.workers)]
        for name, argv in (("encode-dir", ["encode-dir", src, mid] + common), ("decode-dir", ["decode-dir", mid, dst] + common)):
            for rep in range(2):                      # second run: files in the page cache, library warm
                buf = io.StringIO()
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(buf):
                    rc = cli.main(argv)
                dt = time.perf_counter() - t0
                assert rc == 0, buf.getvalue()
            print(f"{name}: {px / dt:.0f} MP/s file to file ({dt:.2f} s)   [{buf.getvalue().strip().splitlines()[-1]}]")
        a = np.asarray(Image.open(os.path.join(src, "img0000.png")))
        b = np.asarray(Image.open(os.path.join(dst, sorted(os.listdir(dst))[0])))
        print("first image lossless:", bool(np.array_equal(a, b)))


if __name__ == "__main__":
    main()
