#!/usr/bin/env python
"""Short-train a checkpoint with THIS repo's `mode: train` (the library's training step) on the recipe
tools/train_reference_ckpt.py runs through the unmodified reference on the CPU: the same synthetic PNGs (384 train images of
192 x 192 at four noise levels, 8 validation images of 160 x 160), patches of 128, batches of 8, Adam at 1e-3, six epochs.

GPU box only:

    python tools/train_b200_ckpt.py [--epochs 6] [--out gpurun_out/ckpt_A_b200_trained.npz]

Prints the validation rate after every epoch, then the rate `mode: validate` gives the REFERENCE-trained checkpoint
(tests/golden/ckpt_A_trained.npz) on the same validation images: two trainings of the same network on the same data, by the
reference's autograd on the CPU and by llicti_backward_dev on the GPU, should end at about the same rate.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def run_main(cfg, work, tag):
    path = os.path.join(work, tag + ".json")
    with open(path, "w") as f:
        json.dump(cfg, f, indent=1)
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), path], cwd=work, env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=1500)
    if r.returncode != 0:
        raise SystemExit(f"main.py ({tag}) failed:\n{r.stderr[-3000:]}")
    exp = os.path.join(work, "experiments", cfg["multi_exp_name"], "exp_0")
    return exp, open(os.path.join(exp, "logs", "exp_debug.log")).read(), time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--images", type=int, default=384)
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=1, help="seed of the crops / epoch orders and (torch.manual_seed) of the initial weights")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ckpt_A_b200_trained.npz"))
    args = ap.parse_args()

    import numpy as np
    import torch
    from PIL import Image
    from llicti_b200.synth import synthetic_image

    work = tempfile.mkdtemp(prefix="llicti_b200_train_")
    dirs = {k: os.path.join(work, k) for k in ("train", "valid", "test")}
    counts = {"train": args.images, "valid": 8, "test": 2}
    idx = 5000
    t0 = time.time()
    for k, d in dirs.items():                                       # the image recipe of tools/train_reference_ckpt.py
        os.makedirs(d, exist_ok=True)
        for i in range(counts[k]):
            idx += 1
            noise = (0.7, 1.5, 2.5, 4.0)[i % 4]
            size = args.size if k == "train" else 160
            Image.fromarray(np.transpose(synthetic_image(size, size, idx, noise=noise), (1, 2, 0))).save(os.path.join(d, f"{k}_{i:04d}.png"))
    print(f"{sum(counts.values())} synthetic images in {time.time() - t0:.1f} s")

    with open(os.path.join(ROOT, "configs", "llicti_A.json")) as f:
        base = json.load(f)
    common = {"val_batch_size": 1, "val_patch_size": 0, "valid_data": dirs["valid"], "test_data": dirs["test"], "seed": args.seed}
    cfg = dict(base, **common, mode="train", resume_training=False, batch_size=args.batch, patch_size=args.patch, patches_per_img=1,
               grad_acc_iters=1, loss_prnt_iters=10 ** 9, learning_rate=args.lr, max_epoch=args.epochs, validate_every=1,
               num_train_dirs=1, train_data_1=dirs["train"], multi_exp_name="b200_train")
    exp, log, secs = run_main(cfg, work, "train")
    train, valid, last = [], [], None
    for m in re.finditer(r"(Train Epoch|Valid Epoch|Train Itera)|\(\(([0-9.]+)\)\)", log):     # a table's total follows its header
        if m.group(1):
            last = m.group(1)
        elif last == "Train Epoch":
            train.append(float(m.group(2)))
        elif last == "Valid Epoch":
            valid.append(float(m.group(2)))
    steps = args.epochs * -(-args.images // args.batch)
    print(f"B200 training: {args.epochs} epochs = {steps} steps in {secs:.1f} s wall (process start, loaders, validation and checkpoints included)")
    print("  train rate per epoch (bpp):", " ".join(f"{v:.3f}" for v in train))
    print("  valid rate per epoch (bpp):", " ".join(f"{v:.3f}" for v in valid))

    sd = torch.load(os.path.join(exp, "checkpoints", "model_best.pth.tar"), map_location="cpu", weights_only=False)["state_dict"]
    arrays = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in sd.items()
              if k.startswith("entropymodel.entmdls_scale_band.0.") and "conditional_prob_model" not in k}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, **arrays)

    # the reference-trained checkpoint on the same validation images, through the same validate()
    ref_npz = os.path.join(ROOT, "tests", "golden", "ckpt_A_trained.npz")
    with np.load(ref_npz) as z:
        ref_sd = {k: torch.from_numpy(z[k]) for k in z.files}
    cfg2 = dict(base, **common, mode="validate", multi_exp_name="ref_ckpt")
    ck_dir = os.path.join(work, "experiments", "ref_ckpt", "exp_0", "checkpoints")
    os.makedirs(ck_dir, exist_ok=True)
    torch.save({"epoch": 0, "iteration": 0, "best_valid_loss": 1e9, "state_dict": ref_sd}, os.path.join(ck_dir, "model_best.pth.tar"))
    _, log2, _ = run_main(cfg2, work, "validate_ref")
    ref_valid = float(re.findall(r"\(\(([0-9.]+)\)\)", log2)[-1])
    print(f"reference-trained checkpoint (CPU, the reference's own mode: train, same recipe): valid rate {ref_valid:.3f} bpp")
    print(f"B200-trained checkpoint: best valid rate {min(valid):.3f} bpp  -> wrote {args.out}")


if __name__ == "__main__":
    main()
