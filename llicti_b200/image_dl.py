"""Test-image loading for eval_model (mirrors dataloaders/image_dl.py:40-45, 60, 106-111:
every *.png / *.jpg of `config.test_data`, RGB, float32 [1,3,H,W] in [0,1], batch 1)."""
import os

import numpy as np
import torch


def list_images(root):
    if not os.path.isdir(root):
        raise SystemExit(f"Dataset could not be found: {root}")
    return [os.path.join(root, f) for f in sorted(os.listdir(root)) if f.endswith(".png") or f.endswith(".jpg")]


class TestImageLoader:
    def __init__(self, root):
        self.files = list_images(root)

    def __len__(self):
        return len(self.files)

    def __iter__(self):
        from PIL import Image
        for path in self.files:
            with open(path, "rb") as f:
                img = np.asarray(Image.open(f).convert("RGB"))
            x = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).to(torch.float32) / 255
            yield x[None]
