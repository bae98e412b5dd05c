"""Image loading for eval_model and validate (mirrors dataloaders/image_dl.py:40-51, 60-78, 106-111: every *.png / *.jpg of
`config.test_data` as float32 [1,3,H,W] in [0,1], batch 1; `config.valid_data` centre-cropped to `val_patch_size` in
batches of `val_batch_size`)."""
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


class ValidImageLoader:
    """The reference's valid_loader (image_dl.py:27-29, 46-51): centre crops of `size` x `size` (zero-padded when the
    image is smaller, as torchvision's CenterCrop), batches of `batch_size`, file order.  The reference also applies
    RandomHorizontalFlip to validation crops (:76); a validation loss that depends on a coin is not reproduced here."""

    def __init__(self, root, size, batch_size):
        self.files = list_images(root)
        self.size, self.batch_size = int(size), max(int(batch_size), 1)

    def __len__(self):
        return -(-len(self.files) // self.batch_size) if self.size > 0 else len(self.files)

    def _load(self, path):
        from PIL import Image
        with open(path, "rb") as f:
            img = np.asarray(Image.open(f).convert("RGB"))
        if self.size > 0:
            h, w, s = img.shape[0], img.shape[1], self.size
            canvas = np.zeros((max(h, s), max(w, s), 3), img.dtype)
            top, left = (canvas.shape[0] - h) // 2, (canvas.shape[1] - w) // 2
            canvas[top:top + h, left:left + w] = img
            ct, cl = int(round((canvas.shape[0] - s) / 2.0)), int(round((canvas.shape[1] - s) / 2.0))
            img = canvas[ct:ct + s, cl:cl + s]
        return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).to(torch.float32) / 255

    def __iter__(self):
        if self.size <= 0:                       # whole images: sizes differ, one per batch
            for path in self.files:
                yield self._load(path)[None]
            return
        for k in range(0, len(self.files), self.batch_size):
            yield torch.stack([self._load(p) for p in self.files[k:k + self.batch_size]])
