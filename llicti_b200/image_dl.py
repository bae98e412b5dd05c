"""Image loading for eval_model, validate and train (mirrors dataloaders/image_dl.py:40-51, 60-78, 106-111: every *.png / *.jpg of
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


class TrainImageLoader:
    """The reference's train_loader (image_dl.py:18-39, 54-104): every epoch visits the images of the training
    directories in a new random order; each yields `patches_per_img` random crops of `size` x `size`, flipped left-right
    with probability 1/2; images smaller than the crop are resized to fit it (ImageOps.fit); batches of `batch_size`
    images (so [B, 3, size, size], or [B, patches_per_img, 3, size, size] like the reference, which the training loop
    flattens).  The random stream is a numpy Generator seeded by the caller (the reference relies on torch's global
    seed and the DataLoader workers' own).  With `world` > 1 (data-parallel training under torchrun) a rank visits every
    world-th image of the epoch's shared order, `batch_size` of them per step: the global batch is world x batch_size."""

    def __init__(self, roots, size, batch_size, patches_per_img=1, seed=0, rank=0, world=1):
        roots = [roots] if isinstance(roots, str) else list(roots)
        self.rank, self.world = int(rank), max(int(world), 1)
        self.files = [f for r in roots for f in list_images(r)]
        self.size, self.batch_size, self.patches = int(size), max(int(batch_size), 1), max(int(patches_per_img), 1)
        if self.size <= 0:
            raise ValueError("training needs patch_size > 0")
        self.order_rng = np.random.default_rng(seed)                 # the epoch's order: the same stream on every rank
        self.rng = np.random.default_rng([seed, self.rank])          # crops and flips: a stream per rank

    def __len__(self):
        from .shard import rank_share
        return -(-len(rank_share(range(len(self.files)), self.rank, self.world)) // self.batch_size)

    def _patches(self, path):
        from PIL import Image, ImageOps
        with open(path, "rb") as f:
            img = Image.open(f).convert("RGB")
        w, h = img.size
        if w < self.size or h < self.size:
            img = ImageOps.fit(img, (max(w, self.size) if w >= self.size else self.size,
                                     max(h, self.size) if h >= self.size else self.size))
        a = np.asarray(img)
        out = []
        for _ in range(self.patches):
            top = int(self.rng.integers(0, a.shape[0] - self.size + 1))
            left = int(self.rng.integers(0, a.shape[1] - self.size + 1))
            p = a[top:top + self.size, left:left + self.size]
            if self.rng.random() < 0.5:
                p = p[:, ::-1]
            out.append(torch.from_numpy(np.ascontiguousarray(p.transpose(2, 0, 1))).to(torch.float32) / 255)
        return out[0] if self.patches == 1 else torch.stack(out)

    def __iter__(self):
        from .shard import rank_share
        order = rank_share(self.order_rng.permutation(len(self.files)), self.rank, self.world)
        for k in range(0, len(order), self.batch_size):
            yield torch.stack([self._patches(self.files[i]) for i in order[k:k + self.batch_size]])
