"""Batched image ingest / egress around the hot path.

The reference feeds `compress` one PIL-decoded image at a time from the main thread
(dataloaders/image_dl.py:40-45 `num_workers=0`, :106-111 `pil_loader`; agents/llicti_agent.py:129-149 loops
with batch 1).  At the speed of the CUDA path the PNG decode is what the job waits for, so here

  * files are decoded by a pool of host threads (PIL releases the GIL inside its codecs) straight into pinned
    uint8 staging buffers `[n, 3, H, W]`,
  * images of equal size are grouped into batches (the C ABI codes uniform batches; the order of `paths` is kept
    inside a group and every result is returned under its original index),
  * the decode of the next batch overlaps the GPU work on the current one (a bounded queue of staged batches).

Nothing here touches pixel values: a staged batch holds exactly the bytes `pil_loader(path)` yields, in the
planar layout `llicti_encode_host` takes.  The staging logic runs without a GPU (`pin=False`), which is how the
CPU tests exercise it.
"""
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterator, List, Sequence, Tuple

import numpy as np
import torch


def load_rgb(path: str) -> np.ndarray:
    """uint8 [3, H, W], the pixels of the reference's pil_loader (image_dl.py:106-111)."""
    from PIL import Image
    with open(path, "rb") as f:
        img = np.asarray(Image.open(f).convert("RGB"))
    return np.ascontiguousarray(img.transpose(2, 0, 1))


def image_size(path: str) -> Tuple[int, int]:
    """(H, W) from the file header only."""
    from PIL import Image
    with Image.open(path) as im:
        w, h = im.size
    return h, w


def plan_batches(sizes: Sequence[Tuple[int, int]], max_batch: int) -> List[List[int]]:
    """Group indices of equally sized images into batches of at most `max_batch`, keeping the input order
    inside a group; groups appear in the order of their first image."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    groups: Dict[Tuple[int, int], List[int]] = {}
    for i, s in enumerate(sizes):
        groups.setdefault((int(s[0]), int(s[1])), []).append(i)
    out = []
    for idx in groups.values():
        for k in range(0, len(idx), max_batch):
            out.append(idx[k:k + max_batch])
    return out


class StagedBatches:
    """Iterate `(indices, uint8 tensor [n, 3, H, W])` over `paths`, decoded `depth` batches ahead by `workers`
    host threads.  With `pin=True` the tensors live in pinned memory (ready for cudaMemcpyAsync); every yielded
    tensor is the caller's until the next `depth` batches have been taken (the buffers rotate)."""

    def __init__(self, paths: Sequence[str], max_batch: int = 16, workers: int = 0, depth: int = 2, pin: bool = True,
                 loader: Callable[[str], np.ndarray] = load_rgb, sizes: Sequence[Tuple[int, int]] = None):
        self.paths = list(paths)
        self.loader = loader
        self.workers = workers or min(32, (os.cpu_count() or 4))
        self.depth = max(1, int(depth))
        self.pin = bool(pin) and torch.cuda.is_available()
        with ThreadPoolExecutor(self.workers) as ex:
            self.sizes = list(sizes) if sizes is not None else list(ex.map(image_size, self.paths))
        self.batches = plan_batches(self.sizes, max_batch)

    def __len__(self):
        return len(self.batches)

    def _stage(self, ex: ThreadPoolExecutor, idx: List[int]) -> torch.Tensor:
        H, W = self.sizes[idx[0]]
        buf = torch.empty((len(idx), 3, H, W), dtype=torch.uint8, pin_memory=self.pin)
        view = buf.numpy()

        def one(slot_i):
            slot, i = slot_i
            img = self.loader(self.paths[i])
            if img.shape != (3, H, W):
                raise ValueError(f"{self.paths[i]}: decoded to {img.shape}, header said {(3, H, W)}")
            view[slot] = img

        list(ex.map(one, enumerate(idx)))
        return buf

    def __iter__(self) -> Iterator[Tuple[List[int], torch.Tensor]]:
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def produce():
            try:
                with ThreadPoolExecutor(self.workers) as ex:
                    for idx in self.batches:
                        if stop.is_set():
                            return
                        q.put((idx, self._stage(ex, idx)))
                q.put(None)
            except BaseException as e:      # surfaces in the consumer
                q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while t.is_alive():          # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.05)


def compress_files(codec, paths: Sequence[str], max_batch: int = 16, workers: int = 0) -> List[list]:
    """`bytestream_list` of every file, in the order of `paths` (Codec.compress_images per staged batch)."""
    out: List[list] = [None] * len(paths)
    for idx, batch in StagedBatches(paths, max_batch=max_batch, workers=workers):
        for i, bsl in zip(idx, codec.compress_images(batch.numpy())):
            out[i] = bsl
    return out


def decompress_to_files(codec, bsls: Sequence[list], dst_paths: Sequence[str], max_batch: int = 16, workers: int = 0):
    """Decode `bsls` in batches of equal size and write PNGs with a pool of host threads while the GPU decodes
    the next batch."""
    from PIL import Image
    assert len(bsls) == len(dst_paths)
    from . import container
    sizes = [container.stream_size(b) for b in bsls]
    workers = workers or min(32, (os.cpu_count() or 4))

    def save(arr_path):
        arr, path = arr_path
        # (zlib level 1: the PNG is a lossless container of pixels that exist, smaller, as the .llicti stream; at the default
        #  level 6 the directory egress ran at 39 MP/s on 16 host threads, all of it inside zlib)
        Image.fromarray(np.ascontiguousarray(arr.transpose(1, 2, 0)), "RGB").save(path, compress_level=1)

    pending = []
    with ThreadPoolExecutor(workers) as ex:
        for idx in plan_batches(sizes, max_batch):
            rec = codec.decompress_images([bsls[i] for i in idx])
            pending.append(ex.map(save, [(rec[k].copy(), dst_paths[i]) for k, i in enumerate(idx)]))
        for p in pending:
            list(p)
