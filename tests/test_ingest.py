"""Host-side staging around the hot path (llicti_b200/ingest.py): grouping, ordering, staging buffers and the
look-ahead queue -- no GPU needed.  The GPU test codes a directory of mixed-size PNGs through the CLI."""
import os
import threading

import numpy as np
import pytest

from llicti_b200 import container, ingest
from oracle import llicti_oracle as O


def test_plan_batches_groups_by_size_and_keeps_order():
    sizes = [(4, 6), (8, 8), (4, 6), (4, 6), (8, 8), (2, 2), (4, 6)]
    b = ingest.plan_batches(sizes, max_batch=2)
    assert b == [[0, 2], [3, 6], [1, 4], [5]]
    assert sorted(i for g in b for i in g) == list(range(len(sizes)))
    assert ingest.plan_batches([], 4) == []
    with pytest.raises(ValueError):
        ingest.plan_batches(sizes, 0)


def _write_pngs(tmp_path, shapes):
    from PIL import Image
    paths, imgs = [], []
    for i, (h, w) in enumerate(shapes):
        img = O.synthetic_image(h, w, 60 + i)
        p = os.path.join(tmp_path, f"img_{i:02d}.png")
        Image.fromarray(np.ascontiguousarray(img.transpose(1, 2, 0)), "RGB").save(p)
        paths.append(p)
        imgs.append(img)
    return paths, imgs


def test_staged_batches_deliver_every_pixel_once(tmp_path):
    shapes = [(33, 47), (64, 96), (33, 47), (64, 96), (33, 47), (17, 17)]
    paths, imgs = _write_pngs(str(tmp_path), shapes)
    assert [ingest.image_size(p) for p in paths] == shapes
    st = ingest.StagedBatches(paths, max_batch=2, workers=3, depth=1, pin=False)
    seen = {}
    for idx, batch in st:
        assert batch.dtype.is_floating_point is False and tuple(batch.shape[1:]) == (3,) + shapes[idx[0]]
        for k, i in enumerate(idx):
            assert i not in seen
            seen[i] = batch[k].numpy().copy()
    assert sorted(seen) == list(range(len(paths)))
    for i, img in enumerate(imgs):
        assert np.array_equal(seen[i], img)
    assert len(st) == 4


def test_staged_batches_surface_loader_errors_and_stop_cleanly(tmp_path):
    paths, _ = _write_pngs(str(tmp_path), [(16, 16)] * 6)

    def bad_loader(p):
        if p.endswith("03.png"):
            raise OSError("truncated file")
        return ingest.load_rgb(p)

    with pytest.raises(OSError):
        for _ in ingest.StagedBatches(paths, max_batch=2, workers=2, pin=False, loader=bad_loader):
            pass
    # leaving the loop early must not leave the producer thread blocked on its queue
    before = threading.active_count()
    it = iter(ingest.StagedBatches(paths, max_batch=1, workers=2, depth=1, pin=False))
    next(it)
    it.close()
    assert threading.active_count() <= before + 1


def test_stream_size_reads_the_header_row():
    ocfg = O.OracleConfig()
    codec = O.OracleCodec(ocfg, O.synthetic_state_dict(ocfg))
    for H, W in ((33, 47), (64, 96), (53, 77)):
        assert container.stream_size(codec.compress(O.synthetic_image(H, W, 3))) == (H, W)


@pytest.mark.gpu
def test_cli_codes_a_directory_of_mixed_sizes(tmp_path, built_lib):
    from PIL import Image
    from llicti_b200 import cli
    src, mid, dst = (os.path.join(str(tmp_path), d) for d in ("in", "streams", "out"))
    os.makedirs(src)
    shapes = [(64, 96), (48, 80), (64, 96), (64, 96), (48, 80)]
    paths, imgs = _write_pngs(src, shapes)
    cfg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "llicti_B.json")
    for sub_len in (0, 256):
        assert cli.main(["encode-dir", "--config", cfg, "--sub-len", str(sub_len), "--batch", "2", src, mid]) == 0
        assert sorted(os.listdir(mid)) == [os.path.splitext(os.path.basename(p))[0] + ".llicti" for p in paths]
        assert cli.main(["decode-dir", "--config", cfg, "--batch", "3", mid, dst]) == 0
        for p, img in zip(paths, imgs):
            rec = np.asarray(Image.open(os.path.join(dst, os.path.basename(p))).convert("RGB")).transpose(2, 0, 1)
            assert np.array_equal(rec, img), p
