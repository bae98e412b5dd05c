"""Fuzzing of the host-side parsers (no GPU): whatever bytes arrive, `.llicti` files and bytestream_lists are either
parsed to something that re-serialises identically or rejected with ValueError -- never another exception, never a
silently different image size."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from llicti_b200 import container, fileformat
from oracle import llicti_oracle as O


def _valid_file(sub_len=0, H=37, W=53, seed=5):
    rng = np.random.default_rng(seed)
    S = 2
    planes, _, pad_int = O.pyramid_split(np.zeros((3, H, W), dtype=np.int16), (0, 1))
    h_last, w_last = planes[-1].shape[1:]
    rgb = rng.integers(0, 256, size=(1, 3, H, W), dtype=np.uint8)
    lens = rng.integers(0, 40, size=9 * S)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    blob = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    mm = rng.integers(-255, 256, size=(1, 6)).astype(np.int16)
    bsl = container.assemble(S, sub_len, h_last, w_last, pad_int, rgb, blob, off, mm)[0]
    return fileformat.dumps(bsl, sub_len, H, W)


@settings(max_examples=300, deadline=None)
@given(st.data())
def test_corrupted_llicti_files_are_rejected_or_round_trip(data):
    good = bytearray(_valid_file(data.draw(st.sampled_from([0, 256]))))
    kind = data.draw(st.sampled_from(["flip", "truncate", "insert", "random"]))
    if kind == "flip":
        for _ in range(data.draw(st.integers(1, 4))):
            i = data.draw(st.integers(0, len(good) - 1))
            good[i] ^= 1 << data.draw(st.integers(0, 7))
        blob = bytes(good)
    elif kind == "truncate":
        blob = bytes(good[:data.draw(st.integers(0, len(good) - 1))])
    elif kind == "insert":
        i = data.draw(st.integers(0, len(good)))
        blob = bytes(good[:i]) + data.draw(st.binary(min_size=1, max_size=8)) + bytes(good[i:])
    else:
        blob = data.draw(st.binary(max_size=200))
    try:
        bsl, sub_len, H, W = fileformat.loads(blob)
    except ValueError:
        return
    # accepted: it must be self-consistent and re-serialise to the same bytes
    assert fileformat.dumps(bsl, sub_len, H, W) == blob
    assert container.stream_size(bsl) == (H, W)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.lists(st.binary(max_size=12), min_size=0, max_size=10), min_size=0, max_size=5), st.integers(1, 5),
       st.sampled_from([0, 64]))
def test_parse_rejects_arbitrary_bytestream_lists(rows, S, sub_len):
    try:
        blob, off, mm, x00, n, H, W = container.parse(S, sub_len, [rows])
    except ValueError:
        return
    assert n == 1 and off[-1] == sum(len(e) for r in rows[1:] for e in r) and x00.shape[1] == 3
