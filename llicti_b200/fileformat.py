"""On-disk / wire form of one compressed image.

The reference never serialises its result: `compress` returns a list of lists of `bytes`
(`bytestream_list`, graphs/models/LLICTI_nets.py:346-354, 409-411) that lives in RAM until
`decompres` consumes it.  A `.llicti` file is exactly that list, length-prefixed:

    magic   8 B   b"LLICTI" + u8 format version (1) + u8 container mode (0 torchac streams, 1 substreams)
    u8            number of scales S
    u32 LE        sub_len (0 in mode 0)
    u32, u32 LE   image height, width (redundant with the header row; checked on read)
    (1 + S) rows of 9 entries:  u32 LE length + payload
      row 0       the reference's header row [dims, minmax, pad word, raw coarsest band, tag, b'' x 4]
      row 1..S    the 9 streams (3 bands x 3 channels) of scale S-1 .. 0

Pure host logic; no GPU involved.
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

from . import container

MAGIC = b"LLICTI"
VERSION = 1


def dumps(bsl: Sequence[Sequence[bytes]], sub_len: int, H: int, W: int) -> bytes:
    S = len(bsl) - 1
    if S < 1 or any(len(r) != 9 for r in bsl):
        raise ValueError("bytestream_list must have 1 + S rows of 9 entries")
    if container.parse_mode_tag(bytes(bsl[0][4])) != max(sub_len, 0):
        raise ValueError("sub_len does not match the container tag of the header row")
    out = [MAGIC, bytes([VERSION, 1 if sub_len > 0 else 0, S]), struct.pack("<III", max(sub_len, 0), H, W)]
    for row in bsl:
        for e in row:
            e = bytes(e)
            out.append(struct.pack("<I", len(e)))
            out.append(e)
    return b"".join(out)


def loads(data: bytes) -> Tuple[List[List[bytes]], int, int, int]:
    """-> (bytestream_list, sub_len, H, W); raises ValueError on anything malformed."""
    mv = memoryview(data)
    if len(mv) < 21 or bytes(mv[:6]) != MAGIC:
        raise ValueError("not a .llicti stream (bad magic)")
    version, mode, S = mv[6], mv[7], mv[8]
    if version != VERSION:
        raise ValueError(f"unsupported .llicti version {version}")
    sub_len, H, W = struct.unpack_from("<III", mv, 9)
    if (mode == 0) != (sub_len == 0) or mode > 1 or S < 1:
        raise ValueError("inconsistent container mode / sub_len / scale count")
    pos = 21
    bsl = []
    for _ in range(1 + S):
        row = []
        for _ in range(9):
            if pos + 4 > len(mv):
                raise ValueError("truncated .llicti stream")
            (n,) = struct.unpack_from("<I", mv, pos)
            pos += 4
            if pos + n > len(mv):
                raise ValueError("truncated .llicti stream")
            row.append(bytes(mv[pos:pos + n]))
            pos += n
        bsl.append(row)
    if pos != len(mv):
        raise ValueError("trailing bytes after the last stream")
    hdr = bsl[0]
    if len(hdr[0]) != 3 or hdr[0][0] != S or container.parse_mode_tag(hdr[4]) != sub_len:
        raise ValueError("header row disagrees with the file header")
    pad_int = int.from_bytes(hdr[2], "little") if len(hdr[2]) == 2 else -1
    if pad_int < 0 or container.image_size_from_header(S, hdr[0][1], hdr[0][2], pad_int) != (H, W):
        raise ValueError("image size in the file header disagrees with the header row")
    return bsl, sub_len, H, W


def write(path: str, bsl, sub_len: int, H: int, W: int) -> int:
    data = dumps(bsl, sub_len, H, W)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def read(path: str):
    with open(path, "rb") as f:
        return loads(f.read())
