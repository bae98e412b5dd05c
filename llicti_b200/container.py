"""Pure host logic around the reference's `bytestream_list` (no GPU needed):
header assembly / parsing (LLICTI_nets.py:346-354, 420-431, 533-542) and the flat
(blob, offsets) form the C ABI works on."""
import zlib
from typing import Sequence

import numpy as np


CONTAINER_VERSION = 2     # 2: substreams of sub_len symbols at the finest scale, sub_len / 2 at the coarser ones (1: sub_len everywhere)


def mode_tag(sub_len: int) -> bytes:
    """Header slot 4 (b'' in the reference): b'' = torchac-compatible streams, otherwise
    [container version, sub_len as u32 LE] = interleaved-substream container."""
    if sub_len <= 0:
        return b""
    return bytes([CONTAINER_VERSION]) + int(sub_len).to_bytes(4, "little")


def parse_mode_tag(tag: bytes) -> int:
    if len(tag) == 0:
        return 0
    if len(tag) == 5 and tag[0] == 1:
        raise ValueError("substream container version 1 (one substream length for every scale) is no longer read")
    if len(tag) != 5 or tag[0] != CONTAINER_VERSION:
        raise ValueError("unknown container tag in header slot 4")
    return int.from_bytes(tag[1:5], "little")


CNN_ARITHMETIC = {0: "fp32", 1: "tcgen05/bf16", 2: "tcgen05/fp16"}       # llicti_cnn_operands()


def fingerprint(cnn_operands: int, numerics: int, weights_crc: int) -> bytes:
    """Header slot 5 (b'' in the reference): what the decoder must share with the encoder for the CDFs to agree --
    [2, CNN arithmetic (0 fp32, 1 tcgen05 with bf16 operands, 2 tcgen05 with fp16 operands), numerics profile of the
    CDF stage, crc32 of the fp32 weights as u32 LE].  Streams are decodable only by a codec with the same fingerprint:
    other operand types give other network outputs, hence other tables."""
    return bytes([2, cnn_operands & 0xFF, numerics & 0xFF]) + int(weights_crc & 0xFFFFFFFF).to_bytes(4, "little")


def describe_fingerprint(fp: bytes) -> str:
    if len(fp) != 7 or fp[0] != 2:
        return "unknown fingerprint"
    return (f"cnn={CNN_ARITHMETIC.get(fp[1], fp[1])}, numerics={fp[2]}, "
            f"weights crc32={int.from_bytes(fp[3:7], 'little'):08x}")


def check_fingerprint(hdr, expected: bytes):
    """Raise ValueError when a stream carries a fingerprint other than this codec's.  Streams without one (made by
    the reference itself, or by an older version of this library) are not checked."""
    got = bytes(hdr[5]) if len(hdr) > 5 else b""
    if len(got) == 0 or not expected:
        return
    if len(got) != 7 or got[0] != 2:
        raise ValueError("unknown fingerprint in header slot 5")
    if got != expected:
        raise ValueError(f"stream was coded with {describe_fingerprint(got)}, this codec has {describe_fingerprint(expected)}: "
                         "the CDF tables would differ and the decode would be garbage")


def image_checksum(hdr):
    """crc32 of the original uint8 RGB planes (header slot 6), or None."""
    c = bytes(hdr[6]) if len(hdr) > 6 else b""
    if len(c) == 0:
        return None
    if len(c) != 4:
        raise ValueError("malformed image checksum in header slot 6")
    return int.from_bytes(c, "little")


def image_size_from_header(num_scales: int, h_last: int, w_last: int, pad_int: int):
    """Undo the pyramid's size bookkeeping: H_{s-1} = 2*H_s - padH_s (pad word has level 0 in
    its most significant bit pair, LLICTI_nets.py:230)."""
    h, w = h_last, w_last
    for s in range(num_scales - 1, -1, -1):
        bits = (pad_int >> (2 * (num_scales - 1 - s))) & 3
        h, w = 2 * h - (bits >> 1), 2 * w - (bits & 1)
    return h, w


def _header_row(bsl):
    """The header row of a bytestream_list, shape-checked (LLICTI_nets.py:346-354)."""
    if len(bsl) < 2 or len(bsl[0]) < 4:
        raise ValueError("bytestream_list has no header row")
    hdr = bsl[0]
    if len(hdr[0]) != 3 or len(hdr[1]) != 12 or len(hdr[2]) != 2:
        raise ValueError("malformed header row (expected 3 + 12 + 2 bytes of dims, min/max and pad word)")
    return hdr


def stream_size(bsl):
    """(H, W) of the image a bytestream_list codes, from its header row alone."""
    hdr = _header_row(bsl)
    ns, h_last, w_last = (int(v) for v in np.frombuffer(hdr[0], dtype=np.uint8))
    if len(bsl) != ns + 1:
        raise ValueError(f"header says {ns} scales, list has {len(bsl) - 1}")
    pad_int = int(np.frombuffer(hdr[2], dtype=np.uint16)[0])
    return image_size_from_header(ns, h_last, w_last, pad_int)


def assemble(num_scales: int, sub_len: int, h_last: int, w_last: int, pad_int: int, rgb: np.ndarray,
             blob, off: np.ndarray, minmax: np.ndarray, fp: bytes = b"", checksum: bool = False):
    """(blob, off, minmax) of a batch -> list of bytestream_lists
    [[hdr(3B), minmax(12B), pad(2B), x00 raw RGB, tag, fingerprint, crc32(rgb), b'' x2], 9 streams per scale S-1..0].
    Slots 0-3 are the reference's (LLICTI_nets.py:346-354, which leaves slots 4-8 empty and never reads them back)."""
    n = rgb.shape[0]
    S = num_scales
    st = 2 ** S
    blob_b = blob.tobytes() if isinstance(blob, np.ndarray) else bytes(blob)
    out = []
    for i in range(n):
        hdr = [bytes([S, h_last, w_last]),
               np.asarray(minmax[i], dtype=np.int16).tobytes(),
               np.array([pad_int & 0xFFFF], dtype=np.uint16).tobytes(),
               np.ascontiguousarray(rgb[i, :, 0::st, 0::st]).tobytes(),
               mode_tag(sub_len), bytes(fp),
               (zlib.crc32(np.ascontiguousarray(rgb[i]).data) & 0xFFFFFFFF).to_bytes(4, "little") if checksum else b"",
               b"", b""]
        rows = [hdr]
        for r in range(S):
            base = i * 9 * S + r * 9
            rows.append([blob_b[int(off[base + j]):int(off[base + j + 1])] for j in range(9)])
        out.append(rows)
    return out


def parse(num_scales: int, sub_len: int, bsls: Sequence, fp: bytes = b""):
    """list of bytestream_lists -> (blob u8, off u64 [n*9S+1], minmax i16 [n,6], x00 u8
    [n,3,h,w], n, H, W).  Raises ValueError on inconsistent headers, and when a stream's fingerprint
    (header slot 5) differs from `fp`."""
    S = num_scales
    n = len(bsls)
    if n == 0:
        raise ValueError("empty batch")
    dims = None
    mm = np.empty((n, 6), dtype=np.int16)
    x00s, parts, offs, pos = [], [], [0], 0
    for i, bsl in enumerate(bsls):
        hdr = _header_row(bsl)
        ns, h_last, w_last = (int(v) for v in np.frombuffer(hdr[0], dtype=np.uint8))
        if ns != S or len(bsl) != S + 1:
            raise ValueError(f"stream has {ns} scales, model has {S}")               # LLICTI_nets.py:424
        got = parse_mode_tag(hdr[4] if len(hdr) > 4 else b"")
        if got != sub_len:
            raise ValueError(f"stream coded with sub_len={got}, codec configured with {sub_len}")
        check_fingerprint(hdr, fp)
        mm[i] = np.frombuffer(hdr[1], dtype=np.int16)
        pad_int = int(np.frombuffer(hdr[2], dtype=np.uint16)[0])
        hw = image_size_from_header(S, h_last, w_last, pad_int)
        if dims is None:
            dims = hw
        elif dims != hw:
            raise ValueError("all images of a batch must have the same size")
        if len(hdr[3]) != 3 * h_last * w_last:
            raise ValueError("raw coarsest band has the wrong length")
        x00s.append(np.frombuffer(hdr[3], dtype=np.uint8).reshape(3, h_last, w_last))
        for r in range(1, S + 1):
            if len(bsl[r]) != 9:
                raise ValueError("expected 9 streams per scale")
            for j in range(9):
                parts.append(bsl[r][j])
                pos += len(bsl[r][j])
                offs.append(pos)
    blob = np.frombuffer(b"".join(parts), dtype=np.uint8)
    if blob.size == 0:
        blob = np.zeros(1, dtype=np.uint8)
    return blob, np.array(offs, dtype=np.uint64), mm, np.stack(x00s), n, dims[0], dims[1]
