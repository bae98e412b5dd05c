"""Stand-in for torchac==0.9.3's two entry points the reference calls
(graphs/models/LLICTI_nets.py:406-407, 492-493), routed to the C restatement in
oracle/torchac_port.c.  Mirrors the real package's argument checks.  Test
infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.abspath(os.path.join(_HERE, "..", ".."))
while not os.path.exists(os.path.join(_ORACLE, "torchac_port.c")) and os.path.dirname(_ORACLE) != _ORACLE:
    _ORACLE = os.path.dirname(_ORACLE)          # the copy under oracle/_ref/shims sits one level deeper


def _lib():
    so = os.path.join(_ORACLE, "_build", "libtorchac_port.so")
    src = os.path.join(_ORACLE, "torchac_port.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.oracle_ac_encode_table.restype = ctypes.c_size_t
    lib.oracle_ac_encode_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_size_t]
    lib.oracle_ac_decode_table.restype = None
    lib.oracle_ac_decode_table.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_void_p]
    return lib


_L = None


def _check(cdf, sym=None):
    if cdf.is_cuda or (sym is not None and sym.is_cuda):
        raise ValueError("CUDA tensors are not supported")
    if cdf.dtype != torch.int16 or (sym is not None and sym.dtype != torch.int16):
        raise ValueError("expected int16 tensors")
    if sym is not None and cdf.shape[:-1] != sym.shape:
        raise ValueError(f"cdf.shape[:-1] {tuple(cdf.shape[:-1])} != sym.shape {tuple(sym.shape)}")


def encode_int16_normalized_cdf(cdf_int, sym):
    global _L
    _L = _L or _lib()
    _check(cdf_int, sym)
    Lp = cdf_int.shape[-1]
    cdf = np.ascontiguousarray(cdf_int.reshape(-1, Lp).numpy())
    s = np.ascontiguousarray(sym.reshape(-1).numpy())
    n = s.shape[0]
    cap = 2 * n + 16
    out = np.empty(cap, dtype=np.uint8)
    ln = _L.oracle_ac_encode_table(cdf.ctypes.data, s.ctypes.data, n, Lp, out.ctypes.data, cap)
    assert ln <= cap
    return bytes(out[:ln])


def decode_int16_normalized_cdf(cdf_int, byte_stream):
    global _L
    _L = _L or _lib()
    _check(cdf_int)
    Lp = cdf_int.shape[-1]
    cdf = np.ascontiguousarray(cdf_int.reshape(-1, Lp).numpy())
    n = cdf.shape[0]
    buf = np.frombuffer(byte_stream, dtype=np.uint8)
    out = np.empty(n, dtype=np.int16)
    _L.oracle_ac_decode_table(cdf.ctypes.data, n, Lp, buf.ctypes.data if len(buf) else None, len(buf),
                              out.ctypes.data)
    return torch.from_numpy(out).reshape(cdf_int.shape[:-1])
