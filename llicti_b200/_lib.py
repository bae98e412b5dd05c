"""ctypes binding of libllicti_b200.so (include/llicti.h).

The library is built in-tree by `python -m llicti_b200.build` (or __graft_entry__.build()).
There is no fallback: if the shared object is missing, loading raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libllicti_b200.so")

MAX_SCALES = 8
OK, E_ARG, E_CUDA, E_NOMEM, E_STREAM, E_NODEVICE, E_TIMEOUT = 0, -1, -2, -3, -4, -5, -6
NUM_TORCH_CUDA, NUM_TORCH_CPU = 0, 1
CNN_FP32, CNN_TCGEN05 = 0, 1
KERNEL_CLASSES = ("split", "cnn", "bounds", "encode", "compact", "index", "decode", "merge", "window")


class Config(C.Structure):
    _fields_ = [("num_scales", C.c_int32), ("chs", C.c_int32), ("num_mixtures", C.c_int32),
                ("sub_len", C.c_int32), ("numerics", C.c_int32), ("cnn_impl", C.c_int32),
                ("device", C.c_int32), ("decode_impl", C.c_int32)]


class Weights(C.Structure):
    _fields_ = [("l0_w", C.c_void_p * 6), ("l0_b", C.c_void_p * 6),
                ("l1_w", C.c_void_p * 3), ("l1_b", C.c_void_p * 3),
                ("l2_w", C.c_void_p * 3), ("l2_b", C.c_void_p * 3)]


class Geom(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("num_scales", C.c_int32),
                ("Hs", C.c_int32 * MAX_SCALES), ("Ws", C.c_int32 * MAX_SCALES),
                ("padH", C.c_int32 * MAX_SCALES), ("padW", C.c_int32 * MAX_SCALES),
                ("pad_int", C.c_int32),
                ("crop_h", (C.c_int32 * 3) * MAX_SCALES), ("crop_w", (C.c_int32 * 3) * MAX_SCALES),
                ("num_sub", (C.c_int32 * 3) * MAX_SCALES),
                ("positions", C.c_int64), ("symbols", C.c_int64), ("substreams", C.c_int64),
                ("max_stream_bytes", C.c_int64)]


class EncodeItem(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("out", C.c_void_p), ("out_cap", C.c_size_t),
                ("stream_off", C.c_void_p), ("minmax", C.c_void_p)]


class DecodeItem(C.Structure):
    _fields_ = [("blob", C.c_void_p), ("stream_off", C.c_void_p), ("minmax", C.c_void_p), ("x00_rgb", C.c_void_p),
                ("H", C.c_int32), ("W", C.c_int32), ("rgb_out", C.c_void_p)]


class LlictiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libllicti_b200 error {code}: {msg}")
        self.code = code


_PROTOS = {
    # name: (restype, argtypes)
    "llicti_abi_version": (C.c_int, []),
    "llicti_last_error": (C.c_char_p, []),
    "llicti_device_count": (C.c_int, []),
    "llicti_geometry": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.POINTER(Geom)]),
    "llicti_create": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(C.c_void_p)]),
    "llicti_destroy": (None, [C.c_void_p]),
    "llicti_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "llicti_color_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.c_void_p, C.c_void_p]),
    "llicti_merge_color": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "llicti_cnn_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "llicti_cdf_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "llicti_cdf_bounds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "llicti_ac_encode_bounds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p]),
    "llicti_ac_decode_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "llicti_encode_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "llicti_encode_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "llicti_decode_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p]),
    "llicti_decode_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "llicti_forward_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.c_void_p]),
    "llicti_set_weights_dev": (C.c_int, [C.c_void_p, C.POINTER(Weights), C.c_void_p]),
    "llicti_train_forward_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "llicti_backward_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(Weights), C.c_void_p]),
    "llicti_cnn_operands": (C.c_int, [C.c_void_p]),
    "llicti_encode_batch_host": (C.c_int, [C.c_void_p, C.POINTER(EncodeItem), C.c_int, C.c_void_p]),
    "llicti_decode_batch_host": (C.c_int, [C.c_void_p, C.POINTER(DecodeItem), C.c_int, C.c_void_p]),
    "llicti_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llicti_launch_count": (C.c_int64, [C.c_void_p]),
    "llicti_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "llicti_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "llicti_decode_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "llicti_selftest_fdiv": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.POINTER(C.c_uint64)]),
}

EXPORTS = tuple(_PROTOS)

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m llicti_b200.build` "
                "(llicti_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.llicti_abi_version() != 1:
            raise ImportError("libllicti_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise LlictiError(rc, load().llicti_last_error().decode("utf-8", "replace"))
