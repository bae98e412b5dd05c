"""llicti_b200 -- B200-native compress / decompress path of LLICTI behind the reference's
`main.py configs/*.json` (eval_model) entry point and `LLICTI` model interface.

    from llicti_b200 import LLICTI            # drop-in for graphs.models.LLICTI_nets.LLICTI
    from llicti_b200 import Codec, CodecConfig  # batch API over the C ABI (include/llicti.h)

All computation lives in libllicti_b200.so (hand-written sm_100a CUDA, llicti_b200/csrc);
build it with `python -m llicti_b200.build`.  There is no CPU / PyTorch fallback.
"""
from . import _lib
from .codec import Codec, CodecConfig
from .model import LLICTI

__all__ = ["LLICTI", "Codec", "CodecConfig", "_lib"]
