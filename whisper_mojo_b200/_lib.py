"""ctypes binding of libwhisper_b200.so (include/whisper_b200.h).

The product path has NO CPU fallback: importing this module needs the built shared library, and
every compute call raises `WhisperB200Error` when the CUDA side fails (no device, launch error...).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# WB_PRECISION=bf16 in the environment selects the bf16 build of the library (A/B measurements; the
# default fp16 build is the one that meets the parity bar, see csrc/dtype.h)
PRECISION = os.environ.get("WB_PRECISION", "fp16").lower()
if PRECISION not in ("fp16", "bf16"):
    raise ImportError(f"WB_PRECISION={PRECISION!r}: expected fp16 or bf16")
LIB_PATH = os.path.join(HERE, "libwhisper_b200.so" if PRECISION == "fp16" else "libwhisper_b200_bf16.so")

WB_OK, WB_ERR_ARG, WB_ERR_CUDA, WB_ERR_IO, WB_ERR_STATE = 0, 1, 2, 3, 4


class WhisperB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libwhisper_b200 error {code}: {msg}")
        self.code = code


# name -> (restype, argtypes); every symbol the header declares (tests check the export list against this)
_PF, _PI32 = POINTER(c_float), POINTER(c_int32)
SIGNATURES = {
    "wb_last_error": (c_int, [ctypes.c_char_p, c_size_t]),
    "wb_abi_version": (c_int, []),
    "wb_kernel_launch_count": (c_int64, []),
    "wb_precision": (c_char_p, []),
    "wt_tensor_alloc": (c_int, [c_int64, c_int64, POINTER(c_uint64)]),
    "wt_tensor_view": (c_int, [c_uint64, c_int64, c_int64, c_int64, POINTER(c_uint64)]),
    "wt_tensor_free": (c_int, [c_uint64]),
    "wt_tensor_shape": (c_int, [c_uint64, POINTER(c_int64), POINTER(c_int64)]),
    "wt_tensor_data": (c_int, [c_uint64, POINTER(c_void_p)]),
    "wt_tensor_upload": (c_int, [c_uint64, c_int64, c_void_p, c_int64]),
    "wt_tensor_download": (c_int, [c_uint64, c_int64, c_void_p, c_int64]),
    "wt_tensor_copy": (c_int, [c_uint64, c_int64, c_uint64, c_int64, c_int64]),
    "wt_matmul": (c_int, [c_uint64, c_uint64, c_uint64, c_uint64]),
    "wt_layer_norm": (c_int, [c_uint64, c_uint64, c_uint64, c_uint64, c_float]),
    "wt_gelu": (c_int, [c_uint64]),
    "wt_softmax": (c_int, [c_uint64]),
    "wt_transpose_conv_weights": (c_int, [c_uint64, c_int, c_int, c_int, POINTER(c_uint64)]),
    "wt_conv1d": (c_int, [c_uint64, c_uint64, c_uint64, c_uint64, c_int, c_int, c_int]),
    "wt_argmax": (c_int, [c_uint64, POINTER(c_int64)]),
    "wt_add": (c_int, [c_uint64, c_uint64, c_uint64]),
    "wt_scale_mask": (c_int, [c_uint64, c_float, c_int, c_int64]),
    "wt_embed": (c_int, [c_uint64, c_uint64, c_uint64, c_void_p, c_int, c_int]),
    "wt_transpose": (c_int, [c_uint64, c_uint64]),
    "wm_create": (c_int, [c_void_p, c_void_p, POINTER(c_uint64)]),
    "wm_destroy": (c_int, [c_uint64]),
    "wm_stream": (c_int, [c_uint64, POINTER(c_void_p)]),
    "wm_synchronize": (c_int, [c_uint64]),
    "wm_weight_count": (c_int64, [c_void_p]),
    "wm_load_weights_file": (c_int, [c_uint64, c_char_p]),
    "wm_load_weights": (c_int, [c_uint64, c_void_p, c_int64]),
    "wm_weight_tensor": (c_int, [c_uint64, c_int, POINTER(c_void_p), POINTER(c_int64)]),
    "wm_set_option": (c_int, [c_uint64, c_char_p, c_int64]),
    "wm_logmel": (c_int, [c_uint64, c_void_p, c_int, c_void_p]),
    "wm_logmel_dev": (c_int, [c_uint64, c_void_p, c_int, c_void_p]),
    "wm_encode": (c_int, [c_uint64, c_void_p, c_int, c_void_p]),
    "wm_encode_dev": (c_int, [c_uint64, c_void_p, c_int, c_void_p]),
    "wm_kvcache_create": (c_int, [c_uint64, c_int, c_int, POINTER(c_uint64)]),
    "wm_kvcache_destroy": (c_int, [c_uint64]),
    "wm_kvcache_reset": (c_int, [c_uint64]),
    "wm_kvcache_len": (c_int, [c_uint64, POINTER(c_int)]),
    "wm_kvcache_set_encoder_dev": (c_int, [c_uint64, c_uint64, c_void_p]),
    "wm_decode_step": (c_int, [c_uint64, c_uint64, c_void_p, c_int, c_void_p, c_void_p]),
    "wm_transcribe": (c_int, [c_uint64, c_void_p, c_int, c_void_p, c_void_p]),
    "wm_transcribe_dev": (c_int, [c_uint64, c_void_p, c_int, c_void_p, c_void_p]),
    "wm_transcribe_pcm": (c_int, [c_uint64, c_void_p, c_int, c_void_p, c_void_p]),
    "wm_transcribe_pcm_dev": (c_int, [c_uint64, c_void_p, c_int, c_void_p, c_void_p]),
    "wm_set_stop_lengths": (c_int, [c_uint64, c_void_p, c_int]),
    "wm_teacher_forced": (c_int, [c_uint64, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "wm_last_timing": (c_int, [c_uint64, POINTER(c_float)]),
    "wm_last_kernel_timing": (c_int, [c_uint64, c_char_p, POINTER(c_float), POINTER(c_int64)]),
    "wb_debug_gemm": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                              c_int, c_void_p, c_int, c_void_p]),
    "wb_debug_decode_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "wb_debug_encoder_attention": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "wb_debug_cross_attention_absorbed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (building it is the job of `python -m whisper_mojo_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m whisper_mojo_b200.build` "
                "(there is no CPU fallback for this package)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().wb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc != WB_OK:
        raise WhisperB200Error(rc, last_error())


def precision() -> str:
    """"fp16" or "bf16": the 16-bit operand type of the loaded library."""
    return load().wb_precision().decode()


def launch_count() -> int:
    return int(load().wb_kernel_launch_count())
