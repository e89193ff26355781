"""Shared helpers for the GPU parity tests (all calls go through the C ABI)."""
import ctypes
from ctypes import c_void_p

import numpy as np

from whisper_mojo_b200 import _lib


def bf16_round(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).bfloat16().float().numpy()


def debug_gemm(impl, A, W, bias, epi, batches=1, taps=1, conv_stride=1, pad=0, rows_per_batch=None, out0=None):
    A = np.ascontiguousarray(A, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    src_rows, lda = A.shape[-2], A.shape[-1]
    N = W.shape[0]
    rows = rows_per_batch or src_rows
    out = np.zeros((batches * rows, N), np.float32) if out0 is None else np.ascontiguousarray(out0, np.float32).copy()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    _lib.check(_lib.load().wb_debug_gemm(impl, A.ctypes.data_as(c_void_p), batches, src_rows, lda, lda, taps,
                                         conv_stride, pad, rows, W.ctypes.data_as(c_void_p), N,
                                         b.ctypes.data_as(c_void_p) if b is not None else None, epi,
                                         out.ctypes.data_as(c_void_p)))
    return out


def debug_decode_attention(q, K, V, H, splits):
    B, ln, D = K.shape
    out = np.zeros((B, D), np.float32)
    _lib.check(_lib.load().wb_debug_decode_attention(q.ctypes.data_as(c_void_p), K.ctypes.data_as(c_void_p),
                                                     V.ctypes.data_as(c_void_p), B, H, ln, splits,
                                                     out.ctypes.data_as(c_void_p)))
    return out


def tokens_agree_up_to_margin(got, ref, margins, tau):
    """Greedy ids must equal the oracle's up to the first step whose fp32 top-1/top-2 margin is below
    `tau` (a bf16 pipeline cannot resolve such a step); returns (ok, message)."""
    k = min(len(got), len(ref))
    bad = np.nonzero(np.asarray(got[:k]) != np.asarray(ref[:k]))[0]
    if len(bad) == 0:
        return len(got) == len(ref), f"length {len(got)} vs {len(ref)}"
    i = int(bad[0])
    if i < 4:
        return False, f"prompt differs at {i}"
    low = np.nonzero(margins[: i - 4 + 1] < tau)[0]
    return len(low) > 0, f"first mismatch at {i}, oracle margin there {margins[i - 4]:.4f}, min margin before {margins[:i - 3].min():.4f}"


def debug_encoder_attention(impl, qkv, B, S, H):
    D = H * 64
    qkv = np.ascontiguousarray(qkv, np.float32)
    out = np.zeros((B * S, D), np.float32)
    _lib.check(_lib.load().wb_debug_encoder_attention(impl, qkv.ctypes.data_as(c_void_p), B, S, H,
                                                      out.ctypes.data_as(c_void_p)))
    return out


def debug_cross_attention_absorbed(qp, enc, H):
    B, S, D = enc.shape
    qp = np.ascontiguousarray(qp, np.float32)
    enc = np.ascontiguousarray(enc, np.float32)
    out = np.zeros((B, H * D), np.float32)
    _lib.check(_lib.load().wb_debug_cross_attention_absorbed(qp.ctypes.data_as(c_void_p), enc.ctypes.data_as(c_void_p),
                                                             B, S, D, H, out.ctypes.data_as(c_void_p)))
    return out
