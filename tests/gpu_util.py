"""Shared helpers for the GPU parity tests (all calls go through the C ABI)."""
import ctypes
from ctypes import c_void_p

import numpy as np

from whisper_mojo_b200 import _lib


def h16_round(a):
    """Round to the 16-bit operand type of the loaded library (fp16 by default, bf16 for the WB_PRECISION=bf16 build)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a, np.float32))
    return (t.half() if _lib.precision() == "fp16" else t.bfloat16()).float().numpy()


bf16_round = h16_round  # old name (tools/)


def tolerances():
    """Stated parity bounds of the fast path against the fp32 oracle, by library precision:
    (enc_out max-abs, enc_out mean-abs, logits max-abs, logits median-abs, margin tau).  fp16 is north_star's
    max-abs <= 1e-2 on enc_out and logits; tau is the oracle top-1/top-2 margin below which a differing argmax is
    within the logit bound (2 x LOGIT_MAX: both candidates may move by the bound)."""
    if _lib.precision() == "fp16":
        return 1e-2, 1.5e-3, 1e-2, 2e-3, 2e-2
    return 4e-2, 6e-3, 1e-1, 1e-2, 0.1


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
    """Greedy ids must equal the oracle's.  The only excuse for a first mismatch at sequence index i is that the
    oracle's own top-1/top-2 logit margin AT THAT STEP, margins[i - 4], is below `tau` (the histories are identical
    up to i, so no earlier margin matters).  Returns (ok, message); the message always carries matched/total."""
    got, ref = np.asarray(got), np.asarray(ref)
    k = min(len(got), len(ref))
    bad = np.nonzero(got[:k] != ref[:k])[0]
    if len(bad) == 0:
        return len(got) == len(ref), f"matched {k}/{len(ref)} ids (length {len(got)} vs {len(ref)})"
    i = int(bad[0])
    if i < 4:
        return False, f"prompt differs at {i}"
    m = float(margins[i - 4])
    return m < tau, f"matched {i}/{len(ref)} ids, first mismatch at {i} where the oracle margin is {m:.4f} (tau {tau})"


def token_report(got_list, ref_list, margin_list, tau):
    """Batch form: (all_ok, n_identical, text) over chunks."""
    ok_all, ident, lines = True, 0, []
    for c, (g, r, mg) in enumerate(zip(got_list, ref_list, margin_list)):
        ok, msg = tokens_agree_up_to_margin(g, r, mg, tau)
        same = len(g) == len(r) and np.array_equal(np.asarray(g), np.asarray(r))
        ident += int(same)
        ok_all &= ok
        if not same:
            lines.append(f"chunk {c}: {msg}")
    return ok_all, ident, f"{ident}/{len(got_list)} chunks identical to the oracle" + ("; " + "; ".join(lines) if lines else "")


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
