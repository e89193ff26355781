"""ctypes binding of oracle/libwhisper_oracle.so (CPU fp32 restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of whisper_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs; never from the
product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_longlong, c_void_p, c_size_t

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwhisper_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "whisper_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _Cfg(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("d_model", "n_heads", "n_layers", "vocab", "n_audio_ctx", "n_text_ctx", "n_mels")]


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(POINTER(c_float))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(POINTER(c_int))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        PF, PI = POINTER(c_float), POINTER(c_int)
        L.wo_matmul.argtypes = [PF, PF, PF, PF, c_int, c_int, c_int]
        L.wo_layer_norm.argtypes = [PF, PF, PF, PF, c_int, c_int, c_float]
        L.wo_gelu.argtypes = [PF, c_size_t]
        L.wo_softmax.argtypes = [PF, c_int, c_int]
        L.wo_transpose_conv_weights.argtypes = [PF, PF, c_int, c_int, c_int]
        L.wo_conv1d.argtypes = [PF, PF, PF, PF, c_int, c_int, c_int, c_int, c_int, c_int]
        L.wo_argmax.argtypes = [PF, c_int]
        L.wo_argmax.restype = c_int
        L.wo_weight_count.argtypes = [POINTER(_Cfg)]
        L.wo_weight_count.restype = c_longlong
        L.wo_create.argtypes = [POINTER(_Cfg), PF, c_longlong]
        L.wo_create.restype = c_void_p
        L.wo_destroy.argtypes = [c_void_p]
        L.wo_kvcache_create.argtypes = [c_void_p, c_int]
        L.wo_kvcache_create.restype = c_void_p
        L.wo_kvcache_destroy.argtypes = [c_void_p]
        L.wo_kvcache_len.argtypes = [c_void_p]
        L.wo_kvcache_len.restype = c_int
        L.wo_kvcache_ptr.argtypes = [c_void_p, c_int, c_int]
        L.wo_kvcache_ptr.restype = PF
        L.wo_encode_taps.argtypes = [c_void_p, PF, PF, PF, PF]
        L.wo_encode.argtypes = [c_void_p, PF, PF]
        L.wo_decoder_forward.argtypes = [c_void_p, c_void_p, PI, c_int, PF, c_int, c_int, PF, PF]
        L.wo_greedy.argtypes = [c_void_p, PF, PI, c_int, c_int, PI, c_int, PF]
        L.wo_greedy.restype = c_int
        L.wo_transcribe.argtypes = [c_void_p, PF, PI, c_int, c_int, PI, c_int]
        L.wo_transcribe.restype = c_int
        L.wo_teacher_forced.argtypes = [c_void_p, PF, PI, c_int, c_int, PF]
        L.wo_num_threads.restype = c_int
        L.wo_set_num_threads.argtypes = [c_int]
        _lib = L
    return _lib


# ---- op-level -------------------------------------------------------------------------------

def matmul(A: np.ndarray, B: np.ndarray, bias=None) -> np.ndarray:
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    M, K = A.shape
    N = B.shape[0]
    C = np.empty((M, N), np.float32)
    b = None if bias is None else _fp(np.ascontiguousarray(bias, np.float32))
    lib().wo_matmul(_fp(C), _fp(A), _fp(B), b, M, N, K)
    return C


def layer_norm(x, gamma, beta, eps=1e-5):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().wo_layer_norm(_fp(out), _fp(x), _fp(np.ascontiguousarray(gamma, np.float32)),
                        _fp(np.ascontiguousarray(beta, np.float32)), x.shape[0], x.shape[1], eps)
    return out


def gelu(x):
    y = np.array(x, np.float32, copy=True, order="C")
    lib().wo_gelu(_fp(y), y.size)
    return y


def softmax(x):
    y = np.array(x, np.float32, copy=True, order="C")
    lib().wo_softmax(_fp(y), y.shape[0], y.shape[1])
    return y


def transpose_conv_weights(w, C_out, C_in, K=3):
    w = np.ascontiguousarray(w, np.float32)
    out = np.empty((C_out * K, C_in), np.float32)
    lib().wo_transpose_conv_weights(_fp(out), _fp(w), C_out, C_in, K)
    return out


def conv1d(inp, weight_T, bias, stride, padding, out_T=False):
    inp = np.ascontiguousarray(inp, np.float32)
    C_in, L_in = inp.shape
    C_out = weight_T.shape[0] // 3
    L_out = (L_in + 2 * padding - 3) // stride + 1
    out = np.empty((L_out, C_out) if out_T else (C_out, L_out), np.float32)
    lib().wo_conv1d(_fp(out), _fp(inp), _fp(np.ascontiguousarray(weight_T, np.float32)),
                    _fp(np.ascontiguousarray(bias, np.float32)), C_in, L_in, C_out, stride, padding, int(out_T))
    return out


def argmax(x) -> int:
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    return int(lib().wo_argmax(_fp(x), x.size))


# ---- model-level ----------------------------------------------------------------------------

class OracleWhisper:
    """Mirrors Whisper (whisper.mojo:170-223) on the CPU oracle."""

    def __init__(self, cfg, flat_weights: np.ndarray):
        self.cfg = cfg
        self._w = np.ascontiguousarray(flat_weights, np.float32)  # must outlive the model
        c = _Cfg(cfg.d_model, cfg.n_heads, cfg.n_layers, cfg.vocab_size, cfg.n_audio_ctx, cfg.n_text_ctx, cfg.n_mels)
        self._m = lib().wo_create(ctypes.byref(c), _fp(self._w), self._w.size)
        if not self._m:
            raise ValueError("oracle: weight count does not match config (got %d)" % self._w.size)

    def __del__(self):
        if getattr(self, "_m", None):
            lib().wo_destroy(self._m)
            self._m = None

    def encode(self, mel: np.ndarray, taps: bool = False):
        cfg = self.cfg
        mel = np.ascontiguousarray(mel, np.float32)
        assert mel.shape == (cfg.n_mels, cfg.n_frames)
        out = np.empty((cfg.n_audio_ctx, cfg.d_model), np.float32)
        if not taps:
            lib().wo_encode(self._m, _fp(mel), _fp(out))
            return out
        c1 = np.empty((cfg.d_model, cfg.n_frames), np.float32)
        c2 = np.empty((cfg.n_audio_ctx, cfg.d_model), np.float32)
        lib().wo_encode_taps(self._m, _fp(mel), _fp(out), _fp(c1), _fp(c2))
        return out, c1, c2

    def greedy(self, enc_out: np.ndarray, pos_quirk=None, max_iters=None, margins: bool = False):
        cfg = self.cfg
        pos_quirk = cfg.pos_quirk if pos_quirk is None else pos_quirk
        max_iters = cfg.max_iters if max_iters is None else max_iters
        toks = np.zeros(5 + max_iters, np.int32)
        prompt = np.array(cfg.prompt, np.int32)
        mg = np.zeros(1 + max_iters, np.float32)
        n = lib().wo_greedy(self._m, _fp(np.ascontiguousarray(enc_out, np.float32)), _ip(toks), pos_quirk,
                            max_iters, _ip(prompt), cfg.eot, _fp(mg))
        return (toks[:n].copy(), mg[: n - 4].copy()) if margins else toks[:n].copy()

    def transcribe(self, mel: np.ndarray, pos_quirk=None, max_iters=None) -> np.ndarray:
        cfg = self.cfg
        pos_quirk = cfg.pos_quirk if pos_quirk is None else pos_quirk
        max_iters = cfg.max_iters if max_iters is None else max_iters
        toks = np.zeros(5 + max_iters, np.int32)
        prompt = np.array(cfg.prompt, np.int32)
        n = lib().wo_transcribe(self._m, _fp(np.ascontiguousarray(mel, np.float32)), _ip(toks), pos_quirk,
                                max_iters, _ip(prompt), cfg.eot)
        return toks[:n].copy()

    def decoder_forward_sequence(self, enc_out: np.ndarray, calls) -> np.ndarray:
        """WhisperDecoder.forward (whisper.mojo:130-167) called once per entry of `calls` = [(tokens, start_pos), ...]
        on one fresh KVCache; returns the logits of the LAST call's last position.  q_len > 1 takes the reference's
        block path with the causal fill (layers.mojo:273-342), q_len == 1 the cached decode path (:186-272)."""
        cfg = self.cfg
        enc = np.ascontiguousarray(enc_out, np.float32)
        cache = lib().wo_kvcache_create(self._m, cfg.n_text_ctx)
        logits = np.empty(cfg.vocab_size, np.float32)
        try:
            for toks, start_pos in calls:
                t = np.ascontiguousarray(toks, np.int32)
                lib().wo_decoder_forward(self._m, cache, _ip(t), t.size, _fp(enc), 1, int(start_pos), _fp(logits), None)
        finally:
            lib().wo_kvcache_destroy(cache)
        return logits

    def teacher_forced(self, enc_out: np.ndarray, forced: np.ndarray, pos_quirk=None) -> np.ndarray:
        """forced[0:4] is the prefill; returns logits [len(forced)-3, vocab]."""
        cfg = self.cfg
        pos_quirk = cfg.pos_quirk if pos_quirk is None else pos_quirk
        forced = np.ascontiguousarray(forced, np.int32)
        out = np.empty((forced.size - 3, cfg.vocab_size), np.float32)
        lib().wo_teacher_forced(self._m, _fp(np.ascontiguousarray(enc_out, np.float32)), _ip(forced), forced.size,
                                pos_quirk, _fp(out))
        return out


def num_threads() -> int:
    return int(lib().wo_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP threads of the oracle's parallel loops (the CPU arms of bench.py set this explicitly)."""
    lib().wo_set_num_threads(int(n))
