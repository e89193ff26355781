"""Host-side mirror of the reference's L2 layer (whisper_tensor.mojo): `Tensor` and the op set,
each a thin call into the C ABI (wt_* in include/whisper_b200.h).  Same names, argument order and
"empty tensor = absent" convention as the reference, so code written against whisper_tensor.mojo
reads the same here.  All storage is fp32 on the GPU.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int64, c_uint64, c_void_p

import numpy as np

from . import _lib


class Tensor:
    """whisper_tensor.mojo:10-69.  Tensor(rows, cols) owns zero-filled device storage;
    Tensor.view(...) is a non-owning window; Tensor(0, 0) means "absent"."""

    def __init__(self, rows: int, cols: int, _handle: int | None = None, _view: bool = False, _keep=None):
        self.rows, self.cols, self.size, self.is_view = int(rows), int(cols), int(rows) * int(cols), _view
        self._keep = _keep  # a view keeps its base alive
        if _handle is not None:
            self._h = _handle
        elif self.size == 0:
            self._h = 0
        else:
            h = c_uint64(0)
            _lib.check(_lib.load().wt_tensor_alloc(self.rows, self.cols, ctypes.byref(h)))
            self._h = h.value

    @staticmethod
    def view(base: "Tensor", rows: int, cols: int, offset: int = 0) -> "Tensor":
        """Tensor.view(data_ptr, rows, cols) (whisper_tensor.mojo:25-33); the pointer is (base, offset)."""
        h = c_uint64(0)
        _lib.check(_lib.load().wt_tensor_view(base._h, int(offset), int(rows), int(cols), ctypes.byref(h)))
        return Tensor(rows, cols, _handle=h.value, _view=True, _keep=base)

    @staticmethod
    def from_numpy(a: np.ndarray) -> "Tensor":
        a = np.ascontiguousarray(a, np.float32)
        a2 = a.reshape(1, -1) if a.ndim == 1 else a.reshape(a.shape[0], -1)
        t = Tensor(a2.shape[0], a2.shape[1])
        if t.size:
            _lib.check(_lib.load().wt_tensor_upload(t._h, 0, a2.ctypes.data_as(c_void_p), t.size))
        return t

    def numpy(self) -> np.ndarray:
        out = np.empty((self.rows, self.cols), np.float32)
        if self.size:
            _lib.check(_lib.load().wt_tensor_download(self._h, 0, out.ctypes.data_as(c_void_p), self.size))
        return out

    def copy(self) -> "Tensor":
        """__copyinit__ deep copy (whisper_tensor.mojo:35-44)."""
        t = Tensor(self.rows, self.cols)
        if self.size:
            _lib.check(_lib.load().wt_tensor_copy(t._h, 0, self._h, 0, self.size))
        return t

    def load(self, idx: int) -> float:
        v = np.empty(1, np.float32)
        _lib.check(_lib.load().wt_tensor_download(self._h, int(idx), v.ctypes.data_as(c_void_p), 1))
        return float(v[0])

    def store(self, idx: int, val: float) -> None:
        v = np.array([val], np.float32)
        _lib.check(_lib.load().wt_tensor_upload(self._h, int(idx), v.ctypes.data_as(c_void_p), 1))

    def get(self, r: int, c: int) -> float:
        return self.load(r * self.cols + c)

    def set(self, r: int, c: int, val: float) -> None:
        self.store(r * self.cols + c, val)

    @property
    def data_ptr(self) -> int:
        p = c_void_p(0)
        _lib.check(_lib.load().wt_tensor_data(self._h, ctypes.byref(p)))
        return p.value or 0

    def __del__(self):
        h, self._h = getattr(self, "_h", 0), 0
        if h:
            try:
                _lib.load().wt_tensor_free(h)
            except Exception:
                pass


def memcpy(dest: Tensor, dest_off: int, src: Tensor, src_off: int, count: int) -> None:
    _lib.check(_lib.load().wt_tensor_copy(dest._h, int(dest_off), src._h, int(src_off), int(count)))


def matmul(C: Tensor, A: Tensor, B: Tensor, bias: Tensor) -> None:
    """C = A @ B.T + bias (whisper_tensor.mojo:151-246; the MAX wrappers :74-146 share the contract)."""
    _lib.check(_lib.load().wt_matmul(C._h, A._h, B._h, bias._h if bias is not None else 0))


# The reference's statically shaped MAX wrappers are the same contraction (layers.mojo:120-123 falls
# back from one to the other); they are kept as aliases so call sites read like the reference.
def matmul_384x384(C, A, B, bias): matmul(C, A, B, bias)  # noqa: E704
def matmul_384x1536(C, A, B, bias): matmul(C, A, B, bias)  # noqa: E704
def matmul_1536x384(C, A, B, bias): matmul(C, A, B, bias)  # noqa: E704
def matmul_384xVocab(C, A, B): matmul(C, A, B, Tensor(0, 0))  # noqa: E704
def matmul_Q_K(C, A, B, bias): matmul(C, A, B, bias)  # noqa: E704
def matmul_S_V(C, A, B, bias): matmul(C, A, B, bias)  # noqa: E704


def layer_norm(out: Tensor, inp: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> None:
    _lib.check(_lib.load().wt_layer_norm(out._h, inp._h, gamma._h, beta._h, eps))


def gelu(t: Tensor) -> None:
    _lib.check(_lib.load().wt_gelu(t._h))


def softmax(t: Tensor) -> None:
    _lib.check(_lib.load().wt_softmax(t._h))


def transpose_conv_weights(w: Tensor, C_out: int, C_in: int, K: int) -> Tensor:
    h = c_uint64(0)
    _lib.check(_lib.load().wt_transpose_conv_weights(w._h, C_out, C_in, K, ctypes.byref(h)))
    return Tensor(C_out * K, C_in, _handle=h.value)


def conv1d(out: Tensor, inp: Tensor, weight: Tensor, bias: Tensor, stride: int, padding: int,
           out_T: bool = False) -> None:
    _lib.check(_lib.load().wt_conv1d(out._h, inp._h, weight._h, bias._h, stride, padding, int(out_T)))


def argmax(t: Tensor) -> int:
    idx = c_int64(0)
    _lib.check(_lib.load().wt_argmax(t._h, ctypes.byref(idx)))
    return int(idx.value)


def add(out: Tensor, a: Tensor, b: Tensor) -> None:
    _lib.check(_lib.load().wt_add(out._h, a._h, b._h))


def scale_mask(scores: Tensor, scale: float, mask: bool, base: int) -> None:
    _lib.check(_lib.load().wt_scale_mask(scores._h, scale, int(mask), int(base)))


def embed(out: Tensor, token_emb: Tensor, pos_emb: Tensor, tokens, start_pos: int) -> None:
    toks = np.ascontiguousarray(tokens, np.int32)
    _lib.check(_lib.load().wt_embed(out._h, token_emb._h, pos_emb._h, toks.ctypes.data_as(c_void_p), toks.size,
                                    int(start_pos)))


def transpose(out: Tensor, inp: Tensor) -> None:
    _lib.check(_lib.load().wt_transpose(out._h, inp._h))
