"""GPU: op-level C ABI (wt_*) against the CPU oracle's restatement of the same reference routine."""
import numpy as np
import pytest

from oracle import oracle as O
from whisper_mojo_b200 import Tensor, _lib
from whisper_mojo_b200 import whisper_tensor as wt

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(0)


def T(a):
    return Tensor.from_numpy(a)


def test_tensor_semantics():
    t = Tensor(3, 5)
    assert np.array_equal(t.numpy(), np.zeros((3, 5), np.float32))  # zero filled (whisper_tensor.mojo:23)
    t.set(1, 2, 7.5)
    assert t.get(1, 2) == 7.5 and t.load(7) == 7.5
    v = Tensor.view(t, 1, 5, offset=5)
    assert v.is_view and v.get(0, 2) == 7.5
    v.store(0, -1.0)
    assert t.get(1, 0) == -1.0  # views alias
    c = t.copy()
    c.set(0, 0, 3.0)
    assert t.get(0, 0) == 0.0  # copies do not
    assert Tensor(0, 0).size == 0
    with pytest.raises(_lib.WhisperB200Error):
        Tensor.view(t, 4, 5)  # window outside the base


@pytest.mark.parametrize("M,N,K", [(1, 384, 384), (1, 51865, 384), (4, 51, 37), (5, 19, 70), (150, 384, 64), (7, 64, 1500), (1500, 384, 384)])
def test_matmul(M, N, K):
    A = rng.standard_normal((M, K), dtype=np.float32)
    B = rng.standard_normal((N, K), dtype=np.float32)
    b = rng.standard_normal(N, dtype=np.float32)
    for bias in (None, b):
        C = Tensor(M, N)
        wt.matmul(C, T(A), T(B), T(bias) if bias is not None else Tensor(0, 0))
        ref = O.matmul(A, B, bias)
        assert np.abs(C.numpy() - ref).max() <= 2e-4 * max(1.0, np.sqrt(K) / 8)  # fp32, summation order only
    with pytest.raises(_lib.WhisperB200Error):
        wt.matmul(Tensor(M, N + 1), T(A), T(B), Tensor(0, 0))


def test_layer_norm_gelu_softmax_add():
    x = rng.standard_normal((9, 384), dtype=np.float32) * 3 + 1
    g, b = rng.standard_normal(384, dtype=np.float32), rng.standard_normal(384, dtype=np.float32)
    out = Tensor(9, 384)
    wt.layer_norm(out, T(x), T(g), T(b))
    assert np.abs(out.numpy() - O.layer_norm(x, g, b)).max() <= 2e-5
    y = rng.standard_normal((3, 1536), dtype=np.float32) * 3
    t = T(y)
    wt.gelu(t)
    assert np.abs(t.numpy() - O.gelu(y)).max() <= 2e-6
    for cols in (4, 13, 1500):
        s = rng.standard_normal((6, cols), dtype=np.float32) * 4
        s[0, -1] = -1e10
        t = T(s)
        wt.softmax(t)
        assert np.abs(t.numpy() - O.softmax(s)).max() <= 1e-6
    a, c = rng.standard_normal((4, 8), dtype=np.float32), rng.standard_normal((4, 8), dtype=np.float32)
    o = Tensor(4, 8)
    wt.add(o, T(a), T(c))
    assert np.array_equal(o.numpy(), a + c)


@pytest.mark.parametrize("C_in,L,C_out,stride,out_T", [(80, 3000, 384, 1, False), (384, 3000, 384, 2, True), (8, 7, 8, 2, True)])
def test_conv1d_and_weight_transpose(C_in, L, C_out, stride, out_T):
    x = rng.standard_normal((C_in, L), dtype=np.float32)
    w = rng.standard_normal((C_out, C_in * 3), dtype=np.float32) / np.sqrt(3 * C_in)
    b = rng.standard_normal(C_out, dtype=np.float32)
    wT = wt.transpose_conv_weights(T(w), C_out, C_in, 3)
    ref_wT = O.transpose_conv_weights(w, C_out, C_in)
    assert np.array_equal(wT.numpy(), ref_wT)
    L_out = (L + 2 - 3) // stride + 1
    out = Tensor(L_out, C_out) if out_T else Tensor(C_out, L_out)
    wt.conv1d(out, T(x), wT, T(b), stride, 1, out_T)
    assert np.abs(out.numpy() - O.conv1d(x, ref_wT, b, stride, 1, out_T)).max() <= 1e-4


def test_argmax_first_max_and_scale_mask_embed_transpose():
    x = np.zeros((1, 51865), np.float32)
    x[0, [17, 40000]] = 3.0
    assert wt.argmax(T(x)) == 17
    assert wt.argmax(T(np.full((1, 5), -1.0, np.float32))) == 0
    big = rng.standard_normal((1, 51865), dtype=np.float32)
    assert wt.argmax(T(big)) == O.argmax(big)
    s = rng.standard_normal((4, 4), dtype=np.float32)
    t = T(s)
    wt.scale_mask(t, 0.125, True, 0)
    ref = s * np.float32(0.125)
    ref[np.triu_indices(4, 1)] = -1e10
    assert np.array_equal(t.numpy(), ref)
    te, pe = rng.standard_normal((50, 128), dtype=np.float32), rng.standard_normal((16, 128), dtype=np.float32)
    o = Tensor(3, 128)
    wt.embed(o, T(te), T(pe), [7, 0, 49], 5)
    assert np.array_equal(o.numpy(), te[[7, 0, 49]] + pe[5:8])
    with pytest.raises(_lib.WhisperB200Error):
        wt.embed(o, T(te), T(pe), [7, 0, 50], 5)
    m = rng.standard_normal((37, 70), dtype=np.float32)
    o = Tensor(70, 37)
    wt.transpose(o, T(m))
    assert np.array_equal(o.numpy(), m.T)
