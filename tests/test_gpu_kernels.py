"""GPU: the fast path's kernels in isolation -- tcgen05/TMA GEMM (all epilogues, conv taps) against
the CUDA-core kernel and fp64 numpy on bf16-rounded operands; decode attention; log-mel frontend."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gpu_util import (h16_round, debug_cross_attention_absorbed, debug_decode_attention, debug_encoder_attention,
                      debug_gemm)
from oracle import logmel_oracle as LM
from whisper_mojo_b200 import Whisper, WhisperConfig, synth

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(1)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1, 1000, 128), (4, 3241, 384), (200, 1152, 384), (300, 384, 1536), (129, 130, 192)])
def test_gemm_tc_all_epilogues(M, N, K):
    A = rng.standard_normal((M, K), dtype=np.float32)
    W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
    b = rng.standard_normal(N, dtype=np.float32)
    ref = h16_round(A).astype(np.float64) @ h16_round(W).astype(np.float64).T + b
    tol = 2e-5 * max(1.0, K / 64)
    for impl in (0, 1):
        assert np.abs(debug_gemm(impl, A, W, b, 3) - ref).max() <= tol  # fp32 store: bit-level bf16 x bf16 products
    assert np.array_equal(debug_gemm(1, A, W, b, 0), h16_round(debug_gemm(1, A, W, b, 3)))  # bf16 store = rounded fp32
    g = torch.nn.functional.gelu(torch.from_numpy(ref), approximate="tanh").numpy()
    assert np.abs(debug_gemm(1, A, W, b, 1) - g).max() <= 2e-2  # bf16 output, |g| <~ 5
    x0 = rng.standard_normal((M, N), dtype=np.float32)
    assert np.abs(debug_gemm(1, A, W, b, 2, out0=x0) - (x0 + ref)).max() <= tol
    lg = h16_round(A).astype(np.float64) @ h16_round(W).astype(np.float64).T
    for impl in (0, 1):
        am = debug_gemm(impl, A, W, None, 4)
        assert np.array_equal(am[:, -1].astype(np.int64), lg.argmax(1))
        assert np.abs(am[:, :-1] - lg[:, :-1]).max() <= tol


def test_gemm_argmax_ties_pick_lowest_index():
    A = np.zeros((3, 64), np.float32)
    A[:, 0] = 1.0
    W = np.zeros((700, 64), np.float32)
    W[[5, 300, 650], 0] = 2.0  # equal maxima in three different 128-column tiles
    W[[130, 131], 0] = 2.0
    am = debug_gemm(1, A, W, None, 4)
    assert np.all(am[:, -1] == 5)


@pytest.mark.parametrize("cs,Cin,L,N,B", [(1, 128, 300, 128, 2), (2, 128, 301, 256, 3), (2, 384, 3000, 384, 2), (1, 128, 3000, 384, 1)])
def test_gemm_conv_taps(cs, Cin, L, N, B):
    A = rng.standard_normal((B, L, Cin), dtype=np.float32)
    W = rng.standard_normal((N, 3 * Cin), dtype=np.float32) / np.sqrt(3 * Cin)
    Lo = (L + 2 - 3) // cs + 1
    Ab, Wb = h16_round(A).astype(np.float64), h16_round(W).astype(np.float64)
    ref = np.zeros((B, Lo, N))
    for t in range(3):
        rows = np.arange(Lo) * cs + t - 1
        ok = (rows >= 0) & (rows < L)
        ref[:, ok] += Ab[:, rows[ok]] @ Wb[:, t * Cin:(t + 1) * Cin].T  # zero padding outside the chunk
    for impl in (0, 1):
        out = debug_gemm(impl, A, W, None, 3, batches=B, taps=3, conv_stride=cs, pad=1, rows_per_batch=Lo)
        assert np.abs(out.reshape(B, Lo, N) - ref).max() <= 1e-4, impl


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 1152, 384), (700, 384, 1536), (129, 130, 192), (1, 1000, 128),
                                   (2500, 1536, 384), (40000, 384, 384)])
def test_gemm_cta_pair_kernel_all_epilogues(M, N, K):
    """cta_group::2 kernel (impl 3 forces it): 256 x {128,192,256} pair tiles, ragged M / N, many tiles per pair."""
    A = rng.standard_normal((M, K), dtype=np.float32)
    W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
    b = rng.standard_normal(N, dtype=np.float32)
    ref = h16_round(A).astype(np.float64) @ h16_round(W).astype(np.float64).T + b
    tol = 2e-5 * max(1.0, K / 64)
    f32 = debug_gemm(3, A, W, b, 3)
    assert np.abs(f32 - ref).max() <= tol
    assert np.array_equal(f32, debug_gemm(2, A, W, b, 3))  # same k order as the single-CTA kernel: bit-identical
    assert np.array_equal(debug_gemm(3, A, W, b, 0), h16_round(f32))
    g = torch.nn.functional.gelu(torch.from_numpy(ref), approximate="tanh").numpy()
    assert np.abs(debug_gemm(3, A, W, b, 1) - g).max() <= 2e-2
    x0 = rng.standard_normal((M, N), dtype=np.float32)
    assert np.abs(debug_gemm(3, A, W, b, 2, out0=x0) - (x0 + ref)).max() <= tol
    lg = h16_round(A).astype(np.float64) @ h16_round(W).astype(np.float64).T
    am = debug_gemm(3, A, W, None, 4)  # fused logits + argmax partials (first maximum wins)
    assert np.array_equal(am[:, -1].astype(np.int64), lg.argmax(1))
    assert np.abs(am[:, :-1] - lg[:, :-1]).max() <= tol


def test_gemm_cta_pair_argmax_ties_and_vocab_width():
    A = np.zeros((3, 64), np.float32)
    A[:, 0] = 1.0
    W = np.zeros((700, 64), np.float32)
    W[[5, 300, 650], 0] = 2.0  # equal maxima in three different tiles
    W[[130, 131], 0] = 2.0
    assert np.all(debug_gemm(3, A, W, None, 4)[:, -1] == 5)
    A = rng.standard_normal((300, 384), dtype=np.float32)
    W = rng.standard_normal((51865 // 8, 384), dtype=np.float32) / 20  # ragged last tile like the real vocabulary
    lg = h16_round(A).astype(np.float64) @ h16_round(W).astype(np.float64).T
    assert np.array_equal(debug_gemm(3, A, W, None, 4)[:, -1].astype(np.int64), lg.argmax(1))


@pytest.mark.parametrize("cs,Cin,L,N,B", [(1, 128, 300, 128, 2), (2, 128, 301, 256, 3), (2, 384, 3000, 384, 2), (1, 128, 3000, 384, 3)])
def test_gemm_cta_pair_conv_taps(cs, Cin, L, N, B):
    A = rng.standard_normal((B, L, Cin), dtype=np.float32)
    W = rng.standard_normal((N, 3 * Cin), dtype=np.float32) / np.sqrt(3 * Cin)
    Lo = (L + 2 - 3) // cs + 1
    Ab, Wb = h16_round(A).astype(np.float64), h16_round(W).astype(np.float64)
    ref = np.zeros((B, Lo, N))
    for t in range(3):
        rows = np.arange(Lo) * cs + t - 1
        ok = (rows >= 0) & (rows < L)
        ref[:, ok] += Ab[:, rows[ok]] @ Wb[:, t * Cin:(t + 1) * Cin].T
    out = debug_gemm(3, A, W, None, 3, batches=B, taps=3, conv_stride=cs, pad=1, rows_per_batch=Lo)
    assert np.abs(out.reshape(B, Lo, N) - ref).max() <= 1e-4


@pytest.mark.parametrize("B,H,ln,splits", [(3, 6, 1500, 1), (3, 6, 1500, 11), (2, 2, 96, 1), (2, 12, 200, 1), (1, 6, 1, 1),
                                           (2, 6, 2, 1), (2, 6, 3, 1), (2, 6, 17, 1), (2, 6, 19, 1), (1, 6, 1500, 7)])
def test_decode_attention(B, H, ln, splits):
    D = H * 64
    q = h16_round(rng.standard_normal((B, D), dtype=np.float32) * 1.5)
    K = h16_round(rng.standard_normal((B, ln, D), dtype=np.float32) * 1.5)
    V = h16_round(rng.standard_normal((B, ln, D), dtype=np.float32))
    out = debug_decode_attention(q, K, V, H, splits)
    qh, Kh, Vh = (x.astype(np.float64) for x in (q.reshape(B, H, 64), K.reshape(B, ln, H, 64), V.reshape(B, ln, H, 64)))
    s = np.einsum("bhd,bjhd->bhj", qh, Kh) * 0.125
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    ref = np.einsum("bhj,bjhd->bhd", p, Vh).reshape(B, D)
    assert np.abs(out - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max() / 2)  # bf16 output rounding


@pytest.mark.parametrize("B,S,H", [(2, 1500, 6), (1, 96, 2), (3, 128, 2), (2, 129, 2), (1, 300, 12), (2, 1, 2)])
def test_encoder_attention_tc_vs_fp64(B, S, H):
    """tcgen05 flash attention == softmax(q k^T / 8) v on the same bf16-rounded inputs; covers a ragged
    last key block (1500 = 11*128 + 92), a single block, an exact multiple and S = 1."""
    D = H * 64
    qkv = h16_round(rng.standard_normal((B * S, 3 * D), dtype=np.float32) * np.array([1.5] * (2 * D) + [1.0] * D, np.float32))
    x = qkv.reshape(B, S, 3, H, 64).astype(np.float64)
    s = np.einsum("bihd,bjhd->bhij", x[:, :, 0], x[:, :, 1]) * 0.125
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    ref = np.einsum("bhij,bjhd->bihd", p, x[:, :, 2]).reshape(B * S, D)
    for impl in (0, 1):
        out = debug_encoder_attention(impl, qkv, B, S, H)
        # P is rounded to bf16 before the PV product on the tensor-core path, outputs are bf16
        assert np.abs(out - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max()), (impl, np.abs(out - ref).max())


@pytest.mark.parametrize("B,S,D,H", [(3, 1500, 384, 6), (1, 96, 128, 2), (2, 128, 128, 2), (5, 129, 256, 4), (300, 200, 384, 6),
                                     (2, 1, 128, 2), (3, 1500, 768, 12), (2, 65, 512, 8), (160, 130, 640, 10), (1, 1, 768, 12),
                                     (170, 260, 768, 12), (151, 129, 512, 8)])
def test_cross_attention_absorbed_vs_fp64(B, S, D, H):
    """ctx[b][h] = sum_j softmax2_j(q'_h . enc_j) enc_j (base-2 softmax: log2(e)/8 is folded into q');
    covers the ragged last key block, a single block, more chunks than SMs, S = 1, the 64-key-block form (M = 64
    score accumulators, d_model 640) and the CTA-pair form of d_model 512 / 768 (two CTAs per chunk swap partial
    scores through distributed shared memory), there also with more chunks than clusters and several key blocks."""
    qp = h16_round(rng.standard_normal((B, H * D), dtype=np.float32) * 0.15)
    enc = h16_round(rng.standard_normal((B, S, D), dtype=np.float32))
    out = debug_cross_attention_absorbed(qp, enc, H)
    q = qp.reshape(B, H, D).astype(np.float64)
    e = enc.astype(np.float64)
    s = np.einsum("bhc,bjc->bhj", q, e)
    p = np.exp2(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    ref = np.einsum("bhj,bjc->bhc", p, e).reshape(B, H * D)
    assert np.abs(out - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()


@pytest.mark.parametrize("impl", [1, 0], ids=["tf32x3_tensor_core", "fp32_fma"])
def test_logmel_frontend_matches_hf_golden_and_oracle(impl):
    m = Whisper(WhisperConfig.tiny())
    m.set_option("frontend_impl", impl)
    g = np.load(os.path.join(GOLDEN, "logmel_hf.npz"))
    a = synth.make_audio(int(g["n_chunks"]), seed=int(g["seed"]))
    mel = m.log_mel(a)
    rngv = float(g["mel_max"].max() - g["mel_min"].min())
    # tolerance: north_star "log-mel within 1e-4 relative"; formula max|a-b| / (max(b) - min(b))
    assert np.abs(mel[:, :, :64] - g["mel_first_frames"]).max() / rngv <= 1e-4
    assert np.abs(mel[:, :, -64:] - g["mel_last_frames"]).max() / rngv <= 1e-4
    assert np.abs(mel.astype(np.float64).sum(axis=2) - g["mel_row_sums"]).max() / 3000 <= 1e-5
    ref = LM.log_mel(a)
    assert np.abs(mel - ref).max() / (ref.max() - ref.min()) <= 1e-4
    # edge cases: silence (everything at the 1e-10 floor), a full-scale click, 5 s audio zero padded
    edge = np.zeros((3, 480000), np.float32)
    edge[1, 1234] = 1.0
    edge[2, :80000] = np.random.default_rng(3).standard_normal(80000).astype(np.float32)
    me, re_ = m.log_mel(edge), LM.log_mel(edge)
    assert np.abs(me - re_).max() / max(re_.max() - re_.min(), 1.0) <= 1e-4
    assert np.all(me[0] == me[0, 0, 0]) and abs(me[0, 0, 0] - (-10 + 4) / 4) < 1e-6
    assert np.abs(me[2, :, :64] - g["short_first_frames"]).max() <= 1e-4


def test_logmel_tensor_core_frontend_ragged_frames_and_loud_tone():
    """Tensor-core frontend on a config whose 192 frames do not fill a 256-frame pair tile, and on a
    full-scale tone over a 1e-4 noise floor (worst case for the split-precision DFT: weak bins next to a
    dominant one, right at the -8 clamp)."""
    cfg = WhisperConfig.micro()
    m = Whisper(cfg)
    a = np.random.default_rng(5).standard_normal((3, cfg.n_samples)).astype(np.float32) * 0.3
    ref = LM.log_mel(a, n_samples=cfg.n_samples)
    assert np.abs(m.log_mel(a) - ref).max() / (ref.max() - ref.min()) <= 1e-4
    t = np.arange(480000) / 16000.0
    tone = (0.9 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * np.random.default_rng(0).standard_normal(480000)).astype(np.float32)
    mt = Whisper(WhisperConfig.tiny())
    ref = LM.log_mel(tone[None])
    got = mt.log_mel(tone[None])
    assert np.abs(got - ref).max() / (ref.max() - ref.min()) <= 1e-4
    mt.set_option("frontend_impl", 0)
    assert np.abs(mt.log_mel(tone[None]) - got).max() <= 2e-4  # the two device frontends agree


def test_cross_attention_cta_pair_is_deterministic():
    """The CTA-pair form swaps partial scores between two CTAs through distributed shared memory every block; a
    protocol race (a row read before it arrived, a buffer overwritten early) would show up as run-to-run differences.
    Several chunks per cluster, several blocks per chunk, five runs: identical bits."""
    r = np.random.default_rng(5)
    B, S, D, H = 222, 700, 768, 12
    qp = h16_round(r.standard_normal((B, H * D), dtype=np.float32) * 0.15)
    enc = h16_round(r.standard_normal((B, S, D), dtype=np.float32))
    first = debug_cross_attention_absorbed(qp, enc, H)
    assert np.isfinite(first).all()
    for _ in range(4):
        assert np.array_equal(debug_cross_attention_absorbed(qp, enc, H), first)
