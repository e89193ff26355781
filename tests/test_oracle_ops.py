"""CPU: the oracle's op-level routines against plain numpy/torch fp32 math and the edge cases the
reference's loops have (tails, M<=4 vs tiled path, mask fill, first-max argmax)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

rng = np.random.default_rng(0)


@pytest.mark.parametrize("M,N,K", [(1, 384, 384), (4, 51, 37), (5, 19, 70), (1500 // 10, 384, 64), (7, 64, 1500), (3, 8, 5)])
def test_matmul(M, N, K):
    A = rng.standard_normal((M, K), dtype=np.float32)
    B = rng.standard_normal((N, K), dtype=np.float32)
    b = rng.standard_normal(N, dtype=np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(O.matmul(A, B) - ref).max() <= 1e-3
    assert np.abs(O.matmul(A, B, b) - (ref + b)).max() <= 1e-3


def test_layer_norm_and_gelu_and_softmax():
    x = rng.standard_normal((7, 384), dtype=np.float32) * 3 + 1
    g = rng.standard_normal(384, dtype=np.float32)
    b = rng.standard_normal(384, dtype=np.float32)
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (384,), torch.from_numpy(g), torch.from_numpy(b), 1e-5)
    assert np.abs(O.layer_norm(x, g, b) - ref.numpy()).max() <= 1e-4
    y = rng.standard_normal(1536 * 3, dtype=np.float32) * 3
    ref = torch.nn.functional.gelu(torch.from_numpy(y), approximate="tanh").numpy()
    assert np.abs(O.gelu(y) - ref).max() <= 1e-5
    # gelu ignores the tail beyond size//8 vectors (whisper_tensor.mojo:308)
    z = np.full(11, 2.0, np.float32)
    out = O.gelu(z)
    assert np.all(out[8:] == 2.0) and np.all(out[:8] != 2.0)
    for cols in (4, 8, 13, 1500):
        s = rng.standard_normal((5, cols), dtype=np.float32) * 4
        s[0, -1] = -1e10
        ref = torch.softmax(torch.from_numpy(s), dim=1).numpy()
        assert np.abs(O.softmax(s) - ref).max() <= 1e-6


@pytest.mark.parametrize("C_in,L,C_out,stride,out_T", [(80, 64, 16, 1, False), (16, 64, 24, 2, True), (8, 7, 8, 2, True)])
def test_conv1d(C_in, L, C_out, stride, out_T):
    x = rng.standard_normal((C_in, L), dtype=np.float32)
    w = rng.standard_normal((C_out, C_in, 3), dtype=np.float32)
    b = rng.standard_normal(C_out, dtype=np.float32)
    wT = O.transpose_conv_weights(w, C_out, C_in)
    assert wT.shape == (C_out * 3, C_in) and wT[1 * 3 + 2, 5] == w[1, 5, 2]
    ref = torch.nn.functional.conv1d(torch.from_numpy(x)[None], torch.from_numpy(w), torch.from_numpy(b),
                                     stride=stride, padding=1)[0].numpy()
    out = O.conv1d(x, wT, b, stride, 1, out_T)
    assert np.abs((out.T if out_T else out) - ref).max() <= 1e-4


def test_argmax_first_max_wins():
    x = np.zeros(100, np.float32)
    x[[17, 40]] = 3.0
    assert O.argmax(x) == 17
    assert O.argmax(np.full(5, -1.0, np.float32)) == 0


def test_block_prefill_with_causal_fill_equals_cached_single_token_steps():
    """whisper.mojo:195-197 runs the 4 prompt ids as ONE decoder.forward through the block path, where keys j > i are
    filled with -1e10 before the softmax (layers.mojo:304-320): in fp32 exp(-1e10 - max) is exactly 0, so the result
    must equal feeding the ids one by one through the cached decode path (layers.mojo:186-272) up to summation order.
    This is the property the GPU prefill rests on (one q_len = 4 forward == four cached steps, tests/test_gpu_model.py)."""
    from whisper_mojo_b200 import WhisperConfig, synth

    cfg = WhisperConfig.micro()
    w = synth.make_weights(cfg, seed=3)
    om = O.OracleWhisper(cfg, w)
    enc = om.encode(synth.make_mel(1, cfg, 5)[0])
    p = list(cfg.prompt)
    block = om.decoder_forward_sequence(enc, [(p, 0)])
    steps = om.decoder_forward_sequence(enc, [([p[0]], 0), ([p[1]], 1), ([p[2]], 2), ([p[3]], 3)])
    mixed = om.decoder_forward_sequence(enc, [(p[:2], 0), (p[2:], 2)])  # a block call on a non-empty cache (past_len = 2)
    scale = float(np.abs(block).max())
    assert np.abs(block - steps).max() <= 2e-5 * max(scale, 1.0), np.abs(block - steps).max()
    assert np.abs(block - mixed).max() <= 2e-5 * max(scale, 1.0), np.abs(block - mixed).max()
    assert int(block.argmax()) == int(steps.argmax()) == int(mixed.argmax())
    # and the mask matters: without it (prompt order permuted) the last position's logits change
    other = om.decoder_forward_sequence(enc, [([p[1], p[0], p[2], p[3]], 0)])
    assert np.abs(block - other).max() > 1e-3 * max(scale, 1.0)
