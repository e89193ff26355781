"""Seeded synthetic assets in the reference's wire formats.

The reference's real assets (whisper_tiny_weights.bin, sample_input.bin) are produced by
export_weights.py from the HF hub and a downloaded wav (export_weights.py:13-14,95-120) and are not
shipped, so tests and the bench use seeded stand-ins with the same shapes and byte layout:
a headerless little-endian fp32 stream in `WhisperConfig.weight_layout()` order.

numpy's `default_rng` (PCG64) is used so the bytes are identical on every machine with this image.
"""
from __future__ import annotations

import numpy as np

from .config import WhisperConfig, HOP


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """Whisper's fixed encoder positional embedding (what model.encoder.embed_positions holds)."""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def make_weights(cfg: WhisperConfig, seed: int = 0, lin_scale: float = 1.5, emb_std: float = 0.1) -> np.ndarray:
    """Random-init weights of the named shapes as one flat fp32 array in file order.

    Linear / conv weights ~ N(0, lin_scale / sqrt(fan_in)) so every block changes the residual
    stream by O(1) (a network that is close to the identity maps every token to itself through the
    tied embedding and makes greedy decoding a fixed point); biases ~ N(0, 0.02); LayerNorm
    gamma ~ 1 + N(0, 0.1), beta ~ N(0, 0.1) so the affine terms are exercised; encoder positions are
    Whisper's sinusoids; token / decoder-position embeddings ~ N(0, emb_std).
    """
    rng = np.random.default_rng(seed)
    parts = []
    for name, shape in cfg.weight_layout():
        n = int(np.prod(shape))
        if name == "enc.pos":
            w = sinusoids(shape[0], shape[1]).reshape(-1)
        elif name in ("dec.token_emb", "dec.pos"):
            w = rng.standard_normal(n, dtype=np.float32) * np.float32(emb_std)
        elif name.endswith("ln.w") or name.endswith("ln_post.w"):
            w = np.float32(1.0) + rng.standard_normal(n, dtype=np.float32) * np.float32(0.1)
        elif name.endswith("ln.b") or name.endswith("ln_post.b"):
            w = rng.standard_normal(n, dtype=np.float32) * np.float32(0.1)
        elif name.endswith(".b"):
            w = rng.standard_normal(n, dtype=np.float32) * np.float32(0.02)
        else:
            fan_in = int(np.prod(shape[1:]))
            w = rng.standard_normal(n, dtype=np.float32) * np.float32(lin_scale / np.sqrt(fan_in))
        parts.append(w.astype(np.float32, copy=False))
    flat = np.concatenate(parts)
    assert flat.size == cfg.weight_count()
    return flat


def make_weights_hf_init(cfg: WhisperConfig, seed: int = 0, std: float = 0.02) -> np.ndarray:
    """HF `_init_weights` statistics: linear / conv / embedding weights ~ N(0, 0.02), biases 0,
    LayerNorm 1 / 0, sinusoidal encoder positions.  Used for the encoder-only tolerance check: activations stay small,
    so this is the init under which north_star's 1e-2 bound is meant to be read."""
    rng = np.random.default_rng(seed)
    parts = []
    for name, shape in cfg.weight_layout():
        n = int(np.prod(shape))
        if name == "enc.pos":
            w = sinusoids(shape[0], shape[1]).reshape(-1)
        elif name.endswith("ln.w") or name.endswith("ln_post.w"):
            w = np.ones(n, np.float32)
        elif name.endswith(".b"):
            w = np.zeros(n, np.float32)
        else:
            w = rng.standard_normal(n, dtype=np.float32) * np.float32(std)
        parts.append(w)
    flat = np.concatenate(parts)
    assert flat.size == cfg.weight_count()
    return flat


def write_weights(path: str, flat: np.ndarray) -> None:
    flat.astype("<f4", copy=False).tofile(path)


def make_audio(n_chunks: int, cfg: WhisperConfig = WhisperConfig.tiny(), seed: int = 0) -> np.ndarray:
    """Synthetic 16 kHz PCM, f32 [n_chunks, n_samples]: 0.1*N(0,1) noise plus two sine sweeps and a
    quiet tail per chunk, so the -8 dB clamp of the log-mel is exercised."""
    rng = np.random.default_rng(seed + 1000003)
    n = cfg.n_samples
    t = np.arange(n, dtype=np.float64) / 16000.0
    out = np.empty((n_chunks, n), dtype=np.float32)
    for c in range(n_chunks):
        x = 0.1 * rng.standard_normal(n)
        f0, f1 = 100.0 + 50.0 * (c % 7), 3000.0 + 400.0 * (c % 5)
        T = max(t[-1], 1e-9)
        x += 0.5 * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / T * t * t))
        x += 0.25 * np.sin(2 * np.pi * (7000.0 - 100.0 * (c % 11)) * t)
        q = int(n * 0.8)
        x[q:] *= 1e-3  # quiet tail -> values below the clamp
        out[c] = x.astype(np.float32)
    return out


def make_mel(n_chunks: int, cfg: WhisperConfig = WhisperConfig.tiny(), seed: int = 0) -> np.ndarray:
    """Synthetic log-mel-like input, f32 [n_chunks, n_mels, n_frames] in the value range the
    frontend produces (roughly [-1, 1.5]) with smooth structure along time."""
    rng = np.random.default_rng(seed + 7)
    x = rng.standard_normal((n_chunks, cfg.n_mels, cfg.n_frames), dtype=np.float32)
    k = np.ones(9, dtype=np.float32) / 9.0
    x = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 2, x) * 2.0
    return np.clip(x, -1.0, 1.5).astype(np.float32)
