"""CPU: audio ingest (wav decode, mono, Fourier resample, chunking: export_weights.py:98-114) against
scipy / the stdlib wave writer, and the checkpoint exporter (export_weights.py:11-92) as the inverse of the
oracle-side HF loader."""
import os
import struct
import wave

import numpy as np
import pytest

from whisper_mojo_b200 import WhisperConfig, audio, export, synth


@pytest.mark.parametrize("nx,num", [(44100, 16000), (48000, 16000), (8000, 16000), (1001, 364), (1000, 363), (999, 2000),
                                    (1000, 1000), (7, 3), (2, 5)])
def test_resample_matches_scipy(nx, num):
    from scipy.signal import resample

    x = np.random.default_rng(nx + num).standard_normal(nx)
    ref = resample(x, num)
    got = audio.resample_fourier(x, num)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def _write_wav_int(path, x, sr, width, ch):
    with wave.open(path, "wb") as w:
        w.setnchannels(ch), w.setsampwidth(width), w.setframerate(sr)
        w.writeframes(x.tobytes())


def test_wav_reader_formats(tmp_path):
    rng = np.random.default_rng(0)
    p = str(tmp_path / "a.wav")
    # 16-bit stereo
    x16 = rng.integers(-32768, 32767, size=(500, 2), dtype=np.int16)
    _write_wav_int(p, x16, 22050, 2, 2)
    y, sr = audio.load_wav(p)
    assert sr == 22050 and y.shape == (500, 2) and np.array_equal(y, x16.astype(np.float64) / 32768.0)
    assert np.allclose(audio.to_mono(y), y.mean(axis=1))
    # 8-bit mono (unsigned), 32-bit mono
    x8 = rng.integers(0, 255, size=300, dtype=np.uint8)
    _write_wav_int(p, x8, 8000, 1, 1)
    y, sr = audio.load_wav(p)
    assert sr == 8000 and np.array_equal(y, (x8.astype(np.float64) - 128) / 128)
    x32 = rng.integers(-2**31, 2**31 - 1, size=100, dtype=np.int32)
    _write_wav_int(p, x32, 16000, 4, 1)
    assert np.array_equal(audio.load_wav(p)[0], x32.astype(np.float64) / 2147483648.0)
    # 24-bit mono, written by hand
    v = rng.integers(-2**23, 2**23 - 1, size=64)
    raw = b"".join(struct.pack("<i", int(t))[:3] for t in v)
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 48000, 3, 24)
    open(p, "wb").write(hdr + b"data" + struct.pack("<I", len(raw)) + raw)
    assert np.array_equal(audio.load_wav(p)[0], v.astype(np.float64) / 8388608.0)
    # IEEE float32 with an odd-sized LIST chunk in front of the data
    xf = rng.standard_normal(50).astype("<f4")
    body = (b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32) + b"LIST" + struct.pack("<I", 3) + b"abc\0"
            + b"data" + struct.pack("<I", xf.nbytes) + xf.tobytes())
    open(p, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    y, sr = audio.load_wav(p)
    assert sr == 16000 and np.array_equal(y, xf.astype(np.float64))
    open(p, "wb").write(b"nope")
    with pytest.raises(ValueError):
        audio.load_wav(p)


def test_prepare_and_chunk():
    sr = 44100
    t = np.arange(int(sr * 1.5)) / sr
    x = np.stack([np.sin(2 * np.pi * 440 * t), np.sin(2 * np.pi * 440 * t)], axis=1)
    y = audio.prepare_audio(x, sr)
    assert y.dtype == np.float32 and len(y) == int(len(t) * 16000 / sr)
    t16 = np.arange(len(y)) / 16000.0
    assert np.abs(y[200:-200] - np.sin(2 * np.pi * 440 * t16)[200:-200]).max() < 5e-3  # same tone after resampling
    c = audio.chunk_audio(np.ones(480000 + 5, np.float32))
    assert c.shape == (2, 480000) and c[1, :5].sum() == 5 and c[1, 5:].sum() == 0
    assert audio.chunk_audio(np.zeros(0, np.float32)).shape == (1, 480000)


def test_exporter_is_inverse_of_hf_loader():
    from oracle.hf_crosscheck import build_hf

    cfg = WhisperConfig.micro()
    w = synth.make_weights(cfg, seed=3)
    hf = build_hf(cfg, w)
    flat = export.export_state_dict(hf.state_dict(), cfg.n_layers)
    assert flat.dtype == np.float32 and flat.shape == w.shape and np.array_equal(flat, w)
    cfg2 = export.config_from_hf(hf.config, prompt=cfg.prompt, eot=cfg.eot, max_iters=cfg.max_iters)
    assert (cfg2.d_model, cfg2.n_heads, cfg2.n_layers, cfg2.vocab_size, cfg2.n_audio_ctx, cfg2.n_text_ctx) == \
        (cfg.d_model, cfg.n_heads, cfg.n_layers, cfg.vocab_size, cfg.n_audio_ctx, cfg.n_text_ctx)
    order = export.tensor_order(4)
    assert len(order) == 167 and not any("k_proj.bias" in k or "proj_out" in k for k in order)


def test_vocab_writer_roundtrips_through_tokenizer(tmp_path):
    from whisper_mojo_b200 import Tokenizer

    p = str(tmp_path / "vocab.txt")
    export.write_vocab(p, {"Ġhello": 1, "<|en|>": 0, "a\nb": 2, "Ġworld": 3})
    tok = Tokenizer(p)
    assert tok.decode([0, 1, 3, 2]) == " hello worlda\nb"
