"""Audio ingest ahead of the log-mel frontend: what export_weights.py:98-114 does on the
host before it calls the feature extractor -- decode a wav file (the reference uses `soundfile`), mix to mono
(`audio_data.mean(axis=1)`, :102-103), resample to 16 kHz with the Fourier method (`scipy.signal.resample`,
:106-110) -- generalised from "first 30 s of one clip" (:112-114) to consecutive zero-padded 30 s chunks that
feed `Whisper.transcribe_pcm_batch`.

Host-side numpy only (no scipy / soundfile dependency at run time); the GPU work starts at the frontend.
"""
from __future__ import annotations

import struct
from typing import List, Tuple

import numpy as np

SAMPLE_RATE = 16000
CHUNK_SECONDS = 30

__all__ = ["load_wav", "to_mono", "resample_fourier", "prepare_audio", "chunk_audio", "transcribe_audio"]


def load_wav(path: str) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE reader: PCM 8/16/24/32-bit and IEEE float 32/64, any channel count (also
    WAVE_FORMAT_EXTENSIBLE).  Returns (float64 [n] or [n, channels], sample rate), integer formats scaled
    to [-1, 1) as `soundfile.read` does (export_weights.py:99)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            if size < 16:
                raise ValueError(f"{path}: short fmt chunk")
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and size >= 26:  # WAVE_FORMAT_EXTENSIBLE: the real tag leads the sub-format GUID
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)  # chunks are word aligned
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if ch < 1:
        raise ValueError(f"{path}: no channels")
    bps = bits // 8
    n = len(pcm) // (bps * ch) * ch
    raw = pcm[:n * bps]
    if tag == 1:  # integer PCM
        if bits == 8:
            x = (np.frombuffer(raw, np.uint8).astype(np.float64) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, "<i2").astype(np.float64) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (v - ((v & 0x800000) << 1)).astype(np.float64) / 8388608.0
        elif bits == 32:
            x = np.frombuffer(raw, "<i4").astype(np.float64) / 2147483648.0
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3 and bits in (32, 64):  # IEEE float
        x = np.frombuffer(raw, "<f4" if bits == 32 else "<f8").astype(np.float64)
    else:
        raise ValueError(f"{path}: unsupported wav format tag {tag} / {bits} bits")
    return (x.reshape(-1, ch) if ch > 1 else x), int(sr)


def to_mono(x: np.ndarray) -> np.ndarray:
    """export_weights.py:102-103: average the channels."""
    x = np.asarray(x)
    return x.mean(axis=1) if x.ndim > 1 else x


def resample_fourier(x: np.ndarray, num: int) -> np.ndarray:
    """`scipy.signal.resample(x, num)` for a real 1-D signal (export_weights.py:106-110): truncate or
    zero-pad the rFFT spectrum to the new length, halve / double the shared Nyquist bin, inverse rFFT,
    rescale by num / len(x).  float64 in, float64 out."""
    x = np.asarray(x, np.float64)
    nx = x.shape[0]
    if num <= 0 or nx == 0:
        return np.zeros(max(num, 0), np.float64)
    X = np.fft.rfft(x)
    n = min(num, nx)
    nyq = n // 2 + 1
    Y = np.zeros(num // 2 + 1, X.dtype)
    Y[:nyq] = X[:nyq]
    if n % 2 == 0:
        if num < nx:  # downsampling: the kept Nyquist bin stands for +f and -f of the original
            Y[n // 2] *= 2.0
        elif nx < num:  # upsampling: the original Nyquist bin is split between +f and -f
            Y[n // 2] *= 0.5
    return np.fft.irfft(Y, num) * (float(num) / float(nx))


def prepare_audio(x: np.ndarray, sample_rate: int) -> np.ndarray:
    """mono -> 16 kHz -> float32, exactly the order of export_weights.py:102-110
    (`num_samples = int(len * 16000 / sr)`)."""
    x = to_mono(x)
    if sample_rate != SAMPLE_RATE:
        x = resample_fourier(x, int(len(x) * SAMPLE_RATE / sample_rate))
    return np.asarray(x, np.float32)


def chunk_audio(x: np.ndarray, n_samples: int = SAMPLE_RATE * CHUNK_SECONDS) -> np.ndarray:
    """16 kHz mono f32 [n] -> f32 [n_chunks, n_samples]: consecutive windows, the last one zero padded
    (the extractor pads to 30 s; the reference keeps only the first window, export_weights.py:112-114)."""
    x = np.asarray(x, np.float32).reshape(-1)
    n_chunks = max(1, -(-x.shape[0] // n_samples))
    out = np.zeros((n_chunks, n_samples), np.float32)
    out.reshape(-1)[:x.shape[0]] = x
    return out


def transcribe_audio(model, x: np.ndarray, sample_rate: int) -> List[List[int]]:
    """Whole path for one recording: ingest -> 30 s chunks -> batched GPU transcription; one id list per chunk."""
    pcm = chunk_audio(prepare_audio(x, sample_rate), model.config.n_samples)
    toks, lens = model.transcribe_pcm_batch(pcm)
    return [[int(t) for t in toks[i, :lens[i]]] for i in range(pcm.shape[0])]
