"""CPU restatement of the log-mel frontend the reference computes offline.

TEST INFRASTRUCTURE ONLY (same rule as whisper_oracle.c).

The reference obtains its [80, 3000] input from HF `WhisperProcessor(...)(audio, sampling_rate=16000,
return_tensors="pt").input_features` (export_weights.py:116, fallback :149).  The arithmetic lives in
the third-party `transformers` package (pinned 4.57.3 in /root/reference/uv.lock:2525-2526; class
WhisperFeatureExtractor, `_torch_extract_fbank_features`), not under /root/reference, so it is
restated here from its published recipe (which is OpenAI Whisper's `log_mel_spectrogram`):

  pad / truncate to 480 000 samples -> reflect-pad 200 each side (torch.stft center=True) ->
  3001 frames of 400 at hop 160 -> periodic Hann(400) -> rDFT (201 bins) -> drop the last frame ->
  |.|^2 -> slaney mel filterbank (80 x 201, 0..8000 Hz, slaney area norm) -> log10(max(., 1e-10)) ->
  max(., chunk_max - 8) -> (. + 4) / 4.

Pinned by tests/golden/logmel_*.npz, produced by oracle/make_golden.py from the transformers
package installed in this image (5.5.0; 4.57.3 is not installable offline -- stated version drift).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_SAMPLES = 480000


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filters(n_mels: int = 80, n_fft: int = N_FFT, sr: int = SAMPLE_RATE) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular filterbank, float32 [n_mels, n_fft//2+1]
    (the transpose of HF's `mel_filters` attribute), built in float64 then cast."""
    n_freqs = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sr / 2.0, n_freqs)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    f_pts = _mel_to_hz(mel_pts)
    fdiff = np.diff(f_pts)
    ramps = f_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def pad_or_trim(audio: np.ndarray, n_samples: int = N_SAMPLES) -> np.ndarray:
    audio = np.asarray(audio, np.float32)
    if audio.shape[-1] >= n_samples:
        return audio[..., :n_samples]
    pad = [(0, 0)] * (audio.ndim - 1) + [(0, n_samples - audio.shape[-1])]
    return np.pad(audio, pad)


def log_mel(audio: np.ndarray, n_mels: int = 80, n_samples: int = N_SAMPLES) -> np.ndarray:
    """audio f32 [n] or [B, n] -> f32 [B?, n_mels, n_samples // HOP]."""
    x = pad_or_trim(audio, n_samples)
    squeeze = x.ndim == 1
    if squeeze:
        x = x[None]
    n_frames = n_samples // HOP
    win = hann_periodic().astype(np.float64)
    filt = mel_filters(n_mels).astype(np.float64)
    out = np.empty((x.shape[0], n_mels, n_frames), np.float32)
    for b in range(x.shape[0]):
        xp = np.pad(x[b].astype(np.float64), (N_FFT // 2, N_FFT // 2), mode="reflect")
        idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]  # last (3001st) frame dropped
        frames = xp[idx] * win[None, :]
        power = np.abs(np.fft.rfft(frames, axis=1)) ** 2  # [frames, 201]
        mel = filt @ power.T  # [n_mels, frames]
        logs = np.log10(np.maximum(mel.astype(np.float32), np.float32(1e-10)))
        logs = np.maximum(logs, logs.max() - np.float32(8.0))
        out[b] = (logs + np.float32(4.0)) / np.float32(4.0)
    return out[0] if squeeze else out
