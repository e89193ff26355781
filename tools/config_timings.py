"""Device timings of the BASELINE.json parity configs that are not the bench line:
   configs[1] frontend only (1 h of audio = 120 chunks), configs[2] encoder only (256 log-mel chunks),
   configs[4] Small-shaped end to end.   python tools/config_timings.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes
from ctypes import c_void_p
import numpy as np
import torch
from bench import synth_pcm_gpu
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, _lib, synth

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    e0.record(st)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

cfg = WhisperConfig.tiny()
m = Whisper(cfg)
m.load(WeightLoader(data=synth.make_weights(cfg, seed=0)))
lib = _lib.load()
dev = torch.device("cuda")
pcm = synth_pcm_gpu(0, 120, cfg.n_samples, dev, 3)
mel = torch.empty((120, cfg.n_mels, cfg.n_frames), dtype=torch.float32, device=dev)
ms = timed(lambda: _lib.check(lib.wm_logmel_dev(m._h, c_void_p(pcm.data_ptr()), 120, c_void_p(mel.data_ptr()))))
print(f"configs[1] frontend, 1 h of 16 kHz audio (120 chunks) on one GPU: {ms:.3f} ms -> {3600 / (ms * 1e-3):.3e} audio-s/s")
mel256 = torch.from_numpy(synth.make_mel(8, cfg, 5)).to(dev).repeat(32, 1, 1).contiguous()
enc = torch.empty((256, cfg.n_audio_ctx, cfg.d_model), dtype=torch.float32, device=dev)
ms = timed(lambda: _lib.check(lib.wm_encode_dev(m._h, c_void_p(mel256.data_ptr()), 256, c_void_p(enc.data_ptr()))))
print(f"configs[2] encoder only, 256 chunks: {ms:.2f} ms -> {256 * 36.937728e9 / (ms * 1e-3) / 1e12:.0f} TFLOP/s bf16")
del m, pcm, mel, mel256, enc
scfg = WhisperConfig.small_shaped()
ms_ = Whisper(scfg)
ms_.load(WeightLoader(data=synth.make_weights(scfg, seed=1)))
C = 1024
pcm = synth_pcm_gpu(0, C, scfg.n_samples, dev, 4)
ms_.transcribe_pcm_batch(pcm)
t = timed(lambda: ms_.transcribe_pcm_batch(pcm), reps=2)
print(f"configs[4] Small-shaped (12 layers, d 768) end to end, {C} chunks on one GPU: {t:.0f} ms -> {C * 30 / (t * 1e-3):.0f} audio-s/s  {ms_.last_timing()}")
