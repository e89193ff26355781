"""Decode time of the fused chain kernels (decode_fused=1) against the kernel-per-op form (0) over wave sizes: the two
produce the same bits, so the library picks per wave size (decode_fused=2, `auto`).
    python tools/fused_ab.py [chunks ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pcm_gpu
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth
sizes = [int(a) for a in sys.argv[1:]] or [64, 256, 512, 1024, 1536, 2048]
cfg = WhisperConfig.tiny()
w = synth.make_weights(cfg, seed=0)
for C in sizes:
    pcm = synth_pcm_gpu(0, C, cfg.n_samples, torch.device("cuda"), 1)
    row = {}
    for f in (0, 1):
        m = Whisper(cfg)
        m.set_option("decode_fused", f)
        m.load(WeightLoader(data=w))
        best = 1e9
        for _ in range(3):
            m.transcribe_pcm_batch(pcm)
            best = min(best, m.last_timing()["decode_ms"])
        row[f] = best
        del m
    print(f"chunks {C:5d}: decode kernel-per-op {row[0]:8.2f} ms, fused chains {row[1]:8.2f} ms  ->  {'fused' if row[1] < row[0] else 'per-op'}", flush=True)
