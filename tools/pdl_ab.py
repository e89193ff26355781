"""Programmatic dependent launch of the decode-step kernels (option pdl) on / off, per wave size.
    python tools/pdl_ab.py [chunks ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import synth_pcm_gpu
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth
sizes = [int(a) for a in sys.argv[1:]] or [1, 256, 2048]
cfg = WhisperConfig.tiny()
w = synth.make_weights(cfg, seed=0)
for C in sizes:
    pcm = synth_pcm_gpu(0, C, cfg.n_samples, torch.device("cuda"), 1)
    row, ids = {}, {}
    for pdl in (0, 1, 0, 1):
        m = Whisper(cfg)
        m.set_option("pdl", pdl)
        m.load(WeightLoader(data=w))
        best = 1e9
        for _ in range(3):
            t, l = m.transcribe_pcm_batch(pcm)
            best = min(best, m.last_timing()["decode_ms"])
        row.setdefault(pdl, []).append(round(best, 2))
        ids[pdl] = t.cpu().numpy() if hasattr(t, "cpu") else t
        del m
    print(f"chunks {C:5d}: decode pdl off {row[0]} ms, pdl on {row[1]} ms, ids equal {np.array_equal(ids[0], ids[1])}", flush=True)
