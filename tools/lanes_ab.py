"""Two decode lanes (half-batches on two streams inside one graph) against one, per wave size.
    [WB_FUSED_LANES=1] python tools/lanes_ab.py [chunks ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import synth_pcm_gpu
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth
sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
cfg = WhisperConfig.tiny()
w = synth.make_weights(cfg, seed=0)
for C in sizes:
    pcm = synth_pcm_gpu(0, C, cfg.n_samples, torch.device("cuda"), 1)
    row, ids = {}, {}
    for lanes in (1, 2, 1, 2):
        m = Whisper(cfg)
        m.set_option("decode_lanes", lanes)
        m.load(WeightLoader(data=w))
        best = 1e9
        for _ in range(3):
            t, l = m.transcribe_pcm_batch(pcm)
            best = min(best, m.last_timing()["decode_ms"])
        row.setdefault(lanes, []).append(best)
        ids[lanes] = t.cpu().numpy() if hasattr(t, "cpu") else t
        del m
    print(f"chunks {C:5d}: decode 1 lane {row[1]} ms, 2 lanes {row[2]} ms, ids equal {np.array_equal(ids[1], ids[2])}", flush=True)
