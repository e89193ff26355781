"""Batch-1 latency of `Whisper.transcribe(mel)` -- the reference's own published scenario (readme.md:82: 0.74 s for the
Mojo binary, 0.78 s for HF Python on the author's machine; main.mojo:29-31 times exactly this call) -- on one B200,
host log-mel in, host ids out, random-init Tiny weights (196 decoder forwards, EOT never fires).
    python tools/batch1_latency.py [option=value ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth
cfg = WhisperConfig.tiny()
m = Whisper(cfg)
for kv in sys.argv[1:]:  # options as key=value, e.g. pdl=1
    k, v = kv.split("=")
    m.set_option(k, int(v))
m.load(WeightLoader(data=synth.make_weights(cfg, seed=0)))
mel = synth.make_mel(1, cfg, 0)[0]
pcm = synth.make_audio(1, cfg, 0)
for name, fn in (("transcribe(mel [80, 3000])", lambda: m.transcribe(mel)), ("transcribe_pcm_batch(pcm [1, 480000])", lambda: m.transcribe_pcm_batch(pcm))):
    fn(); fn()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    tm = m.last_timing()
    print(f"{name}: median {1e3 * np.median(ts):.2f} ms, min {1e3 * min(ts):.2f} ms  ({30.0 / np.median(ts):.0f} audio-s/s at batch 1)  device phases {tm}")
