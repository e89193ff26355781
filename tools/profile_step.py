"""Short profiling workload for ncu: whole hot path at full batch but only a few greedy steps.
    python tools/profile_step.py [--chunks 2048] [--iters 8] [--reps 2]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import synth_pcm_gpu  # noqa: E402
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, _lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=2048)
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--lanes", type=int, default=1)
ap.add_argument("--enc-batch", type=int, default=0)
ap.add_argument("--config", default="tiny", choices=["tiny", "small"])
ap.add_argument("--cross-impl", type=int, default=-1)
a = ap.parse_args()
base = WhisperConfig.tiny() if a.config == "tiny" else WhisperConfig.small_shaped()
cfg = WhisperConfig(**{**base.__dict__, "max_iters": a.iters})
m = Whisper(cfg, stream=torch.cuda.current_stream().cuda_stream)
m.set_option("use_graph", a.graph)
m.set_option("decode_lanes", a.lanes)
if a.enc_batch:
    m.set_option("enc_batch", a.enc_batch)
if a.cross_impl >= 0:
    m.set_option("cross_impl", a.cross_impl)
m.load(WeightLoader(data=synth.make_weights(cfg, seed=0)))
pcm = synth_pcm_gpu(0, a.chunks, cfg.n_samples, torch.device("cuda"), 1234)
for r in range(a.reps):
    torch.cuda.synchronize()
    t = time.time()
    toks, lens = m.transcribe_pcm_batch(pcm)
    torch.cuda.synchronize()
    print(f"rep {r}: {time.time() - t:.3f}s timing {m.last_timing()} launches {_lib.launch_count()}", flush=True)
