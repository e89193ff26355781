"""Encoder sub-batch size sweep: device time of frontend + encoder for 2048 chunks as a function of `enc_batch`
(activations of a sub-batch vs the 126 MB L2).   python tools/enc_batch_sweep.py [chunks]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_pcm_gpu
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = WhisperConfig.tiny()
m = Whisper(cfg)
m.load(WeightLoader(data=synth.make_weights(cfg, seed=0)))
pcm = synth_pcm_gpu(0, C, cfg.n_samples, torch.device("cuda"), 1)
ref = None
for eb in (128, 16, 24, 32, 48, 64, 96, 192, 256, 128):
    m.set_option("enc_batch", eb)
    toks, lens = m.transcribe_pcm_batch(pcm)
    toks, lens = m.transcribe_pcm_batch(pcm)
    tm = m.last_timing()
    same = True if ref is None else bool((toks == ref).all())
    ref = toks if ref is None else ref
    print(f"enc_batch {eb:4d}: frontend {tm['frontend_ms']:7.2f} ms  encoder {tm['encoder_ms']:7.2f} ms  decode {tm['decode_ms']:7.2f} ms  ids identical {same}", flush=True)
