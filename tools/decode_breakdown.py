"""Per-kernel-category device time of the decode loop, warm caches, eager launches with an event pair around
every kernel (profile_attn = 2).   python tools/decode_breakdown.py [chunks]"""
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
m.transcribe_pcm_batch(pcm)
print("graph run:", m.last_timing())
m.set_option("profile_attn", 2)
m.transcribe_pcm_batch(pcm)
tm = m.last_timing()
print("eager + events run:", tm)
tot = 0.0
for k in ["cross_attention", "self_attention", "gemm_qkv", "gemm_o", "gemm_cross_q", "gemm_cross_o", "gemm_fc1", "gemm_fc2",
          "layer_norm", "gemm_logits", "misc", "chain_first", "chain_b", "chain_ca"]:
    ms, n = m.last_kernel_timing(k)
    tot += ms
    print(f"{k:16s} {ms:8.2f} ms  {n:5d} launches  {1e3 * ms / max(n, 1):7.2f} us each")
print(f"sum of kernels {tot:.1f} ms of decode {tm['decode_ms']:.1f} ms")
