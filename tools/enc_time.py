"""Encoder-only device time (BASELINE configs[2]: 256 log-mel chunks).   [WB_ATTN_POLY=n] python tools/enc_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctypes import c_void_p
import torch
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, _lib, synth
cfg = WhisperConfig.tiny()
m = Whisper(cfg, stream=torch.cuda.current_stream().cuda_stream)
m.load(WeightLoader(data=synth.make_weights(cfg, seed=0)))
lib = _lib.load()
dev = torch.device("cuda")
mel = torch.from_numpy(synth.make_mel(8, cfg, 5)).to(dev).repeat(32, 1, 1).contiguous()
enc = torch.empty((256, cfg.n_audio_ctx, cfg.d_model), dtype=torch.float32, device=dev)
f = lambda: _lib.check(lib.wm_encode_dev(m._h, c_void_p(mel.data_ptr()), 256, c_void_p(enc.data_ptr())))
f(); f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f()
torch.cuda.synchronize()  # the library runs on its own stream
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"WB_ATTN_POLY={os.environ.get('WB_ATTN_POLY', 'default')}: encoder 256 chunks {ms:.3f} ms -> {256 * 36.937728e9 / (ms * 1e-3) / 1e12:.0f} TFLOP/s; checksum {float(enc[::37].abs().mean()):.6f}")
