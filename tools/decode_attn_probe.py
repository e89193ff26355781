"""One launch shape of the decode self-attention kernel for ncu: B chunks, H heads, `len` cached rows (splits = 1, the
device-side length path of the decode step).   python tools/decode_attn_probe.py [B] [len]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gpu_util import debug_decode_attention
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
H, D = 6, 384
rng = np.random.default_rng(0)
q = rng.standard_normal((B, D), dtype=np.float32)
K = rng.standard_normal((B, L, D), dtype=np.float32)
V = rng.standard_normal((B, L, D), dtype=np.float32)
for _ in range(3):
    out = debug_decode_attention(q, K, V, H, 1)
print("ok", out.shape, float(np.abs(out).mean()), "algorithmic bytes per launch", 2 * B * L * D * 2 + 2 * B * D * 2)
