"""Timestamp dump of cross_attn_absorbed_kernel's CTA 0 (WB_XA_DBG).   python tools/xa_dbg.py [B] [D]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["WB_XA_DBG"] = "1"
from gpu_util import debug_cross_attention_absorbed
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 384
H = D // 64
rng = np.random.default_rng(0)
qp = (rng.standard_normal((B, H * D), dtype=np.float32) * 0.15)
enc = rng.standard_normal((B, 1500, D), dtype=np.float32)
debug_cross_attention_absorbed(qp, enc, H)
