"""Timestamp dump of two CTAs of the encoder attention kernel.  WB_EA_DBG=<linear id of 2nd CTA> python tools/ea_dbg.py [B]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("WB_EA_DBG", "148")
from gpu_util import debug_encoder_attention
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(0)
qkv = rng.standard_normal((B * 1500, 3 * 384), dtype=np.float32)
debug_encoder_attention(1, qkv, B, 1500, 6)
