#!/usr/bin/env python
"""Timeline of the fused decode-step chain kernel (decode_chain.cu): run with WB_CHAIN_DBG=1; at model destruction the
library prints CTA 0's timestamps of the LAST chain launch (the last layer's cross-o -> LN -> fc1 -> fc2 -> LN chain).

    WB_CHAIN_DBG=1 python tools/chain_probe.py [chunks ...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth  # noqa: E402

cfg = WhisperConfig.tiny()
w = synth.make_weights(cfg, seed=0)
for n in [int(a) for a in sys.argv[1:]] or [2048, 256, 8]:
    m = Whisper(cfg)
    m.load(WeightLoader(data=w))
    mel = torch.from_numpy(synth.make_mel(8, cfg, 0)).cuda().repeat((n + 7) // 8, 1, 1)[:n].contiguous()
    for _ in range(2):
        m.transcribe_batch(mel)
    print(f"== {n} chunks: {m.last_timing()}", file=sys.stderr, flush=True)
    del m
