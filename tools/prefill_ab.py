import sys, numpy as np, torch
sys.path.insert(0, ".")
from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, synth
from bench import synth_pcm_gpu
cfg = WhisperConfig.tiny()
w = synth.make_weights(cfg, seed=0)
res = {}
for pf in (0, 1):
    m = Whisper(cfg); m.set_option("prefill_impl", pf); m.load(WeightLoader(data=w))
    for C in (256, 2048):
        pcm = synth_pcm_gpu(0, C, cfg.n_samples, torch.device("cuda"), 1)
        best = 1e9
        for _ in range(3):
            t, l = m.transcribe_pcm_batch(pcm); best = min(best, m.last_timing()["decode_ms"])
        res[(pf, C)] = (best, t.cpu().numpy() if hasattr(t, "cpu") else t)
        print("prefill_impl", pf, "chunks", C, "decode_ms", best, flush=True)
for C in (256, 2048):
    print("ids equal at", C, np.array_equal(res[(0, C)][1], res[(1, C)][1]))
