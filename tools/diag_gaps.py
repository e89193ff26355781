"""Where does wall time go beyond the event-timed phases?  python tools/diag_gaps.py [chunks]"""
import os, sys, time
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
for _ in range(3):
    m.transcribe_pcm_batch(pcm)
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter()
    m.transcribe_pcm_batch(pcm)
    t1 = time.perf_counter()
    tm = m.last_timing()
    print(f"C={C} wall {1e3*(t1-t0):.1f} ms  phases {tm['total_ms']:.1f} (fe {tm['frontend_ms']:.1f} enc {tm['encoder_ms']:.1f} dec {tm['decode_ms']:.1f})  gap {1e3*(t1-t0)-tm['total_ms']:.1f}")
t0 = time.perf_counter(); free, total = torch.cuda.mem_get_info(); t1 = time.perf_counter()
print(f"mem_get_info {1e3*(t1-t0):.2f} ms")
x = torch.empty(2_500_000_000, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); x.zero_(); torch.cuda.synchronize(); print(f"memset 2.5 GB {1e3*(time.perf_counter()-t0):.2f} ms")
