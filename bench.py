#!/usr/bin/env python
"""Benchmark of the Whisper-Tiny batched greedy transcription path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chunks C] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the whole hot path (log-mel frontend -> conv stem -> encoder -> cross-K/V ->
prefill + 195 greedy steps with fused logits+argmax) over ONE GLOBAL seeded batch of C = 2048 synthetic
30 s chunks (BASELINE.json configs[3]; random-init Whisper-Tiny weights in the reference's file format).
Chunks are independent, so the batch is sharded over the ranks with `dist.shard_range` (chunk i -> rank
i*G // C) and no data-path collective; only the final gather of token ids crosses NVLink (NCCL
all_gather), inside the timed region.  Default `--scaling strong`: C chunks in total (2048 / G per GPU);
`--scaling weak` keeps C chunks PER GPU (round 1's line).  Chunk i's audio depends only on i, so the gathered
ids must be identical for every G: rank 0 prints their sha256 (`ids_sha256`) and re-runs 8 chunks spread over
all shards alone on its own GPU to check it (`parity.cross_g`), and checks 4 chunks against the CPU oracle
(`parity.oracle`) -- all outside the timed region.

  value  audio-seconds per second, whole job, PCM already resident in HBM, CUDA events on the stream
  e2e    same through the public host API (Whisper.transcribe_pcm_batch) with pinned HOST pcm:
         host->device copy of the pcm and device->host read of the tokens inside the timed region
  roofline      decode cross-attention kernel (dominant): algorithmic K/V bytes / event-timed duration
  cpu_baseline  oracle/ (CPU restatement of the reference) timed on this box's host cores, rank 0, N=1

`--impl reference` times the reference's own algorithm on the host CPU (the Mojo binary cannot run
here: Mach-O arm64, no Mojo toolchain -> the C restatement in oracle/ is used, kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio_sec_per_sec_whisper_tiny_batched_greedy"
WORKLOAD = "whisper-tiny batched greedy transcription (BASELINE.json configs[3]): %d synthetic 30 s chunks sharded over the GPUs"
UNIT = "audio-s/s"
CHUNK_SECONDS = 30.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunks", type=int, default=2048, help="30 s chunks per step: in total (strong) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong = --chunks in total, sharded over the ranks (BASELINE configs[3]); weak = --chunks per GPU")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run id checks (cross-G subset, CPU oracle)")
    ap.add_argument("--parity-chunks", type=int, default=4, help="chunks checked against the CPU oracle after the timed run")
    ap.add_argument("--no-bf16", action="store_true", help="skip the bf16-build comparison run (N=1 only)")
    ap.add_argument("--eot-schedule", default="40:195", help="lo:hi -- one extra untimed pass pair with declared chunk lengths ~U(lo, hi) "
                    "(EOT never fires with random weights): decode time with / without skipping finished chunks; '' = off")
    ap.add_argument("--breakdown", action="store_true", help="add a per-kernel-category decode breakdown (one extra eager pass)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="tiny", choices=["tiny", "small"],
                    help="tiny = BASELINE.json configs[3] (the bench line); small = configs[4], Small-shaped 12 layers d 768")
    ap.add_argument("--cpu-chunks", type=int, default=16, help="chunks per CPU-baseline sample (about 10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hf-baseline", action="store_true", help="skip the HF model.generate CPU baseline (benchmark_python.py analogue)")
    ap.add_argument("--lanes", type=int, default=1, help="decode lanes (2 = two half-batches on two streams)")
    ap.add_argument("--sampler", default="nvml", choices=["nvml", "smi", "none"], help="clock sampler during the timed region")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on host cores (oracle port)
# ---------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Host threads the CPU arms may use: every core this process is allowed on."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def force_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arms must use all host cores whatever
    the launcher set (round 1: the reference arm timed out at N = 2, 4, 8 for this reason).  Called before libgomp is
    loaded (oracle .so / torch import), and the oracle's thread count is also set explicitly after loading."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def cpu_transcribe_rate(n_chunks: int, repeats: int = 1, small: bool = False, budget_s: float = 0.0):
    """audio-s/s of the CPU restatement (all host threads via OpenMP), batch 1 like the reference.
    budget_s > 0 bounds the sample: the chunk count is cut so that the timed work stays near the budget."""
    from oracle import oracle as O
    from whisper_mojo_b200 import WhisperConfig, synth

    O.set_num_threads(host_threads())
    cfg = WhisperConfig.small_shaped() if small else WhisperConfig.tiny()
    w = synth.make_weights(cfg, seed=1 if small else 0)
    mel = synth.make_mel(n_chunks, cfg, 0)
    om = O.OracleWhisper(cfg, w)
    t0 = time.perf_counter()
    om.transcribe(mel[0])  # warm-up (mirrors benchmark_python.py:25-26)
    t1 = time.perf_counter() - t0
    if budget_s > 0:
        n_chunks = max(1, min(n_chunks, int(budget_s / max(t1, 1e-3))))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        n_tok = 0
        for i in range(n_chunks):
            n_tok += len(om.transcribe(mel[i]))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_chunks * CHUNK_SECONDS / best, best, O.num_threads(), n_tok, n_chunks


def hf_generate_baseline():
    """benchmark_python.py analogue (CPU baseline 3): HF `model.generate` on the host cores, one chunk."""
    try:
        import torch

        torch.set_num_threads(host_threads())
        from oracle import hf_crosscheck as H
        from whisper_mojo_b200 import WhisperConfig, synth

        cfg = WhisperConfig.tiny()
        dt, threads = H.hf_generate_rate(cfg, synth.make_weights(cfg, seed=0), synth.make_mel(1, cfg, 0)[0], 4 + cfg.max_iters - 3)
        return {"value": CHUNK_SECONDS / dt, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "HF transformers WhisperForConditionalGeneration.generate (the reference's benchmark_python.py:25-30: "
                          f"1 warm-up + 1 timed call), same random weights, 1 chunk, 196 decoder forwards, {dt:.2f} s"}
    except Exception as ex:
        return {"error": str(ex)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    force_host_threads()
    hf = None if args.no_hf_baseline else hf_generate_baseline()  # before the OpenMP port: its spinning worker threads would slow torch's down
    # each step = a bounded sample of the workload: at most --cpu-chunks chunks, cut so that the whole run
    # (warm-up + K steps) stays near two minutes of CPU work whatever K the driver passes
    steps = max(args.steps, 1)
    budget = 120.0 / steps
    times, cores, n_used = [], 1, args.cpu_chunks
    for _ in range(steps):
        _, dt, cores, _, n_used = cpu_transcribe_rate(args.cpu_chunks, budget_s=budget)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_used * CHUNK_SECONDS / (ms / 1e3)
    sample = (f"{n_used} synthetic 30 s chunks per step, batch 1, precomputed log-mel "
              f"(the reference's timed region, main.mojo:29-31), encoder + prefill + 195 greedy steps, {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.chunks,
                   "sample_chunks_per_step": n_used,
                   "reference_arm": "CPU restatement of whisper.Mojo (oracle/whisper_oracle.c); the shipped ./main is "
                                    "Mach-O arm64 and no Mojo toolchain exists in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline_hf_generate": hf,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed region, in-process through NVML
    (nvidia_ml_py) from a background thread; `nvidia-smi -lms` in a subprocess is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, mode: str = "nvml", period_s: float = 0.25):
        self.idx, self.rows, self.proc, self.mode, self.period = gpu_index, [], None, mode, period_s
        self.stop_flag = threading.Event()
        self.thread = None

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                import pynvml as N

                N.nvmlInit()
                h = N.nvmlDeviceGetHandleByIndex(self.idx)
                mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
                bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown,
                        "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                        "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown,
                        "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}

                def loop():
                    while not self.stop_flag.is_set():
                        try:
                            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                            r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                            pw = N.nvmlDeviceGetPowerUsage(h) / 1e3
                            self.rows.append((float(sm), float(mx), pw, [k for k, b in bits.items() if r & b]))
                        except Exception:
                            pass
                        self.stop_flag.wait(self.period)

                self.thread = threading.Thread(target=loop, daemon=True)
                self.thread.start()
                return
            except Exception:
                self.mode = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(int(self.period * 1e3)), "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((float(r[1]), float(r[2]), float(r[3]),
                                  [n for n, v in zip(names, r[5:9]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        pw = [r[2] for r in self.rows]
        reasons = sorted({x for r in self.rows for x in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "sm_mhz_min": min(sm) if sm else None,
                "power_w_max": max(pw) if pw else None, "sampler": self.mode}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
PCM_BLOCK = 64  # chunks per generation block: chunk i's audio depends only on (seed, i // 64, i % 64)


def synth_pcm_gpu(lo, hi, n_samples, device, seed):
    """Synthetic 16 kHz audio for the GLOBAL chunk indices [lo, hi), made on the device: 0.1*N(0,1) + a sine sweep +
    a high tone, quiet tail.  Generated in blocks of 64 chunks seeded by the block index, so a chunk's samples do not
    depend on which rank (or batch) it lands in -- the precondition for comparing ids across GPU counts."""
    import torch

    out = torch.empty((hi - lo, n_samples), dtype=torch.float32, device=device)
    t = torch.arange(n_samples, device=device, dtype=torch.float32) / 16000.0
    T = float(n_samples) / 16000.0
    for blk in range(lo // PCM_BLOCK, (hi + PCM_BLOCK - 1) // PCM_BLOCK):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + blk)
        i0 = blk * PCM_BLOCK
        k = torch.arange(i0, i0 + PCM_BLOCK, device=device, dtype=torch.float32)[:, None]
        f0, f1 = 100.0 + 50.0 * (k % 7), 3000.0 + 400.0 * (k % 5)
        x = 0.1 * torch.randn((PCM_BLOCK, n_samples), generator=g, device=device)
        x += 0.5 * torch.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / T * t * t))
        x += 0.25 * torch.sin(2 * np.pi * (7000.0 - 100.0 * (k % 11)) * t)
        x[:, int(n_samples * 0.8):] *= 1e-3
        a, b = max(lo, i0), min(hi, i0 + PCM_BLOCK)
        out[a - lo:b - lo] = x[a - i0:b - i0]
    return out


def ncu_traffic(chunks):
    """dram__bytes_read.sum + dram__bytes_write.sum per cross-attention launch from the committed
    `ncu --set full` capture (profiles/), valid for the chunk count it was captured at (2048); else None."""
    for name in ("r02_cross_attn_absorbed_ncu_full.json", "r01_cross_attn_absorbed_ncu_full_v3.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            if chunks == 2048 and "2048" in d["command"]:
                return d["dram_traffic_bytes_per_launch_mean"]
        except Exception:
            pass
    return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def parity_checks(args, model, cfg, toks_all, lens_all, n_total, dev, seed):
    """Outside the timed region, rank 0: (a) cross-G / batch identity -- 8 global chunks spread over every shard are
    regenerated and transcribed alone on this GPU; their ids must equal the gathered ids of the timed run; (b) the
    first `--parity-chunks` of those against the fp32 CPU oracle (numpy log-mel -> C restatement) under the strict
    rule of tests/gpu_util.py: identical, or first mismatch at a step whose oracle top-2 margin is < 2e-2."""
    import torch

    out = {}
    picks = sorted({int(round(x)) for x in np.linspace(0, n_total - 1, 8)})
    pcm_sub = torch.cat([synth_pcm_gpu(i, i + 1, cfg.n_samples, dev, seed) for i in picks])
    t_sub, l_sub = model.transcribe_pcm_batch(pcm_sub)
    t_sub, l_sub = t_sub.cpu().numpy(), l_sub.cpu().numpy()
    same = [bool(np.array_equal(t_sub[k], toks_all[i]) and l_sub[k] == lens_all[i]) for k, i in enumerate(picks)]
    out["cross_g"] = {"chunks": picks, "identical_to_solo_run_on_rank0": same, "ok": all(same)}
    try:
        from oracle import logmel_oracle as LM
        from oracle import oracle as O
        from whisper_mojo_b200 import synth

        O.set_num_threads(host_threads())
        om = O.OracleWhisper(cfg, synth.make_weights(cfg, seed=0))
        rows = []
        for k, i in list(enumerate(picks))[: max(args.parity_chunks, 0)]:
            mel = LM.log_mel(pcm_sub[k].cpu().numpy())
            ref, mg = om.greedy(om.encode(mel), margins=True)
            got = toks_all[i, : lens_all[i]]
            n = min(len(got), len(ref))
            bad = np.nonzero(got[:n] != ref[:n])[0]
            first = int(bad[0]) if len(bad) else None
            rows.append({"chunk": i, "ids": int(len(ref)), "matched": int(first if first is not None else n),
                         "identical": bool(first is None and len(got) == len(ref)),
                         "oracle_margin_at_first_mismatch": None if first is None else float(mg[first - 4]),
                         "ok": bool((first is None and len(got) == len(ref)) or (first is not None and first >= 4 and mg[first - 4] < 2e-2))})
        out["oracle"] = {"rule": "ids identical to the fp32 oracle, or first mismatch where the oracle top-1/top-2 margin < 2e-2",
                         "chunks": rows, "ok": all(r["ok"] for r in rows), "identical": sum(r["identical"] for r in rows)}
    except Exception as ex:
        out["oracle"] = {"error": str(ex)[:300]}
    return out


def run_b200(args):
    import hashlib

    import torch
    import torch.distributed as dist

    from whisper_mojo_b200 import WeightLoader, Whisper, WhisperConfig, _lib, synth
    from whisper_mojo_b200.dist import gather_tokens, shard_range, wait_for_rank0

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: keep a private handle to it and point fd 1 at stderr for everything
    # else (NCCL prints its version banner straight to fd 1 when the first communicator is created)
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    small = args.model == "small"
    cfg = WhisperConfig.small_shaped() if small else WhisperConfig.tiny()
    strong = args.scaling == "strong"
    n_total = args.chunks if strong else args.chunks * world
    lo, hi = shard_range(n_total, rank, world)  # contiguous shard of the global batch
    C = hi - lo
    seed = 1234
    stream = torch.cuda.current_stream()
    model = Whisper(cfg, stream=stream.cuda_stream)
    model.load(WeightLoader(data=synth.make_weights(cfg, seed=1 if small else 0)))
    if args.lanes != 1:
        model.set_option("decode_lanes", args.lanes)
    pcm = synth_pcm_gpu(lo, hi, cfg.n_samples, dev, seed)  # 3.9 GB at 2048 chunks: larger than the 126 MB L2

    def step():
        toks, lens = model.transcribe_pcm_batch(pcm)
        if world > 1:
            toks, lens = gather_tokens(toks, lens, n_total, dst=None)
        return toks, lens

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        toks, lens = step()
    sync_all()
    sampler = ClockSampler(local_rank, args.sampler)
    sampler.start()
    launches0 = _lib.launch_count()
    phases = {"frontend_ms": 0.0, "encoder_ms": 0.0, "decode_ms": 0.0}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms, step_decode_ms = [], []
    sync_all()
    e0.record()
    for _ in range(args.steps):
        toks, lens = step()
        tm = model.last_timing()
        step_ms.append(tm["total_ms"])
        step_decode_ms.append(tm["decode_ms"])
        for k in phases:
            phases[k] += tm[k] / args.steps
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = n_total * CHUNK_SECONDS / (ms / 1e3)
    toks_all, lens_all = toks.cpu().numpy(), lens.cpu().numpy()  # the timed run's ids, global chunk order
    mean_len = float(lens_all.mean())
    ids_sha = hashlib.sha256(np.ascontiguousarray(toks_all).tobytes() + np.ascontiguousarray(lens_all).tobytes()).hexdigest()

    # ---- end-to-end through the public host API: host pcm in, host tokens out ---------------------------------
    e2e = e2e_pageable = None
    if not args.no_e2e:
        def time_host_api(arr):
            model.transcribe_pcm_batch(arr)  # warm-up of the host path
            sync_all()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                th, lh = model.transcribe_pcm_batch(arr)  # H2D pcm + compute + D2H tokens, synchronous
                if world > 1:  # the gather of ids stays inside the timed region, as in `value`
                    gather_tokens(torch.from_numpy(th).to(dev), torch.from_numpy(lh).to(dev), n_total, dst=None)
            sync_all()
            dt = (time.perf_counter() - t0) / args.steps
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt, th, lh

        pcm_host = torch.empty((C, cfg.n_samples), dtype=torch.float32, pin_memory=True)
        pcm_host.copy_(pcm)
        pcm_np = pcm_host.numpy()
        dt, th, lh = time_host_api(pcm_np)
        e2e = {"value": n_total * CHUNK_SECONDS / dt, "unit": UNIT, "h2d_bytes_per_step": int(pcm_np.nbytes) * world,
               "d2h_bytes_per_step": int(th.nbytes + lh.nbytes) * world, "ms_per_step": dt * 1e3, "host_memory": "pinned",
               "api": "Whisper.transcribe_pcm_batch (wm_transcribe_pcm)",
               "ids_equal_device_run": bool(np.array_equal(th, toks_all[lo:hi]) and np.array_equal(lh, lens_all[lo:hi]))}
        if world == 1:  # the same call with an ordinary (pageable) numpy array, e.g. what examples/main.cpp reads into
            pcm_pg = np.array(pcm_np, copy=True)
            dtp, _, _ = time_host_api(pcm_pg)
            e2e_pageable = {"value": n_total * CHUNK_SECONDS / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3,
                            "host_memory": "pageable (driver-staged copies, issued one encoder sub-batch ahead)"}
            del pcm_pg
        del pcm_host, pcm_np

    # ---- roofline of the dominant kernel: decode cross-attention, event-timed per launch -------
    roofline = None
    peaks = load_peaks()
    try:
        # one extra, untimed pass with CUDA events around every cross-attention launch; it runs a single
        # decode lane and no graph so the launches neither overlap other kernels nor hide inside a graph
        model.set_option("profile_attn", 1)
        model.set_option("decode_lanes", 1)
        # every timed launch is the decode-step kernel (one query row per chunk): the prompt goes through four cached
        # steps in this pass (same ids; the one-forward prefill attends for two prompt rows per pass)
        model.set_option("prefill_impl", 0)
        model.transcribe_pcm_batch(pcm)
        torch.cuda.synchronize()
        tot_ms, n_launch = model.last_cross_attention_timing()
        prof_decode_ms = model.last_timing()["decode_ms"]
        model.set_option("profile_attn", 0)
        model.set_option("prefill_impl", 1)
        model.set_option("decode_lanes", args.lanes)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes per launch of the absorbed cross-attention: every chunk's 16-bit enc_out read once
        # (it stands in for both K and V of the layer), plus the folded queries and the per-head contexts:
        # B * (S*D + 2*H*D) * 2 bytes.  (The K/V-cache form, cross_impl=0, reads B * (2*S*D + 2*D) * 2.)
        alg = C * (cfg.n_audio_ctx * cfg.d_model + 2 * cfg.n_heads * cfg.d_model) * 2
        if n_launch > 0:
            avg_ms = tot_ms / n_launch
            ach = alg / (avg_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "cross_attn_absorbed_kernel (cross-attention of one layer, one decode step)",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s",
                        "traffic": None if small else ncu_traffic(C), "avg_launch_ms": avg_ms, "launches_timed": n_launch,
                        "algorithmic_bytes_per_launch": alg, "chunks_per_launch": C,
                        "share_of_decode": tot_ms / max(prof_decode_ms, 1e-9)}
    except Exception as ex:  # never lose the headline number to the profiling pass
        roofline = {"error": str(ex)}

    breakdown = None
    if args.breakdown:
        try:
            model.set_option("profile_attn", 2)
            model.transcribe_pcm_batch(pcm)
            torch.cuda.synchronize()
            breakdown = {"chunks": C, "decode_ms_eager": model.last_timing()["decode_ms"], "kernels": {}}
            for k in ("cross_attention", "self_attention", "gemm_qkv", "gemm_o", "gemm_cross_q", "gemm_cross_o", "gemm_fc1",
                      "gemm_fc2", "layer_norm", "gemm_logits", "misc", "chain_first", "chain_b", "chain_ca"):
                t_ms, n_l = model.last_kernel_timing(k)
                if n_l:
                    breakdown["kernels"][k] = {"ms": t_ms, "launches": n_l, "us_each": 1e3 * t_ms / max(n_l, 1)}
            model.set_option("profile_attn", 0)
        except Exception as ex:
            breakdown = {"error": str(ex)}

    # ---- finished chunks (whisper.mojo:206-207): declared lengths, decode time with / without skipping them -------
    eot_variant = None
    if args.eot_schedule and not small:
        try:
            lo_l, hi_l = [int(x) for x in args.eot_schedule.split(":")]
            lens_decl = np.random.default_rng(99).integers(lo_l, hi_l + 1, n_total).astype(np.int32)[lo:hi]
            model.set_stop_lengths(lens_decl)
            res = {}
            for skip in (1, 0):
                model.set_option("skip_done", skip)
                model.transcribe_pcm_batch(pcm)
                tk, lk = model.transcribe_pcm_batch(pcm)
                torch.cuda.synchronize()
                res[skip] = (model.last_timing()["decode_ms"], tk.cpu().numpy(), lk.cpu().numpy())
            model.set_option("skip_done", 1)
            model.set_stop_lengths(None)
            eot_variant = {"declared_lengths": f"U({lo_l},{hi_l}) ids per chunk (wm_set_stop_lengths)", "mean_len": float(lens_decl.mean()),
                           "live_chunk_steps_frac": float((lens_decl - 4).sum() / ((cfg.max_tokens - 4) * len(lens_decl))),
                           "decode_ms_skip_done": res[1][0], "decode_ms_stream_all": res[0][0], "decode_ms_all_live": phases["decode_ms"],
                           "ids_identical": bool(np.array_equal(res[1][1], res[0][1]) and np.array_equal(res[1][2], res[0][2])),
                           "lengths_as_declared": bool(np.array_equal(res[1][2], lens_decl))}
        except Exception as ex:
            eot_variant = {"error": str(ex)[:300]}

    # ---- phase rooflines: encoder on the tensor pipe, whole decode step and frontend on HBM ----
    phase_roof = None
    try:
        D, S, L, H, F, V = cfg.d_model, cfg.n_audio_ctx, cfg.n_layers, cfg.n_heads, 4 * cfg.d_model, cfg.vocab_size
        enc_flop = (2 * cfg.n_frames * 3 * cfg.n_mels * D + 2 * S * 3 * D * D
                    + L * (8 * S * D * D + 4 * S * S * D + 4 * S * D * F))  # 36.94 GFLOP for Tiny
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1340.0))
        enc_tf = enc_flop * C / (phases["encoder_ms"] * 1e-3) / 1e12
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # decode: bytes one greedy step must move (16-bit): weights touched once + per chunk the encoder output once
        # per layer (absorbed cross-attention) + the self K/V rows written so far; averaged over the decoder forwards.
        # n_fwd = 1 + max_iters = 196: the reference's q_len = 4 prefill (whisper.mojo:195-197) runs as ONE
        # forward (prefill_impl = 1), then max_iters cached steps; step k attends over 5 + k self K/V rows.
        n_fwd = 1 + cfg.max_iters
        w_step = L * (4 * D * D + 2 * H * D * D + 2 * D * F) + V * D  # self qkv/o, folded cross q'/o', mlp, logits
        t_avg = (cfg.max_iters * (5 + (cfg.max_iters - 1) / 2.0) + 10) / n_fwd
        step_bytes = 2 * w_step + C * 2 * (L * S * D + 2 * L * t_avg * D)
        dec_gbs = step_bytes * n_fwd / (phases["decode_ms"] * 1e-3) / 1e9
        fe_bytes = C * (cfg.n_samples * 4 + cfg.n_mels * cfg.n_frames * 4)
        phase_roof = {
            "encoder": {"bound": "tensor", "achieved": enc_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": enc_tf / tf_peak,
                        "flop_per_chunk": enc_flop, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (fp16 runs at the same kind::f16 rate)"},
            "decode_step": {"bound": "hbm", "achieved": dec_gbs, "peak": hbm_peak, "unit": "GB/s",
                            "frac": dec_gbs / hbm_peak, "bytes_per_step": step_bytes, "forwards": n_fwd,
                            "note": "196 forwards = one q_len = 4 prompt forward + 195 greedy steps (SURVEY 8d); algorithmic bytes: 16-bit weights once, enc_out once per layer, the self K/V rows written so far"},
            "frontend": {"bound": "hbm", "achieved": fe_bytes / (phases["frontend_ms"] * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": fe_bytes / (phases["frontend_ms"] * 1e-3) / 1e9 / hbm_peak,
                         "note": "pcm f32 in + log-mel f32 out; the TF32x3 DFT adds 2.9 GFLOP-equivalent per chunk",
                         "dft_tflops_tf32": 3 * 2 * cfg.n_frames * 416 * 416 * C / (phases["frontend_ms"] * 1e-3) / 1e12},
        }
    except Exception as ex:
        phase_roof = {"error": str(ex)}

    parity = None
    if rank == 0 and not args.no_parity and not small:
        try:
            parity = parity_checks(args, model, cfg, toks_all, lens_all, n_total, dev, seed)
        except Exception as ex:
            parity = {"error": str(ex)[:300]}
    if world > 1:
        # the other ranks wait for rank 0's checks before tearing the group down -- asleep on the store's socket, not
        # spinning in a NCCL barrier (7 polling ranks on 16 cores made the OpenMP oracle several times slower)
        wait_for_rank0("parity", timeout_s=900.0)

    cpu = cpu_hf = bf16_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        force_host_threads()
        cpu_hf = None if (small or args.no_hf_baseline) else hf_generate_baseline()
        n_cpu = 1 if small else args.cpu_chunks
        rate, dt, cores, _, n_cpu = cpu_transcribe_rate(n_cpu, small=small)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} of the same synthetic-weight 30 s chunks, batch 1, precomputed log-mel "
                         f"(reference's timed region), {dt:.1f} s of CPU work"}
    if rank == 0 and world == 1 and not args.no_bf16 and not small and _lib.precision() == "fp16":
        # the bf16 build of the same sources, same workload, device-resident number only (A/B: the operand type costs nothing)
        try:
            env = dict(os.environ, WB_PRECISION="bf16")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(args.steps), "--warmup", str(args.warmup),
                                "--chunks", str(args.chunks), "--scaling", args.scaling, "--no-e2e", "--no-cpu-baseline",
                                "--no-parity", "--no-bf16", "--sampler", "none", "--eot-schedule", ""], env=env, capture_output=True, text=True, timeout=300)
            b = json.loads(r.stdout.strip().splitlines()[-1])
            bf16_line = {"value": b["value"], "unit": UNIT, "ms_per_step": b["ms_per_step"], "dtype": b["dtype"],
                         "note": "libwhisper_b200_bf16.so (-DWB_BF16): enc_out / logits max-abs 2.7e-2 / 5.4e-2 vs the oracle, above north_star's 1e-2"}
        except Exception as ex:
            bf16_line = {"error": str(ex)[:200]}

    if rank == 0:
        prec = _lib.precision()
        line = {
            "metric": METRIC.replace("tiny", "small_shaped") if small else METRIC, "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": ("whisper-small-shaped (12 layers, d 768, 12 heads) batched greedy transcription "
                                    "(BASELINE.json configs[4]): " if small else
                                    "whisper-tiny batched greedy transcription (BASELINE.json configs[3]): ") +
                                   f"{n_total} synthetic 30 s chunks per step sharded over {world} GPU(s), pcm -> log-mel -> encoder -> "
                                   "196 decoder forwards: the 4-id prompt as one q_len = 4 forward + 195 greedy steps (EOT never fires with random weights)",
                       "chunks_per_gpu": C, "global_chunks": n_total, "weights": "random-init whisper-%s shapes" % ("small" if small else "tiny"),
                       "precision": f"{prec} weights / GEMM operands / KV cache, fp32 accumulation, fp32 residual stream, LayerNorm, softmax and logits",
                       "l2": "inputs (pcm %.1f GB per GPU) larger than the 126 MB L2; no explicit flush" % (pcm.numel() * 4 / 1e9),
                       "parallelism": f"chunk i -> rank i*{world} // {n_total} (contiguous shards), no data-path collective, final NCCL all_gather of ids inside the timed region",
                       "mean_tokens_per_chunk": mean_len},
            "clocks": clocks, "e2e": e2e, "e2e_pageable": e2e_pageable, "gpu_launches": int(launches), "roofline": roofline,
            "phase_rooflines": phase_roof, "parity": parity, "ids_sha256": ids_sha,
            "cpu_baseline": cpu, "cpu_baseline_hf_generate": cpu_hf, "bf16_variant": bf16_line, "decode_breakdown": breakdown,
            "eot_schedule_variant": eot_variant,
            "phases_ms_per_step": phases, "device_ms_each_step": step_ms, "decode_ms_each_step": step_decode_ms,
        }
        def plain(o):  # numpy scalars that slipped into the record
            return o.item() if hasattr(o, "item") else str(o)

        print(json.dumps(line, default=plain), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
