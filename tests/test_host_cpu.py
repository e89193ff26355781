"""CPU: host-side logic -- synthetic assets, loader, tokenizer, sharding, and the world_size-2
gather over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

from whisper_mojo_b200 import Tokenizer, WhisperConfig, synth
from whisper_mojo_b200.dist import gather_tokens, shard_range


def test_synth_weights_deterministic_and_sized():
    cfg = WhisperConfig.micro()
    a, b = synth.make_weights(cfg, seed=3), synth.make_weights(cfg, seed=3)
    assert a.dtype == np.float32 and a.size == cfg.weight_count() and np.array_equal(a, b)
    assert not np.array_equal(a, synth.make_weights(cfg, seed=4))
    off, shape = cfg.weight_offsets()["enc.pos"]
    assert np.allclose(a[off:off + shape[0] * shape[1]].reshape(shape), synth.sinusoids(*shape))


def test_weight_file_roundtrip(tmp_path):
    cfg = WhisperConfig.micro()
    w = synth.make_weights(cfg, seed=0)
    p = tmp_path / "w.bin"
    synth.write_weights(str(p), w)
    assert os.path.getsize(p) == 4 * cfg.weight_count()
    assert np.array_equal(np.fromfile(p, "<f4"), w)


def test_audio_shapes():
    cfg = WhisperConfig.tiny()
    a = synth.make_audio(2, cfg, seed=0)
    assert a.shape == (2, 480000) and a.dtype == np.float32
    assert np.abs(a[:, int(480000 * 0.8):]).max() < 1e-2  # quiet tail exercises the -8 clamp


def test_tokenizer_matches_reference_rules(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("Hello\nĠworld\n<|endoftext|>\nline\\nbreak\n", encoding="utf-8")
    t = Tokenizer(str(p))
    assert t.decode([0, 1, 2, 3, 99, -1]) == "Hello worldline\nbreak"


def test_tokenizer_on_reference_golden():
    vocab = "/root/reference/vocab.txt"
    if not os.path.exists(vocab):
        pytest.skip("reference vocab not present on this machine")
    import json

    ids = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_expected_tokens.json")))["ids"]
    text = Tokenizer(vocab).decode([50258, 50259, 50359, 50363] + ids + [50257])
    assert text.startswith(" This is my voice on the left.") and text.endswith("out of phase on three.")


@pytest.mark.parametrize("n,world", [(2048, 8), (2048, 1), (7, 4), (3, 8), (0, 2)])
def test_shard_range_partitions(n, world):
    parts = [shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_total, T, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    toks = (torch.arange(lo, hi, dtype=torch.int32)[:, None] * 1000 + torch.arange(T, dtype=torch.int32)[None, :])
    lens = torch.arange(lo, hi, dtype=torch.int32) % T + 1
    out_t, out_l = gather_tokens(toks, lens, n_total, dst=0)
    all_t, all_l = gather_tokens(toks, lens, n_total, dst=None)
    ok = True
    if rank == 0:
        exp_t = (torch.arange(n_total, dtype=torch.int32)[:, None] * 1000 + torch.arange(T, dtype=torch.int32)[None, :])
        exp_l = torch.arange(n_total, dtype=torch.int32) % T + 1
        ok = torch.equal(out_t, exp_t) and torch.equal(out_l, exp_l)
    else:
        ok = out_t is None and out_l is None
    ok = ok and all_t.shape == (n_total, T) and int(all_l[n_total - 1]) == (n_total - 1) % T + 1
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gather_tokens_world2_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _wait_worker(rank, world, port, q):
    import time

    from whisper_mojo_b200.dist import wait_for_rank0

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t0 = time.time()
    if rank == 0:
        time.sleep(1.5)  # rank 0's host-side work
    wait_for_rank0("t", timeout_s=60.0)
    q.put((rank, time.time() - t0))
    dist.barrier()
    dist.destroy_process_group()


def test_wait_for_rank0_releases_the_other_ranks_only_after_rank0_world2_gloo():
    """bench.py's end-of-run wait: ranks != 0 sleep on the rendezvous store until rank 0 is through its checks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wait_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[1] >= 1.0, res  # rank 1 was held until rank 0 arrived
    assert res[0] < 60 and res[1] < 60


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm): rank 0 prints ONE JSON line carrying the bench
    contract's keys with the CPU restatement's own value; other ranks exit 0 silently; the B200 arm refuses to run
    without a device (no CPU fallback)."""
    import json
    import subprocess
    import sys

    import torch

    bench = os.path.join(ROOT, "bench.py")
    cmd = [sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-chunks", "1", "--no-hf-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: rank 0 of the reference arm must still use
    # every host core (round 1's arm timed out at N > 1 for this reason)
    r3 = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                        env={**os.environ, "RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    assert r3.returncode == 0, r3.stderr[-2000:]
    d3 = json.loads([ln for ln in r3.stdout.splitlines() if ln.strip()][0])
    assert d3["cpu_baseline"]["cores"] == max(1, len(os.sched_getaffinity(0)))
    if not torch.cuda.is_available():
        r2 = subprocess.run([sys.executable, bench, "--steps", "1"], capture_output=True, text=True, timeout=600)
        assert r2.returncode != 0 and "no CUDA device" in (r2.stderr + r2.stdout)


def test_bench_chunk_audio_depends_only_on_the_global_index():
    """bench.py shards ONE global seeded batch: chunk i's samples must not depend on the shard it is generated in,
    otherwise ids could not be compared across GPU counts (strong scaling)."""
    import importlib.util

    import torch

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    dev = torch.device("cpu")
    full = b.synth_pcm_gpu(0, 192, 1600, dev, 7)
    for lo, hi in ((0, 96), (96, 192), (64, 128), (130, 131)):
        assert torch.equal(b.synth_pcm_gpu(lo, hi, 1600, dev, 7), full[lo:hi]), (lo, hi)
    assert not torch.equal(full[0], full[64])
