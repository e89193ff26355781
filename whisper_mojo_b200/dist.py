"""Multi-GPU plumbing: shard independent 30 s chunks across ranks, gather token ids at the end.

Each chunk is a self-contained `transcribe` (own mel, own KV cache; whisper.mojo:184-223), so the
path shards with no data-path collective: rank r gets the contiguous block
[n*r//world, n*(r+1)//world) and the only traffic is one gather of int32 tokens + lengths
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of chunk indices owned by `rank` (chunk i -> rank i*G // N)."""
    assert 0 <= rank < world
    return (n_chunks * rank) // world, (n_chunks * (rank + 1)) // world


def gather_tokens(tokens, lens, n_total: int, group=None, dst: Optional[int] = 0):
    """tokens int32 [n_local, T], lens int32 [n_local] (torch tensors, CPU for gloo / CUDA for nccl)
    -> on rank `dst` (or on every rank when dst is None): (tokens [n_total, T], lens [n_total]) in
    global chunk order; other ranks get (None, None).  Shards may differ in size by one chunk, so
    every rank pads to the largest shard before the all_gather."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    T = tokens.shape[1]
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert tokens.shape[0] == sizes[rank] and lens.shape[0] == sizes[rank]
    mx = max(sizes)
    buf = torch.full((mx, T + 1), -1, dtype=torch.int32, device=tokens.device)
    buf[: sizes[rank], :T] = tokens
    buf[: sizes[rank], T] = lens
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    if dst is not None and rank != dst:
        return None, None
    cat = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    return cat[:, :T].contiguous(), cat[:, T].contiguous()


def wait_for_rank0(tag: str, timeout_s: float = 3600.0) -> None:
    """Every rank calls this; ranks != 0 return once rank 0 has called it.  The wait SLEEPS on the rendezvous store's
    socket (a c10d key wait) instead of spinning in a collective: a NCCL barrier keeps one host thread per waiting rank
    polling at 100 %, which starves host-side work rank 0 still has to do (bench.py's CPU-oracle parity check ran
    several times slower at 8 ranks on 16 cores for that reason).  Falls back to `dist.barrier()` if the process
    group exposes no store."""
    import datetime

    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    try:
        store = dist.distributed_c10d._get_default_store()
        key = f"wb_rank0_done/{tag}"
        if dist.get_rank() == 0:
            store.set(key, "1")
        else:
            store.wait([key], datetime.timedelta(seconds=timeout_s))
    except AttributeError:  # no store on this process group: the collective form
        dist.barrier()
    except RuntimeError:  # rank 0 never arrived within timeout_s: give up waiting, the caller tears the group down
        pass


def transcribe_sharded(model, mel_or_pcm: np.ndarray, rank: int, world: int, pcm: bool = False):
    """Run this rank's shard of a host-resident batch; returns (tokens, lens, (lo, hi))."""
    lo, hi = shard_range(mel_or_pcm.shape[0], rank, world)
    if hi == lo:
        T = model.config.max_tokens
        return np.zeros((0, T), np.int32), np.zeros((0,), np.int32), (lo, hi)
    part = mel_or_pcm[lo:hi]
    toks, lens = model.transcribe_pcm_batch(part) if pcm else model.transcribe_batch(part)
    return toks, lens, (lo, hi)
