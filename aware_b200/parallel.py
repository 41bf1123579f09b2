"""Clip-level data parallelism: every clip is independent, so the
clips of a job are split contiguously over the ranks of one box and the only
communication is one all-reduce of the BER / quality counters.  Backend-agnostic
(`nccl` on the B200 box, `gloo` in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `n_items` for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_counters(counters: torch.Tensor, sums: torch.Tensor | None = None):
    """Sum int64 {bit errors, bits, clips} counters (and optional float64 aggregates such as
    sum-of-SNR) over all ranks, in place.  No-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
        if sums is not None:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return counters, sums


def ber_percent(counters: torch.Tensor) -> float:
    """counters[..., 0] errors / counters[..., 1] bits, in percent (metrics/audio.py:15 upstream)."""
    c = counters.reshape(-1, 3).sum(0)
    return 100.0 * float(c[0]) / max(float(c[1]), 1.0)
