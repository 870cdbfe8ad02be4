"""Partitioning of independent streams / frame batches over the GPUs of one box.

There is no collective on the data path: every (stream, frame) is
independent and an overlay belongs to exactly one stream, so rank r simply owns the
streams with id % world == r (config 5) or its own frame batches (config 3, weak scaling).
torch.distributed is only the harness plumbing: a barrier around the timed region, the MAX
of the per-rank device times, and a gather of the per-rank result records.
"""
from __future__ import annotations

from typing import List, Sequence


def shard_streams(n_streams: int, world: int, rank: int) -> List[int]:
    """Stream ids owned by `rank`: static id % world partition."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return [s for s in range(n_streams) if s % world == rank]


def shard_batches(n_batches: int, world: int, rank: int) -> List[int]:
    """Round-robin frame batches of one fat stream (strong-scaling split)."""
    return shard_streams(n_batches, world, rank)


def reduce_max(value: float, dist=None, device=None) -> float:
    """MAX over ranks of a per-rank scalar (device time in ms)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_records(record: Sequence[float], dist=None, device=None) -> List[List[float]]:
    """All ranks' result records (frames, ms, kernel_ms, ...), rank order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(map(float, record))]
    import torch
    t = torch.tensor(list(map(float, record)), dtype=torch.float64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def aggregate_fps(records: Sequence[Sequence[float]]) -> float:
    """Whole-job frames/s: all ranks' frames over the slowest rank's time.
    record = (frames, milliseconds, ...)."""
    frames = sum(r[0] for r in records)
    ms = max(r[1] for r in records)
    return frames / (ms * 1e-3) if ms > 0 else 0.0
