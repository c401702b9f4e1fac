"""Multi-GPU plumbing: independent streams one per GPU (SURVEY.md §8e — replicas only, no data-path collective).

The only communication is the reduction of throughput counters: the slowest rank's elapsed time (MAX) and the total
number of frames (SUM).  Backend "nccl" on the GPU box (NVLink/NVSwitch, a few bytes — latency bound), "gloo" in the
CPU tests.
"""
from __future__ import annotations


def stream_seed(rank: int) -> int:
    """Each rank tracks its own synthetic sequence (BASELINE.json config 5: seeds 1..8)."""
    return rank + 1


def reduce_counters(elapsed_ms: float, frames: int, device=None):
    """(max over ranks of elapsed_ms, sum over ranks of frames).  Works without an initialised process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), int(frames)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([int(frames)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


def aggregate_fps(elapsed_ms: float, frames: int, device=None) -> float:
    """Whole-job frames/sec: all ranks' frames over the slowest rank's time."""
    ms, total = reduce_counters(elapsed_ms, frames, device)
    return total / (ms * 1e-3)
