"""Host-side logic for running independent sequences one-per-GPU ("replicas only").

A single camera stream has no intra-frame partition worth its latency, so the multi-GPU mode is N independent
ElasticFusion instances, one process and one sequence per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests)
is used only for the start/end barriers and to combine per-rank timings. No data-path collective exists.
"""
from __future__ import annotations

from typing import List, Sequence


def shard_sequences(n_sequences: int, world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of sequence ids to ranks (every rank gets floor or ceil of n/world)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, n_sequences, world_size))


def sequence_seed(base_seed: int, sequence_id: int) -> int:
    """Seed of the synthetic trajectory / noise of one sequence (BASELINE config 5: seeds 42..49)."""
    return base_seed + sequence_id


def barrier(dist, device=None) -> None:
    """Launch barrier: a 1-element all-reduce (works for both nccl and gloo)."""
    import torch

    t = torch.zeros(1, device=device) if device is not None else torch.zeros(1)
    dist.all_reduce(t)
    if device is not None and getattr(device, "type", "cpu") == "cuda":
        torch.cuda.synchronize(device)


def aggregate_throughput(dist, frames_local: int, seconds_local: float, device=None) -> dict:
    """Whole-job frames/sec: total frames over all ranks divided by the slowest rank's time."""
    import torch

    kw = {"device": device} if device is not None else {}
    f = torch.tensor([float(frames_local)], dtype=torch.float64, **kw)
    s = torch.tensor([float(seconds_local)], dtype=torch.float64, **kw)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    total, slowest = float(f.item()), float(s.item())
    return {"frames": total, "seconds": slowest, "fps": total / slowest if slowest > 0 else float("nan")}


def local_frames(frames_per_sequence: int, seqs: Sequence[int]) -> int:
    return frames_per_sequence * len(seqs)
