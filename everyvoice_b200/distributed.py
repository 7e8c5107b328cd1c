"""Multi-GPU plumbing for the one exchange step of the path: corpus statistics.

The corpus is sharded by utterance, one process per GPU; nothing crosses GPUs except the
five-number summaries ``{count, sum, sumsq, min, max}`` (float64) per feature, which are
all-reduced (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests) and then
turned into the reference's ``Scaler`` statistics (preprocessor/helpers.py:86-106).
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist


def shard_utterances(lengths, world_size: int) -> list[list[int]]:
    """Greedy longest-first partition of utterance indices into ``world_size`` shards of
    near-equal total length (replaces the loky pool of preprocessor.py:1197-1209).
    Deterministic; every rank computes the same partition."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    loads = [0] * world_size
    shards: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(int(i))
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def allreduce_stats(stats5: torch.Tensor, sample_size: int, group=None) -> tuple[torch.Tensor, int]:
    """All-reduce ``{count, sum, sumsq}`` with SUM and ``{min, max}`` with MIN / MAX, and the
    number of files with SUM.  ``stats5`` is a float64 tensor of 5 values on the device the
    process group's backend expects (CUDA for nccl, CPU for gloo)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats5, sample_size
    backend = dist.get_backend(group)
    work = stats5.detach().clone()
    if backend == "gloo" and work.is_cuda:
        work = work.cpu()
    sums = torch.cat([work[:3], work.new_tensor([float(sample_size)])])
    ext = torch.stack([-work[3], work[4]])  # one MAX all-reduce covers min and max
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    out = torch.stack([sums[0], sums[1], sums[2], -ext[0], ext[1]]).to(stats5.device)
    return out, int(round(float(sums[3])))


def finalize_stats(stats5, sample_size: int) -> dict:
    """``{count, sum, sumsq, min, max}`` -> the dict ``Scaler.calculate_stats`` returns
    (helpers.py:86-106): ``mean = nanmean``, ``std = torch.std`` (unbiased), float32 values."""
    n, s, sq, mn, mx = (float(v) for v in stats5)
    if n < 1:
        raise ValueError("no non-NaN values to compute statistics from")
    mean = s / n
    var = (sq - s * s / n) / (n - 1) if n > 1 else float("nan")
    std = math.sqrt(max(var, 0.0)) if n > 1 else float("nan")
    f32 = lambda v: float(np.float32(v))  # noqa: E731
    mean32, std32, mn32, mx32 = np.float32(mean), np.float32(std), np.float32(mn), np.float32(mx)
    return {
        "sample_size": int(sample_size),
        "norm_min": f32((mn32 - mean32) / std32),
        "norm_max": f32((mx32 - mean32) / std32),
        "min": f32(mn32),
        "max": f32(mx32),
        "mean": f32(mean32),
        "std": f32(std32),
    }
