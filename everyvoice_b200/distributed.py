"""Multi-GPU plumbing for the one exchange step of the path: corpus statistics.

The corpus is sharded by utterance, one process per GPU; nothing crosses GPUs except the
five-number summaries ``{count, sum, sumsq, min, max}`` (float64) per feature, which are
all-gathered in ONE collective (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU
tests; 40 bytes per rank) and merged where they are needed: on the device by the normalisation
kernel, on the host for the reference's ``Scaler`` statistics (preprocessor/helpers.py:86-106).
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


def allgather_stats(stats5: torch.Tensor, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """ONE collective for the exchange step: all-gather every rank's five-number summary
    ``{count, sum, sumsq, min, max}`` (float64) into ``[world, 5]``.  Asynchronous on the current
    stream for NCCL (no host synchronisation); ``evf_normalize_by_gathered_stats`` /
    ``evf_stats_merge`` consume the result directly on the device.  Without a process group the
    result is ``stats5[None]``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats5.reshape(1, -1)
    world = dist.get_world_size(group)
    src = stats5.detach().reshape(-1)[:5].contiguous()
    if dist.get_backend(group) == "gloo" and src.is_cuda:
        src = src.cpu()
    if out is None or out.numel() != world * 5 or out.device != src.device or not out.is_contiguous():
        out = torch.empty((world, 5), dtype=torch.float64, device=src.device)
    dist.all_gather_into_tensor(out.view(-1), src, group=group)  # flat output: accepted by nccl and gloo
    return out.view(world, 5)


def merge_stats(parts: torch.Tensor) -> torch.Tensor:
    """``[world, 5]`` -> ``[5]`` (SUM, SUM, SUM, MIN, MAX) with torch ops on whatever device the
    parts live on; the host-side twin of ``evf_stats_merge`` used when the statistics are needed
    as Python numbers anyway (``Scaler.calculate_stats`` -> stats.json)."""
    parts = parts.reshape(-1, 5)
    return torch.cat([parts[:, :3].sum(dim=0), parts[:, 3].min().reshape(1), parts[:, 4].max().reshape(1)])


def allreduce_stats(stats5: torch.Tensor, sample_size: int, group=None) -> tuple[torch.Tensor, int]:
    """Corpus-wide ``{count, sum, sumsq, min, max}`` and number of files on every rank.  Returns
    Python-side values (synchronises); the hot loop uses ``allgather_stats`` instead."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats5, sample_size
    parts = allgather_stats(stats5, group)
    n = torch.tensor([float(sample_size)], dtype=torch.float64, device=parts.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return merge_stats(parts).to(stats5.device), int(round(float(n[0])))


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


def bind_to_gpu_numa_node(local_rank: int) -> list[int] | None:
    """Pin this process to the CPU cores NVML reports as local to GPU ``local_rank`` (its NUMA node / PCIe root).
    Call it BEFORE allocating pinned host buffers: page-locked memory is placed by first touch, and a rank whose
    staging buffers sit on the other socket pushes every H2D / D2H byte over the inter-socket link -- with 4-8 ranks
    per box that link, not PCIe, bounds the end-to-end rate.  Returns the CPU list, or ``None`` if NVML or the
    affinity call is unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(
                f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0".encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1 and 64 * i + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
