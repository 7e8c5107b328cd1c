"""The one exchange step of the path on N > 1 ranks, on CPU: world_size-2 `gloo` process
groups run the same host code the NCCL ranks run (shard by utterance -> per-rank five-number
summary -> all-reduce -> finalise), and every rank must end up with the statistics the
reference's Scaler computes over the whole corpus (preprocessor/helpers.py:86-106)."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _corpus():
    """Per-utterance value vectors (phone-level energies) with NaNs, ragged."""
    rng = np.random.default_rng(77)
    utts = [rng.normal(30.0, 9.0, size=int(n)).astype(np.float32) for n in rng.integers(5, 90, size=41)]
    utts[3][2] = np.nan
    utts[17][[0, 4]] = np.nan
    return utts


def _five(x: np.ndarray) -> torch.Tensor:
    """What evf_stats_partial leaves in out5 for this shard (the test plays the kernel's part;
    the kernel itself is checked on the GPU in test_gpu_parity.py)."""
    v = x[~np.isnan(x)].astype(np.float64)
    if v.size == 0:
        return torch.tensor([0.0, 0.0, 0.0, float("inf"), float("-inf")], dtype=torch.float64)
    return torch.tensor([v.size, v.sum(), (v * v).sum(), v.min(), v.max()], dtype=torch.float64)


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from everyvoice_b200.distributed import (allgather_stats, allreduce_stats, finalize_stats, merge_stats,
                                                 shard_utterances)

        utts = _corpus()
        mine = shard_utterances([len(u) for u in utts], world)[rank]
        local = np.concatenate([utts[i] for i in mine]) if mine else np.zeros(0, np.float32)
        stats5, n_files = allreduce_stats(_five(local), len(mine))
        st = finalize_stats(stats5.tolist(), n_files)
        # a rank whose shard is empty still takes part and gets the same answer
        empty5, n2 = allreduce_stats(_five(local if rank == 0 else np.zeros(0, np.float32)), len(mine) if rank == 0 else 0)
        # the hot-loop form: ONE all-gather, merged later (on the GPU: evf_normalize_by_gathered_stats)
        parts = allgather_stats(_five(local))
        assert tuple(parts.shape) == (world, 5) and torch.equal(parts[rank], _five(local))
        assert torch.equal(merge_stats(parts), stats5)
        q.put((rank, st, len(mine), empty5.tolist(), n2))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_stats_allreduce_world2_matches_reference_scaler():
    from oracle import ev_oracle as O

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0

    utts = _corpus()
    s = O.Scaler()
    for u in utts:
        s.append(torch.from_numpy(u))
    ref = s.calculate_stats()
    assert sum(r[2] for r in results) == len(utts)
    for rank, st, n_mine, empty5, n2 in results:
        assert st["sample_size"] == ref["sample_size"] == len(utts)  # number of FILES, summed over ranks
        assert st["min"] == ref["min"] and st["max"] == ref["max"]
        for k in ("mean", "std", "norm_min", "norm_max"):
            assert st[k] == pytest.approx(ref[k], rel=2e-6), (rank, k)
    assert results[0][1] == results[1][1]  # every rank holds identical statistics
    assert results[0][3] == results[1][3] and results[0][4] == results[1][4] == results[0][2]


def test_allreduce_is_identity_without_a_process_group():
    from everyvoice_b200.distributed import allreduce_stats

    assert not dist.is_initialized()
    t = torch.tensor([3.0, 6.0, 14.0, 1.0, 3.0], dtype=torch.float64)
    out, n = allreduce_stats(t, 5)
    assert out is t and n == 5
    from everyvoice_b200.distributed import allgather_stats, merge_stats

    parts = allgather_stats(t)
    assert tuple(parts.shape) == (1, 5) and torch.equal(merge_stats(parts), t)
    two = torch.stack([t, torch.tensor([2.0, 10.0, 52.0, 4.0, 6.0], dtype=torch.float64)])
    assert merge_stats(two).tolist() == [5.0, 16.0, 66.0, 1.0, 6.0]
