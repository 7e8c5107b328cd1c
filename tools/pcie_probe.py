#!/usr/bin/env python
"""Host <-> device copy ceiling of this box, with no kernels: what the end-to-end path (bench.py `e2e`) can at best
reach.  One process per GPU (launch with torch.distributed.run for N > 1, like bench.py); every rank copies the byte
counts of one e2e step of the default workload -- H2D alone, D2H alone, and both directions at once on two streams --
from / to pinned host buffers, in chunks of --chunk-mb, and rank 0 prints ONE JSON line with the per-rank and
aggregate GB/s (max time over ranks).

    python tools/pcie_probe.py [--h2d-mb 241] [--d2h-mb 153] [--chunk-mb 16] [--reps 10]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d-mb", type=float, default=241.0)   # int16 samples of the 1 000-utterance step
    ap.add_argument("--d2h-mb", type=float, default=153.0)   # log-mel + energy + phone values
    ap.add_argument("--chunk-mb", type=float, default=16.0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-bind", action="store_true")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = None
    if not args.no_bind:
        from everyvoice_b200.distributed import bind_to_gpu_numa_node

        cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local])
        finally:
            sys.stdout.flush()
            os.dup2(fd, 1)
            os.close(fd)
    n_in, n_out, ck = int(args.h2d_mb * 1e6), int(args.d2h_mb * 1e6), int(args.chunk_mb * 1e6)
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n_out, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d():
        with torch.cuda.stream(s_in):
            for a in range(0, n_in, ck):
                d_in[a:a + ck].copy_(h_in[a:a + ck], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s_out):
            for a in range(0, n_out, ck):
                h_out[a:a + ck].copy_(d_out[a:a + ck], non_blocking=True)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            fn()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / args.reps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return dt

    t_in = timed(h2d)
    t_out = timed(d2h)
    t_both = timed(lambda: (h2d(), d2h()))
    if rank == 0:
        gb = 1e-9
        print(json.dumps({
            "n_gpus": world, "chunk_mb": args.chunk_mb, "h2d_mb": args.h2d_mb, "d2h_mb": args.d2h_mb,
            "h2d_alone": {"ms": t_in * 1e3, "gbs_per_rank": n_in * gb / t_in, "gbs_total": world * n_in * gb / t_in},
            "d2h_alone": {"ms": t_out * 1e3, "gbs_per_rank": n_out * gb / t_out, "gbs_total": world * n_out * gb / t_out},
            "duplex": {"ms": t_both * 1e3, "h2d_gbs_per_rank": n_in * gb / t_both, "d2h_gbs_per_rank": n_out * gb / t_both,
                       "gbs_total": world * (n_in + n_out) * gb / t_both},
            "bound_ms_full_duplex": max(t_in, t_out) * 1e3,
            "host_affinity": len(cpus) if cpus else None,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
