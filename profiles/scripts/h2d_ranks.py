#!/usr/bin/env python
"""Measurement aid (not part of the product): concurrent host->device copies from pinned memory on N ranks of one box —
what limits the end-to-end leg when several GPUs upload at once.  Under torchrun:
    python -m torch.distributed.run --nproc-per-node N profiles/scripts/h2d_ranks.py
Every rank copies a pinned buffer of `size` bytes to its GPU `reps` times after a barrier; rank 0 prints per-rank and
aggregate GB/s for (a) ranks bound to disjoint core sets of their GPU's NUMA node, buffers first-touched there, and
(b) unbound ranks.  Alone (N = 1) the same copy is the PCIe floor of one GPU."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                                        # noqa: E402


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    out = {"world": world}
    all_cpus = os.sched_getaffinity(0)
    for mode in ("bound", "unbound"):
        if mode == "bound":
            desc = bench.bind_to_gpu_numa_node(local, world)
        else:
            os.sched_setaffinity(0, all_cpus)
            desc = "all %d cpus" % len(all_cpus)
        for size in (265_000_000, 33_000_000):
            src = torch.empty(size, dtype=torch.uint8).pin_memory()
            src.fill_(1)                                              # first touch under the current affinity
            dst = torch.empty(size, dtype=torch.uint8, device=dev)
            reps = 20 if size > 100_000_000 else 100
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            gbps = torch.tensor([size * reps / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device=dev)
            per = [gbps]
            if world > 1:
                per = [torch.zeros_like(gbps) for _ in range(world)]
                dist.all_gather(per, gbps)
            out["%s_%dMB" % (mode, size // 1_000_000)] = {"affinity": desc, "per_rank_GBps": [round(float(p.item()), 1) for p in per],
                                                         "aggregate_GBps": round(sum(float(p.item()) for p in per), 1)}
            del src, dst
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
