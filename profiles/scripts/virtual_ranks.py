"""What each of N position-sharded ranks does per step, measured on ONE GPU (no 8-GPU box needed for the kernels' side
of strong scaling): the C2 batch is cut with dist.balanced_cuts (or by read count: --cut reads) into N ranges, and every
range's step — tile index, tiles kernel, overflow jobs, region sums over the range — is timed alone with CUDA events.
The all-reduce and the eight-way PCIe sharing are NOT in here (profiles/bench_r02_n8_*.json have them).
Output: one JSON object on stdout."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from plastid_b200 import _lib, dist as pdist  # noqa: E402
from plastid_b200.genome_array import CountPlanes, map_batch, region_sums  # noqa: E402


def main():
    world = 8
    cut = "cost"
    argv = sys.argv[1:]
    if "--world" in argv:
        i = argv.index("--world"); world = int(argv[i + 1]); del argv[i:i + 2]
    if "--cut" in argv:
        i = argv.index("--cut"); cut = argv[i + 1]; del argv[i:i + 2]
    sys.argv = [sys.argv[0]] + argv
    args = bench.parse_args()
    dev = "cuda:0"
    W = bench.build_world(args, 0, dev)
    layout, table, dbatch = W["layout"], W["table"], W["dbatch"]
    table.device(dev)
    L = _lib.lib()
    weights = (1.0, 1.0) if cut == "cost" else (1.0, 0.0)
    out = {"world": world, "cut": cut, "ranks": []}
    for rank in range(world):
        sub, lo, hi, _cuts = pdist.shard_positions_device(dbatch, layout, rank, world, weights=weights)
        planes = CountPlanes(layout, "u32", dev, (lo, hi))
        planes.alloc(("+", "-"))

        def step():
            map_batch(sub, layout, W["fac"], W["sf"], strands=("+", "-"), planes=planes, sync_stats=False, bin_range=(lo, hi))
            return region_sums(planes, table)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        L.pb_enable_kernel_timing(1)
        n = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        for k in range(n):
            ev[k].record()
            step()
        ev[n].record()
        torch.cuda.synchronize()
        kms, kn = C.c_float(0), C.c_int(0)
        _lib.check(L.pb_tiles_kernel_ms_total(C.byref(kms), C.byref(kn)))
        L.pb_enable_kernel_timing(0)
        out["ranks"].append({"rank": rank, "reads_incl_halo": int(sub.n_reads), "bins": int(hi - lo),
                             "tiles_kernel_ms": round(kms.value / max(kn.value, 1), 4),
                             "step_ms": round(ev[0].elapsed_time(ev[n]) / n, 4)})
        del planes, sub
        torch.cuda.empty_cache()
    out["step_ms_max"] = max(r["step_ms"] for r in out["ranks"])
    out["tiles_kernel_ms_max"] = max(r["tiles_kernel_ms"] for r in out["ranks"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
