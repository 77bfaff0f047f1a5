"""pb_region_sums on the resident C2 planes: plain 16-byte loads (default) vs blocks staged by cp.async.bulk + mbarrier
(PB_REGION_TMA=1).  CUDA events around 50 calls each after 5 warm-up calls; tables must be identical."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from plastid_b200.genome_array import CountPlanes, map_batch, region_sums  # noqa: E402


def timed(fn, n=50, warm=5):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, out


def main():
    args = bench.parse_args()
    dev = "cuda:0"
    W = bench.build_world(args, 0, dev)
    planes = CountPlanes(W["layout"], "u32", dev, None)
    planes.alloc(("+", "-"))
    map_batch(W["dbatch"], W["layout"], W["fac"], W["sf"], strands=("+", "-"), planes=planes)
    table = W["table"]
    table.device(dev)
    res = {}
    os.environ.pop("PB_REGION_TMA", None)
    us, (s0, l0) = timed(lambda: region_sums(planes, table))
    res["plain_loads_us"] = round(us, 2)
    s0, l0 = s0.clone(), l0.clone()
    os.environ["PB_REGION_TMA"] = "1"
    us, (s1, l1) = timed(lambda: region_sums(planes, table))
    res["tma_staged_us"] = round(us, 2)
    res["identical"] = bool(torch.equal(s0, s1) and torch.equal(l0, l1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
