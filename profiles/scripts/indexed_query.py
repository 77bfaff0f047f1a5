"""Latency of one-region queries through the drop-in container: BAMGenomeArray(path, indexed=True)[seg] (seek through the
.bai, upload the region's reads, per-segment operator, read-back) against the same query on the container that decoded
the whole file and mapped whole-genome planes first."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import plastid_b200 as pb  # noqa: E402
from plastid_b200 import bam_io  # noqa: E402


def main():
    rng = np.random.default_rng(1)
    n = 400_000
    lens = {"c1": 20_000_000, "c2": 8_000_000}
    pos = np.sort(rng.integers(0, 19_999_000, n))
    recs = [(0, int(p), 16 if i % 2 else 0, [(0, 30)]) for i, p in enumerate(pos)]
    recs += [(1, int(p), 0, [(0, 28)]) for p in np.sort(rng.integers(0, 7_999_000, n // 2))]
    path = os.path.join(tempfile.mkdtemp(), "q.bam")
    bam_io.write_bam(path, lens, recs, payload_rng=np.random.default_rng(2))
    bam_io.build_index(path)
    segs = []
    for k in range(300):
        c = "c1" if k % 2 else "c2"
        a = int(rng.integers(0, lens[c] - 5000))
        segs.append(pb.GenomicSegment(c, a, a + 2000, "+-"[k % 2]))
    out = {"reads": len(recs), "bam_bytes": os.path.getsize(path)}
    t0 = time.perf_counter()
    lazy = pb.BAMGenomeArray(path, mapping=pb.FivePrimeMapFactory(12), device="cuda:0", indexed=True)
    out["open_indexed_ms"] = 1e3 * (time.perf_counter() - t0)
    lat = []
    for s in segs:
        torch.cuda.synchronize()
        t = time.perf_counter()
        v = lazy[s]
        lat.append(time.perf_counter() - t)
    out["lazy_query_ms_median"] = 1e3 * float(np.median(lat[20:]))
    out["lazy_query_ms_max"] = 1e3 * float(np.max(lat[20:]))
    assert lazy.is_lazy
    t0 = time.perf_counter()
    eager = pb.BAMGenomeArray(path, mapping=pb.FivePrimeMapFactory(12), device="cuda:0")
    first = eager[segs[0]]
    torch.cuda.synchronize()
    out["eager_open_decode_map_first_query_ms"] = 1e3 * (time.perf_counter() - t0)
    lat = []
    for s in segs:
        t = time.perf_counter()
        w = eager[s]
        lat.append(time.perf_counter() - t)
    out["eager_query_ms_median"] = 1e3 * float(np.median(lat[20:]))
    same = all((lazy[s] == eager[s]).all() for s in segs[:50])
    out["same_vectors"] = bool(same)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
