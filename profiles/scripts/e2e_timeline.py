"""Where the end-to-end C2 step goes (host clock around the API calls; count_planes ends with the statistics read-back,
i.e. a device synchronisation): container construction, upload + expansion + mapping, region table + read-back."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import plastid_b200 as pb  # noqa: E402
from plastid_b200 import synth  # noqa: E402


def main():
    args = bench.parse_args()
    dev = "cuda:0"
    W = bench.build_world(args, 0, dev)
    hb = synth.device_batch_to_host(W["dbatch"], W["chroms"], W["lens"])
    W["dbatch"] = None
    torch.cuda.empty_cache()
    hb.pack()
    hb.transfer_pinned()
    table = W["table"]
    table.device(dev)
    rows = []
    for it in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ga = pb.BAMGenomeArray(hb, mapping=W["fac"], device=dev, shard=None)
        ga.add_filter("size", W["sf"])
        t1 = time.perf_counter()
        ga.count_planes(("+", "-"))
        t2 = time.perf_counter()
        sums, live = ga.count_chains(table, planes=True)
        t3 = time.perf_counter()
        rows.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0)])
        del ga
    r = np.asarray(rows[3:])
    print(json.dumps({"construct_ms": float(r[:, 0].mean()), "count_planes_ms": float(r[:, 1].mean()),
                      "count_chains_ms": float(r[:, 2].mean()), "total_ms": float(r[:, 3].mean()),
                      "transfer_bytes": int(hb.transfer.nbytes)}))


if __name__ == "__main__":
    main()
