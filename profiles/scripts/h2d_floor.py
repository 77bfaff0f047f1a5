#!/usr/bin/env python
"""Measurement aid (not part of the product): how long does the PCIe leg of the C2 end-to-end step take on
its own?  (a) one cudaMemcpyAsync of a pinned buffer of the delta3 batch's size, (b) the chunked receive
pattern of the e2e leg (7 arrays per chunk) without any mapping, (c) the delta3 expansion alone.
    python profiles/scripts/h2d_floor.py [n_reads]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import plastid_b200 as pb                                           # noqa: E402
from plastid_b200 import synth                                      # noqa: E402
from plastid_b200.batch import Delta3Batch, Delta3Receiver          # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
    dev = "cuda:0"
    chroms, lens = synth.human_like_genome(1.0)
    ann = synth.make_annotation(chroms, lens, 60_000, seed=0, exons=(1, 3), exon_len=(150, 600), intron_len=(100, 3000))
    dbatch = synth.riboseq_reads(ann, n_reads, seed=0, device=dev, frac_in=0.85)
    layout = pb.GenomeLayout(chroms, lens)
    wire = Delta3Batch.from_batch(synth.device_batch_to_host(dbatch, chroms, lens))
    pinned = wire.pinned()
    rx = Delta3Receiver(wire, dev)
    out = {"n_reads": n_reads, "bytes": wire.nbytes}
    big = torch.empty(wire.nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(wire.nbytes, dtype=torch.uint8, device=dev)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return 1000.0 * (time.perf_counter() - t0) / reps

    out["one_copy_ms"] = timed(lambda: dst.copy_(big, non_blocking=True))
    out["one_copy_GBps"] = wire.nbytes / out["one_copy_ms"] / 1e6
    for n_chunks in (1, 8, 16):
        chunks = Delta3Receiver.plan_chunks(wire, layout, n_chunks)

        def receive():
            rx._receive_tables(pinned)
            for a, b, _x, _y in chunks:
                rx._copy_range(pinned, a, b)
        out["chunked_copy_ms_%d" % n_chunks] = timed(receive)
    rx.receive(pinned)
    out["unpack_ms"] = timed(lambda: rx._unpack(0, len(wire)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
