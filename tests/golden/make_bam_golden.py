#!/usr/bin/env python
"""Regenerates tests/golden/htslib_*.{bam,dump.txt}: a seeded SAM with every CIGAR op, unmapped and
placed-unmapped records, converted to BAM and dumped back by the REFERENCE's vendored htslib 1.3
(oracle/_ref/ref_bam_tool, built by `make -C oracle ref`; needs /root/reference).  The committed
outputs pin plastid_b200/csrc/pb_bam.cpp against the reference's own BAM writer/reader."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import random_cigar_reads          # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_bam_tool")
OPS = "MIDNSHP=X"


def main():
    rng = np.random.default_rng(2024)
    lens = {"chrA": 60000, "chrB": 25000, "chrEmpty": 500, "chrC": 9000}
    lines = ["@HD\tVN:1.4\tSO:coordinate"] + ["@SQ\tSN:%s\tLN:%d" % kv for kv in lens.items()]
    n = 0
    for chrom, count in (("chrA", 1500), ("chrB", 600), ("chrC", 200)):
        reads = sorted(random_cigar_reads(rng, count, lens[chrom], lens[chrom] - 2000), key=lambda r: r.reference_start)
        for r in reads:
            qlen = sum(k for op, k in r.cigartuples if op in (0, 1, 4, 7, 8))
            cigar = "".join("%d%s" % (k, OPS[op]) for op, k in r.cigartuples)
            flag = (16 if r.is_reverse else 0) | (256 if n % 37 == 0 else 0) | (1024 if n % 53 == 0 else 0)
            lines.append("r%d\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (n, flag, chrom, r.reference_start + 1, cigar, "A" * qlen))
            if n % 101 == 0:      # an unmapped mate placed at the same coordinate
                lines.append("u%d\t4\t%s\t%d\t0\t*\t*\t0\t0\tACGT\t*" % (n, chrom, r.reference_start + 1))
            n += 1
    for k in range(5):
        lines.append("x%d\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*" % k)
    sam = os.path.join(HERE, "_tmp.sam")
    with open(sam, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    bam = os.path.join(HERE, "htslib_allops.bam")
    subprocess.check_call([TOOL, "sam2bam", sam, bam])
    with open(os.path.join(HERE, "htslib_allops.dump.txt"), "w") as fh:
        subprocess.check_call([TOOL, "dump", bam], stdout=fh)
    os.remove(sam)
    print("wrote", bam, os.path.getsize(bam), "bytes")
    write_positions(bam)
    write_index_and_fetches(bam)


def write_index_and_fetches(bam):
    """htslib_allops.bam.bai: the index htslib's own indexer writes for the golden BAM; htslib_allops.fetch.txt: what
    htslib's iterator (pysam's ``AlignmentFile.fetch``) returns for 300 seeded regions (and a few whole chromosomes) — the records of every region.  Pins
    pb_bam_fetch / pb_bam_build_index (plastid_b200/csrc/pb_bam.cpp) against the reference tree's own code."""
    subprocess.check_call([TOOL, "index", bam])
    rng = np.random.default_rng(77)
    lens = [60000, 25000, 500, 9000]
    args = []
    for k in range(300):
        tid = int(rng.integers(0, 4))
        beg = int(rng.integers(0, lens[tid]))
        width = int(rng.choice([1, 10, 100, 1000]))
        args += [str(tid), str(beg), str(min(lens[tid], beg + width))]
    args += ["0", "0", "60000", "1", "16383", "16385", "3", "8999", "9000", "1", "3000", "25000", "2", "0", "500"]
    out = subprocess.check_output([TOOL, "fetch", bam] + args)
    import gzip
    with open(os.path.join(HERE, "htslib_allops.fetch.txt.gz"), "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fh:
        fh.write(out)
    print("wrote htslib_allops.bam.bai and htslib_allops.fetch.txt.gz:", out.count(b"#"), "regions")


def write_positions(bam):
    """htslib_allops.positions.txt: per read, the reference positions carrying one of its aligned bases
    according to the reference's vendored htslib PILEUP engine (`ref_bam_tool positions`), folded into
    runs: ``qname tid start-end,start-end,...`` (half-open).  Pins AlignedSegment.positions (SURVEY 8a
    row a1) for every CIGAR op against code of the reference tree itself."""
    out = subprocess.check_output([TOOL, "positions", bam]).decode().split("\n")
    per, order = {}, []
    for line in out:
        if not line:
            continue
        name, tid, pos = line.split()
        if name not in per:
            per[name] = (int(tid), [])
            order.append(name)
        per[name][1].append(int(pos))
    with open(os.path.join(HERE, "htslib_allops.positions.txt"), "w") as fh:
        for name in sorted(order, key=lambda q: int(q[1:])):
            tid, pos = per[name]
            pos = np.asarray(sorted(pos))
            cut = np.flatnonzero(np.diff(pos) != 1)
            starts = np.concatenate(([pos[0]], pos[cut + 1]))
            ends = np.concatenate((pos[cut], [pos[-1]])) + 1
            fh.write("%s %d %s\n" % (name, tid, ",".join("%d-%d" % ab for ab in zip(starts, ends))))
    print("wrote htslib_allops.positions.txt:", len(order), "reads")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "index":           # only the index / fetch fixtures of the committed BAM
        write_index_and_fetches(os.path.join(HERE, "htslib_allops.bam"))
    elif len(sys.argv) > 1 and sys.argv[1] == "positions":      # only the pileup-derived positions of the committed BAM
        write_positions(os.path.join(HERE, "htslib_allops.bam"))
    else:
        main()
