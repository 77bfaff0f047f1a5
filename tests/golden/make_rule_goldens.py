#!/usr/bin/env python
"""Golden vectors of the reference's OWN mapping rules and BAMGenomeArray (build container only).

  python tests/golden/make_rule_goldens.py      ->  tests/golden/ref_rules.json.gz

Runs plastid's unmodified ``map_factories.pyx`` (compiled by oracle/build_pyref.py against the stand-in pysam) and
``genome_array.py`` on seeded reads carrying every CIGAR op, and records for a set of query segments: the count
vector and the number of reads every rule returns (``map_fn(reads, seg)``), what ``BAMGenomeArray`` returns for the
same segments with a size filter / normalisation, and ``to_genome_array`` (its last-base quirk).
tests/test_ref_goldens.py replays them through the oracle (CPU) and tests/test_gpu_ref_goldens.py through the CUDA path.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CHROM, CHROM_LEN = "chrR", 6000
OFFSETS_DEFAULT = {25: 12, 26: 12, 27: 13, 28: 13, 29: 14, 30: 14, 31: 14, "default": 13}
OFFSETS_PLAIN = {26: 12, 28: 13, 29: 14, 30: 14, 33: 15}                 # lengths without an entry are dropped (warning)


def make_reads(seed=11, n=600):
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n):
        L = int(rng.integers(18, 42))
        start = int(rng.integers(0, CHROM_LEN - 900))
        u = rng.random()
        k = int(rng.integers(4, L - 3))
        if u < 0.45:
            cigar = "%dM" % L
        elif u < 0.60:
            cigar = "%dM%dN%dM" % (k, int(rng.integers(30, 500)), L - k)
        elif u < 0.68:
            cigar = "%dM%dD%dM" % (k, int(rng.integers(1, 5)), L - k)
        elif u < 0.76:
            cigar = "%dM%dI%dM" % (k, int(rng.integers(1, 4)), L - k)
        elif u < 0.84:
            cigar = "%dS%dM%dS" % (int(rng.integers(1, 5)), L, int(rng.integers(1, 4)))
        elif u < 0.90:
            cigar = "%d=%dX%d=" % (k, 1, max(L - k - 1, 1))
        elif u < 0.95:
            j = int(rng.integers(2, max(k - 1, 3)))
            cigar = "%dM%dN%dM%dN%dM" % (j, int(rng.integers(20, 200)), max(k - j, 1), int(rng.integers(20, 200)), L - k)
        else:
            cigar = "%dH%dM%dP%dM" % (2, k, 1, L - k)
        reads.append((start, "-" if rng.random() < 0.5 else "+", cigar))
    reads.sort(key=lambda r: r[0])
    return reads


SEGMENTS = [(0, CHROM_LEN, "+"), (0, CHROM_LEN, "-"), (0, CHROM_LEN, "."), (1000, 1700, "+"), (2500, 2501, "-"), (3000, 4200, "."),
            (5200, 5990, "-")]


def main():
    from oracle import build_pyref, pyref
    if not build_pyref.build():
        raise SystemExit("the reference cannot be built here")
    m = pyref.modules()
    import pysam
    mf, rt, gam = m["map_factories"], m["roitools"], m["genome_array"]
    reads_in = make_reads()
    reads = [pysam.AlignedSegment(s, pysam.parse_cigar(c), strand == "-", "r%d" % i) for i, (s, strand, c) in enumerate(reads_in)]
    rules = {
        "fiveprime0": mf.FivePrimeMapFactory(0), "fiveprime14": mf.FivePrimeMapFactory(14), "fiveprime30": mf.FivePrimeMapFactory(30),
        "threeprime0": mf.ThreePrimeMapFactory(0), "threeprime15": mf.ThreePrimeMapFactory(15),
        "center0": mf.CenterMapFactory(0), "center12": mf.CenterMapFactory(12),
        "variable_default": mf.VariableFivePrimeMapFactory(dict(OFFSETS_DEFAULT)),
        "variable_plain": mf.VariableFivePrimeMapFactory(dict(OFFSETS_PLAIN)),
        "stratified": mf.StratifiedVariableFivePrimeMapFactory(dict(OFFSETS_DEFAULT), 25, 35),
    }
    out = {"chrom": CHROM, "chrom_len": CHROM_LEN, "reads": reads_in, "segments": SEGMENTS,
           "offsets_default": {str(k): v for k, v in OFFSETS_DEFAULT.items()}, "offsets_plain": {str(k): v for k, v in OFFSETS_PLAIN.items()},
           "operator": {}, "container": {}}
    for name, fn in rules.items():
        rows = []
        for a, b, strand in SEGMENTS:
            seg = rt.GenomicSegment(CHROM, a, b, strand)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kept, counts = fn(list(reads), seg)                     # a bare operator call: no strand pre-filter
            rows.append({"n_reads_out": len(kept), "counts": np.asarray(counts).tolist()})
        out["operator"][name] = rows
    # the container: strand pre-filter, size filter, normalisation, sum, to_genome_array
    bam = pysam.AlignmentFile([CHROM], [CHROM_LEN], {CHROM: reads})
    for name in ("fiveprime14", "threeprime0", "center12", "variable_default"):
        ga = gam.BAMGenomeArray(bam, mapping=rules[name])
        ga.add_filter("size", mf.SizeFilterFactory(min=22, max=36))
        rows = []
        for a, b, strand in SEGMENTS:
            seg = rt.GenomicSegment(CHROM, a, b, strand)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kept, counts = ga.get_reads_and_counts(seg)
                ga.set_normalize(True)
                norm = ga[seg]
                ga.set_normalize(False)
            rows.append({"n_reads_out": len(kept), "counts": np.asarray(counts).tolist(), "normalized": np.asarray(norm).tolist()})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dense = ga.to_genome_array()
            tail = {s: np.asarray(dense[rt.GenomicSegment(CHROM, CHROM_LEN - 400, CHROM_LEN, s)]).tolist() for s in "+-"}
            dsum = float(dense.sum())
        out["container"][name] = {"sum": ga.sum(), "segments": rows, "to_genome_array_tail": tail, "to_genome_array_sum": dsum}
    import gzip
    path = os.path.join(HERE, "ref_rules.json.gz")
    with open(path, "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote %s (%d bytes)" % (path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
