#!/usr/bin/env python
"""Golden outputs of the REFERENCE's own command-line programs, produced by running the reference itself.

  python tests/golden/make_script_goldens.py          # build container only (needs /root/reference)

How: ``oracle/build_pyref.py`` compiles the reference's unmodified Cython sources (map_factories, roitools, c_common)
against a stand-in ``pysam`` (``oracle/ref_stubs``; pysam is absent from this image), ``oracle/pyref.py`` makes
``import plastid...`` resolve to those modules plus the reference's pure-Python modules read in place.  This script
then writes a small seeded data set (BED12 transcripts with gene ids, mask BED, a plain-text alignment listing the
stand-in ``pysam.AlignmentFile`` reads) under ``tests/golden/ref_scripts/in/`` and runs the reference's ``main()``
functions on it:

  counts_in_region, cs generate, cs count, metagene generate, metagene count (median and --use_mean, --keep),
  psite (median and --aggregate, --keep), phase_by_size, make_wiggle (wiggle + bedgraph, 5' and center),
  get_count_vectors

Their output files land in ``tests/golden/ref_scripts/out/`` and are committed.  ``tests/test_ref_goldens.py`` replays
the same inputs through ``oracle/`` (CPU suite) and through ``plastid_b200`` (GPU suite) and compares files.
Lines starting with ``##`` carry dates and command lines and are not compared.
"""
import gzip
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
IN = os.path.join(HERE, "ref_scripts", "in")
OUT = os.path.join(HERE, "ref_scripts", "out")

CHROMS = [("chrA", 150000), ("chrB", 90000)]
OFFSETS = {25: 12, 26: 12, 27: 13, 28: 13, 29: 14, 30: 14, 31: 14, 32: 14, 33: 15, 34: 15, 35: 15}


def make_inputs(seed=20261017):
    """Transcripts (BED12 + gene_id column), masks, reads.  Returns nothing; files under IN."""
    rng = np.random.default_rng(seed)
    os.makedirs(IN, exist_ok=True)
    tx_lines, cds_windows = [], []
    n_gene = 0
    for chrom, clen in CHROMS:
        pos = 500
        while pos < clen - 6000:
            strand = "+" if rng.random() < 0.5 else "-"
            n_ex = int(rng.integers(1, 5))
            ex_len = rng.integers(120, 700, n_ex)
            introns = rng.integers(60, 400, n_ex - 1) if n_ex > 1 else np.zeros(0, dtype=int)
            starts = [pos]
            for k in range(1, n_ex):
                starts.append(starts[-1] + int(ex_len[k - 1]) + int(introns[k - 1]))
            exons = [(s, s + int(n)) for s, n in zip(starts, ex_len)]
            gene = "G%03d" % n_gene
            n_gene += 1
            n_iso = int(rng.integers(1, 4))
            total = int(sum(b - a for a, b in exons))
            for iso in range(n_iso):
                ex = list(exons)
                if iso == 1 and len(ex) > 2:                     # skip an internal exon
                    del ex[1]
                if iso == 2:                                     # alternative 5' end (genomic left end)
                    ex[0] = (ex[0][0] + 30, ex[0][1])
                span0, span1 = ex[0][0], ex[-1][1]
                length = sum(b - a for a, b in ex)
                # coding region in transcript-free genomic terms: pick inside the first/last exons where possible
                utr_l = int(rng.integers(20, 90))
                utr_r = int(rng.integers(20, 90))
                cds_a = min(exons[0][0] + 30 + utr_l, ex[0][1] - 10) if iso != 3 else ex[0][0]
                cds_b = max(ex[-1][1] - utr_r, ex[-1][0] + 10)
                if rng.random() < 0.1:
                    cds_a = cds_b = span0                        # non-coding
                sizes = ",".join(str(b - a) for a, b in ex) + ","
                offs = ",".join(str(a - span0) for a, b in ex) + ","
                tx_lines.append("\t".join([chrom, str(span0), str(span1), "%s_t%d" % (gene, iso), "0", strand, str(cds_a), str(cds_b),
                                           "0,0,0", str(len(ex)), sizes, offs, gene]))
            pos = exons[-1][1] + int(rng.integers(200, 1500))
    with open(os.path.join(IN, "transcripts.bed"), "w") as fh:
        fh.write("\n".join(tx_lines) + "\n")
    # masks: 150-nt blocks on both strands, about 6 % of the genome
    mask_lines = []
    for chrom, clen in CHROMS:
        for k, s in enumerate(sorted(rng.integers(0, clen - 200, clen // 2500))):
            for strand in "+-":
                mask_lines.append("\t".join([chrom, str(int(s)), str(int(s) + 150), "mask_%s_%d%s" % (chrom, k, strand), "0", strand]))
    with open(os.path.join(IN, "masks.bed"), "w") as fh:
        fh.write("\n".join(mask_lines) + "\n")
    # reads: 5' ends concentrated in exons with 3-nt periodicity; a tenth spliced, a few with I / D / S ops
    exon_list = []
    for line in tx_lines:
        f = line.split("\t")
        sizes = [int(x) for x in f[10].strip(",").split(",")]
        offs = [int(x) for x in f[11].strip(",").split(",")]
        for o, n in zip(offs, sizes):
            exon_list.append((f[0], int(f[1]) + o, int(f[1]) + o + n, f[5]))
    reads = []
    n_reads = 90000
    for _ in range(n_reads):
        L = int(rng.choice(np.arange(24, 38), p=_len_probs()))
        if rng.random() < 0.85:
            chrom, a, b, strand = exon_list[int(rng.integers(len(exon_list)))]
            p = int(rng.integers(a, max(b - 3, a + 1)))
            p -= (p - a) % 3 if rng.random() < 0.7 else 0
            strand = strand if rng.random() < 0.9 else ("-" if strand == "+" else "+")
        else:
            chrom, clen = CHROMS[int(rng.integers(len(CHROMS)))]
            p = int(rng.integers(0, clen - 2000))
            strand = "+" if rng.random() < 0.5 else "-"
        clen = dict(CHROMS)[chrom]
        u = rng.random()
        if u < 0.10:
            k = int(rng.integers(5, L - 4))
            cigar = "%dM%dN%dM" % (k, int(rng.integers(60, 400)), L - k)
        elif u < 0.13:
            k = int(rng.integers(5, L - 4))
            cigar = "%dM%dD%dM" % (k, int(rng.integers(1, 4)), L - k)
        elif u < 0.16:
            k = int(rng.integers(5, L - 4))
            cigar = "%dM%dI%dM" % (k, int(rng.integers(1, 3)), L - k)
        elif u < 0.19:
            cigar = "%dS%dM%dS" % (int(rng.integers(1, 4)), L, int(rng.integers(0, 3)))
            cigar = cigar.replace("0S", "")
        else:
            cigar = "%dM" % L
        span = L + sum(int(n) for n, op in __import__("re").findall(r"(\d+)([DN])", cigar))
        p = max(0, min(p, clen - span - 1))
        reads.append((chrom, p, strand, cigar))
    order = {c: i for i, (c, _n) in enumerate(CHROMS)}
    reads.sort(key=lambda r: (order[r[0]], r[1]))
    with open(os.path.join(IN, "reads.aln"), "w") as fh:
        for c, n in CHROMS:
            fh.write("@SQ\t%s\t%d\n" % (c, n))
        for r in reads:
            fh.write("%s\t%d\t%s\t%s\n" % r)
    with open(os.path.join(IN, "p_offsets.txt"), "w") as fh:
        fh.write("length\tp_offset\n")
        for k in sorted(OFFSETS):
            fh.write("%d\t%d\n" % (k, OFFSETS[k]))
        fh.write("default\t14\n")


def _len_probs():
    w = np.array([1, 3, 6, 10, 16, 18, 16, 10, 7, 5, 3, 2, 2, 1], dtype=float)
    return w / w.sum()


def _write_gz(path, data):
    with open(path, "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fh:
        fh.write(data)


def run(label, main, argv):
    print("[ref] %s %s" % (label, " ".join(argv)))
    sys.stdout.flush()
    main(argv)


def main():
    from oracle import build_pyref, pyref
    if not build_pyref.build():
        raise SystemExit("the reference cannot be built here")
    pyref.load()
    import importlib
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    make_inputs()
    tx, masks, aln = (os.path.join(IN, x) for x in ("transcripts.bed", "masks.bed", "reads.aln"))
    ann = ["--annotation_files", tx, "--annotation_format", "BED", "--bed_extra_columns", "gene_id"]
    msk = ["--mask_annotation_files", masks, "--mask_annotation_format", "BED"]
    cnt = ["--count_files", aln, "--countfile_format", "BAM", "--min_length", "25", "--max_length", "35"]
    o = lambda name: os.path.join(OUT, name)      # noqa: E731
    mod = lambda name: importlib.import_module("plastid.bin." + name)     # noqa: E731

    run("counts_in_region", mod("counts_in_region").main, [o("counts_in_region_fiveprime14.txt")] + cnt + ["--fiveprime", "--offset", "14"] + ann + msk)
    run("counts_in_region", mod("counts_in_region").main, [o("counts_in_region_center12.txt")] + cnt + ["--center", "--nibble", "12"] + ann + msk)
    run("counts_in_region", mod("counts_in_region").main, [o("counts_in_region_variable.txt")] + cnt
        + ["--fiveprime_variable", "--offset", os.path.join(IN, "p_offsets.txt")] + ann)
    run("cs generate", mod("cs").main, ["generate", o("cs")] + ann + msk)
    run("cs count", mod("cs").main, ["count", o("cs_gene.positions"), o("cs_count_threeprime")] + cnt + ["--threeprime", "--offset", "0"])
    run("cs count", mod("cs").main, ["count", o("cs_gene.positions"), o("cs_count_fiveprime14")] + cnt + ["--fiveprime", "--offset", "14"])
    run("metagene generate", mod("metagene").main, ["generate", o("mg_start"), "--landmark", "cds_start", "--upstream", "50",
                                                    "--downstream", "100"] + ann + msk)
    run("metagene generate", mod("metagene").main, ["generate", o("mg_stop"), "--landmark", "cds_stop", "--upstream", "100",
                                                    "--downstream", "50"] + ann)
    mg = ["--fiveprime_variable", "--offset", os.path.join(IN, "p_offsets.txt")]
    run("metagene count", mod("metagene").main, ["count", o("mg_start_rois.txt"), o("mg_start_median"), "--keep", "--min_counts", "5",
                                                 "--normalize_over", "20", "80"] + cnt + mg)
    run("metagene count", mod("metagene").main, ["count", o("mg_start_rois.txt"), o("mg_start_mean"), "--use_mean", "--min_counts", "5",
                                                 "--normalize_over", "20", "80"] + cnt + mg)
    run("metagene count", mod("metagene").main, ["count", o("mg_stop_rois.txt"), o("mg_stop_center"), "--min_counts", "5",
                                                 "--normalize_over", "-80", "-20"] + cnt + ["--center", "--nibble", "10"])
    ps = ["--min_counts", "5", "--normalize_over", "20", "80", "--require_upstream"]
    run("psite", mod("psite").main, [o("mg_start_rois.txt"), o("psite_median"), "--keep"] + ps + cnt)
    run("psite", mod("psite").main, [o("mg_start_rois.txt"), o("psite_aggregate"), "--aggregate"] + ps + cnt)
    run("phase_by_size", mod("phase_by_size").main, [o("mg_start_rois.txt"), o("phase"), "--codon_buffer", "5"] + cnt
        + ["--fiveprime", "--offset", "14"])
    run("phase_by_size", mod("phase_by_size").main, [o("phase_ann"), "--codon_buffer", "3"] + cnt + ["--fiveprime_variable", "--offset",
                                                                                                    os.path.join(IN, "p_offsets.txt")] + ann)
    run("make_wiggle", mod("make_wiggle").main, ["-o", o("wig_fiveprime14"), "--output_format", "variable_step"] + cnt + ["--fiveprime", "--offset", "14"])
    run("make_wiggle", mod("make_wiggle").main, ["-o", o("bg_center12"), "--output_format", "bedgraph"] + cnt + ["--center", "--nibble", "12"])
    run("make_wiggle", mod("make_wiggle").main, ["-o", o("bg_threeprime_norm"), "--output_format", "bedgraph", "--normalize"] + cnt + ["--threeprime"])
    os.makedirs(o("count_vectors"), exist_ok=True)
    run("get_count_vectors", mod("get_count_vectors").main, [o("count_vectors")] + cnt + ["--fiveprime", "--offset", "14"] + ann + msk
        + ["--out_prefix", "cv_", "--format", "%d"])
    # committed form: no "##" provenance lines (dates, absolute paths, the whole argparse namespace), gzip members
    # without time stamps — the files are then a pure function of the seed
    for dirpath, _dirs, files in os.walk(OUT):
        for name in sorted(files):
            path = os.path.join(dirpath, name)
            if name.endswith((".png", ".svg", ".pdf")) or "<" in name:
                os.remove(path)
                continue
            if name.endswith(".gz"):
                with gzip.open(path, "rb") as fh:
                    data = fh.read()
                with open(path, "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fh:
                    fh.write(data)
                continue
            with open(path) as fh:
                lines = [ln for ln in fh if not ln.startswith("##") or ln.startswith("## total_dataset_counts")]
            if name.endswith(".wig"):              # tracks are large: committed compressed
                os.remove(path)
                _write_gz(path + ".gz", "".join(lines).encode())
                continue
            with open(path, "w") as fh:
                fh.writelines(lines)
    # get_count_vectors writes one file per region: bundled as "<file name>\t<space-separated values>" lines
    cv = o("count_vectors")
    bundle = []
    for name in sorted(os.listdir(cv)):
        with open(os.path.join(cv, name)) as fh:
            bundle.append("%s\t%s\n" % (name, " ".join(fh.read().split())))
    shutil.rmtree(cv)
    _write_gz(o("count_vectors.txt.gz"), "".join(bundle).encode())
    with open(aln, "rb") as fh:
        data = fh.read()
    os.remove(aln)
    _write_gz(aln + ".gz", data)
    print("[ref] wrote %d files under %s" % (sum(len(f) for _r, _d, f in os.walk(OUT)), OUT))


if __name__ == "__main__":
    main()
