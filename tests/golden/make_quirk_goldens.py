#!/usr/bin/env python
"""Golden outputs of the REFERENCE's own programs on the three inputs round 1 refused (missing 6):

  python tests/golden/make_quirk_goldens.py          # build container only (needs /root/reference)

1. ``metagene generate`` on UNSTRANDED transcripts (strand '.'): the reference lays their window columns in
   reverse while it computes offsets forward (plastid/bin/metagene.py:443-455).
2. ``phase_by_size`` on reads that span exon junctions of the coding region: ``read_dict`` is not reset between
   the exons of a CDS (plastid/bin/phase_by_size.py:186-194), so a read fetched for two exons is mapped twice
   against the later one — with a point rule AND with ``--center``.
3. ``cs generate`` on genes whose transcripts lie on several chromosomes / strands: "Skipping gene ..." is
   printed and the positions of all places are pooled on the first transcript's chromosome and strand
   (plastid/bin/cs.py:324-343).

Same machinery as make_script_goldens.py (the unmodified reference, run through oracle/pyref.py); inputs are
derived from that script's seeded transcripts.  Files land in tests/golden/ref_quirks/{in,out} and are committed.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
SRC = os.path.join(HERE, "ref_scripts", "in")
IN = os.path.join(HERE, "ref_quirks", "in")
OUT = os.path.join(HERE, "ref_quirks", "out")
CHROMS = [("chrA", 150000), ("chrB", 90000)]


def make_inputs(seed=20261018):
    rng = np.random.default_rng(seed)
    os.makedirs(IN, exist_ok=True)
    with open(os.path.join(SRC, "transcripts.bed")) as fh:
        tx = [ln.rstrip("\n").split("\t") for ln in fh if ln.strip()]
    # 1. the same transcripts without a strand
    with open(os.path.join(IN, "transcripts_unstranded.bed"), "w") as fh:
        for f in tx:
            fh.write("\t".join(f[:5] + ["."] + f[6:]) + "\n")
    # (masks stay stranded: the reference's GenomeHash raises KeyError on a '.' feature, genome_hash.py:254)
    # 2. reads across the exon junctions of the transcripts (both strands), and plain reads inside exons
    reads = []
    for f in tx:
        chrom, start, strand = f[0], int(f[1]), f[5]
        sizes = [int(x) for x in f[10].strip(",").split(",")]
        offs = [int(x) for x in f[11].strip(",").split(",")]
        exons = [(start + o, start + o + n) for o, n in zip(offs, sizes)]
        for (a0, b0), (a1, b1) in zip(exons[:-1], exons[1:]):
            for _ in range(12):
                L = int(rng.integers(25, 36))
                k = int(rng.integers(3, L - 2))
                s = strand if rng.random() < 0.9 else ("-" if strand == "+" else "+")
                reads.append((chrom, b0 - k, s, "%dM%dN%dM" % (k, a1 - b0, L - k)))
            if rng.random() < 0.3 and a1 - b0 < 200:                # unspliced across a short intron
                L = 35
                reads.append((chrom, b0 - 5, strand, "%dM" % L))
        for a, b in exons:
            for _ in range(8):
                L = int(rng.integers(25, 36))
                p = int(rng.integers(a, max(b - L, a + 1)))
                reads.append((chrom, p, strand if rng.random() < 0.9 else ("-" if strand == "+" else "+"), "%dM" % L))
    order = {c: i for i, (c, _n) in enumerate(CHROMS)}
    reads.sort(key=lambda r: (order[r[0]], r[1]))
    with open(os.path.join(IN, "junction_reads.aln"), "w") as fh:
        for c, n in CHROMS:
            fh.write("@SQ\t%s\t%d\n" % (c, n))
        for r in reads:
            fh.write("%s\t%d\t%s\t%s\n" % r)
    # 3. the stranded transcripts plus genes in several places: GX on chrA '+' and chrB '-', GY on both strands of
    # chrB (overlapping each other), GZ twice on chrA '+' and once on chrB '+' (longer chromosome first in file order)
    extra = [
        ["chrA", "140000", "141500", "GX_t0", "0", "+", "140100", "141400", "0,0,0", "2", "600,500,", "0,1000,", "GX"],
        ["chrB", "80000", "81800", "GX_t1", "0", "-", "80200", "81700", "0,0,0", "3", "300,400,500,", "0,600,1300,", "GX"],
        ["chrB", "83000", "84000", "GY_t0", "0", "+", "83100", "83900", "0,0,0", "1", "1000,", "0,", "GY"],
        ["chrB", "83500", "84700", "GY_t1", "0", "-", "83600", "84600", "0,0,0", "2", "400,500,", "0,700,", "GY"],
        ["chrA", "143000", "144200", "GZ_t0", "0", "+", "143050", "144100", "0,0,0", "2", "500,400,", "0,800,", "GZ"],
        ["chrB", "86000", "87500", "GZ_t1", "0", "+", "86100", "87400", "0,0,0", "2", "700,600,", "0,900,", "GZ"],
        ["chrA", "143200", "144600", "GZ_t2", "0", "+", "143300", "144500", "0,0,0", "1", "1400,", "0,", "GZ"],
    ]
    with open(os.path.join(IN, "transcripts_multi.bed"), "w") as fh:
        for f in tx + extra:
            fh.write("\t".join(f) + "\n")


def main():
    from oracle import build_pyref, pyref
    import make_script_goldens as msg
    if not build_pyref.build():
        raise SystemExit("the reference cannot be built here")
    pyref.load()
    import importlib
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    make_inputs()
    o = lambda name: os.path.join(OUT, name)      # noqa: E731
    i = lambda name: os.path.join(IN, name)       # noqa: E731
    mod = lambda name: importlib.import_module("plastid.bin." + name)     # noqa: E731
    bed = ["--annotation_format", "BED", "--bed_extra_columns", "gene_id"]
    stranded = ["--annotation_files", os.path.join(SRC, "transcripts.bed")] + bed
    cnt = ["--count_files", i("junction_reads.aln"), "--countfile_format", "BAM", "--min_length", "25", "--max_length", "35"]
    # 1
    msg.run("metagene generate", mod("metagene").main, ["generate", o("mgu_start"), "--landmark", "cds_start", "--upstream", "50",
                                                        "--downstream", "100", "--annotation_files", i("transcripts_unstranded.bed")] + bed
            + ["--mask_annotation_files", os.path.join(SRC, "masks.bed"), "--mask_annotation_format", "BED"])
    msg.run("metagene generate", mod("metagene").main, ["generate", o("mgu_stop"), "--landmark", "cds_stop", "--upstream", "100",
                                                        "--downstream", "50", "--annotation_files", i("transcripts_unstranded.bed")] + bed)
    # 2
    msg.run("phase_by_size", mod("phase_by_size").main, [o("phase_junction_fiveprime"), "--codon_buffer", "3"] + cnt
            + ["--fiveprime", "--offset", "14"] + stranded)
    msg.run("phase_by_size", mod("phase_by_size").main, [o("phase_junction_threeprime"), "--codon_buffer", "0"] + cnt
            + ["--threeprime", "--offset", "3"] + stranded)
    msg.run("phase_by_size", mod("phase_by_size").main, [o("phase_junction_threeprime2"), "--codon_buffer", "2"] + cnt
            + ["--threeprime", "--offset", "3"] + stranded)
    msg.run("phase_by_size", mod("phase_by_size").main, [o("phase_junction_center"), "--codon_buffer", "3"] + cnt
            + ["--center", "--nibble", "10"] + stranded)
    msg.run("phase_by_size", mod("phase_by_size").main, [os.path.join(HERE, "ref_scripts", "out", "mg_start_rois.txt"),
                                                         o("phase_junction_roi_center"), "--codon_buffer", "5"] + cnt + ["--center", "--nibble", "8"])
    # 3
    msg.run("cs generate", mod("cs").main, ["generate", o("csm"), "--annotation_files", i("transcripts_multi.bed")] + bed
            + ["--mask_annotation_files", os.path.join(SRC, "masks.bed"), "--mask_annotation_format", "BED"])
    for dirpath, _dirs, files in os.walk(OUT):
        for name in sorted(files):
            path = os.path.join(dirpath, name)
            if name.endswith((".png", ".svg", ".pdf")) or "<" in name:
                os.remove(path)
                continue
            with open(path) as fh:
                lines = [ln for ln in fh if not ln.startswith("##")]
            with open(path, "w") as fh:
                fh.writelines(lines)
    aln = i("junction_reads.aln")
    with open(aln, "rb") as fh:
        data = fh.read()
    os.remove(aln)
    msg._write_gz(aln + ".gz", data)


if __name__ == "__main__":
    main()
