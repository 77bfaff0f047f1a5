"""The three inputs round 1 refused, pinned to the REFERENCE ITSELF (tests/golden/make_quirk_goldens.py ran plastid's
unmodified programs on them; outputs under tests/golden/ref_quirks/out) and reproduced by the oracle restatements:

* ``metagene generate`` on unstranded transcripts (window columns laid in reverse, offsets forward — metagene.py:443-455);
* ``phase_by_size`` on reads across the exon junctions of a coding region (``read_dict`` is not reset between exons, so
  such a read counts twice in the later exon — phase_by_size.py:186-194), point rules and ``--center``;
* ``cs generate`` on genes whose transcripts lie on several chromosomes / strands (pooled on the first one's place —
  cs.py:324-343).

The product's programs reproduce the same files on the GPU: tests/test_gpu_ref_goldens.py."""
import gzip
import os
import warnings

import numpy as np
import pytest

from oracle import generate as og
from oracle import pyoracle as po
from oracle import scripts as osc
import refgold as rg

QIN = os.path.join(rg.HERE, "golden", "ref_quirks", "in")
QOUT = os.path.join(rg.HERE, "golden", "ref_quirks", "out")


def qout(name):
    return os.path.join(QOUT, name)


def read_bed_transcripts(path):
    txs = []
    with open(path) as fh:
        rows = [ln.rstrip("\n").split("\t") for ln in fh if ln.strip()]
    for f in rows:
        chrom, start, strand = f[0], int(f[1]), f[5]
        sizes = [int(x) for x in f[10].strip(",").split(",")]
        offs = [int(x) for x in f[11].strip(",").split(",")]
        segs = [po.Seg(chrom, start + o, start + o + n, strand) for o, n in zip(offs, sizes)]
        attr = dict(ID=f[3], gene_id=f[12])
        if int(f[6]) < int(f[7]):
            attr.update(cds_genome_start=int(f[6]), cds_genome_end=int(f[7]))
        txs.append(og.Tx(*segs, **attr))
    return txs


def read_masks():
    masks = []
    for f in rg.bed_rows("masks.bed"):
        m = po.Chain(po.Seg(f[0], int(f[1]), int(f[2]), f[5]))
        m.name = f[3]
        masks.append(m)
    return masks


def junction_store():
    chrom_lengths, by_chrom = {}, {}
    with gzip.open(os.path.join(QIN, "junction_reads.aln.gz"), "rt") as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if f[0] == "@SQ":
                chrom_lengths[f[1]] = int(f[2])
                by_chrom[f[1]] = []
            elif line.strip():
                cigar = [(rg._CIGAR_OPS.index(op), int(n)) for n, op in rg._CIGAR_RE.findall(f[3])]
                by_chrom[f[0]].append(po.Read(int(f[1]), cigar, f[2] == "-"))
    return po.ReadStore(chrom_lengths, by_chrom)


@pytest.mark.parametrize("landmark,up,down,masked", [("cds_start", 50, 100, True), ("cds_stop", 100, 50, False)])
def test_oracle_metagene_generate_on_unstranded_transcripts(landmark, up, down, masked):
    txs = read_bed_transcripts(os.path.join(QIN, "transcripts_unstranded.bed"))
    fn = {"cds_start": og.window_cds_start, "cds_stop": og.window_cds_stop}[landmark]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rows = og.group_regions_make_windows(txs, po.GenomeHash(read_masks() if masked else []), up, down, window_func=fn)
    head, want = rg.table(qout("mgu_%s_rois.txt" % landmark.split("_")[1]))
    assert len(want) > 5
    assert [[str(r[h]) for h in head] for r in rows] == want


def check_phasing(sums, fname):
    """The table as phase_by_size.py:216-258 writes it: ``reads_counted`` is an INTEGER column (fractional Center sums
    are truncated on assignment, :218-219) and the fractions are taken from the truncated numbers."""
    head, want = rg.table(qout(fname))
    lengths = sorted(sums)
    counted = np.array([int(sums[k].sum()) for k in lengths])
    fmt = lambda v: "nan" if np.isnan(v) else "%.6f" % v      # noqa: E731
    for i, k in enumerate(lengths):
        with np.errstate(all="ignore"):
            ph = sums[k].astype(float) / sums[k].astype(float).sum()
            frac = float(counted[i]) / counted.sum() if counted.sum() else float("nan")
        got = ["%d" % k, "%d" % counted[i], fmt(frac)] + [fmt(v) for v in ph]
        assert got[:2] == want[i][:2], (fname, got, want[i])
        for g, t in zip(got[2:], want[i][2:]):
            assert g == t or abs(float(g) - float(t)) <= 1.01e-6, (fname, got, want[i])


PHASE_CASES = [("phase_junction_fiveprime", lambda: po.FivePrimeMap(14), 3),
               ("phase_junction_threeprime", lambda: po.ThreePrimeMap(3), 0),      # `[0:-0]` selects nothing: a table of nan
               ("phase_junction_threeprime2", lambda: po.ThreePrimeMap(3), 2),
               ("phase_junction_center", lambda: po.CenterMap(10), 3)]


@pytest.mark.parametrize("tag,rule,buffer", PHASE_CASES)
def test_oracle_phase_by_size_counts_junction_reads_like_the_reference(tag, rule, buffer):
    ga = po.OracleBAMGenomeArray(junction_store(), mapping=rule())
    ga.add_filter("size:25-35", po.SizeFilter(25, 35))
    cds = [og.tx_cds(tx) for tx in read_bed_transcripts(rg.inp("transcripts.bed"))]
    sums = osc.phase_by_size(ga, cds, list(range(25, 36)), buffer, -buffer)
    check_phasing(sums, tag + "_phasing.txt")


def test_oracle_phase_by_size_roi_file_center_rule():
    ga = po.OracleBAMGenomeArray(junction_store(), mapping=po.CenterMap(8))
    ga.add_filter("size:25-35", po.SizeFilter(25, 35))
    cols = rg.columns(rg.out("mg_start_rois.txt"))
    cds = []
    for region, offset, zero in zip(cols["region"], cols["alignment_offset"], cols["zero_point"]):
        chain = po.Chain.from_str(region)
        cds.append(og.get_subchain(chain, int(zero) - int(round(float(offset))), chain.length))
    check_phasing(osc.phase_by_size(ga, cds, list(range(25, 36)), 5, -1), "phase_junction_roi_center_phasing.txt")


def test_oracle_cs_generate_pools_genes_in_several_places_like_the_reference():
    txs = read_bed_transcripts(os.path.join(QIN, "transcripts_multi.bed"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        genes, txrows, merged = og.cs_process_partial_group({tx.get_name(): tx for tx in txs}, po.GenomeHash(read_masks()))
    for fname, rows in (("csm_gene.positions", genes), ("csm_transcript.positions", txrows)):
        head, want = rg.table(qout(fname))
        got = [[r[h] for h in head] for r in rows]
        assert got == want, fname
    gx = [r for r in want if r[0].startswith("GX")]                 # transcript rows: both on the first transcript's place
    assert len(gx) == 2 and all("chrA:" in r[head.index("exon")] and "(+)" in r[head.index("exon")] for r in gx)
