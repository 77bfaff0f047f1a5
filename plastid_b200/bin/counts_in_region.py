"""``counts_in_region``: masked counts, counts per nucleotide and RPKM per region
(plastid/bin/counts_in_region.py:63-131), all regions in one gather launch."""
import argparse
import sys

import numpy as np

from . import _cli


def format_row(name, region, counts, length, normconst):
    """counts_in_region.py:120-124: rpnt = counts/length, rpkm = rpnt * 1e9/sum; '%.8e' x3, '%d'."""
    rpnt = np.nan if length == 0 else float(counts) / length
    rpkm = np.nan if length == 0 else rpnt * normconst
    return [name, region, "%.8e" % counts, "%.8e" % rpnt, "%.8e" % rpkm, "%d" % length]


def count_regions(ga, chains, masks=None, mask_features=None):
    """Rows of the output table for ``chains`` (SegmentChains; ``masks[i]`` = mask segments of
    chain i, added with ``add_masks`` exactly as counts_in_region.py:115-116 does).
    ``mask_features``: the mask annotation itself (SegmentChains) — the overlap query and the
    masking of every region then run on the device (``plastid_b200.masks``), no per-region work.

    Returns ``(ga_sum, rows)``; every row is ``[name, region, counts, rpnt, rpkm, length]`` already
    formatted (``%.8e`` x3, ``%d``) as the reference writes it."""
    ga_sum = ga.sum()
    normconst = 1000.0 * 1e6 / ga_sum
    if masks is not None:
        for ch, m in zip(chains, masks):
            if m:
                ch.add_masks(*m)
    if mask_features is not None:
        from ..masks import MaskIndex, apply_mask_index
        # regions on chromosomes the alignments do not know count zero over their unmasked length
        # (genome_array.py:795-798): their masks never reach the device, so they are applied here
        unknown = [ch for ch in chains if len(ch) and ch.chrom not in ga.layout.index]
        for ch, m in zip(unknown, overlapping_masks(unknown, mask_features)):
            if m:
                ch.add_masks(*m)
        table = ga.chain_table(chains)
        apply_mask_index(table, MaskIndex(mask_features, ga.layout), ga.device)
        sums, live = ga.count_chains(table)
    else:
        sums, live = ga.count_chains(chains)
    rows = []
    for ch, counts, length in zip(chains, sums, live):
        if length == 0 and ch.length > 0:
            counts = np.nan            # nansum of a fully masked vector formats as 'nan' 
        rows.append(format_row(ch.get_name(), str(ch), counts, length, normconst))
    return ga_sum, rows


def write_table(fout, ga_sum, rows):
    fout.write("## total_dataset_counts: %s\n" % ga_sum)
    fout.write("region_name\tregion\tcounts\tcounts_per_nucleotide\trpkm\tlength\n")
    for row in rows:
        fout.write("%s\n" % "\t".join(row))


def main(argv=sys.argv[1:]):
    """Same command line as plastid/bin/counts_in_region.py:63-131 (``--normalize`` is disabled there, :60).
    Under ``torchrun`` every rank counts the positions of its own genome range; rank 0 writes the table."""
    parser = argparse.ArgumentParser(description=__doc__)
    _cli.add_base_args(parser)
    _cli.add_alignment_args(parser, disabled=("normalize",))
    _cli.add_annotation_args(parser)
    _cli.add_mask_args(parser)
    parser.add_argument("outfile", type=str, help="Output filename")
    args = parser.parse_args(argv)
    ga = _cli.genome_array_from_args(args, disabled=("normalize",))
    chains = _cli.chains_from_args(args)
    mask_chains = _cli.chains_from_args(args, prefix="mask_") if args.mask_annotation_files else None
    ga_sum, rows = count_regions(ga, chains, mask_features=mask_chains)
    if _cli.is_writer():
        with open(args.outfile, "w") as fout:
            write_table(fout, ga_sum, rows)
    _cli.finish_distributed()


def overlapping_masks(chains, mask_chains):
    """Host statement of the same query, per chain: the segments of the mask features on the chain's
    chromosome and strand that reach into its span — what ``GenomeHash.get_overlapping_features``
    (plastid/genomics/genome_hash.py:259-436) hands to ``add_masks`` at counts_in_region.py:114-115.
    The hash keeps '+' and '-' tables only (genome_hash.py:236-257): a '.' mask feature is a KeyError
    there and here."""
    by_key = {}
    for mc in mask_chains:
        for seg in mc:
            if seg.strand not in ("+", "-"):
                raise KeyError(seg.strand)
            by_key.setdefault(seg.chrom, []).append(seg)
    for segs in by_key.values():
        segs.sort(key=lambda s: s.start)
    out = []
    from ..roitools import GenomicSegment
    for ch in chains:
        hits = []
        if len(ch):
            lo, hi = ch.spanning_segment.start, ch.spanning_segment.end
            for seg in by_key.get(ch.chrom, ()):
                if seg.start >= hi:
                    break
                if seg.end > lo and seg.strand == ch.strand:
                    hits.append(GenomicSegment(seg.chrom, seg.start, seg.end, ch.strand))
        out.append(hits)
    return out


if __name__ == "__main__":
    main()
