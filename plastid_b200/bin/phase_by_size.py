"""``phase_by_size``: sub-codon phasing of read 5' ends (or any point rule) per read length
(plastid/bin/phase_by_size.py:165-235)."""
import argparse
import sys

import numpy as np

from . import _cli
import torch

from ..genome_array import stratified_windows
from ..map_factories import CenterMapFactory
from ..regions import ChainTable
from ..roitools import SegmentChain


def do_phase(ga, cds_chains, read_lengths, codon_buffer=5, back_buffer=None):
    """Phase sums per read length over the CDS chains: ``{length: float64[3]}``.

    ``back_buffer`` is the python slice stop the reference uses (``-1`` with an ROI file,
    ``-codon_buffer`` with an annotation; note ``-0`` selects nothing there, and here).

    The reference does not reset its per-length read lists between the exons of a coding region
    (phase_by_size.py:186-194).  The lists hold the rule's ``reads_out``: for point rules the reads whose site lies in
    the exon (they have no site in a later exon — nothing changes), for CenterMapFactory every read the fetch returned,
    so a read across an exon junction is mapped once more against the later exon.  The kernel counts such a read with
    that multiplicity under the Center rule (test_gpu_ref_goldens.py::test_phase_by_size_junction_reads).  Point rules
    count sites; CenterMapFactory counts every trimmed position as an integer per read length and the weight
    ``1 / (L - 2 nibble)`` is applied here."""
    back_buffer = -codon_buffer if back_buffer is None else back_buffer
    chains = [c for c in cds_chains if len(c) > 0]
    read_lengths = list(read_lengths)
    if not read_lengths:
        return {}
    table = ChainTable.from_chains(chains, ga.layout, use_masks=False)
    lo, hi = min(read_lengths), max(read_lengths)
    # all lengths and all chains in one launch: [n_len, n_chains, 3] -> sum over chains
    strat, _ = stratified_windows(ga._device_batch(), ga.layout, ga.map_fn, ga._size_filter(), table, None, 3,
                                  lo, hi, phase=(codon_buffer, back_buffer), bin_range=ga.bin_range)
    # multi-GPU: an 11 x 3 table per rank (the sites of its own genome range), completed with one all-reduce
    sums = ga._allreduce(strat.to(torch.float64).sum(dim=1)).cpu().numpy()
    if isinstance(ga.map_fn, CenterMapFactory):
        for k in range(lo, hi + 1):
            m = k - 2 * ga.map_fn.nibble
            sums[k - lo] = sums[k - lo] / float(m) if m > 0 else 0.0           # integer count / m: correctly rounded
    return {k: sums[k - lo] for k in read_lengths}


def phase_table(sums):
    """phase_by_size.py:216-235: ``reads_counted`` is an integer column (the fractional sums of the Center rule are
    truncated when they are assigned to it), and the fractions are taken from it.  Under the Center rule the reference's
    float sum of ``1 / m`` weights sits a rounding error above or below an integer whenever whole reads lie inside the
    region; here the truncation is applied to the exact value (counts / m, a 1e-9 guard against the last bit)."""
    lengths = sorted(sums)
    counted = np.array([int(np.floor(sums[k].sum() + 1e-9)) for k in lengths], dtype=np.int64)
    with np.errstate(all="ignore"):
        frac = counted.astype(float) / counted.sum()
        phases = np.array([sums[k].astype(float) / sums[k].astype(float).sum() for k in lengths])
    return lengths, counted, frac, phases


def main(argv=sys.argv[1:]):
    """``phase_by_size [ROI_FILE] OUTBASE --count_files ... [--annotation_files ...]`` with the reference's flags
    (plastid/bin/phase_by_size.py:79-262): regions from an ROI file of ``metagene generate`` (CDS part of every
    window, last codon dropped) or the coding regions of an annotation (``codon_buffer`` codons dropped at both ends)."""
    parser = argparse.ArgumentParser(description=__doc__)
    _cli.add_base_args(parser)
    _cli.add_alignment_args(parser, disabled=("normalize",))
    _cli.add_annotation_args(parser)
    parser.add_argument("roi_file", type=str, nargs="?", default=None,
                        help="ROI file from `metagene generate` (CDS start windows); else give --annotation_files")
    parser.add_argument("outbase")
    parser.add_argument("--codon_buffer", type=int, default=5)
    args = parser.parse_args(argv)
    ga = _cli.genome_array_from_args(args, disabled=("normalize",))
    if args.roi_file is not None:
        roi = _cli.read_pl_table(args.roi_file)
        cds = []
        for region, offset, zero_point in zip(roi["region"], roi["alignment_offset"], roi["zero_point"]):
            chain = SegmentChain.from_str(region)                              # roi_row_to_cds, :58-77
            cds_start = int(zero_point) - int(round(float(offset)))
            cds.append(chain.get_subchain(cds_start, chain.length))
        back_buffer = -1
    else:
        if len(args.annotation_files) == 0:
            sys.stderr.write("Either an ROI file or at least annotation file must be given.\n")
            sys.exit(1)
        cds = [tx.get_cds() for tx in _cli.chains_from_args(args, as_transcripts=True)]
        back_buffer = -args.codon_buffer
    sums = do_phase(ga, cds, list(range(args.min_length, args.max_length + 1)), args.codon_buffer, back_buffer)
    lengths, counted, frac, phases = phase_table(sums)
    if _cli.is_writer():
        with open("%s_phasing.txt" % args.outbase, "w") as fout:
            fout.write("read_length\treads_counted\tfraction_reads_counted\tphase0\tphase1\tphase2\n")
            for i, k in enumerate(lengths):
                fout.write("%d\t%d\t%s\n" % (k, counted[i], "\t".join("nan" if np.isnan(v) else "%.6f" % v
                                                                      for v in [frac[i]] + list(phases[i]))))
    _cli.finish_distributed()


if __name__ == "__main__":
    main()
