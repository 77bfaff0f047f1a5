"""``cs count``: per-gene exon / utr5 / cds / utr3 counts and RPKM (plastid/bin/cs.py:667-729)."""
import argparse
import sys

import numpy as np

from . import _cli
from ..roitools import SegmentChain

KEYS = ("exon", "utr5", "cds", "utr3")


def do_count(ga, gene_positions):
    """``gene_positions``: dict with ``region`` + one chain-string column per key in KEYS (the
    positions table ``cs generate`` writes).  Returns (column_order, columns dict)."""
    normconst = 1000.0 * 1e6 / ga.sum()
    names = list(gene_positions["region"])
    chains = []
    for k in KEYS:
        chains.extend(SegmentChain.from_str(s) for s in gene_positions[k])
    sums, _live = ga.count_chains(chains, use_masks=False)      # cs applies no masks at count time
    n = len(names)
    cols, order = {"region": names}, ["region"]
    for j, k in enumerate(KEYS):
        total = sums[j * n:(j + 1) * n]
        length = np.asarray([ch.length for ch in chains[j * n:(j + 1) * n]], dtype=np.int64)
        with np.errstate(all="ignore"):
            rpkm = np.where(length > 0, normconst * total / np.maximum(length, 1), np.nan)
        cols["%s_reads" % k] = [0 if L == 0 else t for t, L in zip(total, length)]   # sum([]) == 0 (int)
        cols["%s_length" % k] = list(length)
        cols["%s_rpkm" % k] = list(rpkm)
        order += ["%s_reads" % k, "%s_length" % k, "%s_rpkm" % k]
    return order, cols


def _fmt(v):
    if isinstance(v, (float, np.floating)):
        return "nan" if np.isnan(v) else "%.8f" % v
    return str(v)


def write_table(fout, order, cols):
    fout.write("\t".join(order) + "\n")
    for i in range(len(cols["region"])):
        fout.write("\t".join(_fmt(cols[c][i]) for c in order) + "\n")


def main(argv=sys.argv[1:]):
    parser = argparse.ArgumentParser(description=__doc__)
    sub = parser.add_subparsers(dest="program")
    cp = sub.add_parser("count")
    _cli.add_alignment_args(cp)
    cp.add_argument("position_file")
    cp.add_argument("outbase")
    args = parser.parse_args(argv)
    if args.program != "count":
        parser.error("only the `count` sub-program is on the GPU path")
    ga = _cli.genome_array_from_args(args)
    order, cols = do_count(ga, _cli.read_pl_table(args.position_file))
    with open("%s.txt" % args.outbase, "w") as fout:
        write_table(fout, order, cols)


if __name__ == "__main__":
    main()
