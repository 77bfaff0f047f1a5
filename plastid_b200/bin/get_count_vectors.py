"""``get_count_vectors``: the masked count vector of every region of an annotation, one file per region
(plastid/bin/get_count_vectors.py:34-108).  All vectors are gathered in one launch (``pb_gather_chains_range``); the
files are written as the reference's ``numpy.savetxt`` calls write them."""
import argparse
import os
import sys

import numpy as np

from . import _cli
from ..genome_array import gather_chains
from ..masks import MaskIndex, apply_mask_index


def count_vectors(ga, chains, mask_features=None):
    """-> list of ``numpy.ma.MaskedArray`` (5'->3'), one per chain: ``chain.get_masked_counts(ga)`` after the
    overlapping mask features were added (get_count_vectors.py:98-103)."""
    table = ga.chain_table(chains)
    if mask_features:
        apply_mask_index(table, MaskIndex(mask_features, ga.layout), ga.device)
    need = tuple(sorted(set("+-."[p] for p in np.unique(table.chain_plane)), key="+-.".index)) or ("+",)
    planes = ga.count_planes(need)
    values, masked, row_off = gather_chains(planes, table)
    ga._allreduce(values)
    values, masked = values.cpu().numpy(), masked.cpu().numpy().astype(bool)
    if planes.dtype == "u32":
        values = values.astype(np.int64)                # the reference's point rules count in int
    if ga._normalize is True:
        values = values / float(ga.sum()) * 1e6
    out = []
    for i, ch in enumerate(chains):
        a, b = int(row_off[i]), int(row_off[i + 1])
        if not table.known[i]:                          # chromosome unknown to the alignments: zeros
            out.append(np.ma.MaskedArray(np.zeros(ch.length), mask=ch.get_masked_counts(ga).mask if ch.length else False))
        else:
            out.append(np.ma.MaskedArray(values[a:b], mask=masked[a:b]))
    return out


def main(argv=sys.argv[1:]):
    """``get_count_vectors OUT_FOLDER --count_files ... --annotation_files ...`` with the reference's flags."""
    parser = argparse.ArgumentParser(description=__doc__)
    _cli.add_base_args(parser)
    _cli.add_alignment_args(parser)
    _cli.add_annotation_args(parser)
    _cli.add_mask_args(parser)
    parser.add_argument("out_folder", type=str, help="Folder in which to save output vectors")
    parser.add_argument("--out_prefix", default="", type=str, help="Prefix to prepend to output files (default: no prefix)")
    parser.add_argument("--format", default="%.8f", type=str, help=r"printf-style format string for output (default: '%%.8f')")
    args = parser.parse_args(argv)
    ga = _cli.genome_array_from_args(args)
    chains = _cli.chains_from_args(args)
    masks = _cli.chains_from_args(args, prefix="mask_") if args.mask_annotation_files else None
    vectors = count_vectors(ga, chains, masks)
    if _cli.is_writer():
        if not os.path.isdir(args.out_folder):
            os.mkdir(args.out_folder)
        for ch, vec in zip(chains, vectors):
            # numpy.savetxt writes the data of the masked array (masked positions keep their counts)
            np.savetxt(os.path.join(args.out_folder, "%s%s.txt" % (args.out_prefix, ch.get_name())), np.ma.getdata(vec),
                       fmt=args.format)
    _cli.finish_distributed()


if __name__ == "__main__":
    main()
