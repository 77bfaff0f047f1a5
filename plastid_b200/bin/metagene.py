"""``metagene count``: window x position count matrix -> normalised -> median / mean profile
(plastid/bin/metagene.py:821-1011), as three launches: gather, normalise, column statistics."""
import argparse
import sys

import numpy as np

from . import _cli
from ..genome_array import gather_windows, window_normalize, column_profile
from ..regions import ChainTable
from ..roitools import SegmentChain

_NORM_START_DEFAULT, _NORM_END_DEFAULT = 20, 50      # metagene.py:770-771


def rois_from_table(roi_table):
    """ROI table columns (``region``, ``masked``, ``alignment_offset``, ``window_size``,
    ``zero_point``) -> (windows with masks added, column offsets, window_size, upstream_flank)."""
    wins, cols = [], []
    for region, masked, off in zip(roi_table["region"], roi_table["masked"], roi_table["alignment_offset"]):
        roi = SegmentChain.from_str(region)
        mask = SegmentChain.from_str(masked)
        roi.add_masks(*mask)
        wins.append(roi)
        cols.append(int(round(float(off))))
    window_size = int(roi_table["window_size"][0])
    for w, c in zip(wins, cols):
        assert c + w.length <= window_size
    return wins, cols, window_size, int(roi_table["zero_point"][0])


def do_count(ga, roi_table, norm_start=None, norm_end=None, min_counts=10, use_mean=False, keep=False):
    """Returns dict with ``profile`` / ``regions_counted`` / ``x`` (numpy) and, when ``keep``,
    ``counts`` / ``norm_counts`` as numpy masked arrays like the reference's return values."""
    wins, cols, window_size, flank = rois_from_table(roi_table)
    norm_start = _NORM_START_DEFAULT if norm_start is None else norm_start
    norm_end = _NORM_END_DEFAULT if norm_end is None else norm_end
    table = ChainTable.from_chains(wins, ga.layout)
    need = sorted(set("+-."[p] for p in np.unique(table.chain_plane))) or ["+"]
    planes = ga.count_planes(tuple(need))
    mat, mmask = gather_windows(planes, table, cols, window_size)
    if ga._normalize is True:
        mat = mat / float(ga.sum()) * 1e6
    denom, sel, norm, nmask = window_normalize(mat, mmask, norm_start, norm_end, min_counts)
    profile, n_regions, col_sum = column_profile(norm, nmask, sel, "mean" if use_mean else "median")
    out = {"x": np.arange(-flank, window_size - flank), "metagene_average": profile.cpu().numpy(),
           "regions_counted": n_regions.cpu().numpy(), "row_select": sel.cpu().numpy().astype(bool),
           "denominator": denom.cpu().numpy()}
    if sel.sum().item() == 0:
        # numpy.ma.median of an empty selection raises; the reference falls back to zeros (:940-944)
        out["metagene_average"] = np.zeros(window_size)
    if keep:
        out["counts"] = np.ma.MaskedArray(mat.cpu().numpy(), mask=mmask.cpu().numpy().astype(bool))
        out["norm_counts"] = np.ma.MaskedArray(norm.cpu().numpy(), mask=nmask.cpu().numpy().astype(bool))
    return out


def write_profile(fout, out):
    fout.write("x\tmetagene_average\tregions_counted\n")
    for x, y, n in zip(out["x"], out["metagene_average"], out["regions_counted"]):
        fout.write("%s\t%s\t%s\n" % (x, "nan" if np.isnan(y) else repr(float(y)), n))


def main(argv=sys.argv[1:]):
    parser = argparse.ArgumentParser(description=__doc__)
    sub = parser.add_subparsers(dest="program")
    cp = sub.add_parser("count")
    _cli.add_alignment_args(cp)
    cp.add_argument("roi_file")
    cp.add_argument("outbase")
    cp.add_argument("--normalize_over", type=int, nargs=2, default=None)
    cp.add_argument("--min_counts", type=int, default=10)
    cp.add_argument("--use_mean", action="store_true")
    cp.add_argument("--keep", action="store_true")
    args = parser.parse_args(argv)
    if args.program != "count":
        parser.error("only the `count` sub-program is on the GPU path")
    ga = _cli.genome_array_from_args(args)
    roi = _cli.read_pl_table(args.roi_file)
    ns = ne = None
    if args.normalize_over is not None:
        flank = int(roi["zero_point"][0])
        ns, ne = args.normalize_over[0] + flank, args.normalize_over[1] + flank
    out = do_count(ga, roi, ns, ne, args.min_counts, args.use_mean, args.keep)
    with open("%s_metagene_profile.txt" % args.outbase, "w") as fout:
        write_profile(fout, out)
    if args.keep:
        np.savetxt("%s_rawcounts.txt.gz" % args.outbase, out["counts"].filled(np.nan), delimiter="\t", fmt="%.8f")
        np.savetxt("%s_normcounts.txt.gz" % args.outbase, out["norm_counts"].filled(np.nan), delimiter="\t")
        np.savetxt("%s_mask.txt.gz" % args.outbase, np.ma.getmaskarray(out["norm_counts"]), delimiter="\t")


if __name__ == "__main__":
    main()
