"""``psite``: per-read-length 5' metagene around start codons -> P-site offset per length
(plastid/bin/psite.py:90-238, 462-552)."""
import argparse
import sys
import warnings

import numpy as np

from . import _cli
from .metagene import rois_from_table, _NORM_START_DEFAULT, _NORM_END_DEFAULT
from ..genome_array import stratified_windows, window_normalize, column_profile, count_profiles
from ..map_factories import (FivePrimeMapFactory, CenterMapFactory,
                             StratifiedVariableFivePrimeMapFactory, _MapFactory)
from ..regions import ChainTable


def do_count(ga, roi_table, norm_start=None, norm_end=None, min_counts=10, min_len=25, max_len=35,
             aggregate=False, keep=False):
    """Per read length k in [min_len, max_len]: window matrix of counts of k-mers under ``ga.map_fn``
    (psite.main forces FivePrimeMapFactory(0), psite.py:357-359), then the same normalise / median
    (or ``--aggregate`` nansum) as ``metagene count``.  All lengths are counted in ONE launch
    (``pb_stratified_windows``) straight from the sorted batch; no per-length genome vectors exist."""
    import torch
    wins, cols, window_size, flank = rois_from_table(roi_table)
    norm_start = _NORM_START_DEFAULT if norm_start is None else norm_start
    norm_end = _NORM_END_DEFAULT if norm_end is None else norm_end
    if isinstance(ga.map_fn, (CenterMapFactory, StratifiedVariableFivePrimeMapFactory)) or not isinstance(ga.map_fn, _MapFactory):
        raise TypeError("psite on the GPU path needs a point mapping rule (5'/3'/variable)")
    table = ChainTable.from_chains(wins, ga.layout)
    dbatch = ga._device_batch()
    strat, maskmat = stratified_windows(dbatch, ga.layout, ga.map_fn, ga._size_filter(), table, cols, window_size,
                                        min_len, max_len, bin_range=ga.bin_range)
    ga._allreduce(strat)         # multi-GPU: every rank counted the sites of its own genome range
    dev = strat.device
    n_len, n = max_len - min_len + 1, table.n_chains
    colidx = torch.arange(window_size, device=dev)[None, :]
    c0 = torch.as_tensor(np.asarray(cols), device=dev)[:, None]
    clen = torch.from_numpy(table.chain_len).to(dev)[:, None]
    uncovered = (colidx < c0) | (colidx >= c0 + clen)
    out = {"x": np.arange(-flank, window_size - flank), "profiles": {}, "regions_counted": {}, "raw": {}, "denominator": {}}
    if not aggregate and not keep:
        # the default path: normalisation fused with the key extraction, medians per (length, column) —
        # no float64 / normalised / mask matrices are materialised
        profile, n_regions, sel = count_profiles(strat, maskmat, norm_start, norm_end, min_counts, "median")
        profile, n_regions = profile.cpu().numpy(), n_regions.cpu().numpy()
        any_sel = sel.any(dim=1).cpu().numpy()
        for j, k in enumerate(range(min_len, max_len + 1)):
            out["profiles"][k] = profile[j]          # no selected row: all nan, like numpy.ma.median of an empty selection
            out["regions_counted"][k] = n_regions[j]
        return out
    # all read lengths at once: the matrices are stacked row-wise, normalised in one launch and reduced
    # per (length, column) in one launch
    mat = strat.to(torch.float64)                                    # [n_len, n, W]
    mat.masked_fill_(uncovered[None, :, :], float("nan"))            # cells no chain position reaches (psite.py:153-157)
    mat = mat.view(n_len * n, window_size)
    mask_all = maskmat.repeat(n_len, 1)                              # the position mask is shared by all lengths
    denom, sel, norm, nmask = window_normalize(mat, mask_all, norm_start, norm_end, min_counts)
    if aggregate:
        profile, n_used, _ = column_profile(mat, mask_all, sel, "sum", n_batch=n_len)
        # numpy.nansum over a masked matrix (psite.py:226): a column whose selected cells are all masked is masked itself
        profile = profile.masked_fill(n_used == 0, float("nan"))
        _p, n_regions, _ = column_profile(norm, nmask, sel, "mean", n_batch=n_len)     # regions_counted uses the norm mask
    else:
        profile, n_regions, _ = column_profile(norm, nmask, sel, "median", n_batch=n_len)
    profile = profile.view(n_len, window_size).cpu().numpy()
    n_regions = n_regions.view(n_len, window_size).cpu().numpy()
    any_sel = sel.view(n_len, n).any(dim=1).cpu().numpy()
    for j, k in enumerate(range(min_len, max_len + 1)):
        out["profiles"][k] = profile[j]
        out["regions_counted"][k] = n_regions[j]
        if keep:
            out["raw"][k] = np.ma.MaskedArray(mat[j * n:(j + 1) * n].cpu().numpy(), mask=maskmat.cpu().numpy().astype(bool))
            out["denominator"][k] = denom[j * n:(j + 1) * n].cpu().numpy()
    return out


def pick_offsets(x, profiles, default=13, constrain=None, require_upstream=False):
    """psite.py:462-521: offset per length = -x[argmax(profile)] over the allowed columns."""
    x = np.asarray(x)
    if constrain is not None:
        mask = np.tile(True, len(x))
        zp = (x == 0).argmax()
        l, r = constrain
        mindist, maxdist = min(l, r), max(l, r)
        mask[zp - maxdist:zp - mindist + 1] = False
    elif require_upstream:
        mask = x >= 0
    else:
        mask = np.tile(False, len(x))
    offsets = {}
    for k, y in profiles.items():
        ymask = np.ma.MaskedArray(np.asarray(y, dtype=float), mask=mask)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if (~mask).sum() == np.isnan(ymask).sum() or np.nanmax(ymask) == 0:
                offsets[k] = default
            else:
                offsets[k] = -x[np.ma.argmax(ymask)]
    return offsets


def write_offsets(fout, offsets, default):
    fout.write("length\tp_offset\n")
    for k in offsets:
        fout.write("%s\t%s\n" % (k, offsets[k]))
    fout.write("default\t%s" % default)


def write_profiles(fout, out):
    """``OUTBASE_metagene_profiles.txt`` (psite.py:401-411): x, then one ``N-mers`` column per read length, as pandas
    writes floats (``na_rep='nan'``)."""
    lengths = list(out["profiles"])
    fout.write("\t".join(["x"] + ["%s-mers" % k for k in lengths]) + "\n")
    cols = [np.ma.filled(np.ma.asarray(out["profiles"][k], dtype=float), np.nan) for k in lengths]
    for i, x in enumerate(out["x"]):
        fout.write("\t".join([str(int(x))] + ["nan" if np.isnan(c[i]) else repr(float(c[i])) for c in cols]) + "\n")


def main(argv=sys.argv[1:]):
    """``psite ROI_FILE OUTBASE --count_files ...`` with the reference's flags (plastid/bin/psite.py:241-351; no mapping
    flags: reads are mapped at their 5' ends, :357-359).  Under ``torchrun`` every rank counts the sites of its own
    genome range; the per-length window matrices are all-reduced and rank 0 writes."""
    parser = argparse.ArgumentParser(description=__doc__)
    _cli.add_base_args(parser)
    _cli.add_alignment_args(parser, disabled=("normalize",))
    parser.add_argument("--min_counts", type=int, default=10, metavar="N")
    parser.add_argument("--normalize_over", type=int, nargs=2, default=None, metavar="N")
    parser.add_argument("--norm_region", type=int, nargs=2, default=None, metavar="N", help="Deprecated. Use --normalize_over")
    parser.add_argument("--require_upstream", action="store_true", default=False)
    parser.add_argument("--constrain", type=int, nargs=2, default=None, metavar="X")
    parser.add_argument("--aggregate", action="store_true", default=False)
    parser.add_argument("--keep", action="store_true", default=False)
    parser.add_argument("--default", type=int, default=13)
    parser.add_argument("roi_file")
    parser.add_argument("outbase")
    args = parser.parse_args(argv)
    args.mapping, args.offset = "fiveprime", 0                   # psite.py:357-359
    ga = _cli.genome_array_from_args(args, disabled=("normalize",))
    for name in list(ga._filters):                               # psite.py:380-383
        ga.remove_filter(name)
    roi = _cli.read_pl_table(args.roi_file)
    from .metagene import norm_region_from_args, keep_matrices
    ns, ne = norm_region_from_args(roi, args)
    out = do_count(ga, roi, ns, ne, args.min_counts, args.min_length, args.max_length, args.aggregate, keep=args.keep)
    offsets = pick_offsets(out["x"], out["profiles"], args.default, args.constrain, args.require_upstream)
    if _cli.is_writer():
        with open("%s_metagene_profiles.txt" % args.outbase, "w") as fout:
            write_profiles(fout, out)
        if args.keep:
            for k in out["raw"]:
                raw, norm, mask = keep_matrices(dict(counts=out["raw"][k], denominator=out["denominator"][k]))
                np.savetxt("%s_%s_rawcounts.txt.gz" % (args.outbase, k), raw, delimiter="\t")
                np.savetxt("%s_%s_normcounts.txt.gz" % (args.outbase, k), norm, delimiter="\t")
                np.savetxt("%s_%s_mask.txt.gz" % (args.outbase, k), mask, delimiter="\t")
        with open("%s_p_offsets.txt" % args.outbase, "w") as fout:
            write_offsets(fout, offsets, args.default)
    _cli.finish_distributed()


if __name__ == "__main__":
    main()
