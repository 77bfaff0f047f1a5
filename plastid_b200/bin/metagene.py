"""``metagene``: the ``generate`` geometry and the ``count`` reduction of plastid/bin/metagene.py.

``count`` (:821-1011): window x position count matrix -> normalised -> median / mean profile, as three
launches: gather, normalise, column statistics.  ``generate`` (:180-766): landmark windows of every
transcript and the maximal spanning window of every gene, as two launches (``pb_landmark_windows``,
``pb_spanning_windows``) plus the mask launch (``pb_mask_chains``)."""
import argparse
import sys
import warnings

import numpy as np

from . import _cli
from .. import _lib
from ..genome_array import gather_windows, window_normalize, column_profile
from ..map_factories import DataWarning
from ..masks import GenomeHash, apply_mask_index, mask_intervals_of_chains
from ..regions import ChainTable
from ..roitools import GenomicSegment, SegmentChain
from ..windows import STRAND_CODE, TranscriptTable, landmark_windows, layout_for_features, spanning_windows

_NORM_START_DEFAULT, _NORM_END_DEFAULT = 20, 50      # metagene.py:770-771


# ---------------------------------------------------------------------------------------------
# generate: window functions (host objects, same signatures and return values as the reference)
# ---------------------------------------------------------------------------------------------
def window_landmark(region, flank_upstream=50, flank_downstream=50, ref_delta=0, landmark=0):
    """Window of ``flank_upstream + flank_downstream`` transcript positions around ``landmark + ref_delta`` ->
    ``(window SegmentChain, alignment offset, (chrom, position, strand) of the zero point)``; same contract as
    plastid/bin/metagene.py:180-239.  Host-object form of what ``pb_landmark_windows`` computes for a whole
    transcript table: the window is the transcript-coordinate interval ``[zero - up, zero + down)`` clipped to the
    region, and the offset is the number of leading window columns the region cannot fill."""
    zero = landmark + ref_delta
    lo, hi = max(zero - flank_upstream, 0), min(zero + flank_downstream, region.length)
    # columns missing on the 5' side; the reference measures the shortfall from `landmark` alone (its ref_delta
    # quirk, :223), and so does pb_landmark_windows
    offset = flank_upstream - landmark if zero < flank_upstream else 0
    window = region.get_subchain(lo, hi)
    if zero == region.length:                  # zero point one past the 3' end: not a transcript position
        span = region.spanning_segment
        zero_point = (span.chrom, span.end if span.strand == "+" else span.start - 1, span.strand)
    else:
        zero_point = region.get_genomic_coordinate(zero)
    return window, offset, zero_point


def window_cds_start(transcript, flank_upstream, flank_downstream, ref_delta=0):
    """plastid/bin/metagene.py:241-290: window around the start codon; ``(SegmentChain(), nan, nan)``
    for a transcript without CDS."""
    if transcript.cds_start is None:
        return SegmentChain(), np.nan, np.nan
    return window_landmark(transcript, flank_upstream, flank_downstream, ref_delta=ref_delta,
                           landmark=transcript.cds_start)


def window_cds_stop(transcript, flank_upstream, flank_downstream, ref_delta=0):
    """plastid/bin/metagene.py:293-340: window around the stop codon (landmark ``cds_end - 3``)."""
    if transcript.cds_start is None:
        return SegmentChain(), np.nan, np.nan
    return window_landmark(transcript, flank_upstream, flank_downstream, ref_delta=ref_delta,
                           landmark=transcript.cds_end - 3)


# transcript coordinate of the landmark (None = no landmark): lets the device evaluate the window
# function for all transcripts at once; other callables are evaluated per region on the host
window_cds_start.landmark_of = lambda tx: tx.cds_start
window_cds_stop.landmark_of = lambda tx: None if tx.cds_start is None else tx.cds_end - 3


def _lower_windows(regions, window_func, flank_upstream, flank_downstream, layout, device):
    """-> (TranscriptTable, win, flags) on ``device`` for ``window_func(region, up, down)`` of every region."""
    import torch
    landmark_of = getattr(window_func, "landmark_of", None)
    if landmark_of is not None:
        table = TranscriptTable.from_transcripts(regions, layout, [landmark_of(r) for r in regions])
        win, flags = landmark_windows(table, flank_upstream, flank_downstream, device)
        return table, win, flags
    rois, win, flags = [], np.zeros((max(len(regions), 1), 4), dtype=np.int64), np.zeros(max(len(regions), 1), dtype=np.uint8)
    for n, region in enumerate(regions):
        try:
            roi, offset, refpoint = window_func(region, flank_upstream, flank_downstream)
        except IndexError:
            warnings.warn("IndexError finding common positions at region '%s'. Ignoring region: " % region.get_name())
            rois.append(SegmentChain())
            flags[n] = _lib.PB_WIN_INDEX_ERROR
            continue
        has_ref = not (isinstance(refpoint, float) and np.isnan(refpoint))
        if has_ref:
            if len(roi) > 0:
                assert offset + roi.length <= flank_upstream + flank_downstream
            win[n] = (0, roi.length, int(offset), layout.bin_of(refpoint[0], refpoint[1]))
            flags[n] = _lib.PB_WIN_HAS_REF
        if len(roi) == 0:                                         # keep the region's strand for the landmark test
            roi = SegmentChain()
            roi.strand = region.strand
        rois.append(roi)
    table = TranscriptTable.from_transcripts(rois, layout, [None] * len(rois))
    table.reverse[:] = [STRAND_CODE.get(r.strand, 0) for r in rois]
    return table, torch.from_numpy(win).to(device), torch.from_numpy(flags).to(device)


def _windows_for_groups(groups, mask_hash, flank_upstream, flank_downstream, window_func, device):
    """``groups``: list of (name, [regions]).  -> list of (window SegmentChain with masks added, offset)
    per group; ``(SegmentChain(), nan)`` where the reference returns no window."""
    regions = [r for _, members in groups for r in members]
    mask_hash = GenomeHash([]) if mask_hash is None else mask_hash
    layout = layout_for_features(regions, mask_hash.features)
    table, win, flags = _lower_windows(regions, window_func, flank_upstream, flank_downstream, layout, device)
    grp_off = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([len(m) for _, m in groups], out=grp_off[1:])
    res = spanning_windows(table, win, flags, grp_off, np.arange(len(regions), dtype=np.int64),
                           flank_upstream, flank_downstream, device)
    out = [(SegmentChain(), np.nan)] * len(groups)
    found = []
    for g, (name, members) in enumerate(groups):
        if res["status"][g] == _lib.PB_SPAN_REF_OUTSIDE:          # metagene.py:498 would raise here
            raise KeyError("SegmentChain.get_segmentchain_coordinate: landmark of '%s' is not in its maximal spanning window" % name)
        if res["status"][g] != _lib.PB_SPAN_WINDOW:
            continue
        first = members[0]
        base = int(layout.chrom_bin_off[layout.index[first.chrom]])
        k0, k1 = int(res["out_off"][g]), int(res["out_off"][g + 1])
        segs = [GenomicSegment(first.chrom, int(a) - base, int(b) - base, first.strand)
                for a, b in zip(res["out_bstart"][k0:k1], res["out_bend"][k0:k1])]
        roi = SegmentChain(*segs)
        roi.attr["ID"] = roi.get_name() if name is None else name
        roi.attr["thickstart"] = int(res["refpos"][g]) - base
        roi.attr["thickend"] = int(res["refpos"][g]) - base + 1
        assert roi.length == int(res["n_pos"][g])
        out[g] = (roi, int(res["offset"][g]))
        found.append(g)
    if found and len(mask_hash):
        chains = [out[g][0] for g in found]
        ctable = ChainTable.from_chains(chains, layout)
        bits = apply_mask_index(ctable, mask_hash.mask_index(layout), device)
        for roi, ivs in zip(chains, mask_intervals_of_chains(ctable, bits)):
            roi.add_masks(*[GenomicSegment(roi.chrom, a, b, roi.strand) for a, b in ivs])
    return out


def maximal_spanning_window(regions, mask_hash, flank_upstream, flank_downstream, window_func=window_cds_start,
                            name=None, printer=None, device="cuda"):
    """plastid/bin/metagene.py:343-502 -> (maximal spanning window, alignment offset), or
    ``(SegmentChain(), nan)`` when the regions do not share a landmark and positions around it."""
    return _windows_for_groups([(name, list(regions))], mask_hash, flank_upstream, flank_downstream,
                               window_func, device)[0]


def group_regions_make_windows(source, mask_hash, flank_upstream, flank_downstream, window_func=window_cds_start,
                               is_sorted=False, group_by="gene_id", printer=None, device="cuda"):
    """plastid/bin/metagene.py:511-766: group regions by ``group_by`` and build each group's maximal
    spanning window -> :class:`pandas.DataFrame` with the reference's columns, sorted by ``region_id``.
    ``is_sorted`` only bounded the reference's memory use and is accepted for compatibility."""
    import pandas as pd
    window_size = flank_upstream + flank_downstream
    group_transcript = {}
    for tx_chain in source:
        attr = tx_chain.attr
        if group_by == "gene_id":
            if "gene_id" in attr:
                group_attr = attr["gene_id"]
            else:
                group_attr = tx_chain.get_gene() if hasattr(tx_chain, "get_gene") else "gene_%s" % tx_chain.get_name()
                warnings.warn("Region '%s' has no gene_id. Inferring gene_id to be '%s'" % (tx_chain.get_name(), group_attr),
                              DataWarning)
        elif group_by in attr:
            group_attr = attr[group_by]
        else:
            warnings.warn("Region '%s' has no attribute '%s', and will not be grouped. Using region name as default group."
                          % (tx_chain.get_name(), group_by), DataWarning)
            group_attr = tx_chain.get_name()
        group_transcript.setdefault(group_attr, []).append(tx_chain)

    groups = list(group_transcript.items())
    cols = ["region_id", "region", "region_length", "masked", "alignment_offset", "window_size", "zero_point",
            "region_bed", "threeprime_offset"]
    dtmp = {c: [] for c in cols}
    if groups:
        windows = _windows_for_groups(groups, mask_hash, flank_upstream, flank_downstream, window_func, device)
        for (region_id, _), (roi, offset) in zip(groups, windows):
            if len(roi) > 0:
                dtmp["region_id"].append(region_id)
                dtmp["window_size"].append(window_size)
                dtmp["region"].append(str(roi))
                dtmp["masked"].append(str(roi.get_masks_as_segmentchain()))
                dtmp["alignment_offset"].append(offset)
                dtmp["zero_point"].append(flank_upstream)
                dtmp["region_bed"].append(roi.as_bed())
                dtmp["region_length"].append(roi.length)
                dtmp["threeprime_offset"].append(window_size - offset - roi.length)
    df = pd.DataFrame(dtmp)
    df.sort_values(["region_id"], inplace=True)
    if printer is not None:
        printer.write("Processed %s genes total. Included %s." % (len(groups), len(df)))
    if (df["alignment_offset"] == flank_upstream).all():
        warnings.warn("All maximal spanning windows lack flanks upstream of reference landmark. This occurs e.g. for start codons when annotation files don't contain UTR data. Please check your annotation file.",
                      DataWarning)
    if (df["threeprime_offset"] == flank_downstream).all():
        warnings.warn("All maximal spanning windows lack flanks downstream of reference landmark. This occurs e.g. for stop codons when annotation files don't contain UTR data. Please check your annotation file.",
                      DataWarning)
    return df


def do_generate(transcripts, mask_hash=None, landmark="cds_start", upstream=50, downstream=50, group_by="gene_id",
                device="cuda"):
    """The ``generate`` sub-program's computation (plastid/bin/metagene.py:1196-1266): ROI table for
    ``count`` from an iterable of transcripts."""
    funcs = {"cds_start": window_cds_start, "cds_stop": window_cds_stop}
    return group_regions_make_windows(transcripts, mask_hash, upstream, downstream, window_func=funcs[landmark],
                                      group_by=group_by, device=device)


# ---------------------------------------------------------------------------------------------
# count
# ---------------------------------------------------------------------------------------------
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
    mat, mmask = gather_windows(planes, table, cols, window_size, touched_only=ga._collective and not keep)
    if ga._collective and not keep:
        # several GPUs: every rank filled the cells of its own positions.  The count matrix stays where it is: rows are
        # completed and normalised by the rank that owns their first position, means are all-reduced as column sums,
        # exact medians are taken per column slice after one all-to-all (plastid_b200.dist.window_profile)
        from .. import dist as pdist
        ranges = pdist.all_ranges(*ga.bin_range, device=mat.device)
        profile, n_regions, denom, sel = pdist.window_profile(
            mat, mmask, table, ranges, norm_start, norm_end, min_counts, "mean" if use_mean else "median",
            per_million_of=ga.sum() if ga._normalize is True else None)
        return {"x": np.arange(-flank, window_size - flank), "metagene_average": profile.cpu().numpy(),
                "regions_counted": n_regions.cpu().numpy(), "row_select": sel.cpu().numpy().astype(bool),
                "denominator": denom.cpu().numpy()}
    ga._allreduce(mat)           # --keep wants the whole matrices on the writing rank: cells of other ranks are 0, NaN where no position
    if ga._normalize is True:
        mat = mat / float(ga.sum()) * 1e6
    denom, sel, norm, nmask = window_normalize(mat, mmask, norm_start, norm_end, min_counts)
    profile, n_regions, col_sum = column_profile(norm, nmask, sel, "mean" if use_mean else "median")
    out = {"x": np.arange(-flank, window_size - flank), "metagene_average": profile.cpu().numpy(),
           "regions_counted": n_regions.cpu().numpy(), "row_select": sel.cpu().numpy().astype(bool),
           "denominator": denom.cpu().numpy()}
    # no row reaches min_counts: numpy.ma.median / mean of the empty selection is an all-masked vector, written as
    # nan (observed with the reference run under numpy 2.3; its zeros fallback, :940-944, needs an exception that numpy
    # no longer raises) — which is what the column kernel returns for columns without a usable cell
    if keep:
        out["counts"] = np.ma.MaskedArray(mat.cpu().numpy(), mask=mmask.cpu().numpy().astype(bool))
        out["norm_counts"] = np.ma.MaskedArray(norm.cpu().numpy(), mask=nmask.cpu().numpy().astype(bool))
    return out


def keep_matrices(out):
    """The three ``--keep`` matrices as the reference's ``numpy.savetxt`` calls write them (metagene.py:926-932).
    ``savetxt`` stores the DATA of a masked array, and numpy.ma's division leaves the numerator in every cell it
    masks (masked numerator, masked or zero denominator): so the raw file keeps the counts of masked positions, the
    normalised file holds quotients only where the division was defined, and the mask file flags every other cell
    (``MaskedArray(x, mask=m)`` keeps the mask ``x`` already has)."""
    counts = out["counts"]
    raw = np.ma.getdata(counts).astype(np.float64)
    cmask = np.ma.getmaskarray(counts)
    den = np.asarray(out["denominator"], dtype=np.float64)[:, None]
    with np.errstate(all="ignore"):
        quotient = raw / den
    defined = ~cmask & ~np.isnan(den) & (den != 0) & np.isfinite(quotient)
    norm = np.where(defined, quotient, raw)
    return raw, norm, ~defined | np.isnan(norm) | np.isinf(norm)


def write_profile(fout, out):
    fout.write("x\tmetagene_average\tregions_counted\n")
    for x, y, n in zip(out["x"], out["metagene_average"], out["regions_counted"]):
        fout.write("%s\t%s\t%s\n" % (x, "nan" if np.isnan(y) else repr(float(y)), n))


def main(argv=sys.argv[1:]):
    """``metagene generate OUTBASE ...`` / ``metagene count ROI_FILE OUTBASE ...`` with the reference's flags
    (plastid/bin/metagene.py:1118-1340).  ``count`` under ``torchrun``: every rank gathers the window cells of its own
    genome range, the count matrix is all-reduced, rank 0 writes."""
    parser = argparse.ArgumentParser(description=__doc__)
    sub = parser.add_subparsers(dest="program")
    gp = sub.add_parser("generate")
    _cli.add_base_args(gp)
    _cli.add_annotation_args(gp)
    _cli.add_mask_args(gp)
    gp.add_argument("--landmark", choices=("cds_start", "cds_stop"), default="cds_start")
    gp.add_argument("--upstream", type=int, default=50)
    gp.add_argument("--downstream", type=int, default=50)
    gp.add_argument("--group_by", default="gene_id")
    gp.add_argument("--device", default="cuda")
    gp.add_argument("outbase")
    cp = sub.add_parser("count")
    _cli.add_base_args(cp)
    _cli.add_alignment_args(cp)
    cp.add_argument("roi_file")
    cp.add_argument("outbase")
    cp.add_argument("--min_counts", type=int, default=10, metavar="N")
    cp.add_argument("--normalize_over", type=int, nargs=2, default=None, metavar="N")
    cp.add_argument("--norm_region", type=int, nargs=2, default=None, metavar="N", help="Deprecated. Use --normalize_over")
    cp.add_argument("--landmark", type=str, default=None, help="Name of landmark at zero point, optional.")
    cp.add_argument("--use_mean", action="store_true", default=False)
    cp.add_argument("--keep", action="store_true", default=False)
    args = parser.parse_args(argv)
    if args.program == "generate":
        transcripts = _cli.chains_from_args(args, as_transcripts=True)
        masks = _cli.chains_from_args(args, prefix="mask_")
        roi_table = do_generate(transcripts, GenomeHash(masks), args.landmark, args.upstream, args.downstream,
                                args.group_by, args.device)
        roi_table.to_csv("%s_rois.txt" % args.outbase, sep="\t", header=True, index=False, na_rep="nan",
                         columns=["region_id", "window_size", "region", "masked", "alignment_offset", "zero_point"])
        with open("%s_rois.bed" % args.outbase, "w") as bed_fh:
            for roi in roi_table["region_bed"]:
                bed_fh.write(roi)
        return
    if args.program != "count":
        parser.error("the `generate` and `count` sub-programs are on the GPU path")
    ga = _cli.genome_array_from_args(args)
    roi = _cli.read_pl_table(args.roi_file)
    ns, ne = norm_region_from_args(roi, args)
    out = do_count(ga, roi, ns, ne, args.min_counts, args.use_mean, args.keep)
    if _cli.is_writer():
        with open("%s_metagene_profile.txt" % args.outbase, "w") as fout:
            write_profile(fout, out)
        if args.keep:
            raw, norm, mask = keep_matrices(out)
            np.savetxt("%s_rawcounts.txt.gz" % args.outbase, raw, delimiter="\t", fmt="%.8f")
            np.savetxt("%s_normcounts.txt.gz" % args.outbase, norm, delimiter="\t")
            np.savetxt("%s_mask.txt.gz" % args.outbase, mask, delimiter="\t")
    _cli.finish_distributed()


def norm_region_from_args(roi_table, args):
    """plastid/bin/metagene.py:776-818 (``_get_norm_region``): ``--normalize_over`` counts from the landmark,
    the deprecated ``--norm_region`` and the default (20, 50) from the window's first column."""
    flank = int(roi_table["zero_point"][0])
    if args.normalize_over is not None:
        return args.normalize_over[0] + flank, args.normalize_over[1] + flank
    if getattr(args, "norm_region", None) is not None:
        warnings.warn("`--norm_region` is deprecated. Use `--normalize_over` instead.", DataWarning)
        return tuple(args.norm_region)
    return _NORM_START_DEFAULT, _NORM_END_DEFAULT


if __name__ == "__main__":
    main()
