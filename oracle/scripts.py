"""Restatement of the counting loops of the reference's scripts, on the Python oracle's objects.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function follows the cited script
statement by statement (numpy / numpy.ma calls kept verbatim so numpy's own masked-array semantics
are the oracle), with ``ga`` an :class:`oracle.pyoracle.OracleBAMGenomeArray` and regions
:class:`oracle.pyoracle.Chain` objects.

Parity: the arithmetic + ``%.8e`` formatting of ``counts_in_region_rows`` is pinned by the seven
rows printed in ``docs/source/examples/gene_expression.rst:75-83`` (tests/test_oracle_kat.py); the
other loops have no in-tree golden outputs (external data tarball) — **parity unpinned** beyond
being line-by-line restatements.
"""
import warnings

import numpy as np

from .pyoracle import Chain


def counts_in_region_rows(ga, chains, masks=None, crossmap=None):
    """plastid/bin/counts_in_region.py:107-125 -> list of output rows (lists of str).  ``crossmap``:
    an :class:`oracle.pyoracle.GenomeHash` of mask features, queried per region like :114-115."""
    ga_sum = ga.sum()
    normconst = 1000.0 * 1e6 / ga_sum                                         # :108
    rows = []
    for n, ivc in enumerate(chains):
        name = ivc.name if hasattr(ivc, "name") else str(ivc)
        if crossmap is not None:
            hits = crossmap.get_overlapping_features(ivc)                      # :114
            ivc.add_masks(*[seg for f in hits for seg in f.segments])          # :115
        if masks is not None and masks[n]:
            ivc.add_masks(*masks[n])                                           # :115-116
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            counts = np.nansum(ivc.get_masked_counts(ga))                      # :120
            length = ivc.masked_length
            rpnt = np.nan if length == 0 else float(counts) / length          # :122
            rpkm = np.nan if length == 0 else rpnt * normconst
            rows.append([name, str(ivc), "%.8e" % counts, "%.8e" % rpnt, "%.8e" % rpkm, "%d" % length])
    return rows


def format_counts_row(name, region, counts, length, ga_sum):
    """The arithmetic + formatting of one counts_in_region row (:120-124) from raw numbers."""
    normconst = 1000.0 * 1e6 / ga_sum
    rpnt = np.nan if length == 0 else float(counts) / length
    rpkm = np.nan if length == 0 else rpnt * normconst
    return [name, region, "%.8e" % counts, "%.8e" % rpnt, "%.8e" % rpkm, "%d" % length]


def cs_count(ga, gene_positions):
    """plastid/bin/cs.py:682-714.  ``gene_positions``: dict with keys region, exon, utr5, cds, utr3
    (lists of chain strings).  Returns dict of columns."""
    keys = ("exon", "utr5", "cds", "utr3")
    total_counts = ga.sum()
    normconst = 1000.0 * 1e6 / total_counts
    out = {"region": []}
    for x in keys:
        for y in ("reads", "length", "rpkm"):
            out["%s_%s" % (x, y)] = []
    for i, name in enumerate(gene_positions["region"]):
        out["region"].append(name)
        for k in keys:
            ivc = Chain.from_str(gene_positions[k][i])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                total = sum(ivc.get_counts(ga))                                # :709 python sequential sum
            length = ivc.length
            rpkm = (normconst * total / length) if length > 0 else np.nan
            out["%s_reads" % k].append(total)
            out["%s_length" % k].append(length)
            out["%s_rpkm" % k].append(rpkm)
    return out


def metagene_count(ga, roi_rows, window_size, norm_start, norm_end, min_counts, use_mean=False):
    """plastid/bin/metagene.py:895-953.  ``roi_rows``: list of dicts with ``region``, ``masked``
    (chain strings) and ``alignment_offset``.  Returns counts, norm_counts (MaskedArrays), profile,
    num_genes, row_select."""
    cshape = (len(roi_rows), window_size)
    counts = np.ma.MaskedArray(np.tile(np.nan, cshape), mask=np.tile(True, cshape))
    for i, row in enumerate(roi_rows):
        roi = Chain.from_str(row["region"])
        mask = Chain.from_str(row["masked"])
        roi.add_masks(*mask.segments)
        offset = int(round(row["alignment_offset"]))
        assert offset + roi.length <= window_size
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mvec = roi.get_masked_counts(ga)
        counts.data[i, offset:offset + roi.length] = mvec.data
        counts.mask[i, offset:offset + roi.length] = mvec.mask
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        denominator = np.nansum(counts[:, norm_start:norm_end], axis=1)        # :918
        row_select = denominator >= min_counts                                 # :919
        norm_counts = (counts.T.astype(float) / denominator).T                 # :921
        norm_counts = np.ma.MaskedArray(norm_counts, mask=counts.mask)
        norm_counts.mask[np.isnan(norm_counts)] = True
        norm_counts.mask[np.isinf(norm_counts)] = True
        try:
            pfunc = np.ma.mean if use_mean else np.ma.median
            profile = pfunc(norm_counts[row_select], axis=0)                   # :934-939
        except (IndexError, ValueError):
            profile = np.zeros(norm_counts.shape[0])
        num_genes = ((~norm_counts.mask)[row_select]).sum(0)                   # :953
    return counts, norm_counts, profile, num_genes, row_select


def psite_count(ga, roi_rows, window_size, norm_start, norm_end, min_counts, min_len, max_len, aggregate=False):
    """plastid/bin/psite.py:153-234 -> (raw_count_dict, profiles dict, regions_counted dict)."""
    shape = (len(roi_rows), window_size)
    raw = {}
    for k in range(min_len, max_len + 1):
        raw[k] = np.ma.MaskedArray(np.tile(np.nan, shape), mask=np.tile(True, shape), dtype=float)
    for i, row in enumerate(roi_rows):
        roi = Chain.from_str(row["region"])
        mask = Chain.from_str(row["masked"])
        roi.add_masks(*mask.segments)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            valid_mask = roi.get_masked_counts(ga).mask                        # :166
        offset = int(round(row["alignment_offset"]))
        assert offset + roi.length <= window_size
        count_vectors = {k: [] for k in raw}
        for seg in roi.segments:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                reads = ga.get_reads(seg)                                      # :182
            read_dict = {k: [] for k in raw}
            for read in filter(lambda x: len(x.positions) in read_dict, reads):
                read_dict[len(read.positions)].append(read)
            for k in read_dict:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    count_vector = ga.map_fn(read_dict[k], seg)[1]             # :191
                count_vectors[k].extend(count_vector)
        for k in raw:
            if roi.strand == "-":
                count_vectors[k] = count_vectors[k][::-1]
            raw[k].data[i, offset:offset + roi.length] = np.array(count_vectors[k])
            raw[k].mask[i, offset:offset + roi.length] = valid_mask
    profiles, regions = {}, {}
    for k in raw:
        k_raw = raw[k]
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            denominator = np.nansum(k_raw[:, norm_start:norm_end], axis=1)
            norm = (k_raw.T.astype(float) / denominator).T
            norm_counts = np.ma.MaskedArray(norm, mask=k_raw.mask)
            norm_counts.mask[np.isnan(norm_counts)] = True
            norm_counts.mask[np.isinf(norm_counts)] = True
            try:
                if aggregate is False:
                    profile = np.ma.median(norm_counts[denominator >= min_counts], axis=0)
                else:
                    profile = np.nansum(k_raw[denominator >= min_counts], axis=0)
            except (IndexError, ValueError):
                profile = np.zeros(window_size, dtype=float)
            num_genes = ((~norm_counts.mask)[denominator >= min_counts]).sum(0)
        profiles[k] = profile
        regions[k] = num_genes
    return raw, profiles, regions


def psite_pick_offsets(x, profiles, default=13, constrain=None, require_upstream=False):
    """plastid/bin/psite.py:462-521 offset choice per read length."""
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
    out = {}
    for k, y in profiles.items():
        y = np.ma.filled(np.ma.asarray(y, dtype=float), np.nan)
        ymask = np.ma.MaskedArray(y, mask=mask)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if (~mask).sum() == np.isnan(ymask).sum() or np.nanmax(ymask) == 0:
                out[k] = default
            else:
                out[k] = -x[np.ma.argmax(ymask)]
    return out


def phase_by_size(ga, cds_chains, read_lengths, codon_buffer, back_buffer):
    """plastid/bin/phase_by_size.py:165-214 -> {length: float64[3]} phase sums."""
    phase_sums = {k: np.zeros(3) for k in read_lengths}
    for cds_part in cds_chains:
        if len(cds_part) > 0:
            read_dict = {k: [] for k in read_lengths}
            count_vectors = {k: [] for k in read_lengths}
            for seg in cds_part.segments:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    reads = ga.get_reads(seg)
                # NOTE the reference does not reset read_dict per segment (:186-194): reads of earlier
                # segments are mapped again against later segments (they contribute only where their
                # site lies inside the later segment, i.e. nowhere, since sites are fixed).
                for read in filter(lambda x: len(x.positions) in read_dict, reads):
                    read_dict[len(read.positions)].append(read)
                for read_length in read_dict:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        count_vector = list(ga.map_fn(read_dict[read_length], seg)[1])
                    count_vectors[read_length].extend(count_vector)
            for k, vec in count_vectors.items():
                counts = np.array(vec)
                if cds_part.strand == "-":
                    counts = counts[::-1]
                newlen = int(len(counts) // 3)
                counts = counts[:3 * newlen].reshape(newlen, 3)
                phase_sums[k] += counts[codon_buffer:back_buffer, :].sum(0)
    return phase_sums
