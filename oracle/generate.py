"""TEST INFRASTRUCTURE ONLY — CPU restatement of the geometry of ``metagene generate`` and of the
position-set arithmetic of ``cs generate`` (second half of this file).

Follows, line by line, plastid v0.6.1:

* ``window_landmark`` / ``window_cds_start`` / ``window_cds_stop``   plastid/bin/metagene.py:180-340
* ``maximal_spanning_window``                                        plastid/bin/metagene.py:343-502
* ``group_regions_make_windows``                                     plastid/bin/metagene.py:511-766
* the SegmentChain / Transcript pieces they call: ``c_get_segmentchain_coordinate``
  roitools.pyx:2957-3012, ``c_get_genomic_coordinate`` :3055-3119, ``c_get_subchain`` :3169-3218
  (a python slice of the position hash — no bounds error), ``positions_to_segments``,
  ``Transcript._update_cds`` :3883-3913, ``get_gene`` :2154-2173.

Pinned against the reference's own unit-test tables for these functions
(plastid/test/unit/bin/test_metagene.py, transcribed to tests/golden/metagene_generate.json by
tests/golden/make_metagene_generate_golden.py) in tests/test_oracle_kat.py.  Only tests/ may import it.
"""
import warnings

import numpy as np

from .pyoracle import Chain, Seg, DataWarning, GenomeHash

nan = np.nan


def positions_to_segments(chrom, strand, positions):
    """roitools.pyx ``positions_to_segments``: set of positions -> sorted runs."""
    segs = []
    run_start = last = None
    for p in sorted(set(positions)):
        if run_start is None:
            run_start = last = p
        elif p == last + 1:
            last = p
        else:
            segs.append(Seg(chrom, run_start, last + 1, strand))
            run_start = last = p
    if run_start is not None:
        segs.append(Seg(chrom, run_start, last + 1, strand))
    return segs


def get_segmentchain_coordinate(chain, genomic_x, stranded=True):
    """roitools.pyx:2957-3012 (KeyError when the position is not in the chain)."""
    cum = 0
    if not chain.segments or genomic_x < chain.segments[0].start:
        raise KeyError(genomic_x)
    for seg in chain.segments:
        cum += len(seg)
        if genomic_x < seg.end:
            if genomic_x >= seg.start:
                ret = cum - seg.end + genomic_x
                if chain.strand == "-" and stranded is True:
                    ret = chain.length - ret - 1
                return ret
            raise KeyError(genomic_x)
    raise KeyError(genomic_x)


def get_genomic_coordinate(chain, x, stranded=True):
    """roitools.pyx:3014-3119 -> (chrom, position, strand); IndexError outside [0, length)."""
    if x < 0 or x >= chain.length:
        raise IndexError(x)
    if chain.strand == "-" and stranded is True:
        x = chain.length - x - 1
    return (chain.chrom, chain.position_list[x], chain.strand)


def get_subchain(chain, start, end, stranded=True):
    """roitools.pyx:3169-3218: python slice of the position hash."""
    if start == end:
        return Chain()
    if stranded is True and chain.strand == "-":
        start, end = chain.length - end, chain.length - start
    return Chain(*positions_to_segments(chain.chrom, chain.strand, chain.position_list[start:end]))


class Tx(Chain):
    """Transcript: a Chain with attr and CDS end points (roitools.pyx:3565-3913)."""

    def __init__(self, *segs, **attr):
        Chain.__init__(self, *segs)
        self.attr = dict(attr)
        self.cds_genome_start = attr.get("cds_genome_start", None)
        self.cds_genome_end = attr.get("cds_genome_end", None)
        self.cds_start = self.cds_end = None
        if self.cds_genome_start is not None and self.cds_genome_end is not None:
            if self.strand == "+":                                  # _update_cds :3883-3913 ('.' takes the minus branch)
                self.cds_start = get_segmentchain_coordinate(self, self.cds_genome_start)
                try:
                    self.cds_end = get_segmentchain_coordinate(self, self.cds_genome_end)
                except KeyError:
                    self.cds_end = 1 + get_segmentchain_coordinate(self, self.cds_genome_end - 1)
            else:
                self.cds_start = get_segmentchain_coordinate(self, self.cds_genome_end - 1)
                self.cds_end = 1 + get_segmentchain_coordinate(self, self.cds_genome_start)
        else:
            self.cds_genome_start = self.cds_genome_end = None

    def get_name(self):
        return self.attr.get("ID", str(self))

    def get_gene(self):                                             # :2154-2173
        gene = self.attr.get("gene_id", self.attr.get("Parent", "gene_%s" % self.get_name()))
        if isinstance(gene, list):
            gene = ",".join(sorted(gene))
        return gene


# ---------------------------------------------------------------------------
# plastid/bin/metagene.py:180-340
# ---------------------------------------------------------------------------
def window_landmark(region, flank_upstream=50, flank_downstream=50, ref_delta=0, landmark=0):
    if landmark + ref_delta >= flank_upstream:                       # :219-224
        fiveprime_offset = 0
        my_start = landmark + ref_delta - flank_upstream
    else:
        fiveprime_offset = flank_upstream - landmark                 # (sic: ref_delta not included)
        my_start = 0
    my_end = min(region.length, landmark + ref_delta + flank_downstream)
    roi = get_subchain(region, my_start, my_end)
    if landmark + ref_delta == region.length:                        # :232-236
        if region.strand == "+":
            ref_point = (region.chrom, region.segments[-1].end, region.strand)
        else:
            ref_point = (region.chrom, region.segments[0].start - 1, region.strand)
    else:
        ref_point = get_genomic_coordinate(region, landmark + ref_delta)
    return roi, fiveprime_offset, ref_point


def window_cds_start(transcript, flank_upstream, flank_downstream, ref_delta=0):
    if transcript.cds_start is None:                                 # :277-278
        return Chain(), nan, nan
    return window_landmark(transcript, flank_upstream, flank_downstream, ref_delta=ref_delta,
                           landmark=transcript.cds_start)


def window_cds_stop(transcript, flank_upstream, flank_downstream, ref_delta=0):
    if transcript.cds_start is None:                                 # :331-332
        return Chain(), nan, nan
    return window_landmark(transcript, flank_upstream, flank_downstream, ref_delta=ref_delta,
                           landmark=transcript.cds_end - 3)


# ---------------------------------------------------------------------------
# plastid/bin/metagene.py:343-502
# ---------------------------------------------------------------------------
def maximal_spanning_window(regions, mask_hash, flank_upstream, flank_downstream,
                            window_func=window_cds_start, name=None):
    """-> (window Chain with masks applied, mask Chain, offset); (empty Chain, empty Chain, nan)
    when the regions do not share landmark and positions."""
    refpoints = []
    window_size = flank_upstream + flank_downstream
    position_matrix = np.tile(np.nan, (len(regions), window_size))   # :438
    for n, region in enumerate(regions):
        try:
            my_roi, my_offset, genomic_refpoint = window_func(region, flank_upstream, flank_downstream)
            refpoints.append(genomic_refpoint)
            if genomic_refpoint is not np.nan and len(my_roi) > 0:
                pos_list = my_roi.position_list
                my_len = len(pos_list)
                assert my_offset + my_len <= window_size
                if my_roi.strand == "+":
                    position_matrix[n, my_offset:my_offset + my_len] = pos_list
                else:
                    position_matrix[n, my_offset:my_offset + my_len] = pos_list[::-1]
        except IndexError:
            warnings.warn("IndexError finding common positions at region '%s'. Ignoring region: "
                          % region.get_name())

    if len(set(refpoints)) == 1 and np.nan not in refpoints:        # :465
        new_shared_positions = []
        for i in range(0, position_matrix.shape[1]):
            col = position_matrix[:, i]
            if len(set(col)) == 1 and not np.isnan(col[0]):          # nan != nan: a nan never matches
                new_shared_positions.append(int(col[0]))
        if len(set(new_shared_positions)) > 0:                       # :474
            new_roi = Chain(*positions_to_segments(regions[0].chrom, regions[0].strand, new_shared_positions))
            if flank_upstream - my_offset == my_roi.length:          # :495-496 (last region's window)
                new_offset = my_offset
            else:
                zero_point_roi = get_segmentchain_coordinate(new_roi, genomic_refpoint[1])
                new_offset = flank_upstream - zero_point_roi
            masks = mask_hash.get_overlapping_features(new_roi)      # :501-506
            mask_segs = []
            for mask in masks:
                mask_segs.extend(mask.segments)
            new_roi.add_masks(*mask_segs)
            if new_roi.position_mask is None:
                mask_chain = Chain()
            else:
                masked = [p for p, m in zip(new_roi.position_list, new_roi.position_mask) if m]
                mask_chain = Chain(*positions_to_segments(new_roi.chrom, new_roi.strand, masked))
            return new_roi, mask_chain, new_offset
    return Chain(), Chain(), nan


# ---------------------------------------------------------------------------
# plastid/bin/metagene.py:511-766
# ---------------------------------------------------------------------------
def group_regions_make_windows(source, mask_hash, flank_upstream, flank_downstream,
                               window_func=window_cds_start, group_by="gene_id"):
    """-> list of row dicts sorted by region_id (the reference's DataFrame, ``region_bed`` left out)."""
    window_size = flank_upstream + flank_downstream
    group_transcript = {}
    for tx_chain in source:                                          # :676-700 (unsorted input: one pass)
        attr = tx_chain.attr
        if group_by == "gene_id":
            if "gene_id" in attr:
                group_attr = attr["gene_id"]
            else:
                group_attr = tx_chain.get_gene()
                warnings.warn("Region '%s' has no gene_id. Inferring gene_id to be '%s'"
                              % (tx_chain.get_name(), group_attr), DataWarning)
        else:
            if group_by in attr:
                group_attr = attr[group_by]
            else:
                warnings.warn("Region '%s' has no attribute '%s', and will not be grouped. Using region name as default group."
                              % (tx_chain.get_name(), group_by), DataWarning)
                group_attr = tx_chain.get_name()
        group_transcript.setdefault(group_attr, []).append(tx_chain)

    rows = []
    for region_id, tx_list in group_transcript.items():              # :702-735
        window, mask_chain, offset = maximal_spanning_window(tx_list, mask_hash, flank_upstream,
                                                             flank_downstream, window_func=window_func,
                                                             name=region_id)
        if len(window) > 0:
            rows.append({"region_id": region_id, "window_size": window_size, "region": str(window),
                         "masked": str(mask_chain), "alignment_offset": offset,
                         "zero_point": flank_upstream, "region_length": window.length,
                         "threeprime_offset": window_size - offset - window.length})
    rows.sort(key=lambda r: r["region_id"])
    return rows


# ---------------------------------------------------------------------------
# cs generate: plastid/bin/cs.py:190-496 (merge_genes, process_partial_group), with python position sets
# exactly as the reference; plastid/util/services/sets.py:16-120 merge_sets = connected components of sets
# sharing a member.  PARITY UNPINNED: the reference holds no in-tree known answers for `cs generate`
# (test_cs.py compares against files of its external data bundle), so this restatement is the anchor.
# ---------------------------------------------------------------------------
import itertools


def tx_subchain(tx, start, end):
    return get_subchain(tx, start, end)


def tx_cds(tx):                                                      # roitools.pyx:4005-4046
    return tx_subchain(tx, tx.cds_start, tx.cds_end) if tx.cds_genome_start is not None else Chain()


def tx_utr5(tx):                                                     # :4048-4087
    return tx_subchain(tx, 0, tx.cds_start) if tx.cds_genome_start is not None else Chain()


def tx_utr3(tx):                                                     # :4089-4130
    return tx_subchain(tx, tx.cds_end, tx.length) if tx.cds_genome_start is not None else Chain()


def merge_sets(list_of_sets):
    """sets.py:16-120: merge sets that share a member until none do."""
    groups = []
    for s in {frozenset(x) for x in list_of_sets}:
        s = set(s)
        keep = []
        for g in groups:
            if g & s:
                s |= g
            else:
                keep.append(g)
        keep.append(s)
        groups = keep
    return groups


def merge_genes(tx_ivcs):                                            # cs.py:190-239
    dout = {}
    exondicts = {"+": {}, "-": {}}
    for txid in tx_ivcs.keys():
        chain = tx_ivcs[txid]
        gene = chain.get_gene()
        for iv in chain.segments:
            exondicts[chain.strand].setdefault(chain.chrom, {}).setdefault((iv.start, iv.end), []).append(gene)
    for strand in exondicts:
        for chrom in exondicts[strand]:
            for group in merge_sets([set(v) for v in exondicts[strand][chrom].values()]):
                merged_name = ",".join(sorted(group))
                for gene in group:
                    dout[gene] = merged_name
    return dout


def _chain_of(chrom, strand, positions):
    return Chain(*positions_to_segments(chrom, strand, positions))


def cs_process_partial_group(transcripts, mask_hash):
    """cs.py:242-496 -> (gene rows, transcript rows, merged_genes); rows are dicts of chain strings
    sorted by ``region`` (the ``*_bed`` columns are left out)."""
    keycombos = list(itertools.permutations(("utr5", "cds", "utr3"), 2))
    merged_genes = merge_genes(transcripts)
    merged_gene_tx = {}
    for txid in transcripts:
        merged_gene_tx.setdefault(merged_genes[transcripts[txid].get_gene()], []).append(txid)

    gene_rows, tx_rows = [], []
    raw = []
    for gene_id, my_txids in merged_gene_tx.items():                 # :313-348
        positions = []
        for txid in my_txids:
            positions.extend(transcripts[txid].position_list)
        first = transcripts[my_txids[0]]
        raw.append(_chain_of(first.chrom, first.strand, set(positions)))
    gene_hash = GenomeHash(raw)                                      # :352

    for (gene_id, my_txids), gene_ivc_raw in zip(merged_gene_tx.items(), raw):      # :354-470
        chrom, strand = gene_ivc_raw.chrom, gene_ivc_raw.strand
        masked_positions = []
        raw_positions = set(gene_ivc_raw.position_list)
        nearby = [x for x in gene_hash.get_overlapping_features(gene_ivc_raw) if set(x.position_list) != raw_positions]
        for gene in nearby:
            masked_positions.extend(gene.position_list)
        for mask in mask_hash.get_overlapping_features(gene_ivc_raw):
            masked_positions.extend(mask.position_list)
        masked_positions = set(masked_positions)
        total_mask = _chain_of(chrom, strand, raw_positions & masked_positions)
        post_mask = _chain_of(chrom, strand, raw_positions - masked_positions)
        masked_positions = set(total_mask.position_list)
        tmp_positions = {"utr5": set(), "cds": set(), "utr3": set()}
        txids = sorted(my_txids)
        for txid in txids:
            tx = transcripts[txid]
            tmp_positions["utr5"] |= set(tx_utr5(tx).position_list)
            tmp_positions["cds"] |= set(tx_cds(tx).position_list)
            tmp_positions["utr3"] |= set(tx_utr3(tx).position_list)
        for txid in txids:
            tx = transcripts[txid]
            tpos = {"utr5": set(tx_utr5(tx).position_list), "cds": set(tx_cds(tx).position_list),
                    "utr3": set(tx_utr3(tx).position_list)}
            for key1, key2 in keycombos:
                tpos[key1] -= tmp_positions[key2]
                tpos[key1] -= masked_positions
            row = {"region": txid, "exon": str(_chain_of(chrom, strand, set(tx.position_list) - masked_positions)),
                   "masked": str(total_mask), "exon_unmasked": str(tx), "transcript_ids": txid}
            for k, v in tpos.items():
                row[k] = str(_chain_of(chrom, strand, v))
            tx_rows.append(row)
        tmp2 = {k: set(v) for k, v in tmp_positions.items()}
        for k1, k2 in keycombos:
            tmp_positions[k1] -= tmp2[k2]
            tmp_positions[k1] -= masked_positions
        row = {"region": gene_id, "transcript_ids": ",".join(sorted(my_txids)), "exon_unmasked": str(gene_ivc_raw),
               "masked": str(total_mask), "exon": str(post_mask)}
        for k in tmp_positions:
            row[k] = str(_chain_of(chrom, strand, tmp_positions[k]))
        gene_rows.append(row)
    gene_rows.sort(key=lambda r: r["region"])
    tx_rows.sort(key=lambda r: r["region"])
    return gene_rows, tx_rows, merged_genes
