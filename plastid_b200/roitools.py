"""Host-side region objects on the counting path: ``GenomicSegment`` and ``SegmentChain``.

Only what the hot path touches is provided, with the reference's names and behaviour
(``plastid/genomics/roitools.pyx``): block layout ``_set_segments`` :1388-1448 /
``merge_segments`` :257-306, string forms ``__str__`` :1597-1612 / ``from_str`` :3378-3418,
masks ``add_masks`` :2213-2301, coordinate conversion :2957-3106, ``get_counts`` :3221-3273 and
``get_masked_counts`` :3275-3315.  Masks are kept as sorted intervals (not per-position sets), so
``add_masks`` is O(#intervals) and lowers directly to the bit masks the gather kernels read.
"""
import re
import warnings

import numpy as np

from .map_factories import DataWarning

_segpat = re.compile(r"([^:]*):([0-9]+)-([0-9]+)\(([+-.])\)")
_ivcpat = re.compile(r"([^:]*):([^(]+)\(([+-.])\)")


class GenomicSegment(object):
    """``GenomicSegment(chrom, start, end, strand)``: 0-based half-open interval."""
    __slots__ = ("chrom", "start", "end", "strand")

    def __init__(self, chrom, start, end, strand):
        if strand not in ("+", "-", "."):
            raise ValueError("Strand must be '+', '-', or '.'. Got %r" % (strand,))
        if end < start:
            raise ValueError("GenomicSegment: start coordinate (%s) must be <= end coordinate (%s)." % (start, end))
        self.chrom, self.start, self.end, self.strand = chrom, int(start), int(end), strand

    @staticmethod
    def from_str(inp):
        chrom, s, e, strand = _segpat.search(inp).groups()
        return GenomicSegment(chrom, int(s), int(e), strand)

    def __len__(self):
        return self.end - self.start

    def __str__(self):
        return "%s:%s-%s(%s)" % (self.chrom, self.start, self.end, self.strand)

    def __repr__(self):
        return "<GenomicSegment %s>" % str(self)

    def _key(self):
        return (self.chrom, self.start, self.end, self.strand)

    def __eq__(self, other):
        return isinstance(other, GenomicSegment) and self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._key() < other._key()

    def __hash__(self):
        return hash(self._key())


def _merge_intervals(intervals):
    """Sorted union; touching intervals are merged (``right.start > left.end`` separates)."""
    out = []
    for a, b in sorted(intervals):
        if out and a <= out[-1][1]:
            if b > out[-1][1]:
                out[-1][1] = b
        else:
            out.append([a, b])
    return [(a, b) for a, b in out]


def _intersect_intervals(xs, ys):
    """Intersection of two sorted disjoint interval lists."""
    out, i, j = [], 0, 0
    while i < len(xs) and j < len(ys):
        a, b = max(xs[i][0], ys[j][0]), min(xs[i][1], ys[j][1])
        if a < b:
            out.append((a, b))
        if xs[i][1] < ys[j][1]:
            i += 1
        else:
            j += 1
    return out


def positions_to_segments(chrom, strand, positions):
    """Set of positions -> sorted list of maximal runs (``roitools.pyx`` ``positions_to_segments``)."""
    pos = np.unique(np.fromiter(positions, dtype=np.int64))
    if len(pos) == 0:
        return []
    cut = np.flatnonzero(np.diff(pos) != 1)
    starts = np.concatenate(([pos[0]], pos[cut + 1]))
    ends = np.concatenate((pos[cut], [pos[-1]])) + 1
    return [GenomicSegment(chrom, int(a), int(b), strand) for a, b in zip(starts, ends)]


class SegmentChain(object):
    """``SegmentChain(*segments, **attr)``: sorted, merged exon blocks on one chromosome strand."""

    def __init__(self, *segments, **attr):
        self.attr = dict(attr)
        self._mask_intervals = None
        self._set_segments(list(segments))

    def _set_segments(self, segments):
        if segments:
            chrom, strand = segments[0].chrom, segments[0].strand
            for s in segments:
                if s.chrom != chrom or s.strand != strand:
                    raise ValueError("Not all segments on same strand or chromosome: %s"
                                     % ", ".join(str(x) for x in segments))
            ivs = _merge_intervals([(s.start, s.end) for s in segments])
            self._segments = [GenomicSegment(chrom, a, b, strand) for a, b in ivs]
            self.chrom, self.strand = chrom, strand
            self.spanning_segment = GenomicSegment(chrom, ivs[0][0], ivs[-1][1], strand)
        else:
            self._segments = []
            self.chrom = self.strand = None
            self.spanning_segment = None
        # cumulative chain coordinate of each block's first base (genomic order)
        cum, total = [0], 0
        for s in self._segments:
            total += s.end - s.start
            cum.append(total)
        self.length = total
        self.masked_length = total
        self._cum = np.array(cum, dtype=np.int64)

    # -- container protocol ------------------------------------------------------------------
    def __len__(self):
        return len(self._segments)

    def __iter__(self):
        return iter(self._segments)

    def __getitem__(self, i):
        return self._segments[i]

    @property
    def segments(self):
        return list(self._segments)

    def get_name(self):
        for key in ("ID", "Name", "name"):
            if key in self.attr:
                return self.attr[key]
        return str(self)

    def __str__(self):
        if len(self) == 0:
            return "na"
        return "%s:%s(%s)" % (self.chrom, "^".join("%s-%s" % (s.start, s.end) for s in self), self.strand)

    def __repr__(self):
        return "<SegmentChain segments=%d bounds=%s name=%s>" % (len(self), self.spanning_segment, self.get_name())

    @staticmethod
    def from_str(inp):
        if inp in ("na", "nan", "None:(None)", "None", "none", None) or (isinstance(inp, float) and np.isnan(inp)):
            return SegmentChain()
        chrom, middle, strand = _ivcpat.search(inp).groups()
        segs = []
        for piece in middle.split("^"):
            a, b = piece.split("-")
            segs.append(GenomicSegment(chrom, int(a), int(b), strand))
        return SegmentChain(*segs)

    def as_bed(self, thickstart=None, thickend=None):
        """BED12 line (roitools.pyx:2660-2806 with its defaults: score 0, colour 0,0,0,
        ``thickstart`` / ``thickend`` from ``attr`` else the chain's start); "" for an empty chain."""
        if len(self) == 0:
            return ""
        span = self.spanning_segment
        thickstart = self.attr.get("thickstart", span.start) if thickstart is None else thickstart
        thickend = self.attr.get("thickend", span.start) if thickend is None else thickend
        fields = [span.chrom, span.start, span.end, self.get_name(), 0, span.strand, thickstart, thickend,
                  self.attr.get("color", "0,0,0"), len(self),
                  ",".join(str(len(x)) for x in self) + ",",
                  ",".join(str(x.start - span.start) for x in self) + ","]
        return "\t".join(str(x) for x in fields) + "\n"

    def get_position_list(self):
        out = []
        for s in self._segments:
            out.extend(range(s.start, s.end))
        return out

    def get_position_set(self):
        return set(self.get_position_list())

    # -- coordinates -------------------------------------------------------------------------
    def get_genomic_coordinate(self, x, stranded=True):
        if x < 0 or x >= self.length:
            raise IndexError("Position %s is outside bounds [0,%s) of SegmentChain %s" % (x, self.length, self))
        if stranded and self.strand == "-":
            x = self.length - 1 - x
        k = int(np.searchsorted(self._cum, x, side="right")) - 1
        return self.chrom, self._segments[k].start + int(x - self._cum[k]), self.strand

    def get_segmentchain_coordinate(self, chrom, genomic_x, strand, stranded=True):
        if chrom != self.chrom or strand != self.strand:
            raise ValueError("coordinate is on a different chromosome or strand")
        for k, s in enumerate(self._segments):
            if s.start <= genomic_x < s.end:
                x = int(self._cum[k]) + genomic_x - s.start
                if stranded and self.strand == "-":
                    x = self.length - 1 - x
                return x
        raise KeyError("Position %s:%s(%s) is not in SegmentChain %s" % (chrom, genomic_x, strand, self))

    def get_subchain(self, start, end, stranded=True, **extra_attr):
        """roitools.pyx:3121-3218: a python slice ``[start:end]`` of the position hash (so bounds outside
        the chain clamp instead of raising); ``ID`` becomes ``<name>_subchain``."""
        if start is None or end is None:
            raise TypeError("start and end may not be None")
        attr = dict(self.attr)
        attr.update(extra_attr)
        attr["ID"] = "%s_subchain" % self.get_name()
        if start == end:
            return SegmentChain(**attr)
        if stranded and self.strand == "-":
            start, end = self.length - end, self.length - start
        start, end, _ = slice(start, end).indices(self.length)
        segs = []
        for k, s in enumerate(self._segments):
            a = max(start, int(self._cum[k]))
            b = min(end, int(self._cum[k + 1]))
            if a < b:
                off = int(self._cum[k])
                segs.append(GenomicSegment(self.chrom, s.start + a - off, s.start + b - off, self.strand))
        return SegmentChain(*segs, **attr)

    # -- masks -------------------------------------------------------------------------------
    @property
    def mask_segments(self):
        if not self._mask_intervals:
            return []
        return [GenomicSegment(self.chrom, a, b, self.strand) for a, b in self._mask_intervals]

    def get_masks(self):
        return self.mask_segments

    def add_masks(self, *mask_segments):
        if len(mask_segments) == 0:
            return
        for m in mask_segments:          # check_segments, roitools.pyx:749-784
            if m.chrom != self.chrom or m.strand != self.strand:
                raise ValueError("Cannot add mask %s to chain %s: chromosome or strand mismatch" % (m, self))
        ivs = [(m.start, m.end) for m in mask_segments] + list(self._mask_intervals or [])
        ivs = _intersect_intervals(_merge_intervals(ivs), [(s.start, s.end) for s in self._segments])
        self._mask_intervals = _merge_intervals(ivs)
        self.masked_length = self.length - sum(b - a for a, b in self._mask_intervals)

    def reset_masks(self):
        self._mask_intervals = None
        self.masked_length = self.length

    def get_masks_as_segmentchain(self):
        return SegmentChain(*self.mask_segments)

    def position_mask(self):
        """0/1 per chain position in genomic order (``_position_mask``, roitools.pyx:2290-2295)."""
        m = np.zeros(self.length, dtype=np.uint8)
        for a, b in self._mask_intervals or ():
            for k, s in enumerate(self._segments):
                lo, hi = max(a, s.start), min(b, s.end)
                if lo < hi:
                    off = int(self._cum[k]) - s.start
                    m[lo + off:hi + off] = 1
        return m

    # -- counts ------------------------------------------------------------------------------
    def get_counts(self, ga, stranded=True):
        if len(self) == 0:
            warnings.warn("%s is a zero-length SegmentChain. Returning 0-length count vector." % self.get_name(),
                          DataWarning)
            return np.array([], dtype=float)
        fast = getattr(ga, "_chain_counts", None)
        if fast is not None:
            out = fast(self)                        # one gather for the whole chain, genomic order
        else:
            parts = [ga.get(seg, roi_order=False) for seg in self._segments]
            dims = list(parts[0].shape)
            dims[-1] = self.length
            out = np.empty(dims, dtype=float)
            i = 0
            for seg, part in zip(self._segments, parts):
                out[..., i:i + len(seg)] = part
                i += len(seg)
        if self.strand == "-" and stranded is True:
            out = out[..., ::-1]
        return out

    def get_masked_counts(self, ga, stranded=True, copy=False):
        counts = self.get_counts(ga)                # reference ignores `stranded` here (:3301)
        if self._mask_intervals is None:
            mask = np.zeros_like(counts)
        else:
            m = self.position_mask().astype(np.intc)
            if self.strand == "-":
                m = m[::-1]
            mask = np.empty_like(counts)
            mask[..., :] = m
        return np.ma.MaskedArray(counts, mask=mask.astype(bool), copy=copy)


class Transcript(SegmentChain):
    """``Transcript(*segments, cds_genome_start=None, cds_genome_end=None, **attr)``: a chain with a
    coding region (roitools.pyx:3565-3913).  ``cds_start`` / ``cds_end`` are transcript coordinates
    derived like ``_update_cds`` (:3883-3913), including its end-of-exon handling."""

    def __init__(self, *segments, **attr):
        SegmentChain.__init__(self, *segments, **attr)
        if "type" not in self.attr:
            self.attr["type"] = "mRNA"
        self.cds_genome_start = attr.get("cds_genome_start", None)
        self.cds_genome_end = attr.get("cds_genome_end", None)
        self.cds_start = self.cds_end = None
        if self.cds_genome_start is not None and self.cds_genome_end is not None:
            self._update_cds()
        else:
            self.cds_genome_start = self.cds_genome_end = None

    def _update_cds(self):
        gs, ge = int(self.cds_genome_start), int(self.cds_genome_end)
        coord = lambda x: self.get_segmentchain_coordinate(self.chrom, x, self.strand)   # noqa: E731
        if self.strand == "+":                       # roitools.pyx:3896: an unstranded transcript takes the minus branch
            self.cds_start = coord(gs)
            try:
                self.cds_end = coord(ge)
            except KeyError:                      # half-open end coincides with an exon end (:3900-3905)
                self.cds_end = 1 + coord(ge - 1)
        else:
            self.cds_start = coord(ge - 1)
            self.cds_end = 1 + coord(gs)

    def as_bed(self, thickstart=None, thickend=None):
        """BED12 line with the coding region in the thickStart / thickEnd columns (roitools.pyx:4385-4455)."""
        return SegmentChain.as_bed(self, thickstart=self.cds_genome_start if thickstart is None else thickstart,
                                   thickend=self.cds_genome_end if thickend is None else thickend)

    def get_gene(self):
        """``gene_id``, else ``Parent``, else ``gene_<name>`` (roitools.pyx:2154-2173)."""
        gene = self.attr.get("gene_id", self.attr.get("Parent", "gene_%s" % self.get_name()))
        if isinstance(gene, list):
            gene = ",".join(sorted(gene))
        return gene

    def _sub(self, start, end, suffix, **extra_attr):
        if self.cds_genome_start is None or self.cds_genome_end is None:
            return SegmentChain()
        sub = SegmentChain.get_subchain(self, start, end)
        sub.attr = dict(gene_id=self.get_gene(), transcript_id=self.get_name(), ID="%s_%s" % (self.get_name(), suffix))
        sub.attr.update(extra_attr)
        return sub

    def get_cds(self, **extra_attr):              # roitools.pyx:4005-4046
        return self._sub(self.cds_start, self.cds_end, "CDS", **extra_attr)

    def get_utr5(self, **extra_attr):             # :4048-4087
        return self._sub(0, self.cds_start, "5UTR", type="5UTR", **extra_attr)

    def get_utr3(self, **extra_attr):             # :4089-4130
        return self._sub(self.cds_end, self.length, "3UTR", type="3UTR", **extra_attr)
