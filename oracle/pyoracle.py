"""Pure-Python restatement of plastid's mapping rules and count containers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function follows
the reference's per-read loop one statement at a time, quirks included, and
cites the lines it restates (paths relative to ``/root/reference``).  It works
on duck-typed read objects (``.positions``, ``.is_reverse``), so the
reference's own unit tests can be transcribed without pysam.

Parity: pinned for ``<L>M`` reads by ``test_map_factories.py:17-200``;
**parity unpinned** against pysam for I/D/N/S/H/P ops (pysam is absent).
"""
import warnings

import numpy as np

# CIGAR op codes, kent/src/htslib/htslib/sam.h:64-73
CMATCH, CINS, CDEL, CREF_SKIP, CSOFT_CLIP, CHARD_CLIP, CPAD, CEQUAL, CDIFF, CBACK = range(10)
BAD_OFFSET = -1            # map_factories.pyx: _BAD_OFFSET
LUT_SIZE = 10000           # map_factories.pxd:11-12


class DataWarning(Warning):
    """plastid/util/services/exceptions.py DataWarning stand-in."""


def positions_from_cigar(reference_start, cigartuples):
    """pysam 0.19.0 ``AlignedSegment.get_reference_positions()`` [3rd-party]:
    M/=/X emit ``len`` consecutive reference positions and advance; D/N advance
    only; I/S/H/P neither emit nor advance (consume table:
    kent/src/htslib/htslib/sam.h:79-104)."""
    out = []
    pos = reference_start
    for op, n in cigartuples:
        if op in (CMATCH, CEQUAL, CDIFF):
            out.extend(range(pos, pos + n))
            pos += n
        elif op in (CDEL, CREF_SKIP):
            pos += n
    return out


class Read(object):
    """Minimal stand-in for ``pysam.AlignedSegment`` as used on the hot path
    (call sites: map_factories.pyx:243,349,448,629,769,838; genome_array.py:813-815)."""
    __slots__ = ("reference_start", "cigartuples", "is_reverse", "reference_id", "_pos")

    def __init__(self, reference_start, cigartuples, is_reverse, reference_id=0):
        self.reference_start = int(reference_start)
        self.cigartuples = [(int(a), int(b)) for a, b in cigartuples]
        self.is_reverse = bool(is_reverse)
        self.reference_id = reference_id
        self._pos = None

    @property
    def positions(self):
        if self._pos is None:
            self._pos = positions_from_cigar(self.reference_start, self.cigartuples)
        return self._pos

    @property
    def reference_end(self):
        pos = self.reference_start
        for op, n in self.cigartuples:
            if op in (CMATCH, CEQUAL, CDIFF, CDEL, CREF_SKIP):
                pos += n
        return pos

    def __repr__(self):
        return "Read(%d,%r,%s)" % (self.reference_start, self.cigartuples, "-" if self.is_reverse else "+")


class Seg(object):
    """GenomicSegment stand-in: chrom, start, end (half-open), strand."""
    __slots__ = ("chrom", "start", "end", "strand")

    def __init__(self, chrom, start, end, strand):
        self.chrom, self.start, self.end, self.strand = chrom, int(start), int(end), strand

    def __len__(self):
        return self.end - self.start

    def __repr__(self):
        return "%s:%d-%d(%s)" % (self.chrom, self.start, self.end, self.strand)


# ---------------------------------------------------------------------------
# mapping rules  (map_factories.pyx:167-839)
# ---------------------------------------------------------------------------
class CenterMap(object):
    """map_factories.pyx:167-275"""

    def __init__(self, nibble=0):
        if nibble < 0:
            raise ValueError("nibble must be >= 0")
        self.nibble = nibble

    def __call__(self, reads, seg):
        s, n = seg.start, seg.end - seg.start
        counts = np.zeros(n, dtype=np.float64)                      # :230
        kept, warn = [], False
        for read in reads:
            pos = read.positions                                   # :243
            L = len(pos)
            m = L - 2 * self.nibble                                 # :245
            if m < 0:                                               # :246-248
                warn = True
                continue
            elif m > 0:
                v = 1.0 / m                                         # :250
                for i in range(self.nibble, L - self.nibble):       # :251-254
                    c = pos[i] - s
                    if 0 <= c < n:
                        counts[c] += v
                kept.append(read)                                   # :256 (even if nothing landed)
        if warn:
            warnings.warn("Data contains read alignments shorter than `2*nibble` value of '%s' nt. Ignoring these."
                          % (2 * self.nibble), DataWarning)
        return kept, counts


class _EndMap(object):
    def __init__(self, offset=0):
        if offset < 0:
            raise ValueError("offset must be >= 0")
        self.offset = offset

    def _index(self, seg):
        raise NotImplementedError

    def __call__(self, reads, seg):
        counts = np.zeros(seg.end - seg.start, dtype=np.int64)      # :334 / :433
        idx = self._index(seg)
        kept, warn = [], False
        for read in reads:
            pos = read.positions
            if self.offset >= len(pos):                             # :351-353 / :450-452
                warn = True
                continue
            p = pos[idx]                                            # :355 / :454 (python negative index)
            if seg.start <= p < seg.end:
                kept.append(read)
                counts[p - seg.start] += 1
        if warn:
            warnings.warn("Data contains read alignments shorter than offset (%s nt). Ignoring." % self.offset,
                          DataWarning)
        return kept, counts


class FivePrimeMap(_EndMap):
    """map_factories.pyx:278-374"""

    def _index(self, seg):
        return -self.offset - 1 if seg.strand == "-" else self.offset   # :345-346


class ThreePrimeMap(_EndMap):
    """map_factories.pyx:377-474"""

    def _index(self, seg):
        return -self.offset - 1 if seg.strand != "-" else self.offset   # :444-445


def build_offset_luts(offset_dict):
    """map_factories.pyx:494-543 → (forward int32[10000], reverse int32[10000])."""
    fw = np.full(LUT_SIZE, BAD_OFFSET, dtype=np.int32)              # :511-512
    rc = np.full(LUT_SIZE, BAD_OFFSET, dtype=np.int32)
    if offset_dict is None:                                         # :517-518
        offset_dict = {"default": 0}
    have_default = "default" in offset_dict
    if have_default:                                                # :520-526
        default = int(offset_dict["default"])
        fw[default + 1:] = default
        for i in range(default + 1, LUT_SIZE):
            rc[i] = i - default - 1
    for length, off in offset_dict.items():                         # :530-543
        if length == "default":
            continue
        if off >= length:
            if not have_default:
                raise UnboundLocalError("default")                  # :533 reads an unbound local
            if length >= default:
                warnings.warn("Given offset '%s' longer than read length '%s'. Falling back to default '%s'."
                              % (off, length, default), DataWarning)
            else:
                warnings.warn("Given offset '%s' and default '%s' are longer than read length '%s'. Ignoring %s-mers."
                              % (off, default, length, length), DataWarning)
            continue                                                # :540 entry skipped either way
        fw[length] = off
        rc[length] = length - off - 1
    return fw, rc


class MalformedFileError(Exception):
    """plastid/util/services/exceptions.py MalformedFileError stand-in."""


def parse_offset_file(fh):
    """argparsers.py:2505-2561 ``_parse_variable_offset_file``: tab-separated
    ``length<TAB>offset`` lines (``default`` allowed as a key); header lines
    starting with ``length`` skipped (:2525-2526, :2533-2534); wrong column count,
    non-integer key/value, or a repeated key raise MalformedFileError.  The
    '#'-comment stripping is done by the CommentReader wrapper at
    map_factories.pyx:580."""
    out = {}
    for line in fh:
        if line.startswith("#") or line.startswith("length"):
            continue
        items = line.strip("\n").split("\t")
        if len(items) != 2:
            raise MalformedFileError("More or fewer than two columns on line: %r" % line)
        key = items[0]
        try:
            key = key if key == "default" else int(key)
        except ValueError:
            raise MalformedFileError("Non integer value for key %r" % key)
        if key in out:
            raise MalformedFileError("multiple offsets defined for read length %s" % key)
        try:
            out[key] = int(items[1])
        except ValueError:
            raise MalformedFileError("Non integer value for value %r" % items[1])
    return out


class VariableFivePrimeMap(object):
    """map_factories.pyx:477-650"""

    def __init__(self, offset_dict):
        self.fw, self.rc = build_offset_luts(offset_dict)

    def __call__(self, reads, seg):
        lut = self.rc if seg.strand == "-" else self.fw             # :625-626
        counts = np.zeros(seg.end - seg.start, dtype=np.int64)
        kept, bad = [], None
        for read in reads:
            pos = read.positions
            L = len(pos)
            off = int(lut[L])                                       # :631 (L>=10000 is UB there)
            if off == BAD_OFFSET:                                   # :633-636
                bad = L
                continue
            p = pos[off]
            if seg.start <= p < seg.end:
                kept.append(read)
                counts[p - seg.start] += 1
        if bad is not None:
            warnings.warn("No usable offset for reads of length %s nt in offset dict. Ignoring these." % bad,
                          DataWarning)
        return kept, counts


class StratifiedVariableFivePrimeMap(VariableFivePrimeMap):
    """map_factories.pyx:653-791.  Quirk kept: no BAD_OFFSET test (:773-774), so
    a length without an offset indexes ``positions[-1]``."""

    def __init__(self, offset_dict, min=25, max=35):
        VariableFivePrimeMap.__init__(self, offset_dict)
        if max <= min:
            raise ValueError("Max length must be >= min length")
        self.min_length, self.max_length = min, max
        self.shape = [max - min + 1]
        self.row_keys = np.arange(min, max + 1)

    def __call__(self, reads, seg):
        lut = self.rc if seg.strand == "-" else self.fw
        counts = np.zeros((self.shape[0], seg.end - seg.start), dtype=np.int64)
        kept = []
        for read in reads:
            pos = read.positions
            L = len(pos)
            if self.min_length <= L <= self.max_length:             # :771
                p = pos[int(lut[L])]
                if seg.start <= p < seg.end:
                    kept.append(read)
                    counts[L - self.min_length, p - seg.start] += 1
        return kept, counts


class SizeFilter(object):
    """map_factories.pyx:794-839"""

    def __init__(self, min=1, max=-1):
        if max != -1 and max < min:
            raise ValueError("max read length must be >= min read length")
        if min < 1:
            raise ValueError("min read length must be >= 1")
        self.min_, self.max_ = min, max

    def __call__(self, read):
        L = len(read.positions)
        return L >= self.min_ and (L <= self.max_ or self.max_ == -1)


# ---------------------------------------------------------------------------
# BAMGenomeArray semantics over an in-memory read store (genome_array.py:626-988)
# ---------------------------------------------------------------------------
class ReadStore(object):
    """In-memory stand-in for an indexed BAM: ``fetch`` returns, in coordinate order (stable for
    ties), reads whose reference span overlaps ``[start,end)`` (pysam ``AlignmentFile.fetch``
    [3rd-party]).  A bisect on the sorted starts only narrows the scan; the overlap test decides."""

    def __init__(self, chrom_lengths, reads_by_chrom):
        import bisect
        self._bisect = bisect
        self.lengths = dict(chrom_lengths)
        self.references = list(chrom_lengths)
        self.reads = {c: sorted(r, key=lambda x: x.reference_start) for c, r in reads_by_chrom.items()}
        self.starts = {c: [r.reference_start for r in v] for c, v in self.reads.items()}
        self.max_span = {c: max([r.reference_end - r.reference_start for r in v] + [1]) for c, v in self.reads.items()}
        self.mapped = sum(len(v) for v in self.reads.values())

    def fetch(self, reference, start, end):
        reads = self.reads.get(reference, ())
        if not reads:
            return
        starts = self.starts[reference]
        lo = self._bisect.bisect_left(starts, start - self.max_span[reference])
        hi = self._bisect.bisect_left(starts, end)
        for r in reads[lo:hi]:
            if r.reference_start < end and r.reference_end > start:
                yield r


class OracleBAMGenomeArray(object):
    """genome_array.py:582-988 restated over :class:`ReadStore` objects."""

    def __init__(self, *stores, **kwargs):
        self.stores = list(stores)
        self.map_fn = kwargs.get("mapping", CenterMap())            # :663
        self._normalize = False
        self._chr_lengths = {}
        for st in self.stores:                                      # :667-672
            for k, v in st.lengths.items():
                self._chr_lengths[k] = max(self._chr_lengths.get(k, 0), v)
        self._chroms = sorted(self._chr_lengths)
        self._filters = {}
        self.reset_sum()

    def reset_sum(self):
        self._sum = sum(st.mapped for st in self.stores)            # :690

    def sum(self):
        return self._sum

    def set_sum(self, val):
        self._sum = val

    def set_normalize(self, value=True):
        self._normalize = value

    def set_mapping(self, fn):
        self.map_fn = fn                                            # :962-963
        self.reset_sum()

    def add_filter(self, name, fn):
        self._filters[name] = fn

    def remove_filter(self, name):
        return self._filters.pop(name)

    def chroms(self):
        return self._chroms

    def lengths(self):
        return self._chr_lengths

    def strands(self):
        return ("+", "-", ".")

    def get_reads_and_counts(self, roi, roi_order=True):
        if roi.chrom not in self._chroms:                           # :795-798 (length-1 quirk)
            return [], np.zeros([1] + getattr(self.map_fn, "shape", []))
        reads = []
        for st in self.stores:                                      # :800-809
            reads.extend(st.fetch(roi.chrom, roi.start, roi.end))
        if roi.strand == "+":                                       # :811-815
            reads = [r for r in reads if r.is_reverse is False]
        elif roi.strand == "-":
            reads = [r for r in reads if r.is_reverse is True]
        for f in self._filters.values():                            # :819-820
            reads = [r for r in reads if f(r)]
        reads, counts = self.map_fn(list(reads), roi)               # :823
        if self._normalize is True:                                 # :826-827
            counts = counts / float(self.sum()) * 1e6
        if roi_order and roi.strand == "-":                         # :829-830
            counts = counts[..., ::-1]
        return reads, counts

    def get_reads(self, roi):
        return self.get_reads_and_counts(roi)[0]

    def get(self, roi, roi_order=True):
        if isinstance(roi, Chain):                                  # :923-924
            return roi.get_counts(self)
        return self.get_reads_and_counts(roi, roi_order=roi_order)[1]

    def to_variable_step(self, fh, trackname, strand, window_size=100000, **kwargs):   # :990-1037
        assert strand in self.strands()
        fh.write("track type=wiggle_0 name=%s" % trackname)
        for k, v in sorted(kwargs.items(), key=lambda x: x[0]):
            fh.write(" %s=%s" % (k, v))
        fh.write("\n")
        for chrom in sorted(self.chroms()):
            my_size = self.lengths()[chrom]
            fh.write("variableStep chrom=%s span=1\n" % chrom)
            for my_start in range(0, my_size, window_size):
                my_end = min(my_start + window_size, my_size)
                my_counts = self.get(Seg(chrom, my_start, my_end, strand), roi_order=False)
                if my_counts.sum() > 0:
                    for idx in my_counts.nonzero()[0]:
                        fh.write("%s\t%s\n" % (my_start + idx + 1, my_counts[idx]))

    def to_bedgraph(self, fh, trackname, strand, window_size=100000, **kwargs):        # :1039-1111
        assert strand in self.strands()
        assert window_size > 0
        fh.write("track type=bedGraph name=%s" % trackname)
        for k, v in sorted(kwargs.items(), key=lambda x: x[0]):
            fh.write(" %s=%s" % (k, v))
        fh.write("\n")
        for chrom in sorted(self.chroms()):
            my_size = self.lengths()[chrom]
            for my_start in range(0, my_size, window_size):
                my_end = min(my_start + window_size, my_size)
                my_counts = self.get(Seg(chrom, my_start, my_end, strand), roi_order=False)
                if my_counts.sum() > 0:
                    genomic_start_x = my_start
                    last_val = my_counts[0]
                    for x, val in enumerate(my_counts[1:]):
                        if val != last_val:
                            genomic_end_x = 1 + x + my_start
                            if last_val > 0:
                                fh.write("%s\t%s\t%s\t%s\n" % (chrom, genomic_start_x, genomic_end_x, last_val))
                            last_val = val
                            genomic_start_x = genomic_end_x
                    if last_val > 0:
                        fh.write("%s\t%s\t%s\t%s\n" % (chrom, genomic_start_x, my_end, last_val))

    def __getitem__(self, roi):
        return self.get(roi, roi_order=True)


# ---------------------------------------------------------------------------
# SegmentChain pieces on the path (roitools.pyx:257-306, 1388-1484, 2213-2301, 3221-3315)
# ---------------------------------------------------------------------------
class Chain(object):
    """Sorted, merged exon blocks on one chrom/strand + optional masks."""

    def __init__(self, *segs):
        segs = sorted(segs, key=lambda s: (s.start, s.end))
        merged = []                                                 # merge_segments :257-306
        for s in segs:
            if merged and s.start <= merged[-1].end:
                if s.end > merged[-1].end:
                    merged[-1] = Seg(s.chrom, merged[-1].start, s.end, s.strand)
            else:
                merged.append(Seg(s.chrom, s.start, s.end, s.strand))
        self.segments = merged
        self.chrom = merged[0].chrom if merged else None
        self.strand = merged[0].strand if merged else None
        self.length = sum(len(s) for s in merged)
        self.position_list = [p for s in merged for p in range(s.start, s.end)]   # :1450-1484
        self.position_mask = None
        self.masked_length = self.length

    @classmethod
    def from_str(cls, text):
        """roitools.pyx:3378-3418: ``chrom:s-e^s-e(strand)``; ``na`` = empty."""
        if text == "na":
            return cls()
        body, strand = text.rsplit("(", 1)
        strand = strand.rstrip(")")
        chrom, spans = body.rsplit(":", 1)
        segs = []
        for sp in spans.split("^"):
            a, b = sp.split("-")
            segs.append(Seg(chrom, int(a), int(b), strand))
        return cls(*segs)

    def __str__(self):                                              # :1597-1612
        if not self.segments:
            return "na"
        return "%s:%s(%s)" % (self.chrom, "^".join("%d-%d" % (s.start, s.end) for s in self.segments), self.strand)

    def __len__(self):
        return len(self.segments)

    def add_masks(self, *masks):
        """roitools.pyx:2213-2301: union of mask positions ∩ chain positions."""
        if not masks:
            return
        for m in masks:                                             # check_segments :749-784
            if m.chrom != self.chrom or m.strand != self.strand:
                raise ValueError("mask chrom/strand mismatch")
        covered = set()
        for m in masks:
            covered |= set(range(m.start, m.end))
        old = self.position_mask or [0] * self.length
        new = [1 if (p in covered or old[i]) else 0 for i, p in enumerate(self.position_list)]
        self.position_mask = new
        self.masked_length = self.length - sum(new)                 # :2298

    def get_counts(self, ga, stranded=True):
        if len(self) == 0:                                          # :3248-3253
            return np.array([], dtype=float)
        parts = [ga.get(s, roi_order=False) for s in self.segments]   # :3259
        dims = list(parts[0].shape)
        dims[-1] = self.length
        out = np.empty(dims, dtype=float)                           # :3262
        i = 0
        for s, part in zip(self.segments, parts):
            out[..., i:i + len(s)] = part
            i += len(s)
        if self.strand == "-" and stranded is True:                 # :3270-3271
            out = out[..., ::-1]
        return out

    def get_masked_counts(self, ga, stranded=True, copy=False):
        counts = self.get_counts(ga)                                # :3301 (ignores `stranded`)
        if self.position_mask is None:
            mask = np.zeros_like(counts)
        else:
            m = np.asarray(self.position_mask, dtype=np.intc)
            if self.strand == "-":                                  # :3308-3310
                m = m[::-1]
            mask = np.empty_like(counts)
            mask[..., :] = m
        return np.ma.MaskedArray(counts, mask=mask.astype(bool), copy=copy)


# ---------------------------------------------------------------------------
# GenomeHash overlap query (genome_hash.py:81, 236-436) + SegmentChain.overlaps (roitools.pyx:1811-1876)
# ---------------------------------------------------------------------------
_STRAND_BITS = {"+": 1, "-": 2, ".": 3}                             # c_common.pxd:1-6


def chains_overlap(a, b):
    """SegmentChain.overlaps: same chromosome, strand bits intersect (roitools.pyx:1873), and two
    neighbours of the merged sorted segment list overlap (c_unstranded_overlaps :1811-1836)."""
    if a.chrom != b.chrom or not (_STRAND_BITS[a.strand] & _STRAND_BITS[b.strand]):
        return False
    segs = sorted(a.segments + b.segments, key=lambda s: (s.start, s.end))
    return any(r.start < l.end for l, r in zip(segs[:-1], segs[1:]))


class GenomeHash(object):
    """``GenomeHash(features, binsize=20000)`` restated: features are binned by chromosome, strand
    and ``start // binsize .. end // binsize`` of each segment; only '+' and '-' tables exist
    (:236-257), so a '.' feature raises KeyError exactly like the reference."""

    def __init__(self, features, binsize=20000):
        self.binsize = binsize
        self.features = list(features)
        self._hash = {}
        for fid, f in enumerate(self.features):
            if f.chrom not in self._hash:
                self._hash[f.chrom] = {"+": {}, "-": {}}
            for b in self._bins(f):
                try:
                    self._hash[f.chrom][f.strand][b].append(fid)
                except KeyError:
                    self._hash[f.chrom][f.strand][b] = [fid]        # KeyError again for '.'

    def _bins(self, chain):                                         # :259-291
        bins = []
        for seg in chain.segments:
            bins.extend(range(seg.start // self.binsize, seg.end // self.binsize + 1))
        return list(set(bins))

    def get_overlapping_features(self, roi, stranded=True):         # :300-436, stranded query only
        assert stranded is True
        try:
            nearby = self._hash[roi.chrom][roi.strand]
        except KeyError:
            nearby = {}
        ids = set()
        for b in self._bins(roi):
            ids.update(nearby.get(b, ()))
        return [self.features[i] for i in sorted(ids) if chains_overlap(roi, self.features[i])]
