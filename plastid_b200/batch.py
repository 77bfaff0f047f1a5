"""Packed SoA alignment batches and the dense genome layout (host side).

What pysam hands the reference one ``AlignedSegment`` at a time (``reference_start``,
``positions``, ``is_reverse``; call sites ``plastid/genomics/map_factories.pyx:243,349,448,629``
and ``plastid/genomics/genome_array.py:800-815``) is packed here once into flat arrays the
kernels stream.  Schema: ``include/plastid_b200.h``.
"""
import numpy as np

from . import _lib

# CIGAR consume table: kent/src/htslib/htslib/sam.h:79-104 (bit0 query, bit1 reference)
_CIGAR_ALIGNED = (0, 7, 8)      # M, =, X : emit reference positions
_CIGAR_REF_ONLY = (2, 3)        # D, N    : advance reference only

MAX_ALIGNED_LEN = 0xFFFF
MAX_BLOCKS = 255


def cigar_to_blocks(cigartuples):
    """Maximal runs of reference-aligned bases as ``[(rel_start, length), ...]`` plus the
    reference span.  Same walk as pysam's ``get_reference_positions`` [3rd-party]: M/=/X emit
    and advance, D/N advance, I/S/H/P do neither (so ``10M2I10M`` is ONE 20-base block)."""
    blocks = []
    pos = 0
    for op, n in cigartuples:
        if op in _CIGAR_ALIGNED:
            if n <= 0:
                continue
            if blocks and blocks[-1][0] + blocks[-1][1] == pos:
                blocks[-1][1] += n
            else:
                blocks.append([pos, n])
            pos += n
        elif op in _CIGAR_REF_ONLY:
            pos += n
    return [(a, b) for a, b in blocks], pos


class GenomeLayout(object):
    """Placement of chromosomes in the concatenated dense count planes."""

    def __init__(self, chroms, lengths):
        self.chroms = list(chroms)
        self.chrom_len = np.asarray(lengths, dtype=np.int64)
        if len(self.chroms) != len(self.chrom_len):
            raise ValueError("chroms and lengths differ in length")
        if len(self.chroms) == 0:
            raise ValueError("a genome layout needs at least one chromosome")
        align = _lib.PB_LAYOUT_ALIGN
        padded = np.maximum((self.chrom_len + align - 1) // align, 1) * align
        self.chrom_bin_off = np.zeros(len(self.chroms) + 1, dtype=np.int64)
        np.cumsum(padded, out=self.chrom_bin_off[1:])
        self.total_bins = int(self.chrom_bin_off[-1])
        self.index = {c: i for i, c in enumerate(self.chroms)}
        self._dev = {}

    def bin_of(self, chrom, pos):
        return int(self.chrom_bin_off[self.index[chrom]]) + int(pos)

    def device_tables(self, device):
        import torch
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.chrom_len).to(device),
                              torch.from_numpy(self.chrom_bin_off).to(device))
        return self._dev[key]

    def c_struct(self, device):
        clen, coff = self.device_tables(device)
        return _lib.PbLayout(len(self.chroms), 0, clen.data_ptr(), coff.data_ptr(), self.total_bins)


class AlignmentBatch(object):
    """Host SoA batch, sorted by (chromosome index, ref_start)."""

    def __init__(self, chroms, chrom_len, ref_start, meta, chrom_read_off, blk_off=None, blk=None,
                 max_span=None, mapped=None):
        self.chroms = list(chroms)
        self.chrom_len = np.asarray(chrom_len, dtype=np.int64)
        self.ref_start = np.ascontiguousarray(ref_start, dtype=np.int32)
        self.meta = np.ascontiguousarray(meta, dtype=np.uint32)
        self.chrom_read_off = np.ascontiguousarray(chrom_read_off, dtype=np.int64)
        self.blk_off = None if blk_off is None else np.ascontiguousarray(blk_off, dtype=np.uint32)
        self.blk = None if blk is None else np.ascontiguousarray(blk, dtype=np.int32).reshape(-1, 2)
        n = len(self.ref_start)
        if len(self.meta) != n or len(self.chrom_read_off) != len(self.chroms) + 1:
            raise ValueError("inconsistent batch arrays")
        if self.chrom_read_off[0] != 0 or self.chrom_read_off[-1] != n:
            raise ValueError("chrom_read_off must run from 0 to n_reads")
        if (self.blk_off is None) != (self.blk is None):
            raise ValueError("blk_off and blk go together")
        if max_span is None:
            max_span = self.compute_max_span()
        self.max_span = int(max_span)
        self.mapped = n if mapped is None else int(mapped)   # what `bamfile.mapped` reports
        self.objects = None      # optional list of the original read objects, batch order
        self._dev = {}

    def __len__(self):
        return len(self.ref_start)

    @property
    def aligned_len(self):
        return (self.meta & 0xFFFF).astype(np.int64)

    @property
    def is_reverse(self):
        return ((self.meta >> 16) & 1).astype(bool)

    def compute_max_span(self):
        if len(self.ref_start) == 0:
            return 1
        span = int((self.meta & 0xFFFF).max())
        if self.blk is not None and len(self.blk):
            span = max(span, int((self.blk[:, 0] + self.blk[:, 1]).max()))
        return max(span, 1)

    def check_sorted(self):
        for c in range(len(self.chroms)):
            a, b = self.chrom_read_off[c], self.chrom_read_off[c + 1]
            if b - a > 1 and np.any(np.diff(self.ref_start[a:b]) < 0):
                raise ValueError("reads of chromosome %s are not sorted by ref_start" % self.chroms[c])

    def with_drop_mask(self, drop):
        """Copy of the batch with the host filter verdicts (True = drop) in meta bit 17."""
        meta = (self.meta & ~np.uint32(1 << 17)) | (np.asarray(drop, dtype=np.uint32) << 17)
        out = AlignmentBatch(self.chroms, self.chrom_len, self.ref_start, meta, self.chrom_read_off,
                             self.blk_off, self.blk, self.max_span, self.mapped)
        out.objects = self.objects
        return out

    def pinned(self):
        """Pinned host tensors (for timed H2D copies)."""
        import torch
        out = {}
        for name in ("ref_start", "meta", "blk_off", "blk", "chrom_read_off"):
            a = getattr(self, name)
            if a is None:
                out[name] = None
                continue
            t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a)
            out[name] = t.pin_memory()
        return out

    def to_device(self, device="cuda"):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = DeviceBatch.from_host(self, device)
        return self._dev[key]

    # -- read views (object protocol of the reference) ------------------------------------
    def positions_of(self, i):
        s = int(self.ref_start[i])
        m = int(self.meta[i])
        L, nblk = m & 0xFFFF, m >> 24
        if nblk <= 1 or self.blk_off is None:
            return list(range(s, s + L))
        out = []
        for k in range(int(self.blk_off[i]), int(self.blk_off[i + 1])):
            a, n = int(self.blk[k, 0]), int(self.blk[k, 1])
            out.extend(range(s + a, s + a + n))
        return out

    def read_view(self, i):
        if self.objects is not None:
            return self.objects[i]
        return BatchRead(self, i)


class BatchRead(object):
    """Duck-typed stand-in for ``pysam.AlignedSegment`` built from one batch row."""
    __slots__ = ("batch", "index")

    def __init__(self, batch, index):
        self.batch, self.index = batch, int(index)

    @property
    def reference_start(self):
        return int(self.batch.ref_start[self.index])

    @property
    def is_reverse(self):
        return bool((int(self.batch.meta[self.index]) >> 16) & 1)

    @property
    def positions(self):
        return self.batch.positions_of(self.index)

    def get_reference_positions(self):
        return self.positions

    def __eq__(self, other):
        return isinstance(other, BatchRead) and other.batch is self.batch and other.index == self.index

    def __hash__(self):
        return hash((id(self.batch), self.index))

    def __repr__(self):
        return "<BatchRead #%d start=%d %s>" % (self.index, self.reference_start, "-" if self.is_reverse else "+")


class DeviceBatch(object):
    """Device-resident mirror of an :class:`AlignmentBatch` (torch tensors used as buffers only)."""

    def __init__(self, n_reads, n_chrom, max_span, ref_start, meta, chrom_read_off, blk_off=None, blk=None):
        self.n_reads, self.n_chrom, self.max_span = int(n_reads), int(n_chrom), int(max_span)
        self.ref_start, self.meta, self.chrom_read_off = ref_start, meta, chrom_read_off
        self.blk_off, self.blk = blk_off, blk

    @classmethod
    def from_host(cls, hb, device, non_blocking=False):
        import torch

        def up(a):
            if a is None:
                return None
            t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a)
            return t.to(device, non_blocking=non_blocking)
        return cls(len(hb), len(hb.chroms), hb.max_span, up(hb.ref_start), up(hb.meta),
                   up(hb.chrom_read_off), up(hb.blk_off), up(hb.blk))

    @property
    def device(self):
        return self.ref_start.device

    def c_struct(self):
        return _lib.PbBatch(self.n_reads, self.ref_start.data_ptr(), self.meta.data_ptr(),
                            None if self.blk_off is None else self.blk_off.data_ptr(),
                            None if self.blk is None else self.blk.data_ptr(),
                            self.chrom_read_off.data_ptr(), self.n_chrom, self.max_span)


def pack_reads(reads_by_chrom, chrom_lengths, keep_objects=True, mapped=None):
    """Pack read objects (anything with ``reference_start``, ``cigartuples``, ``is_reverse``) into
    an :class:`AlignmentBatch`.  ``reads_by_chrom``: ``{chrom: [reads]}``; ``chrom_lengths``:
    ordered ``{chrom: length}`` (BAM header order)."""
    chroms = list(chrom_lengths)
    starts, metas, blk_off, blks, objs = [], [], [0], [], []
    chrom_read_off = [0]
    any_multi = False
    for chrom in chroms:
        recs = []
        for r in reads_by_chrom.get(chrom, ()):
            blocks, _span = cigar_to_blocks(r.cigartuples or ())
            L = sum(n for _a, n in blocks)
            if L > MAX_ALIGNED_LEN:
                raise ValueError("aligned length %d exceeds %d" % (L, MAX_ALIGNED_LEN))
            if len(blocks) > MAX_BLOCKS:
                raise ValueError("read has %d aligned blocks (max %d)" % (len(blocks), MAX_BLOCKS))
            start = int(r.reference_start)
            if blocks and blocks[0][0] != 0:       # leading D/N: positions start after it
                shift = blocks[0][0]
                start += shift
                blocks = [(a - shift, n) for a, n in blocks]
            recs.append((start, L | (int(bool(r.is_reverse)) << 16) | (len(blocks) << 24), blocks, r))
        recs.sort(key=lambda rec: rec[0])          # stable: ties keep input (file) order
        for start, meta, blocks, r in recs:
            starts.append(start)
            metas.append(meta)
            if len(blocks) > 1:
                any_multi = True
                blks.extend(blocks)
            blk_off.append(len(blks))
            objs.append(r)
        chrom_read_off.append(len(starts))
    starts = np.asarray(starts, dtype=np.int64)
    batch_kwargs = {}
    if any_multi:
        batch_kwargs = dict(blk_off=np.asarray(blk_off, dtype=np.uint32),
                            blk=np.asarray(blks, dtype=np.int32).reshape(-1, 2))
    hb = AlignmentBatch(chroms, [chrom_lengths[c] for c in chroms], starts.astype(np.int32),
                        np.asarray(metas, dtype=np.uint32), chrom_read_off, mapped=mapped, **batch_kwargs)
    hb.check_sorted()
    if keep_objects:
        hb.objects = objs
    return hb


def batch_from_arrays(chroms, chrom_len, chrom_id, ref_start, aligned_len, is_reverse,
                      blocks=None, mapped=None):
    """Build a batch from flat per-read arrays (any order).  ``blocks``: optional
    ``(n_blocks[N], blk[B,2])`` listing, in input read order, every block of every read."""
    chrom_id = np.asarray(chrom_id, dtype=np.int64)
    ref_start = np.asarray(ref_start, dtype=np.int64)
    aligned_len = np.asarray(aligned_len, dtype=np.int64)
    if len(aligned_len) and aligned_len.max() > MAX_ALIGNED_LEN:
        raise ValueError("aligned length exceeds %d" % MAX_ALIGNED_LEN)
    order = np.lexsort((ref_start, chrom_id))
    nblk = np.ones(len(ref_start), dtype=np.int64)
    blk_off = blk = None
    if blocks is not None:
        nb_in, blk_in = blocks
        nb_in = np.asarray(nb_in, dtype=np.int64)
        blk_in = np.asarray(blk_in, dtype=np.int32).reshape(-1, 2)
        if len(nb_in) and nb_in.max() > MAX_BLOCKS:
            raise ValueError("more than %d aligned blocks in a read" % MAX_BLOCKS)
        nblk = nb_in
        in_off = np.zeros(len(nb_in) + 1, dtype=np.int64)
        np.cumsum(nb_in, out=in_off[1:])
        listed = np.where(nb_in[order] > 1, nb_in[order], 0)
        blk_off = np.zeros(len(order) + 1, dtype=np.int64)
        np.cumsum(listed, out=blk_off[1:])
        blk = np.zeros((int(blk_off[-1]), 2), dtype=np.int32)
        multi = np.nonzero(listed)[0]
        for dst_i in multi:                     # spliced reads only
            src = order[dst_i]
            blk[blk_off[dst_i]:blk_off[dst_i + 1]] = blk_in[in_off[src]:in_off[src + 1]]
        if len(multi) == 0:
            blk_off = blk = None
    meta = (aligned_len[order].astype(np.uint32)
            | (np.asarray(is_reverse, dtype=np.uint32)[order] << 16)
            | (nblk[order].astype(np.uint32) << 24))
    counts = np.bincount(chrom_id, minlength=len(chroms)) if len(chrom_id) else np.zeros(len(chroms), dtype=np.int64)
    chrom_read_off = np.zeros(len(chroms) + 1, dtype=np.int64)
    np.cumsum(counts, out=chrom_read_off[1:])
    return AlignmentBatch(chroms, chrom_len, ref_start[order].astype(np.int32), meta, chrom_read_off,
                          blk_off=blk_off, blk=blk, mapped=mapped)
