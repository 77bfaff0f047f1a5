"""Packed SoA alignment batches and the dense genome layout (host side).

What pysam hands the reference one ``AlignedSegment`` at a time (``reference_start``,
``positions``, ``is_reverse``; call sites ``plastid/genomics/map_factories.pyx:243,349,448,629``
and ``plastid/genomics/genome_array.py:800-815``) is packed here once into flat arrays the
kernels stream.  Schema: ``include/plastid_b200.h``.
"""
import os

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
        key = _lib.device_key(device)
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
        self.max_block_len = self.compute_max_block_len()
        self.mapped = n if mapped is None else int(mapped)   # what `bamfile.mapped` reports
        self.objects = None      # optional list of the original read objects, batch order
        self._dev = {}
        # the batch in the form that crosses PCIe (delta3 streams, plus block words for spliced batches), written
        # once by whoever produced the batch — the BAM decoder does (bam_io.batch_from_bam) — see pack()
        self.transfer = None
        self._transfer_pinned = None

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

    def compute_max_block_len(self):
        """Longest aligned block: the halo of the tile candidate window (single-block reads: L)."""
        if len(self.ref_start) == 0:
            return 1
        out = int((self.meta & 0xFFFF).max())      # >= every block of every read
        return max(out, 1)

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

    def pack(self, threads=0):
        """Attach the transfer format of this batch (:class:`Delta3Batch`, or :class:`Delta3SplicedBatch` when it has
        multi-block reads): 1-1.3 bytes per read instead of the SoA's 8, 4 bytes per aligned block instead of 12.
        Encoded by the library's multithreaded host encoder (``pb_pack_delta3``); done once per batch — the decoder
        calls it — after which :class:`~plastid_b200.genome_array.BAMGenomeArray` ships this form to the device
        and expands it there.  Returns ``self``."""
        if self.transfer is None:
            cls = Delta3Batch if self.blk is None else Delta3SplicedBatch
            self.transfer = cls.from_batch(self, threads=threads)
            if getattr(self.transfer, "length_hist", None) is None:
                # batch metadata like max_span: reads per aligned length, known to whoever produced the batch (the
                # Center rule derives its tables from it without a pass over the reads)
                self.transfer.length_hist = meta_length_hist(self.meta)
            self._transfer_pinned = None
        return self

    def transfer_pinned(self):
        """The transfer format in page-locked host memory (dict of tensors), built on first use."""
        if self.transfer is None:
            raise ValueError("batch has no transfer format: call pack() first")
        if self._transfer_pinned is None:
            self._transfer_pinned = self.transfer.pinned()
        return self._transfer_pinned

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
        key = _lib.device_key(device)
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


class Wire16Batch(object):
    """Compact 4-byte-per-read host format of an unspliced batch (``pb_unpack_wire16`` in
    ``include/plastid_b200.h``): what crosses PCIe instead of the 8-byte SoA."""
    SEG_BITS = 16

    def __init__(self, chroms, chrom_len, chrom_read_off, start_lo, meta16, seg_off, seg_base, max_span, mapped):
        self.chroms, self.chrom_len = list(chroms), np.asarray(chrom_len, dtype=np.int64)
        self.chrom_read_off = np.ascontiguousarray(chrom_read_off, dtype=np.int64)
        self.start_lo = np.ascontiguousarray(start_lo, dtype=np.uint16)
        self.meta16 = np.ascontiguousarray(meta16, dtype=np.uint16)
        self.seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        self.seg_base = np.ascontiguousarray(seg_base, dtype=np.int32)
        self.max_span, self.mapped = int(max_span), int(mapped)

    def __len__(self):
        return len(self.start_lo)

    @property
    def nbytes(self):
        return self.start_lo.nbytes + self.meta16.nbytes + self.seg_off.nbytes + self.seg_base.nbytes \
            + self.chrom_read_off.nbytes

    @classmethod
    def from_batch(cls, hb):
        if hb.blk is not None:
            raise ValueError("wire16 carries single-block reads only")
        L = hb.meta & 0xFFFF
        if len(L) and int(L.max()) >= (1 << 14):
            raise ValueError("wire16 needs aligned lengths < 16384")
        meta16 = (L | (((hb.meta >> 16) & 1) << 14) | (((hb.meta >> 17) & 1) << 15)).astype(np.uint16)
        n_seg_chrom = np.maximum((hb.chrom_len + 65535) >> 16, 1)
        seg_first = np.zeros(len(hb.chroms) + 1, dtype=np.int64)
        np.cumsum(n_seg_chrom, out=seg_first[1:])
        chrom_of_read = np.repeat(np.arange(len(hb.chroms)), np.diff(hb.chrom_read_off))
        seg_of_read = seg_first[chrom_of_read] + (hb.ref_start.astype(np.int64) >> 16)
        n_seg = int(seg_first[-1])
        seg_off = np.zeros(n_seg + 1, dtype=np.int64)
        np.cumsum(np.bincount(seg_of_read, minlength=n_seg), out=seg_off[1:])
        seg_base = np.concatenate([np.arange(n, dtype=np.int64) << 16 for n in n_seg_chrom]).astype(np.int32)
        return cls(hb.chroms, hb.chrom_len, hb.chrom_read_off, (hb.ref_start & 0xFFFF).astype(np.uint16), meta16,
                   seg_off, seg_base, hb.max_span, hb.mapped)

    def pinned(self):
        import torch
        def pin(a, view=None):
            return torch.from_numpy(a.view(view) if view is not None else a).pin_memory()
        return dict(start_lo=pin(self.start_lo, np.int16), meta16=pin(self.meta16, np.int16),
                    seg_off=pin(self.seg_off), seg_base=pin(self.seg_base), chrom_read_off=pin(self.chrom_read_off))


class Wire16Receiver(object):
    """Device-side landing buffers for wire16 transfers + the expanded :class:`DeviceBatch`."""

    def __init__(self, wire, device):
        import torch
        n, n_seg = len(wire), len(wire.seg_base)
        self.n_seg = n_seg
        self.start_lo = torch.empty(n, dtype=torch.int16, device=device)
        self.meta16 = torch.empty(n, dtype=torch.int16, device=device)
        self.seg_off = torch.empty(n_seg + 1, dtype=torch.int64, device=device)
        self.seg_base = torch.empty(n_seg, dtype=torch.int32, device=device)
        self.batch = DeviceBatch(n, len(wire.chroms), wire.max_span, torch.empty(n, dtype=torch.int32, device=device),
                                 torch.empty(n, dtype=torch.int32, device=device),
                                 torch.empty(len(wire.chroms) + 1, dtype=torch.int64, device=device))

    def receive(self, pinned):
        """Enqueue H2D copies of one wire16 batch and its expansion; returns the DeviceBatch."""
        from . import _lib
        self.start_lo.copy_(pinned["start_lo"], non_blocking=True)
        self.meta16.copy_(pinned["meta16"], non_blocking=True)
        self._receive_tables(pinned)
        self._unpack(0, self.batch.n_reads)
        return self.batch

    def _receive_tables(self, pinned):
        self.seg_off.copy_(pinned["seg_off"], non_blocking=True)
        self.seg_base.copy_(pinned["seg_base"], non_blocking=True)
        self.batch.chrom_read_off.copy_(pinned["chrom_read_off"], non_blocking=True)

    def _unpack(self, a, b):
        from . import _lib
        _lib.check(_lib.lib().pb_unpack_wire16(_lib.ptr(self.start_lo), _lib.ptr(self.meta16), _lib.ptr(self.seg_off),
                                               _lib.ptr(self.seg_base), self.n_seg, int(a), int(b),
                                               _lib.ptr(self.batch.ref_start), _lib.ptr(self.batch.meta),
                                               _lib.stream_ptr()))

    @staticmethod
    def plan_chunks(wire, layout, n_chunks):
        """Cut the sorted batch at 65536-position segment boundaries into ``n_chunks`` pieces of about
        equal read count: ``[(read_a, read_b, bin_a, bin_b), ...]`` covering all reads and all bins."""
        n_seg = len(wire.seg_base)
        n_seg_chrom = np.maximum((wire.chrom_len + 65535) >> 16, 1)
        seg_first = np.zeros(len(wire.chroms) + 1, dtype=np.int64)
        np.cumsum(n_seg_chrom, out=seg_first[1:])
        targets = (np.arange(1, n_chunks) * len(wire)) // max(n_chunks, 1)
        cuts = sorted(set(int(x) for x in np.searchsorted(wire.seg_off[:-1], targets, side="left")) - {0, n_seg})
        seg_cuts = [0] + cuts + [n_seg]
        out = []
        for s0, s1 in zip(seg_cuts[:-1], seg_cuts[1:]):
            def bin_of(s):
                if s >= n_seg:
                    return layout.total_bins
                c = int(np.searchsorted(seg_first, s, side="right")) - 1
                return int(layout.chrom_bin_off[c]) + ((s - int(seg_first[c])) << 16)
            out.append((int(wire.seg_off[s0]), int(wire.seg_off[s1]), bin_of(s0), bin_of(s1)))
        return out

    def receive_chunk(self, pinned, a, b, copy_stream):
        """H2D of reads [a,b) on ``copy_stream``; returns the event the compute stream must wait for."""
        import torch
        with torch.cuda.stream(copy_stream):
            self.start_lo[a:b].copy_(pinned["start_lo"][a:b], non_blocking=True)
            self.meta16[a:b].copy_(pinned["meta16"][a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev


class Delta8Batch(object):
    """2-byte-per-read host format of a sorted unspliced batch (``pb_unpack_delta8`` in
    ``include/plastid_b200.h``): one byte of start delta + one byte of meta-dictionary index per
    read, blocks of 128 reads with an absolute base, an exception list for everything that does not
    fit (delta >= 255, chromosome changes, rare meta words)."""
    BLOCK = 128

    def __init__(self, chroms, chrom_len, chrom_read_off, n_reads, dstart, code, blk_base, blk_exc_off,
                 exc_start, exc_meta, meta_dict, read_bin16k, max_span, mapped):
        self.chroms, self.chrom_len = list(chroms), np.asarray(chrom_len, dtype=np.int64)
        self.chrom_read_off = np.ascontiguousarray(chrom_read_off, dtype=np.int64)
        self.n_reads = int(n_reads)
        self.dstart = np.ascontiguousarray(dstart, dtype=np.uint8)
        self.code = np.ascontiguousarray(code, dtype=np.uint8)
        self.blk_base = np.ascontiguousarray(blk_base, dtype=np.int32)
        self.blk_exc_off = np.ascontiguousarray(blk_exc_off, dtype=np.uint32)
        self.exc_start = np.ascontiguousarray(exc_start, dtype=np.int32)
        self.exc_meta = np.ascontiguousarray(exc_meta, dtype=np.uint32)
        self.meta_dict = np.ascontiguousarray(meta_dict, dtype=np.uint32)
        # chromosome index and start of the first read of every block (chunk planning only; stays on the host)
        self.blk_chrom, self.blk_first_start = read_bin16k
        self.max_span, self.mapped = int(max_span), int(mapped)

    def __len__(self):
        return self.n_reads

    @property
    def nbytes(self):
        """Bytes that cross PCIe per batch."""
        return (self.dstart.nbytes + self.code.nbytes + self.blk_base.nbytes + self.blk_exc_off.nbytes
                + self.exc_start.nbytes + self.exc_meta.nbytes + self.meta_dict.nbytes + self.chrom_read_off.nbytes)

    @classmethod
    def from_batch(cls, hb):
        if hb.blk is not None:
            raise ValueError("delta8 carries single-block reads only")
        n, K = len(hb), cls.BLOCK
        n_blk = (n + K - 1) // K
        start = hb.ref_start.astype(np.int64)
        meta = hb.meta.astype(np.uint32)
        chrom_of_read = np.repeat(np.arange(len(hb.chroms), dtype=np.int32), np.diff(hb.chrom_read_off))
        delta = np.zeros(n, dtype=np.int64)
        if n > 1:
            delta[1:] = start[1:] - start[:-1]
        first_of_chrom = np.zeros(n, dtype=bool)
        if n:
            first_of_chrom[0] = True
            first_of_chrom[1:] = chrom_of_read[1:] != chrom_of_read[:-1]
        first_of_blk = np.zeros(n, dtype=bool)
        first_of_blk[::K] = True
        delta[first_of_blk] = 0
        # dictionary: the 255 most frequent meta words (index 255 is never used as a code)
        words, counts = np.unique(meta, return_counts=True)
        order = np.argsort(-counts, kind="stable")[:255]
        dict_words = np.sort(words[order])
        pos = np.searchsorted(dict_words, meta)
        pos_c = np.minimum(pos, max(len(dict_words) - 1, 0))
        in_dict = (dict_words[pos_c] == meta) if len(dict_words) else np.zeros(n, dtype=bool)
        exc = (delta >= 255) | (delta < 0) | (first_of_chrom & ~first_of_blk) | ~in_dict
        dstart = np.zeros(n_blk * K, dtype=np.uint8)
        code = np.zeros(n_blk * K, dtype=np.uint8)
        dstart[:n] = np.where(exc, 255, np.clip(delta, 0, 254)).astype(np.uint8)
        code[:n] = np.where(in_dict, pos_c, 0).astype(np.uint8)
        meta_dict = np.zeros(256, dtype=np.uint32)
        meta_dict[:len(dict_words)] = dict_words
        exc_per_blk = np.add.reduceat(exc.astype(np.int64), np.arange(0, n, K)) if n else np.zeros(0, dtype=np.int64)
        blk_exc_off = np.zeros(n_blk + 1, dtype=np.int64)
        np.cumsum(exc_per_blk, out=blk_exc_off[1:])
        if blk_exc_off[-1] >= (1 << 32):
            raise ValueError("delta8: too many exceptions")
        return cls(hb.chroms, hb.chrom_len, hb.chrom_read_off, n, dstart, code, hb.ref_start[::K],
                   blk_exc_off, hb.ref_start[exc], meta[exc], meta_dict,
                   (chrom_of_read[::K].copy(), hb.ref_start[::K].astype(np.int64)), hb.max_span, hb.mapped)

    def pinned(self):
        import torch

        def pin(a, view=None):
            a = a.view(view) if view is not None else a
            return torch.from_numpy(a if len(a) else np.zeros(1, dtype=a.dtype)).pin_memory()
        return dict(dstart=pin(self.dstart), code=pin(self.code), blk_base=pin(self.blk_base),
                    blk_exc_off=pin(self.blk_exc_off, np.int32), exc_start=pin(self.exc_start),
                    exc_meta=pin(self.exc_meta, np.int32), meta_dict=pin(self.meta_dict, np.int32),
                    chrom_read_off=pin(self.chrom_read_off))


class Delta8Receiver(object):
    """Device-side landing buffers for delta8 transfers + the expanded :class:`DeviceBatch`; same
    interface as :class:`Wire16Receiver` (``plan_chunks`` / ``receive_chunk`` / ``_unpack``)."""

    def __init__(self, wire, device):
        import torch
        n, K = len(wire), Delta8Batch.BLOCK
        self.wire = wire
        n_blk = len(wire.blk_base)
        self.dstart = torch.empty(n_blk * K, dtype=torch.uint8, device=device)
        self.code = torch.empty(n_blk * K, dtype=torch.uint8, device=device)
        self.blk_base = torch.empty(max(n_blk, 1), dtype=torch.int32, device=device)
        self.blk_exc_off = torch.empty(n_blk + 1, dtype=torch.int32, device=device)
        self.exc_start = torch.empty(max(len(wire.exc_start), 1), dtype=torch.int32, device=device)
        self.exc_meta = torch.empty(max(len(wire.exc_meta), 1), dtype=torch.int32, device=device)
        self.meta_dict = torch.empty(256, dtype=torch.int32, device=device)
        self.batch = DeviceBatch(n, len(wire.chroms), wire.max_span, torch.empty(n, dtype=torch.int32, device=device),
                                 torch.empty(n, dtype=torch.int32, device=device),
                                 torch.empty(len(wire.chroms) + 1, dtype=torch.int64, device=device))

    def receive(self, pinned):
        """Enqueue the H2D copies of one whole batch and its expansion; returns the DeviceBatch."""
        self._receive_tables(pinned)
        n = self.batch.n_reads
        self._copy_range(pinned, 0, n)
        self._unpack(0, n)
        return self.batch

    def _receive_tables(self, pinned):
        self.meta_dict.copy_(pinned["meta_dict"], non_blocking=True)
        self.batch.chrom_read_off.copy_(pinned["chrom_read_off"], non_blocking=True)

    def _copy_range(self, pinned, a, b):
        """Everything reads [a,b) need (a a multiple of 128): both byte streams, block tables, exceptions."""
        K = Delta8Batch.BLOCK
        if b <= a:
            return
        b0, b1 = a // K, (b + K - 1) // K
        self.dstart[b0 * K:b1 * K].copy_(pinned["dstart"][b0 * K:b1 * K], non_blocking=True)
        self.code[b0 * K:b1 * K].copy_(pinned["code"][b0 * K:b1 * K], non_blocking=True)
        self.blk_base[b0:b1].copy_(pinned["blk_base"][b0:b1], non_blocking=True)
        self.blk_exc_off[b0:b1 + 1].copy_(pinned["blk_exc_off"][b0:b1 + 1], non_blocking=True)
        e0, e1 = int(self.wire.blk_exc_off[b0]), int(self.wire.blk_exc_off[b1])
        if e1 > e0:
            self.exc_start[e0:e1].copy_(pinned["exc_start"][e0:e1], non_blocking=True)
            self.exc_meta[e0:e1].copy_(pinned["exc_meta"][e0:e1], non_blocking=True)

    def _unpack(self, a, b):
        from . import _lib
        _lib.check(_lib.lib().pb_unpack_delta8(_lib.ptr(self.dstart), _lib.ptr(self.code), _lib.ptr(self.blk_base),
                                               _lib.ptr(self.blk_exc_off), _lib.ptr(self.exc_start),
                                               _lib.ptr(self.exc_meta), _lib.ptr(self.meta_dict),
                                               self.batch.n_reads, int(a), int(b),
                                               _lib.ptr(self.batch.ref_start), _lib.ptr(self.batch.meta),
                                               _lib.stream_ptr()))

    @staticmethod
    def plan_chunks(wire, layout, n_chunks, weights=None):
        """Cut the sorted batch at block (128-read) boundaries into ``n_chunks`` pieces of about equal
        read count — or, with ``weights`` (one positive number per chunk), of read counts in those
        proportions: ``[(read_a, read_b, bin_a, bin_b), ...]``.  bin_b is the global bin of read_b's
        start rounded down to the layout granularity: every read at or beyond read_b starts at or
        beyond it, so bins below bin_b are final once reads [0, read_b) have landed.  Ranges that
        would be empty are merged into the next chunk.  (When the upload is the slower side, the step ends
        one chunk-mapping after the last byte lands: a small LAST chunk shortens that tail; when mapping
        is the slower side a small FIRST chunk starts it earlier.)"""
        from . import _lib
        K, n = Delta8Batch.BLOCK, len(wire)
        n_blk = len(wire.blk_base)
        if weights is not None:
            w = np.asarray(weights, dtype=np.float64)
            if len(w) < 1 or (w <= 0).any():
                raise ValueError("plan_chunks: weights must be positive")
            frac = np.cumsum(w)[:-1] / w.sum()
            cuts = sorted(set(int(f * n_blk) for f in frac) - {0, n_blk})
        else:
            cuts = sorted(set(int((j * n_blk) // n_chunks) for j in range(1, n_chunks)) - {0, n_blk})
        out, read_a, bin_a = [], 0, 0
        for cb in cuts:
            c = int(wire.blk_chrom[cb])
            g = int(layout.chrom_bin_off[c]) + int(wire.blk_first_start[cb])
            bin_b = (g // _lib.PB_LAYOUT_ALIGN) * _lib.PB_LAYOUT_ALIGN
            if bin_b <= bin_a:
                continue
            out.append((read_a, cb * K, bin_a, bin_b))
            read_a, bin_a = cb * K, bin_b
        out.append((read_a, n, bin_a, int(layout.total_bins)))
        return out

    def receive_chunk(self, pinned, a, b, copy_stream):
        """H2D of reads [a,b) on ``copy_stream``; returns the event the compute stream must wait for."""
        import torch
        with torch.cuda.stream(copy_stream):
            self._copy_range(pinned, a, b)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev


class Delta3Batch(object):
    """1-byte-per-read host format of a sorted unspliced batch (``pb_unpack_delta3`` in
    ``include/plastid_b200.h``): 3 bits of start delta + 5 bits of meta-dictionary index per read, a
    byte stream for deltas of 7..261, an exception list for the rest; blocks of 128 reads."""
    BLOCK = 128

    def __init__(self, chroms, chrom_len, chrom_read_off, n_reads, packed, wide, blk_base, blk_wide_off, blk_exc_off,
                 exc_start, exc_meta, meta_dict, blk_first, max_span, mapped):
        self.chroms, self.chrom_len = list(chroms), np.asarray(chrom_len, dtype=np.int64)
        self.chrom_read_off = np.ascontiguousarray(chrom_read_off, dtype=np.int64)
        self.n_reads = int(n_reads)
        self.packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self.wide = np.ascontiguousarray(wide, dtype=np.uint8)
        self.blk_base = np.ascontiguousarray(blk_base, dtype=np.int32)
        self.blk_wide_off = np.ascontiguousarray(blk_wide_off, dtype=np.uint32)
        self.blk_exc_off = np.ascontiguousarray(blk_exc_off, dtype=np.uint32)
        self.exc_start = np.ascontiguousarray(exc_start, dtype=np.int32)
        self.exc_meta = np.ascontiguousarray(exc_meta, dtype=np.uint32)
        self.meta_dict = np.ascontiguousarray(meta_dict, dtype=np.uint32)
        self.blk_chrom, self.blk_first_start = blk_first      # chunk planning only; stays on the host
        self.max_span, self.mapped = int(max_span), int(mapped)
        self.length_hist = None                               # batch metadata, set by AlignmentBatch.pack

    def __len__(self):
        return self.n_reads

    @property
    def nbytes(self):
        """Bytes that cross PCIe per batch."""
        return (self.packed.nbytes + self.wide.nbytes + self.blk_base.nbytes + self.blk_wide_off.nbytes
                + self.blk_exc_off.nbytes + self.exc_start.nbytes + self.exc_meta.nbytes + self.meta_dict.nbytes
                + self.chrom_read_off.nbytes)

    @classmethod
    def from_batch(cls, hb, native=True, threads=0):
        """``native``: encode with the library's multithreaded host encoder (``pb_pack_delta3``, no CUDA
        involved); ``False`` = the numpy statement of the format below (tests compare the two)."""
        if hb.blk is not None:
            raise ValueError("delta3 carries single-block reads only")
        n, K = len(hb), cls.BLOCK
        n_blk = (n + K - 1) // K
        if native:
            import ctypes as C
            from . import _lib
            start32 = np.ascontiguousarray(hb.ref_start, dtype=np.int32)
            meta32 = np.ascontiguousarray(hb.meta, dtype=np.uint32)
            off = np.ascontiguousarray(hb.chrom_read_off, dtype=np.int64)
            packed = np.empty(n_blk * K, dtype=np.uint8)
            wide = np.empty(max(n, 1), dtype=np.uint8)
            blk_base = np.empty(max(n_blk, 1), dtype=np.int32)
            blk_wide_off = np.empty(n_blk + 1, dtype=np.uint32)
            blk_exc_off = np.empty(n_blk + 1, dtype=np.uint32)
            exc_start = np.empty(max(n, 1), dtype=np.int32)
            exc_meta = np.empty(max(n, 1), dtype=np.uint32)
            meta_dict = np.empty(32, dtype=np.uint32)
            n_wide, n_exc = C.c_int64(0), C.c_int64(0)

            def p(a):
                return C.c_void_p(a.ctypes.data)
            _lib.check(_lib.lib().pb_pack_delta3(p(start32), p(meta32), p(off), len(hb.chroms), n, _lib.host_threads(threads), p(packed),
                                                 p(wide), p(blk_base), p(blk_wide_off), p(blk_exc_off), p(exc_start),
                                                 p(exc_meta), p(meta_dict), C.byref(n_wide), C.byref(n_exc)))
            chrom_of_blk = (np.searchsorted(off, np.arange(0, n, K), side="right") - 1).astype(np.int32)
            return cls(hb.chroms, hb.chrom_len, hb.chrom_read_off, n, packed, wide[:n_wide.value].copy(), blk_base[:n_blk],
                       blk_wide_off, blk_exc_off, exc_start[:n_exc.value].copy(), exc_meta[:n_exc.value].copy(), meta_dict,
                       (chrom_of_blk, start32[::K].astype(np.int64)), hb.max_span, hb.mapped)
        start = hb.ref_start.astype(np.int64)
        meta = hb.meta.astype(np.uint32)
        chrom_of_read = np.repeat(np.arange(len(hb.chroms), dtype=np.int32), np.diff(hb.chrom_read_off))
        delta = np.zeros(n, dtype=np.int64)
        if n > 1:
            delta[1:] = start[1:] - start[:-1]
        first_of_chrom = np.zeros(n, dtype=bool)
        if n:
            first_of_chrom[0] = True
            first_of_chrom[1:] = chrom_of_read[1:] != chrom_of_read[:-1]
        first_of_blk = np.zeros(n, dtype=bool)
        first_of_blk[::K] = True
        delta[first_of_blk] = 0
        words, counts = np.unique(meta, return_counts=True)
        dict_words = np.sort(words[np.argsort(-counts, kind="stable")[:31]])
        pos = np.minimum(np.searchsorted(dict_words, meta), max(len(dict_words) - 1, 0))
        in_dict = (dict_words[pos] == meta) if len(dict_words) else np.zeros(n, dtype=bool)
        exc = (delta > 7 + 254) | (delta < 0) | (first_of_chrom & ~first_of_blk) | ~in_dict
        widef = exc | (delta >= 7)
        packed = np.zeros(n_blk * K, dtype=np.uint8)
        packed[:n] = (np.where(widef, 7, delta) | (np.where(exc, 31, pos) << 3)).astype(np.uint8)
        wide = np.where(exc, 255, delta - 7)[widef].astype(np.uint8)
        meta_dict = np.zeros(32, dtype=np.uint32)
        meta_dict[:len(dict_words)] = dict_words

        def block_offsets(flags):
            per = np.add.reduceat(flags.astype(np.int64), np.arange(0, n, K)) if n else np.zeros(0, dtype=np.int64)
            off = np.zeros(n_blk + 1, dtype=np.int64)
            np.cumsum(per, out=off[1:])
            if off[-1] >= (1 << 32):
                raise ValueError("delta3: too many escapes")
            return off
        return cls(hb.chroms, hb.chrom_len, hb.chrom_read_off, n, packed, wide, hb.ref_start[::K], block_offsets(widef),
                   block_offsets(exc), hb.ref_start[exc], meta[exc], meta_dict,
                   (chrom_of_read[::K].copy(), hb.ref_start[::K].astype(np.int64)), hb.max_span, hb.mapped)

    def pinned(self):
        import torch

        def pin(a, view=None):
            a = a.view(view) if view is not None else a
            return torch.from_numpy(a if len(a) else np.zeros(1, dtype=a.dtype)).pin_memory()
        return dict(packed=pin(self.packed), wide=pin(self.wide), blk_base=pin(self.blk_base),
                    blk_wide_off=pin(self.blk_wide_off, np.int32), blk_exc_off=pin(self.blk_exc_off, np.int32),
                    exc_start=pin(self.exc_start), exc_meta=pin(self.exc_meta, np.int32),
                    meta_dict=pin(self.meta_dict, np.int32), chrom_read_off=pin(self.chrom_read_off))


class Delta3Receiver(Delta8Receiver):
    """Device-side landing buffers for delta3 transfers (same interface as :class:`Delta8Receiver`)."""

    def __init__(self, wire, device):
        import torch
        n, K = len(wire), Delta3Batch.BLOCK
        self.wire = wire
        n_blk = len(wire.blk_base)
        self.packed = torch.empty(n_blk * K, dtype=torch.uint8, device=device)
        self.wide = torch.empty(max(len(wire.wide), 1), dtype=torch.uint8, device=device)
        self.blk_base = torch.empty(max(n_blk, 1), dtype=torch.int32, device=device)
        self.blk_wide_off = torch.empty(n_blk + 1, dtype=torch.int32, device=device)
        self.blk_exc_off = torch.empty(n_blk + 1, dtype=torch.int32, device=device)
        self.exc_start = torch.empty(max(len(wire.exc_start), 1), dtype=torch.int32, device=device)
        self.exc_meta = torch.empty(max(len(wire.exc_meta), 1), dtype=torch.int32, device=device)
        self.meta_dict = torch.empty(32, dtype=torch.int32, device=device)
        self.batch = DeviceBatch(n, len(wire.chroms), wire.max_span, torch.empty(n, dtype=torch.int32, device=device),
                                 torch.empty(n, dtype=torch.int32, device=device),
                                 torch.empty(len(wire.chroms) + 1, dtype=torch.int64, device=device),
                                 length_hist=getattr(wire, "length_hist", None))

    def _copy_range(self, pinned, a, b):
        K = Delta3Batch.BLOCK
        if b <= a:
            return
        b0, b1 = a // K, (b + K - 1) // K
        self.packed[b0 * K:b1 * K].copy_(pinned["packed"][b0 * K:b1 * K], non_blocking=True)
        self.blk_base[b0:b1].copy_(pinned["blk_base"][b0:b1], non_blocking=True)
        self.blk_wide_off[b0:b1 + 1].copy_(pinned["blk_wide_off"][b0:b1 + 1], non_blocking=True)
        self.blk_exc_off[b0:b1 + 1].copy_(pinned["blk_exc_off"][b0:b1 + 1], non_blocking=True)
        w0, w1 = int(self.wire.blk_wide_off[b0]), int(self.wire.blk_wide_off[b1])
        if w1 > w0:
            self.wide[w0:w1].copy_(pinned["wide"][w0:w1], non_blocking=True)
        e0, e1 = int(self.wire.blk_exc_off[b0]), int(self.wire.blk_exc_off[b1])
        if e1 > e0:
            self.exc_start[e0:e1].copy_(pinned["exc_start"][e0:e1], non_blocking=True)
            self.exc_meta[e0:e1].copy_(pinned["exc_meta"][e0:e1], non_blocking=True)

    def _unpack(self, a, b):
        from . import _lib
        _lib.check(_lib.lib().pb_unpack_delta3(_lib.ptr(self.packed), _lib.ptr(self.wide), _lib.ptr(self.blk_base),
                                               _lib.ptr(self.blk_wide_off), _lib.ptr(self.blk_exc_off),
                                               _lib.ptr(self.exc_start), _lib.ptr(self.exc_meta), _lib.ptr(self.meta_dict),
                                               self.batch.n_reads, int(a), int(b),
                                               _lib.ptr(self.batch.ref_start), _lib.ptr(self.batch.meta),
                                               _lib.stream_ptr()))


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


class Delta3SplicedBatch(object):
    """PCIe transfer format of a sorted batch WITH multi-block (spliced) reads: the delta3 streams for
    ``ref_start`` / ``meta`` of every read (1-1.3 B/read) plus one 4-byte "block word" per aligned block of
    the multi-block reads (``pb_unpack_blocks`` in ``include/plastid_b200.h``): 12 bits of block length, 20
    bits of gap to the read's previous block, an exception list for blocks that do not fit.  ``blk_off``
    (4 B/read in the SoA) is not shipped at all: the device rebuilds it from the block counts in ``meta``.
    C3 (100 M reads, 33 % spliced): 17.5 B/read as SoA -> about 4 B/read."""

    def __init__(self, base, bwords, bexc_row, bexc, max_block_len):
        self.base = base                                   # Delta3Batch of (ref_start, meta)
        self.length_hist = None                            # set by from_batch: reads per aligned length
        self.row_of_read = None                            # host only (chunk planning): block rows before each read
        self.bwords = np.ascontiguousarray(bwords, dtype=np.uint32)
        self.bexc_row = np.ascontiguousarray(bexc_row, dtype=np.uint32)
        self.bexc = np.ascontiguousarray(bexc, dtype=np.int32).reshape(-1, 2)
        self.max_block_len = int(max_block_len)

    def __len__(self):
        return len(self.base)

    @property
    def nbytes(self):
        return self.base.nbytes + self.bwords.nbytes + self.bexc_row.nbytes + self.bexc.nbytes

    @classmethod
    def from_batch(cls, hb, native=True, threads=0):
        plain = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start, hb.meta, hb.chrom_read_off, None, None,
                               hb.max_span, hb.mapped)
        base = Delta3Batch.from_batch(plain, native=native, threads=threads)
        hist = meta_length_hist(hb.meta)
        if hb.blk is None or len(hb.blk) == 0:
            out = cls(base, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros((0, 2), np.int32), hb.max_block_len)
            out.length_hist = hist
            out.row_of_read = np.zeros(len(hb) + 1, dtype=np.int64)
            return out
        rel, ln = hb.blk[:, 0].astype(np.int64), hb.blk[:, 1].astype(np.int64)
        n_rows = len(rel)
        if n_rows >= (1 << 32):
            raise ValueError("block words: more than 2^32 block rows")
        first = np.zeros(n_rows, dtype=bool)
        counts = np.diff(hb.blk_off.astype(np.int64))
        first[hb.blk_off.astype(np.int64)[:-1][counts > 0]] = True
        prev_end = np.zeros(n_rows, dtype=np.int64)
        prev_end[1:] = rel[:-1] + ln[:-1]
        prev_end[first] = 0
        gap = rel - prev_end
        fits = (ln >= 1) & (ln < 4096) & (gap >= 0) & (gap < (1 << 20) - 1)
        words = np.where(fits, ln | (gap << 12), 0xFFFFFFFF).astype(np.uint32)
        rows = np.flatnonzero(~fits)
        out = cls(base, words, rows.astype(np.uint32), np.stack([gap[rows], ln[rows]], axis=1).astype(np.int32),
                  hb.max_block_len)
        out.length_hist = hist
        out.row_of_read = hb.blk_off.astype(np.int64)
        return out

    def pinned(self):
        import torch

        def pin(a, view=None):
            a = a.view(view) if view is not None else a
            return torch.from_numpy(a if a.size else np.zeros(2, dtype=a.dtype)).pin_memory()
        out = dict(self.base.pinned())
        out.update(bwords=pin(self.bwords, np.int32), bexc_row=pin(self.bexc_row, np.int32), bexc=pin(self.bexc.reshape(-1)))
        return out


class Delta3SplicedReceiver(object):
    """Device-side landing buffers of a :class:`Delta3SplicedBatch` and the expanded :class:`DeviceBatch`
    (``ref_start``, ``meta``, ``blk_off``, ``blk``)."""

    def __init__(self, wire, device):
        import torch
        self.wire = wire
        self.inner = Delta3Receiver(wire.base, device)
        n, rows, n_exc = len(wire), len(wire.bwords), len(wire.bexc_row)
        self.bwords = torch.empty(max(rows, 1), dtype=torch.int32, device=device)
        self.bexc_row = torch.empty(max(n_exc, 1), dtype=torch.int32, device=device)
        self.bexc = torch.empty(max(2 * n_exc, 2), dtype=torch.int32, device=device)
        self.ws_bytes = int(_lib.lib().pb_unpack_blocks_workspace_bytes(n))
        self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=device)
        ib = self.inner.batch
        self.batch = DeviceBatch(n, ib.n_chrom, ib.max_span, ib.ref_start, ib.meta, ib.chrom_read_off,
                                 torch.empty(n + 1, dtype=torch.int32, device=device),
                                 torch.empty((max(rows, 1), 2), dtype=torch.int32, device=device)[:rows],
                                 wire.max_block_len, wire.length_hist)

    def receive(self, pinned):
        """Enqueue the H2D copies of one whole batch and its expansion on the current stream."""
        n = self.batch.n_reads
        self.receive_tables(pinned)
        self._copy_chunk(pinned, 0, n)
        self.unpack_chunk(0, n)
        return self.batch

    # ---- chunked interface (map_center_streamed): same roles as Delta8Receiver's -----------------------------
    @staticmethod
    def plan_chunks(wire, layout, n_chunks, weights=None):
        """``[(read_a, read_b, bin_a, bin_b), ...]`` like :meth:`Delta8Receiver.plan_chunks`: bins below bin_b
        are final once reads [0, read_b) have landed (aligned positions never lie before a read's start)."""
        return Delta3Receiver.plan_chunks(wire.base, layout, n_chunks, weights)

    def receive_tables(self, pinned):
        """Dictionary, chromosome offsets and the (rare) block exceptions: needed by every chunk."""
        n_exc = len(self.wire.bexc_row)
        self.inner._receive_tables(pinned)
        if n_exc:
            self.bexc_row[:n_exc].copy_(pinned["bexc_row"][:n_exc], non_blocking=True)
            self.bexc[:2 * n_exc].copy_(pinned["bexc"][:2 * n_exc], non_blocking=True)

    def _copy_chunk(self, pinned, a, b):
        self.inner._copy_range(pinned, a, b)
        r0, r1 = int(self.wire.row_of_read[a]), int(self.wire.row_of_read[b])
        if r1 > r0:
            self.bwords[r0:r1].copy_(pinned["bwords"][r0:r1], non_blocking=True)

    def receive_chunk(self, pinned, a, b, copy_stream):
        """H2D of reads [a,b) and their block words on ``copy_stream``; returns the event to wait for."""
        import torch
        with torch.cuda.stream(copy_stream):
            self._copy_chunk(pinned, a, b)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def unpack_chunk(self, a, b):
        """Expand reads [a,b) (a a multiple of 128) on the current stream: start / meta, then blk_off / blk."""
        rows, n_exc = len(self.wire.bwords), len(self.wire.bexc_row)
        self.inner._unpack(a, b)
        _lib.check(_lib.lib().pb_unpack_blocks_range(_lib.ptr(self.batch.meta), self.batch.n_reads, int(a), int(b),
                                                     int(self.wire.row_of_read[a]), _lib.ptr(self.bwords), rows,
                                                     _lib.ptr(self.bexc_row), _lib.ptr(self.bexc), n_exc,
                                                     _lib.ptr(self.batch.blk_off), _lib.ptr(self.batch.blk) if rows else None,
                                                     _lib.ptr(self.ws), self.ws_bytes, _lib.stream_ptr()))

    def first_read_reaching(self, bin_a, layout):
        """Index of a read at or before the first read that can put a base at or beyond global bin ``bin_a``:
        reads are sorted, a read spans at most ``max_span`` positions, and the wire keeps the first start of
        every 128-read block — the answer is the start of the block before the first block whose first read
        starts within ``max_span`` of ``bin_a``."""
        base = self.wire.base
        cached = getattr(self, "_block_first_bin", None)
        if cached is None or cached[0] is not layout:          # 0.8 M entries for 100 M reads: build once
            cached = (layout, layout.chrom_bin_off[base.blk_chrom] + base.blk_first_start)      # ascending
            self._block_first_bin = cached
        j = int(np.searchsorted(cached[1], int(bin_a) - int(base.max_span), side="left"))
        return max(j - 1, 0) * Delta3Batch.BLOCK


def meta_length_hist(meta):
    """Reads per aligned length (int64[65536]) of a host ``meta`` array; reads with the drop bit are left out."""
    import ctypes as C
    m = np.asarray(meta).view(np.uint32) if np.asarray(meta).dtype != np.uint32 else np.asarray(meta)
    m = np.ascontiguousarray(m)
    hist = np.zeros(65536, dtype=np.int64)
    _lib.check(_lib.lib().pb_meta_length_hist(C.c_void_p(m.ctypes.data), len(m), _lib.host_threads(0), C.c_void_p(hist.ctypes.data)))
    return hist


# Upload of a large PAGEABLE numpy array (a plain SoA batch built from arrays, nothing precomputed): the driver stages
# such copies through its own bounce buffers at ~11 GB/s (C2: 1.6 GB in 145 ms).  Here a few host threads copy 32 MB
# chunks into a ring of pinned buffers (numpy releases the GIL for the memcpy) while the copy engine drains the buffers
# that are ready, so host memcpy and PCIe overlap.  Batches that carry their transfer format never come this way.
_STAGED_MIN_BYTES = 64 << 20
_STAGED_CHUNK = 32 << 20
_STAGED_RING = 6
_staging = {}


def _staged_upload(src, device):
    import collections
    import torch
    from concurrent.futures import ThreadPoolExecutor
    key = _lib.device_key(device)
    st = _staging.get(key)
    if st is None:
        ring = int(os.environ.get("PB_STAGE_RING", _STAGED_RING))
        bufs = [torch.empty(_STAGED_CHUNK, dtype=torch.uint8).pin_memory() for _ in range(ring)]
        n_threads = int(os.environ.get("PB_STAGE_THREADS", "4"))          # (env: A/B aid)
        st = _staging[key] = dict(bufs=bufs, views=[b.numpy() for b in bufs], pool=ThreadPoolExecutor(n_threads),
                                  stream=torch.cuda.Stream(device=device))
    flat = src.reshape(-1).view(np.uint8)
    n = flat.size
    dst = torch.empty(n, dtype=torch.uint8, device=device)
    chunks = [(lo, min(lo + _STAGED_CHUNK, n)) for lo in range(0, n, _STAGED_CHUNK)]
    n_ring = len(st["bufs"])
    events = [None] * n_ring
    pending = collections.deque()
    state = {"next": 0}

    def fill(k, lo, hi):
        np.copyto(st["views"][k][:hi - lo], flat[lo:hi])

    def submit():
        while state["next"] < len(chunks) and len(pending) < n_ring:
            i = state["next"]
            k = i % n_ring
            if events[k] is not None:                # the copy engine has drained this buffer
                events[k].synchronize()
                events[k] = None
            pending.append((i, k, st["pool"].submit(fill, k, *chunks[i])))
            state["next"] += 1

    st["stream"].wait_stream(torch.cuda.current_stream())
    submit()
    while pending:
        i, k, fut = pending.popleft()
        fut.result()
        lo, hi = chunks[i]
        with torch.cuda.stream(st["stream"]):
            dst[lo:hi].copy_(st["bufs"][k][:hi - lo], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st["stream"])
        events[k] = ev
        submit()
    torch.cuda.current_stream().wait_stream(st["stream"])
    for ev in events:                                # the ring is reused by the next upload
        if ev is not None:
            ev.synchronize()
    out = dst.view(torch.from_numpy(src.reshape(-1)[:0]).dtype)
    return out.view(src.shape) if src.ndim > 1 else out


class DeviceBatch(object):
    """Device-resident mirror of an :class:`AlignmentBatch` (torch tensors used as buffers only)."""

    def __init__(self, n_reads, n_chrom, max_span, ref_start, meta, chrom_read_off, blk_off=None, blk=None,
                 max_block_len=None, length_hist=None):
        self.n_reads, self.n_chrom, self.max_span = int(n_reads), int(n_chrom), int(max_span)
        self.ref_start, self.meta, self.chrom_read_off = ref_start, meta, chrom_read_off
        self.blk_off, self.blk = blk_off, blk
        self.n_blk = 0 if blk is None else int(blk.shape[0])
        # without better knowledge the reference span bounds every block
        self.max_block_len = int(max_span if max_block_len is None else max_block_len)
        # batch-level metadata like max_span: reads per aligned length (int64[65536], drop bit honoured), known
        # to whoever produced the batch (decoder, transfer-format receiver); None = measure on the device.
        # The Center rule derives its table of map lengths from it.
        self.length_hist = None if length_hist is None else np.ascontiguousarray(length_hist, dtype=np.int64)

    @classmethod
    def from_host(cls, hb, device, non_blocking=False):
        import torch

        def up(a):
            if a is None:
                return None
            src = a.view(np.int32) if a.dtype == np.uint32 else a
            if src.nbytes >= _STAGED_MIN_BYTES and str(device).startswith("cuda") and src.flags.c_contiguous:
                return _staged_upload(src, device)
            return torch.from_numpy(src).to(device, non_blocking=non_blocking)
        return cls(len(hb), len(hb.chroms), hb.max_span, up(hb.ref_start), up(hb.meta),
                   up(hb.chrom_read_off), up(hb.blk_off), up(hb.blk), hb.max_block_len,
                   None if hb.transfer is None else getattr(hb.transfer, "length_hist", None))

    @property
    def device(self):
        return self.ref_start.device

    def c_struct(self):
        return _lib.PbBatch(self.n_reads, self.ref_start.data_ptr(), self.meta.data_ptr(),
                            None if self.blk_off is None else self.blk_off.data_ptr(),
                            None if self.blk is None else self.blk.data_ptr(),
                            self.chrom_read_off.data_ptr(), self.n_chrom, self.max_span, self.n_blk,
                            self.max_block_len, 0)


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
        if blk_off[-1] == 0:
            blk_off = blk = None
        else:                                   # rows of the multi-block reads only, one gather for all of them
            take = np.repeat(in_off[:-1][order] - blk_off[:-1], listed) + np.arange(int(blk_off[-1]), dtype=np.int64)
            blk = np.ascontiguousarray(blk_in[take], dtype=np.int32).reshape(-1, 2)
    meta = (aligned_len[order].astype(np.uint32)
            | (np.asarray(is_reverse, dtype=np.uint32)[order] << 16)
            | (nblk[order].astype(np.uint32) << 24))
    counts = np.bincount(chrom_id, minlength=len(chroms)) if len(chrom_id) else np.zeros(len(chroms), dtype=np.int64)
    chrom_read_off = np.zeros(len(chroms) + 1, dtype=np.int64)
    np.cumsum(counts, out=chrom_read_off[1:])
    return AlignmentBatch(chroms, chrom_len, ref_start[order].astype(np.int32), meta, chrom_read_off,
                          blk_off=blk_off, blk=blk, mapped=mapped)
