"""BAM <-> SoA batch on the host.

The reference leaves BAM decoding to pysam (``plastid/genomics/genome_array.py:660, 800-809``),
a third-party dependency.  Here a sorted BAM is decoded ONCE, whole, by the library's own
multithreaded BGZF/BAM reader (``pb_bam_*`` in ``include/plastid_b200.h``, zlib underneath) into an
:class:`~plastid_b200.batch.AlignmentBatch` — no per-region ``fetch``, no index needed.  Like the
reference, no flag other than is_reverse is interpreted: secondary, duplicate and QC-fail
alignments are counted; records without a reference or flagged unmapped have no positions and are
skipped.  Open ``pysam.AlignmentFile`` handles are accepted too when pysam is installed.

``write_bam`` is a small pure-Python BAM writer (tests, synthetic data); it is not on the path.
"""
import ctypes as C
import struct
import zlib

import numpy as np

from . import _lib
from .batch import AlignmentBatch, cigar_to_blocks, MAX_ALIGNED_LEN, MAX_BLOCKS


def batch_from_bam(source, threads=0, pinned=False, pack=True):
    """Decode a coordinate-sorted BAM (path, or an open pysam file) into an AlignmentBatch.  ``pack``: also emit
    the batch's transfer format (``AlignmentBatch.pack``: delta3 streams + block words) — the form
    ``BAMGenomeArray`` uploads — while the decoder's threads are at hand."""
    if not isinstance(source, (str, bytes)):
        out = _batch_from_pysam(source)
        return out.pack(threads) if pack else out
    L = _lib.lib()
    handle = C.c_void_p()
    path = source.encode() if isinstance(source, str) else source
    _lib.check(L.pb_bam_open(path, C.byref(handle)))
    try:
        _lib.check(L.pb_bam_decode(handle, _lib.host_threads(threads)))
        n_ref = L.pb_bam_n_ref(handle)
        chroms = [L.pb_bam_ref_name(handle, i).decode() for i in range(n_ref)]
        lens = [L.pb_bam_ref_len(handle, i) for i in range(n_ref)]
        n, n_blk = L.pb_bam_n_reads(handle), L.pb_bam_n_blk(handle)

        def buf(count, dtype):
            if pinned:
                import torch
                tdt = {np.int32: torch.int32, np.uint32: torch.int32, np.int64: torch.int64}[dtype]
                return torch.empty(max(count, 1), dtype=tdt).pin_memory().numpy().view(dtype)[:count]
            return np.empty(count, dtype=dtype)
        start, meta = buf(n, np.int32), buf(n, np.uint32)
        off = np.empty(n_ref + 1, dtype=np.int64)
        blk_off = blk = None
        if n_blk:
            blk_off, blk = buf(n + 1, np.uint32), buf(2 * n_blk, np.int32)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _lib.check(L.pb_bam_copy(handle, p(start), p(meta), p(blk_off), p(blk), p(off)))
        out = AlignmentBatch(chroms, lens, start, meta, off, blk_off, None if blk is None else blk.reshape(-1, 2),
                             max_span=L.pb_bam_max_span(handle), mapped=L.pb_bam_n_mapped(handle))
        return out.pack(threads) if pack else out
    finally:
        L.pb_bam_close(handle)


def _copy_handle(L, handle, pinned=False):
    """The batch a decoded / fetched ``pb_bam`` handle holds -> :class:`AlignmentBatch`."""
    n_ref = L.pb_bam_n_ref(handle)
    chroms = [L.pb_bam_ref_name(handle, i).decode() for i in range(n_ref)]
    lens = [L.pb_bam_ref_len(handle, i) for i in range(n_ref)]
    n, n_blk = L.pb_bam_n_reads(handle), L.pb_bam_n_blk(handle)
    start, meta = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.uint32)
    off = np.empty(n_ref + 1, dtype=np.int64)
    blk_off = blk = None
    if n_blk:
        blk_off, blk = np.empty(n + 1, dtype=np.uint32), np.empty(2 * n_blk, dtype=np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)     # noqa: E731
    _lib.check(L.pb_bam_copy(handle, p(start), p(meta), p(blk_off), p(blk), p(off)))
    return AlignmentBatch(chroms, lens, start, meta, off, blk_off, None if blk is None else blk.reshape(-1, 2),
                          max_span=L.pb_bam_max_span(handle), mapped=L.pb_bam_n_mapped(handle))


def index_path(bam_path):
    """The index beside a BAM file: ``x.bam.bai`` or ``x.bai`` (what ``pysam.AlignmentFile`` looks for), else None."""
    import os
    for cand in (bam_path + ".bai", os.path.splitext(bam_path)[0] + ".bai"):
        if os.path.exists(cand):
            return cand
    return None


def build_index(bam_path, bai_path=None):
    """``samtools index``: write the ``.bai`` of a coordinate-sorted BAM (``pb_bam_build_index``).  Returns its path."""
    bai_path = bam_path + ".bai" if bai_path is None else bai_path
    _lib.check(_lib.lib().pb_bam_build_index(bam_path.encode(), bai_path.encode()))
    return bai_path


class IndexedBam(object):
    """A sorted BAM with its ``.bai``: header and index statistics without decoding a record, and
    ``fetch(chrom, start, end)`` -> :class:`AlignmentBatch` of the reads whose span overlaps the region — what
    ``pysam.AlignmentFile`` gives the reference (``references`` / ``lengths`` / ``mapped`` / ``fetch``;
    plastid/genomics/genome_array.py:660-690, 800-809).  Only the BGZF members the index points at are read."""

    def __init__(self, path, index=None):
        L = _lib.lib()
        self.path = path
        self._h, self._idx = C.c_void_p(), C.c_void_p()
        index = index_path(path) if index is None else index
        if index is None:
            raise IOError("no .bai index beside %s (plastid_b200.bam_io.build_index writes one)" % path)
        _lib.check(L.pb_bam_open(path.encode(), C.byref(self._h)))
        try:
            _lib.check(L.pb_bam_read_header(self._h))
            _lib.check(L.pb_bai_open(index.encode(), C.byref(self._idx)))
        except Exception:
            self.close()
            raise
        n_ref = L.pb_bam_n_ref(self._h)
        self.references = tuple(L.pb_bam_ref_name(self._h, i).decode() for i in range(n_ref))
        self.lengths = tuple(int(L.pb_bam_ref_len(self._h, i)) for i in range(n_ref))
        self._tid = {c: i for i, c in enumerate(self.references)}
        m = int(L.pb_bai_mapped(self._idx, -1))
        self.mapped = None if m < 0 else m          # None: an index written without the statistics pseudo-bin

    def fetch(self, chrom, start, end):
        L = _lib.lib()
        if self._h is None:
            raise ValueError("I/O operation on a closed IndexedBam")
        _lib.check(L.pb_bam_fetch(self._h, self._idx, self._tid.get(chrom, -1), int(start), int(end)))
        return _copy_handle(L, self._h)

    def close(self):
        L = _lib.lib()
        if getattr(self, "_idx", None):
            L.pb_bai_close(self._idx)
        if getattr(self, "_h", None):
            L.pb_bam_close(self._h)
        self._h = self._idx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _batch_from_pysam(bam):
    chroms, lengths = list(bam.references), list(bam.lengths)
    per_chrom = [[] for _ in chroms]
    for read in bam.fetch(until_eof=True):
        if read.is_unmapped or read.reference_id < 0 or not read.cigartuples:
            continue
        per_chrom[read.reference_id].append(read)
    starts, metas, blk_off, blks, off = [], [], [0], [], [0]
    multi = False
    for reads in per_chrom:
        recs = []
        for r in reads:
            blocks, _ = cigar_to_blocks(r.cigartuples)
            L = sum(n for _a, n in blocks)
            if L > MAX_ALIGNED_LEN or len(blocks) > MAX_BLOCKS:
                raise ValueError("alignment too long / too fragmented for the packed batch")
            s = r.reference_start
            if blocks and blocks[0][0]:
                s += blocks[0][0]
                blocks = [(a - blocks[0][0], n) for a, n in blocks]
            recs.append((s, L | (int(r.is_reverse) << 16) | (len(blocks) << 24), blocks))
        recs.sort(key=lambda x: x[0])
        for s, m, blocks in recs:
            starts.append(s)
            metas.append(m)
            if len(blocks) > 1:
                multi = True
                blks.extend(blocks)
            blk_off.append(len(blks))
        off.append(len(starts))
    kw = {}
    if multi:
        kw = dict(blk_off=np.asarray(blk_off, dtype=np.uint32), blk=np.asarray(blks, dtype=np.int32).reshape(-1, 2))
    return AlignmentBatch(chroms, lengths, np.asarray(starts, dtype=np.int32), np.asarray(metas, dtype=np.uint32),
                          off, mapped=bam.mapped, **kw)


# ------------------------------------------------------------------------------- writer (tests)
def _bgzf_block(data, level=6):
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    header = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, ord("B"), ord("C"), 2, bsize)
    return header + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def write_bam(path, chrom_lengths, records, block_bytes=60000, record_aligned=False, payload_rng=None):
    """Write a BAM file.  ``chrom_lengths``: ordered ``{name: length}``; ``records``: iterable of
    ``(chrom_index or -1, pos, flag, cigartuples)`` already in coordinate order.  ``record_aligned``:
    start a new BGZF member rather than split a record across two, as htslib's writer does
    (``bgzf_flush_try``); otherwise members are cut every ``block_bytes`` bytes wherever that falls.
    ``payload_rng``: a numpy Generator draws bases and (skewed) qualities so that the file compresses like
    sequencing data (about 3-4x); without it both are constant bytes."""
    chroms = list(chrom_lengths)
    text = "@HD\tVN:1.4\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, chrom_lengths[c]) for c in chroms)
    out = [b"BAM\x01", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(chroms))]
    for c in chroms:
        name = c.encode() + b"\0"
        out.append(struct.pack("<i", len(name)) + name + struct.pack("<i", chrom_lengths[c]))
    for i, (tid, pos, flag, cigar) in enumerate(records):
        name = ("r%d" % i).encode() + b"\0"
        qlen = sum(n for op, n in cigar if op in (0, 1, 4, 7, 8))
        span = sum(n for op, n in cigar if op in (0, 2, 3, 7, 8))
        end = pos + max(span, 1)
        bin_ = _reg2bin(max(pos, 0), max(end, 1))
        core = struct.pack("<iiBBHHHiiii", tid, pos, len(name), 60, bin_, len(cigar), flag, qlen, -1, -1, 0)
        body = core + name + b"".join(struct.pack("<I", (n << 4) | op) for op, n in cigar)
        if payload_rng is None:
            body += b"\x11" * ((qlen + 1) // 2) + b"\xff" * qlen
        else:
            seq = (1 << payload_rng.integers(0, 4, (qlen + 1) // 2)) * 16 + (1 << payload_rng.integers(0, 4, (qlen + 1) // 2))
            qual = 40 - np.minimum(payload_rng.geometric(0.35, qlen), 38)
            body += seq.astype(np.uint8).tobytes() + qual.astype(np.uint8).tobytes()
        out.append(struct.pack("<i", len(body)) + body)
    n_head = 4 + len(chroms)
    with open(path, "wb") as fh:
        if record_aligned:
            head = b"".join(out[:n_head])
            for a in range(0, len(head), block_bytes):
                fh.write(_bgzf_block(head[a:a + block_bytes]))
            pending, size = [], 0
            for rec in out[n_head:]:
                if pending and size + len(rec) > block_bytes:
                    fh.write(_bgzf_block(b"".join(pending)))
                    pending, size = [], 0
                if len(rec) > block_bytes:              # a record larger than a member has to span several
                    for a in range(0, len(rec), block_bytes):
                        fh.write(_bgzf_block(rec[a:a + block_bytes]))
                    continue
                pending.append(rec)
                size += len(rec)
            if pending:
                fh.write(_bgzf_block(b"".join(pending)))
        else:
            data = b"".join(out)
            for a in range(0, len(data), block_bytes):
                fh.write(_bgzf_block(data[a:a + block_bytes]))
        fh.write(_bgzf_block(b""))          # EOF marker


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0
