"""BAM -> SoA batch on the host.

The reference leaves BAM decoding to pysam (``plastid/genomics/genome_array.py:660, 800-809``).
pysam is a third-party dependency that is absent from this image; when it is importable a sorted
BAM is decoded once, whole, into an :class:`~plastid_b200.batch.AlignmentBatch` (no per-region
``fetch``).  Like the reference, no flag other than is_reverse is interpreted: secondary,
duplicate and QC-fail alignments are counted; unmapped records (no reference id / no CIGAR) have
no positions and are skipped, as ``fetch`` never returns them.
"""
import numpy as np

from .batch import AlignmentBatch, cigar_to_blocks, MAX_ALIGNED_LEN, MAX_BLOCKS


def batch_from_bam(source):
    try:
        import pysam
    except ImportError:
        raise ImportError("decoding BAM files needs pysam, which is not installed here; "
                          "build an AlignmentBatch with plastid_b200.batch.pack_reads / batch_from_arrays instead")
    bam = pysam.AlignmentFile(source, "rb") if isinstance(source, str) else source
    chroms, lengths = list(bam.references), list(bam.lengths)
    per_chrom = [[] for _ in chroms]
    for read in bam.fetch(until_eof=True):
        if read.is_unmapped or read.reference_id < 0 or not read.cigartuples:
            continue
        per_chrom[read.reference_id].append(read)
    starts, metas, blk_off, blks = [], [], [0], []
    off = [0]
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
