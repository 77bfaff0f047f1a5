"""Seeded synthetic genomes, annotations and alignment batches (SURVEY.md §8d recipes).

One vectorised torch code path, so the same generator serves the CPU tests (device="cpu") and the
benchmark-scale batches (device="cuda", 200 M reads in a few hundred ms).  Nothing here is on the
product's compute path.
"""
import numpy as np
import torch

from .batch import AlignmentBatch, DeviceBatch
from .roitools import GenomicSegment, SegmentChain

# hg38 primary assembly chromosome lengths (chr1..22, X, Y): 3.09 Gb
HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
HG38_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]

# p_site.rst:143-151-style offsets table
RIBO_OFFSETS = {25: 12, 26: 12, 27: 13, 28: 13, 29: 14, 30: 14, 31: 14, 32: 14, 33: 14, 34: 14, 35: 14,
                "default": 14}


def yeast_like_genome(total=12_000_000, n_chrom=16):
    w = np.linspace(0.4, 1.6, n_chrom)
    lens = np.floor(w / w.sum() * total).astype(np.int64)
    lens[-1] += total - lens.sum()
    return ["chr%s" % (i + 1) for i in range(n_chrom)], lens


def human_like_genome(scale=1.0):
    lens = np.maximum((np.asarray(HG38_LENGTHS, dtype=np.float64) * scale).astype(np.int64), 100000)
    return list(HG38_NAMES), lens


class Annotation(object):
    """Flat exon table of synthetic transcripts: transcript t owns exons [tx_off[t], tx_off[t+1])."""

    def __init__(self, chroms, chrom_len, tx_chrom, tx_strand, tx_off, ex_start, ex_end):
        self.chroms, self.chrom_len = chroms, np.asarray(chrom_len, dtype=np.int64)
        self.tx_chrom, self.tx_strand = np.asarray(tx_chrom), np.asarray(tx_strand)
        self.tx_off, self.ex_start, self.ex_end = np.asarray(tx_off), np.asarray(ex_start), np.asarray(ex_end)

    @property
    def n_tx(self):
        return len(self.tx_chrom)

    def chain(self, t):
        strand = "-" if self.tx_strand[t] else "+"
        chrom = self.chroms[int(self.tx_chrom[t])]
        return SegmentChain(*[GenomicSegment(chrom, int(self.ex_start[k]), int(self.ex_end[k]), strand)
                              for k in range(int(self.tx_off[t]), int(self.tx_off[t + 1]))], ID="tx%d" % t)

    def chains(self):
        return [self.chain(t) for t in range(self.n_tx)]


def make_annotation(chroms, chrom_len, n_tx, seed=0, exons=(1, 2), exon_len=(300, 1200), intron_len=(80, 2000)):
    """``n_tx`` non-overlapping transcripts spread over the genome proportionally to chromosome length."""
    rng = np.random.default_rng(seed)
    chrom_len = np.asarray(chrom_len, dtype=np.int64)
    per = np.floor(n_tx * chrom_len / chrom_len.sum()).astype(np.int64)
    per[np.argmax(chrom_len)] += n_tx - per.sum()
    tx_chrom, tx_strand, tx_off, ex_s, ex_e = [], [], [0], [], []
    max_span = exons[1] * exon_len[1] + (exons[1] - 1) * intron_len[1]
    for c, n in enumerate(per):
        if n == 0:
            continue
        slot = int(chrom_len[c] // n)
        if slot <= max_span + 200:
            raise ValueError("chromosome %s too small for %d transcripts" % (chroms[c], n))
        for j in range(int(n)):
            pos = j * slot + int(rng.integers(50, slot - max_span - 50))
            ne = int(rng.integers(exons[0], exons[1] + 1))
            for e in range(ne):
                ln = int(rng.integers(exon_len[0], exon_len[1] + 1))
                ex_s.append(pos)
                ex_e.append(pos + ln)
                pos += ln + int(rng.integers(intron_len[0], intron_len[1] + 1))
            tx_chrom.append(c)
            tx_strand.append(int(rng.integers(0, 2)))
            tx_off.append(len(ex_s))
    return Annotation(chroms, chrom_len, tx_chrom, tx_strand, tx_off, ex_s, ex_e)


def make_masks(annotation, frac=0.10, block=200, seed=1):
    """Per-transcript mask segments (200-nt blocks covering about ``frac`` of exon positions)."""
    rng = np.random.default_rng(seed)
    out = []
    for t in range(annotation.n_tx):
        segs = []
        chrom = annotation.chroms[int(annotation.tx_chrom[t])]
        strand = "-" if annotation.tx_strand[t] else "+"
        for k in range(int(annotation.tx_off[t]), int(annotation.tx_off[t + 1])):
            s, e = int(annotation.ex_start[k]), int(annotation.ex_end[k])
            n_blocks = rng.binomial(max((e - s) // block, 1), frac)
            for _ in range(int(n_blocks)):
                a = int(rng.integers(s - block // 2, e))
                segs.append(GenomicSegment(chrom, max(a, 0), a + block, strand))
        out.append(segs)
    return out


def _length_sampler(gen, n, lengths, weights, device):
    w = torch.tensor(weights, dtype=torch.float64, device=device)
    idx = torch.multinomial(w / w.sum(), n, replacement=True, generator=gen)
    return torch.tensor(lengths, dtype=torch.int64, device=device)[idx]


def riboseq_reads(annotation, n_reads, seed=0, device="cpu", frac_in=0.9, lengths=None, weights=None,
                  spliced_frac=0.0):
    """Unspliced ribo-seq-like reads: ``frac_in`` of them start inside exons with 3-nt periodicity,
    the rest uniformly over the genome; 50/50 strands for background, transcript strand otherwise.
    Returns a :class:`DeviceBatch` on ``device`` (sorted by chromosome, start)."""
    lengths = list(range(25, 36)) if lengths is None else list(lengths)
    if weights is None:
        weights = [np.exp(-0.5 * ((L - 29) / 2.0) ** 2) for L in lengths]
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    chrom_len = torch.tensor(annotation.chrom_len, dtype=torch.int64, device=device)
    n_in = int(n_reads * frac_in)
    n_bg = n_reads - n_in
    ex_s = torch.tensor(annotation.ex_start, dtype=torch.int64, device=device)
    ex_e = torch.tensor(annotation.ex_end, dtype=torch.int64, device=device)
    ex_tx = torch.repeat_interleave(torch.arange(annotation.n_tx, device=device),
                                    torch.tensor(np.diff(annotation.tx_off), device=device))
    ex_chrom = torch.tensor(annotation.tx_chrom, dtype=torch.int64, device=device)[ex_tx]
    ex_rev = torch.tensor(annotation.tx_strand, dtype=torch.int64, device=device)[ex_tx]
    # exon choice weighted by length * a per-transcript expression level (log-normal)
    expr = torch.exp(1.5 * torch.randn(annotation.n_tx, generator=gen, device=device, dtype=torch.float64))
    w = (ex_e - ex_s).double() * expr[ex_tx]
    ex = torch.multinomial(w / w.sum(), n_in, replacement=True, generator=gen)
    L_in = _length_sampler(gen, n_in, lengths, weights, device)
    span = (ex_e - ex_s)[ex]
    codon = (torch.rand(n_in, generator=gen, device=device, dtype=torch.float64) * (span // 3).double()).long()
    frame_noise = (torch.rand(n_in, generator=gen, device=device) < 0.15).long() * \
        torch.randint(1, 3, (n_in,), generator=gen, device=device)
    psite = ex_s[ex] + codon * 3 + frame_noise                       # genomic P-site
    off = torch.clamp(L_in // 2 - 1, 10, 16)
    rev_in = ex_rev[ex]
    start_in = torch.where(rev_in == 1, psite - (L_in - 1 - off), psite - off)
    chrom_in = ex_chrom[ex]
    # background
    cw = chrom_len.double()
    chrom_bg = torch.multinomial(cw / cw.sum(), max(n_bg, 1), replacement=True, generator=gen)[:n_bg]
    L_bg = _length_sampler(gen, max(n_bg, 1), lengths, weights, device)[:n_bg]
    start_bg = (torch.rand(n_bg, generator=gen, device=device, dtype=torch.float64)
                * (chrom_len[chrom_bg] - L_bg - 1).double()).long()
    rev_bg = torch.randint(0, 2, (n_bg,), generator=gen, device=device)
    chrom = torch.cat([chrom_in, chrom_bg])
    L = torch.cat([L_in, L_bg])
    rev = torch.cat([rev_in, rev_bg])
    start = torch.cat([start_in, start_bg])
    start = torch.minimum(torch.clamp(start, min=0), chrom_len[chrom] - L)
    return _finish(annotation.chroms, chrom_len, chrom, start, L, rev, device)


def _finish(chroms, chrom_len, chrom, start, L, rev, device, blocks=None):
    key = chrom * (1 << 32) + start
    key, order = torch.sort(key, stable=True)
    start_s = (key & 0xFFFFFFFF).to(torch.int32)
    chrom_s = (key >> 32)
    nblk = torch.ones_like(L) if blocks is None else blocks[0]
    meta = (L | (rev << 16) | (nblk << 24))[order].to(torch.int32)          # bit pattern of uint32
    counts = torch.bincount(chrom_s, minlength=len(chroms))
    off = torch.zeros(len(chroms) + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(counts, 0)
    blk_off = blk = None
    max_span = int(L.max().item()) if L.numel() else 1
    max_block_len = max_span
    if blocks is not None:
        nb, rel, ln = blocks            # nb[N], rel[N,K], ln[N,K] (K = max blocks; unused entries 0)
        nb_s, rel_s, ln_s = nb[order], rel[order], ln[order]
        listed = torch.where(nb_s > 1, nb_s, torch.zeros_like(nb_s))
        blk_off = torch.zeros(len(order) + 1, dtype=torch.int64, device=device)
        blk_off[1:] = torch.cumsum(listed, 0)
        K = rel.shape[1]
        valid = (torch.arange(K, device=device)[None, :] < listed[:, None])
        blk = torch.stack([rel_s[valid], ln_s[valid]], dim=1).to(torch.int32).contiguous()
        max_span = max(max_span, int((rel + ln).max().item()))
        blk_off = blk_off.to(torch.int32)
    # batch metadata a decoder knows for free: reads per aligned length
    hist = torch.bincount(L.to(torch.int64), minlength=65536).cpu().numpy() if L.numel() else np.zeros(65536, dtype=np.int64)
    return DeviceBatch(len(order), len(chroms), max_span, start_s.contiguous(), meta.contiguous(), off, blk_off, blk,
                       max_block_len, hist)


def rnaseq_reads(chroms, chrom_len, n_reads, seed=0, device="cpu", read_len=100, one_gap=0.30, two_gaps=0.03,
                 intron=(100, 50000), pileup=0):
    """100-nt RNA-seq-like reads, some with one or two ``N`` gaps (BASELINE config 3).  ``pileup``: that many of
    the reads lie in the last 16.5 kb of the LAST chromosome — what the mitochondrial genome (chrM, last in hg38 order, 10-30 %
    of the reads of an RNA-seq library) does to the tiles the mapping kernels reach last."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    clen = torch.tensor(np.asarray(chrom_len), dtype=torch.int64, device=device)
    cw = clen.double()
    chrom = torch.multinomial(cw / cw.sum(), n_reads, replacement=True, generator=gen)
    u = torch.rand(n_reads, generator=gen, device=device)
    nb = torch.ones(n_reads, dtype=torch.int64, device=device)
    nb[u < one_gap + two_gaps] = 2
    nb[u < two_gaps] = 3
    gap1 = torch.randint(intron[0], intron[1], (n_reads,), generator=gen, device=device)
    gap2 = torch.randint(intron[0], intron[1], (n_reads,), generator=gen, device=device)
    cut1 = torch.randint(10, read_len // 2, (n_reads,), generator=gen, device=device)
    cut2 = cut1 + torch.randint(10, read_len // 2 - 10, (n_reads,), generator=gen, device=device)
    zeros = torch.zeros_like(cut1)
    L = torch.full((n_reads,), read_len, dtype=torch.int64, device=device)
    len0 = torch.where(nb == 1, L, cut1)
    len1 = torch.where(nb == 1, zeros, torch.where(nb == 2, L - cut1, cut2 - cut1))
    len2 = torch.where(nb == 3, L - cut2, zeros)
    rel0 = zeros
    rel1 = torch.where(nb >= 2, cut1 + gap1, zeros)
    rel2 = torch.where(nb == 3, cut2 + gap1 + gap2, zeros)
    rel = torch.stack([rel0, rel1, rel2], 1)
    ln = torch.stack([len0, len1, len2], 1)
    span = (rel + ln).max(dim=1).values
    start = (torch.rand(n_reads, generator=gen, device=device, dtype=torch.float64)
             * (clen[chrom] - span - 1).clamp(min=1).double()).long()
    rev = torch.randint(0, 2, (n_reads,), generator=gen, device=device)
    if pileup:
        k = min(int(pileup), n_reads)
        chrom[:k] = len(chroms) - 1
        nb[:k] = 1
        last = int(clen[-1].item())
        start[:k] = last - 16_700 + torch.randint(0, 16_500, (k,), generator=gen, device=device)
    return _finish(chroms, clen, chrom, start, L, rev, device, blocks=(nb, rel, ln))


def device_batch_to_host(db, chroms, chrom_len, mapped=None):
    """Host :class:`AlignmentBatch` copy of a :class:`DeviceBatch` (for oracles and e2e timing)."""
    def h(t, dt):
        return None if t is None else t.cpu().numpy().view(dt) if dt is not None else t.cpu().numpy()
    return AlignmentBatch(chroms, np.asarray(chrom_len), h(db.ref_start, None), h(db.meta, np.uint32),
                          h(db.chrom_read_off, None), h(db.blk_off, np.uint32), h(db.blk, None),
                          max_span=db.max_span, mapped=mapped)


def annotation_table(annotation, layout):
    """Vectorised :class:`~plastid_b200.regions.ChainTable` for every transcript of a synthetic
    annotation (no masks) — same tables ``ChainTable.from_chains(annotation.chains(), layout)`` builds."""
    from .regions import ChainTable
    ex_tx = np.repeat(np.arange(annotation.n_tx), np.diff(annotation.tx_off))
    base = layout.chrom_bin_off[np.asarray([layout.index[c] for c in annotation.chroms])][annotation.tx_chrom]
    bstart = base[ex_tx] + annotation.ex_start
    bend = base[ex_tx] + annotation.ex_end
    length = np.add.reduceat(annotation.ex_end - annotation.ex_start, annotation.tx_off[:-1])
    return ChainTable(layout, bstart, bend, annotation.tx_off, annotation.tx_strand.astype(np.uint8),
                      annotation.tx_strand.astype(np.uint8), length)


def window_table(annotation, layout, width=350, mask_frac=0.05, mask_block=20, seed=7):
    """Metagene-style windows (BASELINE config 4): for every transcript the first ``width`` positions
    of the chain in 5'->3' order (shorter chains give shorter windows, right-aligned like an upstream
    flank that runs off the transcript), about ``mask_frac`` of positions masked in ``mask_block``-nt
    runs.  Returns (ChainTable, row_col int32[n])."""
    from .regions import ChainTable
    rng = np.random.default_rng(seed)
    bstart, bend, chain_off, plane, length, row_col = [], [], [0], [], [], []
    base_of = layout.chrom_bin_off[np.asarray([layout.index[c] for c in annotation.chroms])]
    bits = []
    for t in range(annotation.n_tx):
        a, b = int(annotation.tx_off[t]), int(annotation.tx_off[t + 1])
        ex = [(int(annotation.ex_start[k]), int(annotation.ex_end[k])) for k in range(a, b)]
        rev = bool(annotation.tx_strand[t])
        need, segs = width, []
        for s, e in (reversed(ex) if rev else ex):            # walk 5'->3'
            take = min(need, e - s)
            segs.append((e - take, e) if rev else (s, s + take))
            need -= take
            if need == 0:
                break
        segs.sort()
        base = int(base_of[int(annotation.tx_chrom[t])])
        for s, e in segs:
            bstart.append(base + s)
            bend.append(base + e)
        n = width - need
        chain_off.append(len(bstart))
        plane.append(1 if rev else 0)
        length.append(n)
        row_col.append(width - n)
        m = np.zeros(n, dtype=np.uint8)
        for _ in range(rng.binomial(max(n // mask_block, 1), mask_frac)):
            p = int(rng.integers(0, max(n - mask_block, 1)))
            m[p:p + mask_block] = 1
        bits.append(m)
    mask_off = np.zeros(len(length), dtype=np.int64)
    np.cumsum(length[:-1], out=mask_off[1:])
    flat = np.concatenate(bits) if bits else np.zeros(0, dtype=np.uint8)
    table = ChainTable(layout, bstart, bend, chain_off, plane, plane, length,
                       np.packbits(flat, bitorder="little"), mask_off)
    return table, np.asarray(row_col, dtype=np.int32)


def isoform_table(annotation, layout, n_iso=3, seed=9, max_shift=150, max_landmark=300):
    """``metagene generate`` input at annotation scale: every transcript of ``annotation`` becomes a gene
    with ``n_iso`` isoforms (as is; 5'-most genomic exon extended to the left; 3'-most genomic exon
    extended to the right; further isoforms repeat these with new extents) that share one landmark
    position.  Returns (TranscriptTable, grp_off, grp_tx): gene t = isoforms ``grp_tx[grp_off[t]:grp_off[t+1]]``."""
    from .windows import TranscriptTable
    rng = np.random.default_rng(seed)
    n = annotation.n_tx
    tx_off = np.asarray(annotation.tx_off, dtype=np.int64)
    counts = np.diff(tx_off)
    base = layout.chrom_bin_off[np.asarray([layout.index[c] for c in annotation.chroms])][np.asarray(annotation.tx_chrom)]
    ex_s = np.asarray(annotation.ex_start, dtype=np.int64) + np.repeat(base, counts)
    ex_e = np.asarray(annotation.ex_end, dtype=np.int64) + np.repeat(base, counts)
    rev = np.asarray(annotation.tx_strand, dtype=np.uint8)
    length = np.add.reduceat(ex_e - ex_s, tx_off[:-1])
    lm0 = np.minimum(rng.integers(0, max_landmark, n), length - 1)
    bs, be, rv, lm, ch = [], [], [], [], []
    for k in range(n_iso):
        s, e = ex_s.copy(), ex_e.copy()
        five = np.zeros(n, dtype=np.int64)
        if k % 3 == 1:
            d = rng.integers(1, max_shift, n)
            s[tx_off[:-1]] -= d
            five = np.where(rev == 0, d, 0)
        elif k % 3 == 2:
            d = rng.integers(1, max_shift, n)
            e[tx_off[1:] - 1] += d
            five = np.where(rev == 1, d, 0)
        bs.append(s)
        be.append(e)
        rv.append(rev)
        lm.append(lm0 + five)
        ch.append(np.asarray(annotation.tx_chrom))
    table = TranscriptTable(layout, np.concatenate(bs), np.concatenate(be), _stack_offsets(tx_off, n_iso),
                            np.concatenate(rv), np.concatenate(lm), np.concatenate(ch))
    grp_off = np.arange(n + 1, dtype=np.int64) * n_iso
    grp_tx = (np.arange(n, dtype=np.int64)[:, None] + np.arange(n_iso, dtype=np.int64)[None, :] * n).reshape(-1)
    return table, grp_off, grp_tx


def _stack_offsets(tx_off, copies):
    """Offsets of ``copies`` concatenated copies of a block table with offsets ``tx_off``."""
    total = int(tx_off[-1])
    return np.concatenate([[0]] + [tx_off[1:] + k * total for k in range(copies)]).astype(np.int64)
