"""Lowering of ``SegmentChain`` collections to the flat block / mask tables the gather kernels read
(``pb_region_sums`` / ``pb_gather_windows`` in ``include/plastid_b200.h``)."""
import numpy as np

from . import _lib

# Positions of a region that lie outside its chromosome (an annotation longer than the BAM header says, a window
# flank, a negative start) are lowered to blocks in this coordinate range, beyond every layout: the gather kernels
# clip every block to the bin range they are given, so these positions count zero — the reference returns zeros
# there, `fetch` yields no reads (genome_array.py:800-809) — while keeping their place in the chain (length, mask
# bits and window columns are unchanged).
VIRTUAL_BIN = 1 << 60


def lower_segment(start, end, base, chrom_len):
    """Blocks ``[(bstart, bend), ...]`` in genomic order for the chromosome positions ``[start, end)`` of a chromosome
    whose bin 0 is global bin ``base``: the part inside ``[0, chrom_len)`` in global bins, the parts outside as
    virtual blocks of the same lengths."""
    out = []
    left = min(end, 0)
    if start < left:
        out.append((VIRTUAL_BIN, VIRTUAL_BIN + (left - start)))
    a, b = max(start, 0), min(end, chrom_len)
    if a < b:
        out.append((base + a, base + b))
    right = max(start, chrom_len, 0)
    if right < end:
        out.append((VIRTUAL_BIN + right, VIRTUAL_BIN + end))
    return out


class ChainTable(object):
    """Flat tables for a list of chains: blocks in global-bin coordinates, per-chain plane index and
    orientation, optional mask bits (bit j of chain c = j-th chain position in genomic order)."""

    def __init__(self, layout, bstart, bend, chain_off, chain_plane, chain_reverse, chain_len,
                 mask_bits=None, mask_off=None, known=None):
        self.layout = layout
        self.bstart = np.ascontiguousarray(bstart, dtype=np.int64)
        self.bend = np.ascontiguousarray(bend, dtype=np.int64)
        self.chain_off = np.ascontiguousarray(chain_off, dtype=np.int64)
        self.chain_plane = np.ascontiguousarray(chain_plane, dtype=np.uint8)
        self.chain_reverse = np.ascontiguousarray(chain_reverse, dtype=np.uint8)
        self.chain_len = np.ascontiguousarray(chain_len, dtype=np.int64)
        self.mask_bits = None if mask_bits is None else np.ascontiguousarray(mask_bits, dtype=np.uint8)
        self.mask_off = None if mask_off is None else np.ascontiguousarray(mask_off, dtype=np.int64)
        self.known = np.ones(len(self.chain_len), dtype=bool) if known is None else np.asarray(known, dtype=bool)
        # unmasked length of chains on chromosomes the layout does not know (they count 0 over their
        # full length: ga.get returns a broadcastable zero vector, genome_array.py:795-798)
        self.unknown_live = np.zeros(len(self.chain_len), dtype=np.int64)
        self._dev = {}

    @property
    def n_chains(self):
        return len(self.chain_len)

    @classmethod
    def from_chains(cls, chains, layout, use_masks=True, unstranded=False):
        """Chains on chromosomes missing from ``layout`` get zero blocks (the reference returns a
        zero vector for them, genome_array.py:795-798) and ``known`` False."""
        bstart, bend, chain_off = [], [], [0]
        plane, reverse, length, known = [], [], [], []
        mask_off, bits = [], []
        nbits = 0
        any_mask = False
        for ch in chains:
            ok = len(ch) > 0 and ch.chrom in layout.index
            known.append(ok or len(ch) == 0)
            strand = "." if unstranded or ch.strand not in ("+", "-") else ch.strand
            plane.append(_lib.PLANE_INDEX[strand])
            reverse.append(1 if ch.strand == "-" else 0)
            if ok:
                ci = layout.index[ch.chrom]
                base, clen = int(layout.chrom_bin_off[ci]), int(layout.chrom_len[ci])
                for seg in ch:
                    for piece in lower_segment(seg.start, seg.end, base, clen):
                        bstart.append(piece[0])
                        bend.append(piece[1])
                length.append(ch.length)
            else:
                length.append(0)
            chain_off.append(len(bstart))
            mask_off.append(nbits)
            if ok and use_masks and getattr(ch, "_mask_intervals", None):
                any_mask = True
                bits.append(ch.position_mask())
            else:
                bits.append(np.zeros(length[-1], dtype=np.uint8))
            nbits += length[-1]
        mask_bits = None
        if any_mask:
            flat = np.concatenate(bits) if bits else np.zeros(0, dtype=np.uint8)
            mask_bits = np.packbits(flat, bitorder="little")
            if len(mask_bits) == 0:
                mask_bits = np.zeros(1, dtype=np.uint8)
        out = cls(layout, bstart, bend, chain_off, plane, reverse, length,
                  mask_bits, mask_off if any_mask else None, known)
        for i, ch in enumerate(chains):
            if not out.known[i]:
                out.unknown_live[i] = ch.masked_length if use_masks else ch.length
        return out

    def device(self, device):
        import torch
        key = _lib.device_key(device)
        if key not in self._dev:
            def up(a):
                return None if a is None else torch.from_numpy(a).to(device)
            bits = self.mask_bits
            if bits is not None and len(bits) % 4:          # the kernels read the bit array as whole 32-bit words
                bits = np.concatenate([bits, np.zeros(4 - len(bits) % 4, dtype=np.uint8)])
            # per block: its chain and the chain position (genomic order) of its first base
            n_blk = np.diff(self.chain_off)
            block_chain = np.repeat(np.arange(self.n_chains, dtype=np.int32), n_blk)
            blen = self.bend - self.bstart
            before = np.cumsum(blen) - blen
            first = before[np.minimum(self.chain_off[:-1], max(len(blen) - 1, 0))] if len(blen) else np.zeros(self.n_chains, dtype=np.int64)
            block_pos = (before - np.repeat(first, n_blk)).astype(np.int64)
            self._dev[key] = dict(bstart=up(self.bstart), bend=up(self.bend), chain_off=up(self.chain_off),
                                  chain_plane=up(self.chain_plane), chain_reverse=up(self.chain_reverse),
                                  block_chain=up(block_chain), block_pos=up(block_pos), chain_len=up(self.chain_len),
                                  block_plane=up(np.repeat(self.chain_plane, n_blk)),
                                  mask_bits=up(bits), mask_off=up(self.mask_off))
        return self._dev[key]

    def live_lengths(self):
        """Unmasked length of every chain (``chain.masked_length``): geometry only — what ``pb_region_sums`` returns in
        ``live_len`` whatever bin range it is given."""
        if getattr(self, "_live", None) is None:
            live = self.chain_len.astype(np.int64).copy()
            if self.mask_bits is not None and self.n_chains:
                nbits = int(self.mask_off[-1] + self.chain_len[-1])
                bits = np.unpackbits(self.mask_bits, bitorder="little")[:nbits].astype(np.int64)
                run = np.concatenate([[0], np.cumsum(bits)])
                live -= run[self.mask_off + self.chain_len] - run[self.mask_off]
            self._live = live
        return self._live

    def owned_rows(self, lo, hi):
        """(indices of the blocks that overlap the global bins [lo, hi), chain offsets into that selection)."""
        idx = np.nonzero((self.bend > lo) & (self.bstart < hi))[0]
        chain_of = np.repeat(np.arange(self.n_chains), np.diff(self.chain_off))
        sub_off = np.zeros(self.n_chains + 1, dtype=np.int64)
        np.cumsum(np.bincount(chain_of[idx], minlength=self.n_chains), out=sub_off[1:])
        return idx, sub_off

    def touched_rows(self, lo, hi):
        """Like :meth:`owned_rows`, but EVERY block of a chain that has a block in [lo, hi): what a rank walks when the
        rows it touches must come out whole (window matrices whose rows are completed rank by rank)."""
        chain_of = np.repeat(np.arange(self.n_chains), np.diff(self.chain_off))
        hit = np.zeros(self.n_chains, dtype=bool)
        hit[chain_of[(self.bend > lo) & (self.bstart < hi)]] = True
        idx = np.nonzero(hit[chain_of])[0]
        sub_off = np.zeros(self.n_chains + 1, dtype=np.int64)
        np.cumsum(np.bincount(chain_of[idx], minlength=self.n_chains), out=sub_off[1:])
        return idx, sub_off

    def owned(self, device, lo, hi, whole_chains=False):
        """The rows of the block tables a rank that owns the global bins [lo, hi) has to look at (position sharding):
        blocks that overlap the range, with the chain offsets that go with them; ``block_pos`` / ``block_chain`` keep
        their values from the whole table, so mask bits and chain totals are addressed as before.  At N = 8 a rank
        otherwise walks all blocks of the table to find that 7 / 8 of them lie elsewhere.  Cached per range."""
        import torch
        key = (_lib.device_key(device), int(lo), int(hi), bool(whole_chains))
        cache = self.__dict__.setdefault("_owned", {})
        if key not in cache:
            d = self.device(device)
            idx, sub_off = self.touched_rows(lo, hi) if whole_chains else self.owned_rows(lo, hi)
            t_idx = torch.from_numpy(idx).to(device)
            sub = dict(d)
            for name in ("bstart", "bend", "block_chain", "block_pos", "block_plane"):
                sub[name] = d[name][t_idx].contiguous()
            sub["chain_off"] = torch.from_numpy(sub_off).to(device)
            sub["n_blocks"] = int(len(idx))
            sub["live"] = torch.from_numpy(self.live_lengths()).to(device)
            cache[key] = sub
        return cache[key]

