"""Device mask pipeline: the masked positions of every region of a list in one launch.

The reference masks regions one at a time: ``GenomeHash.get_overlapping_features(region)``
(plastid/genomics/genome_hash.py:259-436) collects the mask features that share a position with the
region on its chromosome and strand, ``SegmentChain.add_masks`` (plastid/genomics/roitools.pyx:2213-2301)
intersects the union of their positions with the region through Python ``set`` arithmetic
(``counts_in_region.py:114-115``).  The result per region is (union of same-strand mask features) ∩
(region positions); here that union is built once as merged, sorted intervals per strand
(:class:`MaskIndex`) and ``pb_mask_chains`` writes the mask bits of all regions at once.
"""
import numpy as np

from . import _lib
from .roitools import _merge_intervals


class MaskIndex(object):
    """Union of mask features per chromosome strand, in the global-bin coordinates of ``layout``.

    ``features``: SegmentChains (or anything iterable over GenomicSegments).  Like the reference's
    ``GenomeHash._make_hash`` (genome_hash.py:236-257), which only creates '+' and '-' tables, a mask
    feature on strand '.' raises ``KeyError``; features on chromosomes the layout does not know
    cannot overlap any countable region and are ignored."""

    def __init__(self, features, layout):
        self.layout = layout
        per = {}
        for f in features:
            for seg in f:
                if seg.strand not in ("+", "-"):
                    raise KeyError(seg.strand)
                if seg.chrom in layout.index and seg.end > seg.start:
                    per.setdefault((_lib.PLANE_INDEX[seg.strand], layout.index[seg.chrom]), []).append((seg.start, seg.end))
        starts, ends, class_off = [], [], [0]
        for cls in range(3):
            for (k, c) in sorted(key for key in per if key[0] == cls):
                base = int(layout.chrom_bin_off[c])
                for a, b in _merge_intervals(per[(k, c)]):
                    starts.append(base + a)
                    ends.append(base + b)
            class_off.append(len(starts))
        self.mask_start = np.asarray(starts, dtype=np.int64)
        self.mask_end = np.asarray(ends, dtype=np.int64)
        self.class_off = np.asarray(class_off, dtype=np.int64)
        self._dev = {}

    def __len__(self):
        return len(self.mask_start)

    def device(self, device):
        import torch
        key = _lib.device_key(device)
        if key not in self._dev:
            def up(a):
                return torch.from_numpy(a if len(a) else np.zeros(1, dtype=a.dtype)).to(device)
            self._dev[key] = (up(self.mask_start), up(self.mask_end), up(self.class_off))
        return self._dev[key]


class GenomeHash(object):
    """``GenomeHash(features)``: the mask container the reference's programs pass around
    (plastid/genomics/genome_hash.py:81-436).  Queries for many regions at once go through
    :meth:`mask_index` + :func:`apply_mask_index` (one launch); ``get_overlapping_features`` /
    ``hash[roi]`` answer for one region on the host with the reference's semantics: same chromosome,
    same strand, at least one shared position.  Like the reference's ``_make_hash`` (:236-257) only '+'
    and '-' features are accepted (a '.' feature raises ``KeyError``)."""

    def __init__(self, features=None, binsize=20000, do_copy=True):
        self.binsize = binsize
        self.features = list(features or [])
        for f in self.features:
            if len(f) and f.strand not in ("+", "-"):
                raise KeyError(f.strand)
        self._index = {}

    def __len__(self):
        return len(self.features)

    def mask_index(self, layout):
        key = id(layout)
        if key not in self._index:
            self._index[key] = (layout, MaskIndex(self.features, layout))
        return self._index[key][1]

    def get_overlapping_features(self, roi, stranded=True):
        out = []
        for f in self.features:
            if len(f) == 0 or len(roi) == 0 or f.chrom != roi.chrom:
                continue
            if stranded and f.strand != roi.strand:
                continue
            if any(a.start < b.end and b.start < a.end for a in f for b in roi):
                out.append(f)
        return out

    __getitem__ = get_overlapping_features


def mask_intervals_of_chains(table, bits):
    """Decode the mask bits of every chain of ``table`` (device tensor written by
    :func:`apply_mask_index`) into per-chain lists of masked ``(start, end)`` intervals in chromosome
    coordinates — what ``SegmentChain.get_masks`` reports after the reference's ``add_masks``."""
    flat = np.unpackbits(bits.cpu().numpy(), bitorder="little")
    out, off = [], 0
    for c in range(table.n_chains):
        ivs = []
        k0, k1 = int(table.chain_off[c]), int(table.chain_off[c + 1])
        if k1 > k0:
            real = [k for k in range(k0, k1) if table.bstart[k] < int(table.layout.total_bins)]    # not beyond the chromosome
            base = int(table.layout.chrom_bin_off[np.searchsorted(table.layout.chrom_bin_off, table.bstart[real[0]], side="right") - 1]) if real else 0
            for k in range(k0, k1):
                n = int(table.bend[k] - table.bstart[k])
                m = flat[off:off + n]
                off += n
                if m.any() and table.bstart[k] < int(table.layout.total_bins):
                    edge = np.flatnonzero(np.diff(np.concatenate(([0], m, [0]))))
                    start = int(table.bstart[k]) - base
                    for a, b in zip(edge[0::2], edge[1::2]):
                        if ivs and ivs[-1][1] == start + int(a):
                            ivs[-1] = (ivs[-1][0], start + int(b))
                        else:
                            ivs.append((start + int(a), start + int(b)))
        out.append(ivs)
    return out


def apply_mask_index(table, index, device):
    """OR the masks of ``index`` into the mask bits of every chain of ``table`` (a
    :class:`~plastid_b200.regions.ChainTable`) on ``device``; masks the chains already carry from
    ``add_masks`` are kept.  Afterwards ``region_sums`` / ``gather_windows`` on this table see the
    combined masks.  Returns the device tensor of mask bits."""
    import torch
    _lib.require_cuda()
    d = table.device(device)
    n = table.n_chains
    mask_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(table.chain_len, out=mask_off[1:])
    n_bytes = ((int(mask_off[-1]) + 31) // 32) * 4 + 4
    bits = torch.zeros(n_bytes, dtype=torch.uint8, device=device)
    if table.mask_bits is not None:
        bits[:len(table.mask_bits)] = torch.from_numpy(table.mask_bits).to(device)
    ms, me, co = index.device(device)
    d_off = torch.from_numpy(mask_off[:-1].copy() if n else np.zeros(1, dtype=np.int64)).to(device)
    _lib.check(_lib.lib().pb_mask_chains(_lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["chain_off"]),
                                         _lib.ptr(d["chain_plane"]), n, _lib.ptr(ms), _lib.ptr(me), _lib.ptr(co),
                                         _lib.ptr(d_off), _lib.ptr(bits), _lib.stream_ptr()))
    d["mask_bits"], d["mask_off"] = bits, d_off
    return bits
