"""Geometry of ``metagene generate`` on the device: landmark windows of all transcripts
and the maximal spanning window of every gene (``pb_landmark_windows`` / ``pb_spanning_windows`` in
``include/plastid_b200.h``).

The reference builds, per gene, a (transcripts x window) matrix of genomic positions with python loops
and keeps the columns on which all rows agree (plastid/bin/metagene.py:343-502, called once per gene from
``group_regions_make_windows`` :702-735).  Here the transcripts of the whole annotation are lowered once
into a flat block table (:class:`TranscriptTable`), and all genes are solved by two launches (count,
fill) with one warp per gene.
"""
import numpy as np

from . import _lib
from .batch import GenomeLayout


def layout_for_features(*feature_lists):
    """A :class:`GenomeLayout` covering every chromosome the given chains touch (annotation-only
    programs have no alignment header to take chromosome lengths from).  One spare position per
    chromosome keeps reference points just past a transcript's end inside the chromosome's bins."""
    extent = {}
    for feats in feature_lists:
        for ch in feats:
            for seg in ch:
                if seg.end + 1 > extent.get(seg.chrom, 0):
                    extent[seg.chrom] = seg.end + 1
    if not extent:
        extent = {"_": 1}
    chroms = sorted(extent)
    return GenomeLayout(chroms, [extent[c] for c in chroms])


# per-transcript orientation byte of the window kernels: 0 '+', 1 '-' (coordinates run right to left), 2 '.' —
# coordinates run left to right like '+', but the reference lays the window's columns right to left like '-'
# (metagene.py:443-455 tests `strand == "+"`), and so do the kernels
STRAND_CODE = {"+": 0, "-": 1, ".": 2}


class TranscriptTable(object):
    """Flat block table of transcripts in the global-bin coordinates of ``layout``: the input of
    ``pb_landmark_windows`` / ``pb_spanning_windows``."""

    def __init__(self, layout, bstart, bend, tx_off, reverse, landmark, tx_chrom):
        self.layout = layout
        self.bstart = np.ascontiguousarray(bstart, dtype=np.int64)
        self.bend = np.ascontiguousarray(bend, dtype=np.int64)
        self.tx_off = np.ascontiguousarray(tx_off, dtype=np.int64)
        self.reverse = np.ascontiguousarray(reverse, dtype=np.uint8)
        self.landmark = np.ascontiguousarray(landmark, dtype=np.int64)
        self.tx_chrom = np.ascontiguousarray(tx_chrom, dtype=np.int64)
        n_tx = len(self.reverse)
        if len(self.tx_off) != n_tx + 1 or len(self.landmark) != n_tx or len(self.tx_chrom) != n_tx:
            raise ValueError("TranscriptTable: per-transcript arrays differ in length")
        if len(self.bstart) != len(self.bend) or (n_tx and int(self.tx_off[-1]) != len(self.bstart)):
            raise ValueError("TranscriptTable: block arrays do not match tx_off")
        # chain coordinate of each block's first base: running sum of block lengths, restarted per transcript
        blen = self.bend - self.bstart
        if len(blen):
            run = np.cumsum(blen) - blen
            first = run[np.minimum(self.tx_off[:-1], len(run) - 1)]
            self.bcum = np.ascontiguousarray(run - np.repeat(first, np.diff(self.tx_off)), dtype=np.int64)
        else:
            self.bcum = np.zeros(0, dtype=np.int64)
        self._dev = {}

    @property
    def n_tx(self):
        return len(self.reverse)

    @classmethod
    def from_transcripts(cls, transcripts, layout, landmarks):
        """``landmarks[i]``: transcript coordinate of transcript i's landmark, or None."""
        bstart, bend, tx_off, reverse, lm, chrom = [], [], [0], [], [], []
        for tx, mark in zip(transcripts, landmarks):
            base = int(layout.chrom_bin_off[layout.index[tx.chrom]]) if len(tx) else 0
            for seg in tx:
                bstart.append(base + seg.start)
                bend.append(base + seg.end)
            tx_off.append(len(bstart))
            reverse.append(STRAND_CODE.get(tx.strand, 0))
            lm.append(-1 if mark is None else int(mark))
            chrom.append(layout.index[tx.chrom] if len(tx) else -1)
        return cls(layout, bstart, bend, tx_off, reverse, lm, chrom)

    def device(self, device):
        import torch
        key = _lib.device_key(device)
        if key not in self._dev:
            def up(a):
                return torch.from_numpy(a if len(a) else np.zeros(1, dtype=a.dtype)).to(device)
            self._dev[key] = dict(bstart=up(self.bstart), bend=up(self.bend), bcum=up(self.bcum),
                                  tx_off=up(self.tx_off), reverse=up(self.reverse), landmark=up(self.landmark))
        return self._dev[key]


def landmark_windows(table, flank_upstream, flank_downstream, device="cuda"):
    """``window_landmark`` of every transcript: device tensors ``win`` int64[n_tx, 4]
    (w_start, w_end, w_off, ref_pos) and ``flags`` uint8[n_tx]."""
    import torch
    _lib.require_cuda()
    d = table.device(device)
    n = table.n_tx
    win = torch.zeros((max(n, 1), 4), dtype=torch.int64, device=device)
    flags = torch.zeros(max(n, 1), dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().pb_landmark_windows(_lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["bcum"]),
                                              _lib.ptr(d["tx_off"]), _lib.ptr(d["reverse"]), _lib.ptr(d["landmark"]),
                                              n, int(flank_upstream), int(flank_downstream),
                                              _lib.ptr(win), _lib.ptr(flags), _lib.stream_ptr()))
    return win, flags


def spanning_windows(table, win, flags, grp_off, grp_tx, flank_upstream, flank_downstream, device="cuda"):
    """Maximal spanning window of every group.  ``grp_off`` int64[n_grp+1], ``grp_tx`` int64: indices
    into ``table`` in iteration order.  Returns numpy arrays: ``status`` (PB_SPAN_*), ``offset``,
    ``n_pos``, ``n_blk``, ``refpos`` per group, and the windows' blocks ``out_off`` int64[n_grp+1],
    ``out_bstart`` / ``out_bend`` in global bins, ascending per group."""
    import torch
    _lib.require_cuda()
    d = table.device(device)
    grp_off = np.ascontiguousarray(grp_off, dtype=np.int64)
    grp_tx = np.ascontiguousarray(grp_tx, dtype=np.int64)
    n_grp = len(grp_off) - 1
    if len(grp_tx) and (grp_tx.min() < 0 or grp_tx.max() >= table.n_tx):
        raise IndexError("spanning_windows: transcript index outside the table")
    d_off = torch.from_numpy(grp_off).to(device)
    d_tx = torch.from_numpy(grp_tx if len(grp_tx) else np.zeros(1, dtype=np.int64)).to(device)
    m = max(n_grp, 1)
    status = torch.zeros(m, dtype=torch.uint8, device=device)
    offset = torch.zeros(m, dtype=torch.int32, device=device)
    n_pos = torch.zeros(m, dtype=torch.int32, device=device)
    n_blk = torch.zeros(m, dtype=torch.int32, device=device)
    refpos = torch.zeros(m, dtype=torch.int64, device=device)

    def launch(out_off, out_bstart, out_bend):
        _lib.check(_lib.lib().pb_spanning_windows(
            _lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["bcum"]), _lib.ptr(d["tx_off"]),
            _lib.ptr(d["reverse"]), _lib.ptr(win), _lib.ptr(flags), _lib.ptr(d_off), _lib.ptr(d_tx), n_grp,
            int(flank_upstream), int(flank_downstream), _lib.ptr(status), _lib.ptr(offset), _lib.ptr(n_pos),
            _lib.ptr(n_blk), _lib.ptr(refpos), _lib.ptr(out_off), _lib.ptr(out_bstart), _lib.ptr(out_bend),
            _lib.stream_ptr()))

    launch(None, None, None)                                              # count
    out_off = torch.zeros(m + 1, dtype=torch.int64, device=device)
    out_off[1:] = torch.cumsum(n_blk.to(torch.int64), 0)
    total = int(out_off[-1].item())
    out_bstart = torch.zeros(max(total, 1), dtype=torch.int64, device=device)
    out_bend = torch.zeros(max(total, 1), dtype=torch.int64, device=device)
    if total:
        launch(out_off, out_bstart, out_bend)                             # fill
    return dict(status=status[:n_grp].cpu().numpy(), offset=offset[:n_grp].cpu().numpy(),
                n_pos=n_pos[:n_grp].cpu().numpy(), n_blk=n_blk[:n_grp].cpu().numpy(),
                refpos=refpos[:n_grp].cpu().numpy(), out_off=out_off[:n_grp + 1].cpu().numpy(),
                out_bstart=out_bstart[:total].cpu().numpy(), out_bend=out_bend[:total].cpu().numpy())
