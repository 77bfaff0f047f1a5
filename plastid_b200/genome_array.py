"""Count containers: ``BAMGenomeArray``, ``GenomeArray``, ``SparseGenomeArray`` — device resident.

Drop-in for the hot-path surface of ``plastid/genomics/genome_array.py``:
``BAMGenomeArray`` :582-988 (``set_mapping``, ``add_filter``, ``get``/``__getitem__``,
``get_reads_and_counts``, ``sum``/``set_sum``/``set_normalize``, ``to_genome_array``) and
``GenomeArray`` :1354-1611 / ``SparseGenomeArray`` :2134-2298 (``get``/``__setitem__``/``sum``).

Where the reference maps reads lazily per queried segment (one ``fetch`` + Python loop per exon),
this container lowers the mapping rule ONCE into whole-genome count planes on the GPU
(``pb_map_point`` / ``pb_map_center``); ``ga[seg]`` is then a slice of a plane and region tables
are one gather launch.  Equivalence: every mapped site is one of the read's aligned positions, so
``plane[s:e]`` equals fetch+map over ``[s,e)``.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .batch import AlignmentBatch, GenomeLayout
from .map_factories import (CenterMapFactory, SizeFilterFactory, StratifiedVariableFivePrimeMapFactory,
                            _MapFactory)
from .regions import ChainTable
from .roitools import SegmentChain

_STRANDS = ("+", "-", ".")


# ---------------------------------------------------------------------------------------------
# engine: batch -> planes, planes -> tables
# ---------------------------------------------------------------------------------------------
class CountPlanes(object):
    """Dense per-strand count vectors on the device (uint32 for point rules, float64 for center)."""

    def __init__(self, layout, dtype, device, bin_range=None):
        """``bin_range=(lo, hi)`` (multiples of PB_LAYOUT_ALIGN): a position-sharded rank holds only
        the bins [lo, hi) of every plane; kernels still address global bins, so they are handed the
        address bin 0 would have (``plane_ptr``) and are only ever asked for bins of the range."""
        self.layout, self.dtype, self.device = layout, dtype, device
        self.bin_lo, self.bin_hi = (0, int(layout.total_bins)) if bin_range is None else (int(bin_range[0]), int(bin_range[1]))
        if self.bin_lo % _lib.PB_LAYOUT_ALIGN or self.bin_hi % _lib.PB_LAYOUT_ALIGN or not 0 <= self.bin_lo <= self.bin_hi <= layout.total_bins:
            raise ValueError("bin_range must be multiples of %d inside the layout" % _lib.PB_LAYOUT_ALIGN)
        self.planes = {}
        self.stats = np.zeros(_lib.PB_NSTATS, dtype=np.int64)

    def alloc(self, strands):
        import torch
        tdtype = torch.float64 if self.dtype == "f64" else torch.int32   # int32 storage viewed as uint32
        for s in strands:
            if s not in self.planes:
                self.planes[s] = torch.empty(max(self.bin_hi - self.bin_lo, 1), dtype=tdtype, device=self.device)

    def plane_ptr(self, strand):
        """Address of global bin 0 of a plane (below the allocation for a range-only plane)."""
        t = self.planes.get(strand)
        if t is None:
            return None
        return C.c_void_p(t.data_ptr() - self.bin_lo * t.element_size())

    def plane_ptrs(self):
        arr = (C.c_void_p * 3)()
        for s in _STRANDS:
            p = self.plane_ptr(s)
            arr[_lib.PLANE_INDEX[s]] = None if p is None else p.value
        return arr

    def bins(self, strand, g0, g1):
        """Device view of global bins [g0, g1) of a plane (must lie inside this rank's range)."""
        if g0 < self.bin_lo or g1 > self.bin_hi:
            raise IndexError("bins [%d, %d) are outside this rank's range [%d, %d)" % (g0, g1, self.bin_lo, self.bin_hi))
        return self.planes[strand][g0 - self.bin_lo:g1 - self.bin_lo]

    def slice_host(self, strand, chrom, start, end):
        base = int(self.layout.chrom_bin_off[self.layout.index[chrom]])
        t = self.bins(strand, base + start, base + end).cpu().numpy()
        if self.dtype == "u32":
            return t.view(np.uint32).astype(np.int64)
        return t


_workspaces = {}


def _workspace(device, nbytes, slot=0):
    import torch
    key = (_lib.device_key(device), slot)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def map_batch(dbatch, layout, factory, size_filter=None, strands=("+", "-"), planes=None, sync_stats=True,
              bin_range=None, length_hist=None):
    """Lower ``factory`` over a whole device batch into dense planes for the query ``strands``.
    ``bin_range=(lo, hi)``: produce only the global bins [lo, hi) (position sharding,
    ``plastid_b200.dist.shard_positions``) into range-only planes.
    ``length_hist`` (Center rule; int64[65536], reads per aligned length after the filters): the histogram
    the slot / fixed-point tables are derived from.  Default: ``dbatch.length_hist`` when the batch
    carries it, else measured on the device (``pb_length_hist`` + one device->host round trip).  Ranks of a position-sharded run pass the histogram of the WHOLE batch (all-reduced), so
    that every rank uses the same tables and the planes match the unsharded ones bit for bit."""
    import torch
    _lib.require_cuda()
    if not isinstance(factory, _MapFactory) or isinstance(factory, StratifiedVariableFivePrimeMapFactory):
        raise TypeError("map_batch needs a FivePrime/ThreePrime/VariableFivePrime/Center factory")
    dev = dbatch.device
    is_center = isinstance(factory, CenterMapFactory)
    if planes is None:
        planes = CountPlanes(layout, "f64" if is_center else "u32", dev, bin_range)
    ranged = (planes.bin_lo, planes.bin_hi) != (0, int(layout.total_bins))
    if bin_range is not None and (planes.bin_lo, planes.bin_hi) != tuple(int(x) for x in bin_range):
        raise ValueError("planes were allocated for another bin range")
    planes.alloc(strands)
    mask = 0
    for s in strands:
        mask |= _lib.STRAND_PLANE[s]
    L = _lib.lib()
    ws_bytes = L.pb_map_workspace_bytes(layout.total_bins, dbatch.n_blk, dbatch.n_reads)
    ws = _workspace(dev, ws_bytes)
    stats = torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=dev)
    b, lay, rule = dbatch.c_struct(), layout.c_struct(dev), factory.pb_rule(dev, size_filter)
    outs = [planes.plane_ptr(s) if s in strands else None for s in _STRANDS]
    if ranged and not is_center:
        _lib.check(L.pb_map_point_range(C.byref(b), C.byref(lay), C.byref(rule), mask, outs[0], outs[1], outs[2],
                                        _lib.ptr(stats), _lib.ptr(ws), ws_bytes, planes.bin_lo, planes.bin_hi,
                                        dbatch.n_reads, _lib.stream_ptr()))
    elif is_center:
        if length_hist is None and getattr(dbatch, "length_hist", None) is not None and not os.environ.get("PB_MEASURE_HIST"):
            # the batch came with its histogram (decoder / receiver metadata): apply the size filter to it
            h_hist = _filtered_hist(dbatch.length_hist, size_filter)
        elif length_hist is None:
            h_hist = length_histogram(dbatch, factory, size_filter)
        else:
            h_hist = np.ascontiguousarray(length_hist, dtype=np.int64)
            if h_hist.shape != (65536,):
                raise ValueError("length_hist must have 65536 entries")
        launch = _center_launcher(factory, h_hist, planes, strands, dev)
        launch(b, lay, rule, mask, outs, stats, ws, ws_bytes, planes.bin_lo, planes.bin_hi, 0, -1)
    else:
        _lib.check(L.pb_map_point(C.byref(b), C.byref(lay), C.byref(rule), mask, outs[0], outs[1], outs[2],
                                  _lib.ptr(stats), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    planes.stats_dev = stats
    if sync_stats:
        planes.stats = stats.cpu().numpy()
    return planes


def map_wire16_streamed(receiver, pinned, chunks, layout, factory, size_filter=None, strands=("+", "-"),
                        planes=None, copy_stream=None):
    """Upload a pinned wire16 / delta8 batch (``Wire16Receiver`` / ``Delta8Receiver``) chunk by chunk on ``copy_stream`` and map each chunk's bin range
    (``pb_map_point_range``) on the current stream as soon as its reads have landed, so the PCIe
    transfer and the mapping overlap.  Point rules only.  Returns the planes (stats on device)."""
    import torch
    _lib.require_cuda()
    if not isinstance(factory, _MapFactory) or isinstance(factory, (StratifiedVariableFivePrimeMapFactory, CenterMapFactory)):
        raise TypeError("streamed mapping supports FivePrime/ThreePrime/VariableFivePrime factories")
    dbatch = receiver.batch
    dev = dbatch.device
    if planes is None:
        planes = CountPlanes(layout, "u32", dev)
    planes.alloc(strands)
    mask = 0
    for s in strands:
        mask |= _lib.STRAND_PLANE[s]
    L = _lib.lib()
    ws_bytes = L.pb_map_workspace_bytes(layout.total_bins, 0, dbatch.n_reads)
    # Two compute lanes (streams) take the chunks alternately, each with its own workspace and statistics:
    # the tail of one chunk's persistent kernel overlaps the expansion and the first tiles of the next.
    n_lanes = int(os.environ.get("PB_POINT_LANES", "2")) if len(chunks) > 1 else 1       # (env: A/B aid)
    ws = [_workspace(dev, ws_bytes, slot=j) for j in range(n_lanes)]
    stats = [torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=dev) for _ in range(n_lanes)]
    copy_stream = copy_stream or torch.cuda.Stream(device=dev)
    compute = torch.cuda.current_stream()
    lanes = [compute] + [_side_stream(dev, j) for j in range(1, n_lanes)]
    copy_stream.wait_stream(compute)          # landing buffers may still be read by earlier work
    for ln in lanes[1:]:
        ln.wait_stream(compute)
    with torch.cuda.stream(copy_stream):
        receiver._receive_tables(pinned)
    events = [receiver.receive_chunk(pinned, a, b, copy_stream) for a, b, _x, _y in chunks]
    b_c, lay, rule = dbatch.c_struct(), layout.c_struct(dev), factory.pb_rule(dev, size_filter)
    outs = [planes.plane_ptr(s) if s in strands else None for s in _STRANDS]
    unpacked = None                           # event: the previous chunk's reads are expanded (halo of this chunk)
    for k, ((a, b, bin_a, bin_b), ev) in enumerate(zip(chunks, events)):
        ln = lanes[k % n_lanes]
        with torch.cuda.stream(ln):
            ln.wait_event(ev)
            receiver._unpack(a, b)
            done = torch.cuda.Event()
            done.record(ln)
            if unpacked is not None:
                ln.wait_event(unpacked)
            unpacked = done
            _lib.check(L.pb_map_point_range(C.byref(b_c), C.byref(lay), C.byref(rule), mask, outs[0], outs[1], outs[2],
                                            _lib.ptr(stats[k % n_lanes]), _lib.ptr(ws[k % n_lanes]), ws_bytes, bin_a, bin_b, b,
                                            _lib.stream_ptr()))
    for ln in lanes[1:]:
        compute.wait_stream(ln)
    total = stats[0]
    for st in stats[1:]:
        dropped_len = torch.maximum(total[_lib.PB_STAT_DROPPED_LEN], st[_lib.PB_STAT_DROPPED_LEN])
        total = total + st
        total[_lib.PB_STAT_DROPPED_LEN] = dropped_len
    planes.stats_dev = total
    return planes


def _center_launcher(factory, h_hist, planes, strands, dev):
    """Slot / fixed-point tables of the Center rule for a batch with length histogram ``h_hist`` (uploaded
    once) -> ``launch(batch, layout, rule, mask, outs, stats, ws, ws_bytes, bin_lo, bin_hi, read_begin,
    read_limit)`` enqueueing ``pb_map_center_range`` / ``pb_map_center_fixed_range`` on the current stream."""
    import torch
    L = _lib.lib()
    fixed = factory.fixed_point_tables(h_hist)
    aligned = all(planes.planes[s].data_ptr() % 32 == 0 for s in strands)
    if fixed is not None and aligned and not os.environ.get("PB_CENTER_EXACT"):
        # many map lengths: one pass with 64-bit fixed-point weights (rel. error <= 1e-9, see the header)
        slot_of_len, w_fix, shift = fixed
        d_slot, d_w = torch.from_numpy(slot_of_len).to(dev), torch.from_numpy(w_fix).to(dev)

        def launch(b, lay, rule, mask, outs, stats, ws, ws_bytes, bin_lo, bin_hi, read_begin, read_limit):
            _lib.check(L.pb_map_center_fixed_range(C.byref(b), C.byref(lay), C.byref(rule), mask, _lib.ptr(d_slot),
                                                   _lib.ptr(d_w), len(w_fix), shift, outs[0], outs[1], outs[2],
                                                   _lib.ptr(stats), _lib.ptr(ws), ws_bytes, int(bin_lo), int(bin_hi),
                                                   int(read_begin), int(read_limit), _lib.stream_ptr()))
    else:
        slot_of_len, inv_m = factory.slot_tables(h_hist)
        d_slot = torch.from_numpy(slot_of_len).to(dev)
        d_inv = torch.from_numpy(inv_m if len(inv_m) else np.zeros(1)).to(dev)

        def launch(b, lay, rule, mask, outs, stats, ws, ws_bytes, bin_lo, bin_hi, read_begin, read_limit):
            _lib.check(L.pb_map_center_range(C.byref(b), C.byref(lay), C.byref(rule), mask, _lib.ptr(d_slot), _lib.ptr(d_inv),
                                             len(inv_m), outs[0], outs[1], outs[2], _lib.ptr(stats), _lib.ptr(ws), ws_bytes,
                                             int(bin_lo), int(bin_hi), int(read_begin), int(read_limit), _lib.stream_ptr()))
    return launch


def _filtered_hist(hist, size_filter):
    """Batch-metadata histogram with the size filter applied (what ``pb_length_hist`` would measure)."""
    h = np.ascontiguousarray(hist, dtype=np.int64).copy()
    if size_filter is not None:
        lens_ = np.arange(65536)
        keep = lens_ >= size_filter.min_
        if size_filter.max_ != -1:
            keep &= lens_ <= size_filter.max_
        h[~keep] = 0
    return h


def map_center_streamed(receiver, pinned, chunks, layout, factory, size_filter=None, strands=("+", "-"),
                        planes=None, copy_stream=None, length_hist=None):
    """Center rule over a batch WITH multi-block reads that is still being uploaded
    (``Delta3SplicedReceiver``): chunk by chunk on ``copy_stream``; as soon as a chunk's reads and block
    words have landed and are expanded, the bins below the next chunk's first read are final and are
    produced with ``pb_map_center_range`` from the reads that can reach them (read window = from one
    128-read block before the first read within ``max_span`` of the range).  Two compute lanes take the
    chunks alternately.  Same result, bit for bit, as :func:`map_batch` on the whole batch.
    ``length_hist``: histogram to derive the slot tables from (already filtered; a position-sharded rank passes the
    whole batch's); default: the receiver batch's own."""
    import torch
    _lib.require_cuda()
    if not isinstance(factory, CenterMapFactory):
        raise TypeError("map_center_streamed is the Center rule's streamed path")
    dbatch = receiver.batch
    if dbatch.length_hist is None:
        raise ValueError("the transfer format must carry the batch's length histogram")
    dev = dbatch.device
    if planes is None:
        planes = CountPlanes(layout, "f64", dev)
    planes.alloc(strands)
    mask = 0
    for s in strands:
        mask |= _lib.STRAND_PLANE[s]
    hist = _filtered_hist(dbatch.length_hist, size_filter) if length_hist is None else np.ascontiguousarray(length_hist, dtype=np.int64)
    launch = _center_launcher(factory, hist, planes, strands, dev)
    L = _lib.lib()
    ws_bytes = L.pb_map_workspace_bytes(layout.total_bins, dbatch.n_blk, dbatch.n_reads)
    n_lanes = int(os.environ.get("PB_CENTER_LANES", "2")) if len(chunks) > 1 else 1      # (A/B aid)
    ws = [_workspace(dev, ws_bytes, slot=j) for j in range(n_lanes)]
    stats = [torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=dev) for _ in range(n_lanes)]
    copy_stream = copy_stream or torch.cuda.Stream(device=dev)
    compute = torch.cuda.current_stream()
    lanes = [compute] + [_side_stream(dev, j) for j in range(1, n_lanes)]
    copy_stream.wait_stream(compute)
    for ln in lanes[1:]:
        ln.wait_stream(compute)
    with torch.cuda.stream(copy_stream):
        receiver.receive_tables(pinned)
    events = [receiver.receive_chunk(pinned, a, b, copy_stream) for a, b, _x, _y in chunks]
    b_c, lay, rule = dbatch.c_struct(), layout.c_struct(dev), factory.pb_rule(dev, size_filter)
    outs = [planes.plane_ptr(s) if s in strands else None for s in _STRANDS]
    unpacked = None
    for k, ((a, b, bin_a, bin_b), ev) in enumerate(zip(chunks, events)):
        ln = lanes[k % n_lanes]
        with torch.cuda.stream(ln):
            ln.wait_event(ev)
            if unpacked is not None:
                # expansions run one after the other (they share the receiver's scan workspace), and the read
                # window of this chunk reaches back into the previous one
                ln.wait_event(unpacked)
            receiver.unpack_chunk(a, b)
            unpacked = torch.cuda.Event()
            unpacked.record(ln)
            launch(b_c, lay, rule, mask, outs, stats[k % n_lanes], ws[k % n_lanes], ws_bytes, bin_a, bin_b,
                   receiver.first_read_reaching(bin_a, layout), b)
    for ln in lanes[1:]:
        compute.wait_stream(ln)
    total = stats[0]
    for st in stats[1:]:
        dropped_len = torch.maximum(total[_lib.PB_STAT_DROPPED_LEN], st[_lib.PB_STAT_DROPPED_LEN])
        total = total + st
        total[_lib.PB_STAT_DROPPED_LEN] = dropped_len
    planes.stats_dev = total
    return planes


def length_histogram(dbatch, factory, size_filter=None):
    """Reads per aligned length (int64[65536]) after the drop bit and the size filter (``pb_length_hist``)."""
    import torch
    _lib.require_cuda()
    hist = torch.zeros(65536, dtype=torch.int64, device=dbatch.device)
    b, rule = dbatch.c_struct(), factory.pb_rule(dbatch.device, size_filter)
    _lib.check(_lib.lib().pb_length_hist(C.byref(b), C.byref(rule), _lib.PB_PLANE_ANY, _lib.ptr(hist), _lib.stream_ptr()))
    return hist.cpu().numpy()


_side_streams = {}


def _side_stream(device, j):
    import torch
    key = (_lib.device_key(device), j)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def region_sums(planes, table, out=None):
    """Masked sums + unmasked lengths for every chain of ``table`` (device tensors returned)."""
    import torch
    _lib.require_cuda()
    dev = planes.device
    d = table.device(dev)
    n = table.n_chains
    sums = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    live = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ptrs = planes.plane_ptrs()
    # positions outside this rank's bins (position sharding; parts of a region beyond its chromosome) count zero
    L = _lib.lib()
    n_blocks = len(table.bstart)
    ranged = (planes.bin_lo, planes.bin_hi) != (0, int(planes.layout.total_bins))
    if ranged and n_blocks:
        # a rank looks at the blocks that overlap its bins only; unmasked lengths are geometry (the same on every rank)
        d = table.owned(dev, planes.bin_lo, planes.bin_hi)
        n_blocks = d["n_blocks"]
        if n_blocks == 0:                       # no block of the table reaches this rank's bins
            return sums.zero_()[:n], d["live"][:n]
    ws_bytes = L.pb_region_sums_workspace_bytes(n_blocks)
    ws = _workspace(dev, ws_bytes, slot="region_sums")
    _lib.check(L.pb_region_sums(ptrs, 1 if planes.dtype == "f64" else 0, _lib.ptr(d["bstart"]),
                                _lib.ptr(d["bend"]), _lib.ptr(d["chain_off"]), _lib.ptr(d["chain_plane"]),
                                _lib.ptr(d["block_chain"]), _lib.ptr(d["block_pos"]), _lib.ptr(d["block_plane"]),
                                n, n_blocks, _lib.ptr(d["mask_bits"]), _lib.ptr(d["mask_off"]),
                                planes.bin_lo, planes.bin_hi, _lib.ptr(sums), _lib.ptr(live), _lib.ptr(ws), ws_bytes,
                                _lib.stream_ptr()))
    if ranged and "live" in d:
        live = d["live"]
    return sums[:n], live[:n]


def chain_counts(dbatch, layout, factory, size_filter, table, bin_range=None, stats=None):
    """Plane-free region counts of a point rule (``pb_chain_counts``): ``(sums float64[n], unmasked lengths
    int64[n])`` straight from the sorted device batch — what :func:`region_sums` gives over the planes
    :func:`map_batch` would write, without the 4 bytes per genome position of the planes.  ``bin_range``: the
    global bins this rank owns (sites elsewhere count zero).  ``stats``: device int64[PB_NSTATS] or None."""
    import torch
    _lib.require_cuda()
    if not isinstance(factory, _MapFactory) or isinstance(factory, (StratifiedVariableFivePrimeMapFactory, CenterMapFactory)):
        raise TypeError("chain_counts needs a FivePrime/ThreePrime/VariableFivePrime factory")
    dev = dbatch.device
    d = table.device(dev)
    n = table.n_chains
    sums = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    live = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    lo, hi = (0, int(layout.total_bins)) if bin_range is None else (int(bin_range[0]), int(bin_range[1]))
    b, lay, rule = dbatch.c_struct(), layout.c_struct(dev), factory.pb_rule(dev, size_filter)
    L = _lib.lib()
    n_blocks = len(table.bstart)
    ws_bytes = L.pb_chain_counts_workspace_bytes(int(layout.total_bins), n_blocks, n)
    ws = _workspace(dev, ws_bytes, slot="chain_counts")
    _lib.check(L.pb_chain_counts(C.byref(b), C.byref(lay), C.byref(rule), _lib.ptr(d["bstart"]), _lib.ptr(d["bend"]),
                                 _lib.ptr(d["chain_off"]), _lib.ptr(d["chain_plane"]), _lib.ptr(d["block_chain"]),
                                 _lib.ptr(d["block_pos"]), _lib.ptr(d["block_plane"]), n, n_blocks, _lib.ptr(d["mask_bits"]),
                                 _lib.ptr(d["mask_off"]),
                                 lo, hi, _lib.ptr(sums), _lib.ptr(live), _lib.ptr(stats), _lib.ptr(ws), ws_bytes,
                                 _lib.stream_ptr()))
    return sums[:n], live[:n]


class GraphedCount(object):
    """One counting pass — ``map_batch`` of a point rule + ``region_sums`` — captured as a CUDA graph.

    Small genomes (BASELINE config 1: 12 Mb, 2 M reads, 6 k regions) are launch-latency bound: the
    pass is five short kernels and two memsets.  Captured once, it replays as a single graph launch;
    inputs (the device batch, the chain table) and outputs (planes, ``sums``, ``live``) are the
    tensors seen at capture time."""

    def __init__(self, dbatch, layout, factory, size_filter, table, strands=("+", "-")):
        import torch
        _lib.require_cuda()
        if isinstance(factory, CenterMapFactory):
            raise TypeError("the Center pass sizes its slot tables on the host and cannot be captured")
        self.planes = map_batch(dbatch, layout, factory, size_filter, strands=strands, sync_stats=False)   # warm-up: caches LUTs, workspace
        region_sums(self.planes, table)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            map_batch(dbatch, layout, factory, size_filter, strands=strands, planes=self.planes, sync_stats=False)
            self.sums, self.live = region_sums(self.planes, table)
        self.stats = self.planes.stats_dev

    def replay(self):
        self.graph.replay()
        return self.sums, self.live


def gather_windows(planes, table, row_col, width, touched_only=False):
    """Window matrix (n_chains x width, NaN-filled) + mask matrix, rows laid 5'->3'.
    ``touched_only`` (position-sharded planes): only the rows with a position in this rank's bins are produced — whole,
    positions of other ranks as zeros — and the other rows are left unwritten (``dist.window_profile`` completes rows rank by
    rank and never looks at them); the rank then walks its own chains' blocks instead of the whole table."""
    import torch
    _lib.require_cuda()
    dev = planes.device
    d = table.device(dev)
    n = table.n_chains
    n_blocks = len(table.bstart)
    if touched_only and (planes.bin_lo, planes.bin_hi) != (0, int(planes.layout.total_bins)) and n_blocks:
        sub = table.owned(dev, planes.bin_lo, planes.bin_hi, whole_chains=True)
        if sub["n_blocks"]:                     # (a rank that touches no row walks the whole table: nothing is its own)
            d, n_blocks = sub, sub["n_blocks"]
    matrix = torch.empty((max(n, 1), width), dtype=torch.float64, device=dev)
    maskmat = torch.empty((max(n, 1), width), dtype=torch.uint8, device=dev)
    cols = row_col if hasattr(row_col, "data_ptr") else torch.from_numpy(np.ascontiguousarray(row_col, dtype=np.int32)).to(dev)
    _lib.check(_lib.lib().pb_gather_windows(planes.plane_ptrs(), 1 if planes.dtype == "f64" else 0,
                                            _lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["chain_off"]),
                                            _lib.ptr(d["chain_plane"]), _lib.ptr(d["chain_reverse"]),
                                            _lib.ptr(d["block_chain"]), _lib.ptr(d["block_pos"]), _lib.ptr(d["block_plane"]),
                                            _lib.ptr(d["chain_len"]), _lib.ptr(cols), n, n_blocks, width,
                                            _lib.ptr(d["mask_bits"]), _lib.ptr(d["mask_off"]), planes.bin_lo, planes.bin_hi,
                                            _lib.ptr(matrix), _lib.ptr(maskmat), _lib.stream_ptr()))
    return matrix[:n], maskmat[:n]


def gather_chains(planes, table):
    """Count vectors of every chain of ``table`` in one launch, ragged: ``(values float64[total], masked uint8[total],
    row_off int64[n + 1])`` — chain ``c`` owns cells ``[row_off[c], row_off[c + 1])``, laid 5'->3'."""
    import torch
    _lib.require_cuda()
    dev = planes.device
    d = table.device(dev)
    n = table.n_chains
    row_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(table.chain_len, out=row_off[1:])
    total = int(row_off[-1])
    values = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
    masked = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(row_off).to(dev)
    _lib.check(_lib.lib().pb_gather_chains(planes.plane_ptrs(), 1 if planes.dtype == "f64" else 0,
                                           _lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["chain_off"]),
                                           _lib.ptr(d["chain_plane"]), _lib.ptr(d["chain_reverse"]),
                                           _lib.ptr(d["block_chain"]), _lib.ptr(d["block_pos"]), _lib.ptr(d["block_plane"]),
                                           _lib.ptr(d["chain_len"]), _lib.ptr(d_off), n, len(table.bstart),
                                           _lib.ptr(d["mask_bits"]), _lib.ptr(d["mask_off"]), planes.bin_lo, planes.bin_hi,
                                           _lib.ptr(values), _lib.ptr(masked), _lib.stream_ptr()))
    return values[:total], masked[:total], row_off


def stratified_windows(dbatch, layout, factory, size_filter, table, row_col, width, min_len, max_len, phase=None,
                       bin_range=None):
    """Per-read-length window matrices in one launch: (int32[n_len, n_chains, width], uint8 mask
    [n_chains, width]) — the inner loops of psite.py:176-199 / phase_by_size.py:186-194.
    ``phase=(codon_front, codon_back)``: per-length sub-codon phase counts instead (width 3).
    ``bin_range``: the global bins this rank owns — sites elsewhere are not counted (position sharding)."""
    import torch
    _lib.require_cuda()
    dev = dbatch.device
    d = table.device(dev)
    n = table.n_chains
    n_len = max_len - min_len + 1
    if phase is not None:
        width, row_col = 3, np.zeros(n, dtype=np.int32)
    alloc = torch.empty if n else torch.zeros          # the kernel writes every cell of every chain
    out = alloc((n_len, max(n, 1), width), dtype=torch.int32, device=dev)
    maskmat = torch.empty((max(n, 1), width), dtype=torch.uint8, device=dev)
    cols = torch.from_numpy(np.ascontiguousarray(row_col, dtype=np.int32)).to(dev)
    b, lay, rule = dbatch.c_struct(), layout.c_struct(dev), factory.pb_rule(dev, size_filter)
    lo, hi = (0, int(layout.total_bins)) if bin_range is None else (int(bin_range[0]), int(bin_range[1]))
    L = _lib.lib()
    n_blocks = len(table.bstart)
    ws_bytes = L.pb_stratified_windows_workspace_bytes(n_blocks)
    ws = _workspace(dev, ws_bytes, slot="stratified")            # read slice of every exon block, found by one launch
    _lib.check(L.pb_stratified_windows_ws(C.byref(b), C.byref(lay), C.byref(rule), int(min_len), int(max_len),
                                          _lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["chain_off"]),
                                          _lib.ptr(d["chain_plane"]), _lib.ptr(d["chain_reverse"]), _lib.ptr(cols),
                                          n, n_blocks, width, int(phase is not None), int(phase[0]) if phase else 0,
                                          int(phase[1]) if phase else 0, _lib.ptr(d["mask_bits"]),
                                          _lib.ptr(d["mask_off"]), lo, hi, _lib.ptr(out), _lib.ptr(maskmat),
                                          _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    return out[:, :n], maskmat[:n]


def phase_sums(planes, table, codon_front, codon_back):
    """Per-chain sub-codon phase sums (n_chains x 3, device tensor)."""
    import torch
    _lib.require_cuda()
    dev = planes.device
    d = table.device(dev)
    n = table.n_chains
    out = torch.zeros((max(n, 1), 3), dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().pb_phase_sums_range(planes.plane_ptrs(), 1 if planes.dtype == "f64" else 0,
                                              _lib.ptr(d["bstart"]), _lib.ptr(d["bend"]), _lib.ptr(d["chain_off"]),
                                              _lib.ptr(d["chain_plane"]), _lib.ptr(d["chain_reverse"]), n,
                                              int(codon_front), int(codon_back), planes.bin_lo, planes.bin_hi,
                                              _lib.ptr(out), _lib.stream_ptr()))
    return out[:n]


def window_normalize(matrix, maskmat, norm_lo, norm_hi, min_counts, want_norm=True):
    import torch
    n, width = matrix.shape
    dev = matrix.device
    denom = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    sel = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    norm = torch.empty_like(matrix) if want_norm else None
    nmask = torch.empty_like(maskmat) if want_norm else None
    _lib.check(_lib.lib().pb_window_normalize(_lib.ptr(matrix), _lib.ptr(maskmat), n, width, int(norm_lo),
                                              int(norm_hi), float(min_counts), _lib.ptr(denom), _lib.ptr(sel),
                                              _lib.ptr(norm), _lib.ptr(nmask), _lib.stream_ptr()))
    return denom[:n], sel[:n], norm, nmask


def column_profile(values, valmask, row_select, mode="median", n_batch=1):
    """Per-column median / mean / sum over unmasked cells of selected rows.  ``n_batch`` > 1: ``values``
    stacks that many equally shaped matrices row-wise (psite: one per read length); the results come
    back as ``[n_batch, width]`` from one launch."""
    import torch
    n_all, width = values.shape
    if n_all % n_batch:
        raise ValueError("rows must divide evenly over the batch")
    n = n_all // n_batch
    dev = values.device
    L = _lib.lib()
    ws_bytes = n_batch * L.pb_column_profile_workspace_bytes(n, width)
    ws = _scratch(dev, int(ws_bytes))
    profile = torch.empty(n_batch * width, dtype=torch.float64, device=dev)
    n_regions = torch.empty(n_batch * width, dtype=torch.int64, device=dev)
    col_sum = torch.empty(n_batch * width, dtype=torch.float64, device=dev)
    _lib.check(L.pb_column_profile_batched(_lib.ptr(values), _lib.ptr(valmask), _lib.ptr(row_select), n_batch, n, width,
                                           {"median": 0, "mean": 1, "sum": 2}[mode], _lib.ptr(profile),
                                           _lib.ptr(n_regions), _lib.ptr(col_sum), _lib.ptr(ws), ws_bytes,
                                           _lib.stream_ptr()))
    if n_batch == 1:
        return profile, n_regions, col_sum
    return profile.view(n_batch, width), n_regions.view(n_batch, width), col_sum.view(n_batch, width)


def count_profiles(strat, maskmat, norm_lo, norm_hi, min_counts, mode="median"):
    """Normalise + per-column median / mean of the integer window matrices ``strat``
    (uint32 ``[n_batch, n_rows, width]``, e.g. from :func:`stratified_windows`) without materialising
    float64 / normalised / mask matrices: ``(profile [n_batch, width], n_regions, row_select [n_batch, n_rows])``.
    ``maskmat`` ``[n_rows, width]`` is the position mask shared by all matrices."""
    import torch
    strat = strat.contiguous()
    maskmat = maskmat.contiguous()
    n_batch, n, width = strat.shape
    dev = strat.device
    L = _lib.lib()
    ws_bytes = n_batch * L.pb_column_profile_workspace_bytes(n, width)
    ws = _scratch(dev, int(ws_bytes))
    sel = torch.empty(n_batch * max(n, 1), dtype=torch.uint8, device=dev)
    profile = torch.empty(n_batch * width, dtype=torch.float64, device=dev)
    n_regions = torch.empty(n_batch * width, dtype=torch.int64, device=dev)
    col_sum = torch.empty(n_batch * width, dtype=torch.float64, device=dev)
    _lib.check(L.pb_count_profiles_u32(_lib.ptr(strat), _lib.ptr(maskmat), 1, n_batch, n, width, int(norm_lo), int(norm_hi),
                                       float(min_counts), {"median": 0, "mean": 1}[mode], _lib.ptr(sel), _lib.ptr(profile),
                                       _lib.ptr(n_regions), _lib.ptr(col_sum), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    return profile.view(n_batch, width), n_regions.view(n_batch, width), sel[:n_batch * n].view(n_batch, n)


_scratch_bufs = {}


def _scratch(device, nbytes):
    """Reusable device scratch (radix-select keys): grown, never shrunk, one per device."""
    import torch
    key = _lib.device_key(device)
    buf = _scratch_bufs.get(key)
    if buf is None or buf.numel() < nbytes:
        _scratch_bufs[key] = buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
    return buf


def merge_batches(batches):
    """One coordinate-sorted batch from several (the reference chains ``fetch`` over all files,
    genome_array.py:800-809; chromosome lengths take the max over files, :667-672)."""
    if len(batches) == 1:
        return batches[0]
    chroms, lens = [], {}
    for b in batches:
        for c, n in zip(b.chroms, b.chrom_len):
            if c not in lens:
                chroms.append(c)
            lens[c] = max(lens.get(c, 0), int(n))
    index = {c: i for i, c in enumerate(chroms)}
    cid, start, meta, nblk_rows = [], [], [], []
    blk_all, any_blk = [], any(b.blk is not None for b in batches)
    for b in batches:
        per_read_chrom = np.repeat(np.arange(len(b.chroms)), np.diff(b.chrom_read_off))
        cid.append(np.asarray([index[c] for c in b.chroms], dtype=np.int64)[per_read_chrom])
        start.append(b.ref_start)
        meta.append(b.meta)
        if any_blk:
            if b.blk_off is None:
                nblk_rows.append(np.zeros(len(b), dtype=np.int64))
            else:
                nblk_rows.append(np.diff(b.blk_off.astype(np.int64)))
                blk_all.append(b.blk)
    cid, start, meta = np.concatenate(cid), np.concatenate(start), np.concatenate(meta)
    # every input is sorted already: a stable sort of one combined key only has to merge a few runs (ties keep
    # the order of the files, as the reference's chained fetch does)
    order = np.argsort((cid << 32) + start.astype(np.int64), kind="stable")
    blk_off = blk = None
    if any_blk:
        rows = np.concatenate(nblk_rows)
        src_off = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(rows, out=src_off[1:])
        blk_src = np.concatenate(blk_all) if blk_all else np.zeros((0, 2), dtype=np.int32)
        blk_off = np.zeros(len(order) + 1, dtype=np.int64)
        np.cumsum(rows[order], out=blk_off[1:])
        # row k of destination read d comes from row src_off[order[d]] + k: one gather for all reads
        take = np.repeat(src_off[:-1][order] - blk_off[:-1], rows[order]) + np.arange(int(blk_off[-1]), dtype=np.int64)
        blk = np.ascontiguousarray(blk_src[take], dtype=np.int32).reshape(-1, 2)
    counts = np.bincount(cid, minlength=len(chroms))
    off = np.zeros(len(chroms) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    out = AlignmentBatch(chroms, [lens[c] for c in chroms], start[order], meta[order], off, blk_off, blk,
                         mapped=sum(b.mapped for b in batches))
    if all(b.objects is not None for b in batches):
        objs = [o for b in batches for o in b.objects]
        out.objects = [objs[i] for i in order]
    return out


def _resolve_shard(shard):
    """``"auto"``: (rank, world) of the initialised ``torch.distributed`` process group, else (0, 1);
    ``None``: never shard; ``(rank, world)``: as given (no collective is issued unless a group is initialised)."""
    if shard is None:
        return 0, 1
    if shard == "auto":
        from . import dist as pdist
        return pdist.world()
    rank, world = shard
    if not 0 <= int(rank) < int(world):
        raise ValueError("shard=(rank, world) needs 0 <= rank < world")
    return int(rank), int(world)


# ---------------------------------------------------------------------------------------------
# BAMGenomeArray
# ---------------------------------------------------------------------------------------------
class BAMGenomeArray(object):
    """``BAMGenomeArray(*sources, mapping=CenterMapFactory())``.

    ``sources`` are sorted BAM paths (decoded once, whole, by the library's own BGZF/BAM decoder — no index, no
    pysam needed), open ``pysam.AlignmentFile`` handles, or :class:`~plastid_b200.batch.AlignmentBatch` objects.
    ``indexed=True`` (paths with a ``.bai`` beside them; ``bam_io.build_index`` writes one): nothing is decoded up
    front — chromosomes and ``sum()`` come from the header and the index statistics, ``ga[seg]`` /
    ``get_reads_and_counts`` seek through the index like the reference's ``fetch`` (genome_array.py:800-809), and the
    files are decoded whole only when a whole-genome consumer (``count_planes``, ``count_chains``, track export) asks.

    How the alignments reach the device: a batch that carries its transfer format (``AlignmentBatch.pack`` — the
    decoder emits it) is uploaded in that form (1-1.3 bytes per read, 4 per aligned block) chunk by chunk on a copy
    stream and expanded on the device; whole-genome planes are mapped range by range while later chunks are still
    landing (``map_wire16_streamed`` / ``map_center_streamed``).  A plain SoA batch is uploaded as it is.

    More than one GPU (``torch.distributed`` initialised, one process per GPU; or ``shard=(rank, world)``): every
    rank owns a contiguous range of the concatenated genome (cuts of equal cost: reads + plane bins; ``sharding="chromosomes"``
    puts them on chromosome boundaries), keeps the reads that can map into it (its own plus a halo of ``max_span``)
    and the planes of that range only.  Region tables, window matrices and ``ga[seg]`` vectors are completed with
    one all-reduce (every position is owned by exactly one rank); no count vector ever crosses NVLink.
    ``bin_range=(lo, hi)``: the sources are already this rank's shard of such a partition.
    """

    def __init__(self, *sources, **kwargs):
        if len(sources) == 1 and isinstance(sources[0], (list, tuple)):
            sources = tuple(sources[0])
        if not sources:
            raise ValueError("BAMGenomeArray needs at least one alignment source")
        self.device = kwargs.get("device", "cuda")
        self.map_fn = kwargs.get("mapping", None) or CenterMapFactory()
        self._strands = _STRANDS
        self._normalize = False
        self._filters = {}
        self._planes = None
        self._dbatch = None
        self._full_dbatch = None
        self._receiver = None
        self._host_batch = None           # this rank's reads with the generic filters' verdicts (evaluated once)
        self._kwargs = kwargs
        self._indexed = None
        if kwargs.get("indexed"):
            # header + index only: nothing is decoded until a whole-genome consumer asks (`_attach`)
            from .bam_io import IndexedBam
            if not all(isinstance(src, str) for src in sources):
                raise TypeError("indexed=True takes paths of sorted, indexed BAM files")
            self._indexed = [IndexedBam(src) for src in sources]
            self._sources = sources
            chroms, lens = [], {}
            for f in self._indexed:                   # lengths take the max over files (genome_array.py:667-672)
                for c, n in zip(f.references, f.lengths):
                    if c not in lens:
                        chroms.append(c)
                    lens[c] = max(lens.get(c, 0), int(n))
            self._set_genome(chroms, [lens[c] for c in chroms])
            self._rank, self._world = _resolve_shard(kwargs.get("shard", "auto"))
            self._collective = False
            self._bin_range = (0, int(self.layout.total_bins))
            if self._world > 1 or any(f.mapped is None for f in self._indexed):
                self._attach_sources()               # sharded arrays and indexes without statistics: decode now
            else:
                self.reset_sum()
            return
        self._sources = sources
        self._attach_sources()

    def _set_genome(self, chroms, chrom_len):
        self._chr_lengths = {c: int(n) for c, n in zip(chroms, chrom_len)}
        self._chroms = sorted(self._chr_lengths)
        self.layout = GenomeLayout(chroms, chrom_len)

    def __getattr__(self, name):
        # an indexed array decodes its files the first time something needs every read
        if name in ("batch", "batches", "_local", "_global_hist") and self.__dict__.get("_indexed") is not None \
                and "batch" not in self.__dict__:
            self._attach_sources()
            return self.__dict__[name]
        raise AttributeError(name)

    @property
    def is_lazy(self):
        """True while an ``indexed=True`` array has not decoded its files (region queries seek through the index)."""
        return self._indexed is not None and "batch" not in self.__dict__

    def _attach_sources(self):
        kwargs = self._kwargs
        batches = []
        for src in self._sources:
            if isinstance(src, AlignmentBatch):
                batches.append(src)
            else:
                from .bam_io import batch_from_bam
                batches.append(batch_from_bam(src))
        self.batches = batches
        self.batch = merge_batches(batches)
        if len(batches) > 1 and all(b.transfer is not None for b in batches):
            self.batch.pack()
        self._set_genome(self.batch.chroms, self.batch.chrom_len)
        # Center rule on a sharded array: the aligned-length histogram of the WHOLE batch (every rank derives the same
        # slot tables from it); known here when this object shards the sources itself, handed in with pre-sharded ones
        self._global_hist = kwargs.get("length_hist")
        # -- position-range ownership ------------------------------------------------------------
        self._rank, self._world = _resolve_shard(kwargs.get("shard", "auto"))
        self._collective = self._world > 1
        self._local = self.batch                      # the reads this rank maps
        self._bin_range = (0, int(self.layout.total_bins))
        if kwargs.get("bin_range") is not None:       # pre-sharded sources
            self._bin_range = (int(kwargs["bin_range"][0]), int(kwargs["bin_range"][1]))
            from . import dist as pdist
            self._collective = pdist.is_distributed() and pdist.world()[1] > 1
        elif self._world > 1:
            from . import dist as pdist
            snap = {"positions": "bins", "chromosomes": "chromosomes"}[kwargs.get("sharding", "positions")]
            sub, lo, hi = pdist.shard_positions(self.batch, self.layout, self._rank, self._world, snap=snap)
            if self.batch.transfer is not None:
                sub.pack()
            self._local, self._bin_range = sub, (lo, hi)
            if self._global_hist is None:
                from .batch import meta_length_hist
                self._global_hist = meta_length_hist(self.batch.meta)     # every rank decoded the whole file
        kept_sum = self.__dict__.get("_sum")          # an indexed array's sum (index statistic or `set_sum`) stays
        self._update()
        if kept_sum is not None:
            self._sum = kept_sum

    # -- bookkeeping (genome_array.py:681-758, 930-963) ----------------------------------------
    def reset_sum(self):
        if self.is_lazy:                              # `bamfile.mapped`: the index statistic (genome_array.py:690)
            self._sum = sum(f.mapped for f in self._indexed)
        else:
            self._sum = sum(b.mapped for b in self.batches)

    def _update(self):
        self.reset_sum()
        self._planes = None

    def sum(self):
        return self._sum

    def set_sum(self, val):
        self._sum = val

    def set_normalize(self, value=True):
        assert value in (True, False)
        self._normalize = value

    def add_filter(self, name, func):
        self._filters[name] = func
        self._planes = None
        if not isinstance(func, SizeFilterFactory):      # size filters are lowered into the kernels; others re-flag reads
            self._dbatch = self._full_dbatch = self._receiver = self._host_batch = None

    def remove_filter(self, name):
        retval = self._filters.pop(name)
        self._planes = None
        if not isinstance(retval, SizeFilterFactory):
            self._dbatch = self._full_dbatch = self._receiver = self._host_batch = None
        return retval

    def chroms(self):
        return self._chroms

    def lengths(self):
        return self._chr_lengths

    def strands(self):
        return self._strands

    def get_mapping(self):
        return self.map_fn.__doc__

    def set_mapping(self, mapping_function):
        self.map_fn = mapping_function
        self._update()

    @property
    def bin_range(self):
        """Global bins ``(lo, hi)`` this rank owns (the whole layout on one GPU)."""
        return self._bin_range

    @property
    def is_sharded(self):
        return self._bin_range != (0, int(self.layout.total_bins))

    # -- lowering ------------------------------------------------------------------------------
    def _size_filter(self):
        """The lowered size filter: intersection of all SizeFilterFactory filters."""
        smin, smax, have = 1, -1, False
        for f in self._filters.values():
            if isinstance(f, SizeFilterFactory):
                have = True
                smin = max(smin, f.min_)
                if f.max_ != -1:
                    smax = f.max_ if smax == -1 else min(smax, f.max_)
        if not have:
            return None
        if smax != -1 and smax < smin:
            smax = 0          # empty interval: nothing passes
            sf = SizeFilterFactory.__new__(SizeFilterFactory)
            sf.min_, sf.max_ = smin, smax
            return sf
        return SizeFilterFactory(smin, smax)

    def _filtered(self, hb):
        """``hb`` with the verdicts of the generic (non-size) filters in the drop bit.  Filters that are not
        SizeFilterFactory objects are arbitrary python predicates over reads (genome_array.py:819-820): they cannot
        be lowered and are evaluated once per read on the host — unless they implement ``batch_mask`` (below)."""
        generic = [f for f in self._filters.values() if not isinstance(f, SizeFilterFactory)]
        if not generic:
            return hb
        keep = np.ones(len(hb), dtype=bool)
        # a filter may offer `batch_mask(batch) -> bool[n_reads]` (True = keep) and is then asked once for the whole
        # batch (numpy over ref_start / aligned_len / is_reverse) instead of once per read
        slow = []
        for f in generic:
            if hasattr(f, "batch_mask"):
                keep &= np.asarray(f.batch_mask(hb), dtype=bool)
            else:
                slow.append(f)
        if slow:
            for i in np.nonzero(keep)[0]:
                read = hb.read_view(int(i))
                keep[i] = all(f(read) for f in slow)
        out = hb.with_drop_mask(~keep)
        if hb.transfer is not None:
            out.pack()
        return out

    def _local_filtered(self):
        if self._host_batch is None:
            self._host_batch = self._filtered(self._local)
        return self._host_batch

    def _make_receiver(self, hb):
        from .batch import Delta3Receiver, Delta3SplicedBatch, Delta3SplicedReceiver
        cls = Delta3SplicedReceiver if isinstance(hb.transfer, Delta3SplicedBatch) else Delta3Receiver
        return cls(hb.transfer, self.device)

    def _device_batch(self):
        """This rank's reads on the device (uploaded once): the transfer format when the batch carries it — expanded
        on the device — else the SoA."""
        if self._dbatch is None:
            hb = self._local_filtered()
            if hb.transfer is not None and len(hb):
                self._receiver = self._make_receiver(hb)
                self._dbatch = self._receiver.receive(hb.transfer_pinned())
            else:
                self._dbatch = hb.to_device(self.device)
        return self._dbatch

    def _whole_device_batch(self):
        """Every read of the sources on this device: what the per-segment operator (``get_reads_and_counts``) reads.
        The same object as :meth:`_device_batch` unless the array is sharded."""
        if not self.is_sharded:
            self._whole_host = None
            return self._device_batch()
        if self._full_dbatch is None:
            self._whole_host = self._filtered(self.batch)
            self._full_dbatch = self._whole_host.to_device(self.device)
        return self._full_dbatch

    def _is_lowerable(self):
        return isinstance(self.map_fn, _MapFactory) and not isinstance(self.map_fn, StratifiedVariableFivePrimeMapFactory)

    def _center_hist(self):
        """Aligned-length histogram the Center rule derives its tables from: of the WHOLE batch on every rank
        (one 512 KB all-reduce), so that sharded planes equal unsharded ones bit for bit."""
        if not self.is_sharded:
            return None
        from .batch import meta_length_hist
        generic = any(not isinstance(f, SizeFilterFactory) for f in self._filters.values())
        if self._global_hist is not None and not generic:
            h = self._global_hist
        elif self._collective:                      # python filters re-flag reads: count what is left, rank by rank
            from . import dist as pdist
            h = pdist.global_length_hist(self._host_batch, self.layout, *self._bin_range)
        else:
            h = meta_length_hist(self._host_batch.meta)
        return _filtered_hist(h, self._size_filter())

    def count_planes(self, strands=("+", "-")):
        """Planes of this rank's bin range for the current mapping rule and filters (computed once, cached)."""
        if not self._is_lowerable():
            raise TypeError("mapping function %r cannot be lowered to whole-genome planes" % (self.map_fn,))
        need = tuple(s for s in strands if self._planes is None or s not in self._planes.planes)
        if not need:
            return self._planes
        is_center = isinstance(self.map_fn, CenterMapFactory)
        if self._planes is None:
            self._planes = CountPlanes(self.layout, "f64" if is_center else "u32", self.device,
                                       self._bin_range if self.is_sharded else None)
        sf = self._size_filter()
        if self._bin_range[0] == self._bin_range[1]:          # a rank that owns no bins (more ranks than chromosomes)
            import torch
            self._planes.alloc(need)
            self._planes.stats_dev = torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=self.device)
            self._planes.stats = np.zeros(_lib.PB_NSTATS, dtype=np.int64)
            return self._planes
        if self._dbatch is None:
            hb = self._local_filtered()
            spliced = hb.blk is not None
            # point rules stream unspliced batches, the Center rule streams spliced ones (the cases the range kernels
            # can start on before every read has landed); the other two combinations upload whole, then map
            if hb.transfer is not None and len(hb) and is_center == spliced:
                # the upload is the long pole: ship the transfer format in chunks on a copy stream and map every
                # chunk's bin range while the next chunks are still on the wire
                self._receiver = self._make_receiver(hb)
                chunks = self._plan_chunks(hb)
                if is_center:
                    map_center_streamed(self._receiver, hb.transfer_pinned(), chunks, self.layout, self.map_fn, sf, need,
                                        self._planes, length_hist=self._center_hist())
                else:
                    map_wire16_streamed(self._receiver, hb.transfer_pinned(), chunks, self.layout, self.map_fn, sf, need,
                                        self._planes)
                self._dbatch = self._receiver.batch
                self._finish_stats()
                return self._planes
        dbatch = self._device_batch()
        map_batch(dbatch, self.layout, self.map_fn, sf, strands=need, planes=self._planes, sync_stats=False,
                  bin_range=self._bin_range if self.is_sharded else None,
                  length_hist=self._center_hist() if is_center else None)
        self._finish_stats()
        return self._planes

    def _plan_chunks(self, hb):
        """Upload chunks ``[(read_a, read_b, bin_a, bin_b)]`` clipped to this rank's bins: at least 8 M reads each, at
        most 8 (Center rule: 4; small batches are launch-bound: fewer, larger chunks)."""
        # the Center rule's per-chunk launch (tile index, two binning passes, scans over the layout, jobs) costs ~0.8 ms,
        # the point rules' 0.1: fewer, larger chunks there (C3 end to end: 17.4 ms with 8 chunks, 16.3 with 4)
        default = 4 if isinstance(self.map_fn, CenterMapFactory) else 8
        n_chunks = max(1, min(int(os.environ.get("PB_UPLOAD_CHUNKS", default)), len(hb) // 8_000_000))      # (env: A/B aid)
        weights = None
        if os.environ.get("PB_CHUNK_WEIGHTS"):               # (env: A/B aid) relative read counts of the chunks; equal chunks measured best, profiles/NOTES_r02.md 7.13
            weights = [float(x) for x in os.environ["PB_CHUNK_WEIGHTS"].split(",")]
        chunks = type(self._receiver).plan_chunks(hb.transfer, self.layout, n_chunks, weights)
        lo, hi = self._bin_range
        out = []
        for a, b, bin_a, bin_b in chunks:
            bin_a, bin_b = max(bin_a, lo), min(bin_b, hi)
            if bin_b > bin_a or not out:
                out.append((a, b, bin_a, max(bin_b, bin_a)))
            else:                                   # nothing to map yet: the reads ride along with the next chunk
                pa, pb_, pbin_a, pbin_b = out[-1]
                out[-1] = (pa, b, pbin_a, pbin_b)
        a, b, bin_a, bin_b = out[-1]
        out[-1] = (a, len(hb), bin_a, hi)
        return out

    def _finish_stats(self):
        st = self._planes.stats = self._planes.stats_dev.cpu().numpy()
        if st[_lib.PB_STAT_DROPPED_ANY]:
            self.map_fn._warn_dropped(int(st[_lib.PB_STAT_DROPPED_ANY]), int(st[_lib.PB_STAT_DROPPED_LEN]))

    def _allreduce(self, t):
        """Complete a table of which every rank holds the part of its own positions."""
        if self._collective:
            from . import dist as pdist
            pdist.allreduce_sum(t)
        return t

    # -- queries -------------------------------------------------------------------------------
    def _read_range(self, hb, chrom, start, end):
        c = self.layout.index[chrom]
        r0, r1 = int(hb.chrom_read_off[c]), int(hb.chrom_read_off[c + 1])
        starts = hb.ref_start[r0:r1]
        lo = r0 + int(np.searchsorted(starts, start - hb.max_span + 1, side="left"))
        hi = r0 + int(np.searchsorted(starts, end, side="left"))
        return lo, max(hi, lo)

    def get_reads_and_counts(self, roi, roi_order=True):
        chrom, strand, start, end = roi.chrom, roi.strand, roi.start, roi.end
        if chrom not in self._chr_lengths:
            shape = [1] + getattr(self.map_fn, "shape", [])
            return [], np.zeros(shape)
        if not isinstance(self.map_fn, _MapFactory):
            raise TypeError("only plastid_b200 map factories can be evaluated on the GPU")
        if self.is_lazy:
            # the reference's own access pattern (genome_array.py:800-809): seek through the .bai, read the region's
            # records of every file, map them — nothing else of the files is inflated
            hb = self._filtered(merge_batches([f.fetch(chrom, start, end) for f in self._indexed]))
            dbatch = hb.to_device(self.device) if len(hb) else None
            lo, hi = 0, len(hb)
        else:
            dbatch = self._whole_device_batch()
            hb = self._whole_host if self.is_sharded else self._host_batch
            lo, hi = self._read_range(hb, chrom, start, end)
        qs = strand if strand in ("+", "-") else "."
        if hi > lo:
            counts, kept = self.map_fn.map_segment(dbatch, lo, hi, start, end, qs, self._size_filter())
            reads = [hb.read_view(lo + int(i)) for i in np.nonzero(kept)[0]]
        else:
            counts = np.zeros(self.map_fn._leading_shape() + [end - start], dtype=self.map_fn.count_dtype)
            reads = []
        if self._normalize is True:
            counts = counts / float(self.sum()) * 1e6
        if roi_order == True and strand == "-":
            counts = counts[..., ::-1]
        return reads, counts

    def get_reads(self, roi):
        reads, _ = self.get_reads_and_counts(roi)
        return reads

    def get(self, roi, roi_order=True):
        if isinstance(roi, SegmentChain):
            return roi.get_counts(self)
        if roi.chrom not in self._chr_lengths or not self._is_lowerable() or (self.is_lazy and self._planes is None):
            return self.get_reads_and_counts(roi, roi_order=roi_order)[1]
        qs = roi.strand if roi.strand in ("+", "-") else "."
        planes = self.count_planes(("+", "-") if qs != "." else (".",))
        counts = self._plane_slice(planes, qs, roi.chrom, roi.start, roi.end)
        if self._normalize is True:
            counts = counts / float(self.sum()) * 1e6
        if roi_order == True and roi.strand == "-":
            counts = counts[..., ::-1]
        return counts

    def _plane_slice(self, planes, strand, chrom, start, end):
        """``plane[start:end]`` of one chromosome as a fresh host vector: positions outside the chromosome are zero
        (the reference's fetch finds no reads there), positions of other ranks arrive through an all-reduce."""
        import torch
        n = max(end - start, 0)
        ci = self.layout.index[chrom]
        base, clen = int(self.layout.chrom_bin_off[ci]), int(self.layout.chrom_len[ci])
        a, b = min(max(start, 0), clen), max(min(end, clen), 0)           # inside the chromosome
        g0, g1 = max(base + a, planes.bin_lo), min(base + b, planes.bin_hi)   # inside this rank's bins
        whole = g0 == base + start and g1 == base + end
        if whole and not self._collective:
            t = planes.bins(strand, g0, g1)
        else:
            t = torch.zeros(n, dtype=planes.planes[strand].dtype, device=planes.device)
            if g1 > g0:
                t[g0 - base - start:g1 - base - start] = planes.bins(strand, g0, g1)
            if self._collective:
                t = t.to(torch.int64) & 0xFFFFFFFF if planes.dtype == "u32" else t
                self._allreduce(t)
        out = t.cpu().numpy()
        if planes.dtype == "u32":
            return (out.view(np.uint32) if out.dtype == np.int32 else out).astype(np.int64)
        return out

    def __getitem__(self, roi):
        return self.get(roi, roi_order=True)

    def _chrom_vector(self, planes, strand, chrom, n=None):
        """Device vector of the first ``n`` bins (default: all) of one chromosome strand, complete on every rank: a
        view of the plane on one GPU; on a sharded array the owned part in a zero vector, completed by an all-reduce
        (what whole-chromosome consumers — track export, ``to_genome_array`` — read)."""
        import torch
        ci = self.layout.index[chrom]
        base = int(self.layout.chrom_bin_off[ci])
        n = int(self.layout.chrom_len[ci]) if n is None else int(n)
        if not self.is_sharded:
            return planes.bins(strand, base, base + n)
        vec = torch.zeros(max(n, 1), dtype=planes.planes[strand].dtype, device=planes.device)[:n]
        g0, g1 = max(base, planes.bin_lo), min(base + n, planes.bin_hi)
        if g1 > g0:
            vec[g0 - base:g1 - base] = planes.bins(strand, g0, g1)
        return self._allreduce(vec)            # int32 storage of uint32 counts: x + 0 + ... + 0 is exact in two's complement

    # -- bulk entry points used by the scripts --------------------------------------------------
    def chain_table(self, chains, use_masks=True):
        return ChainTable.from_chains(chains, self.layout, use_masks=use_masks)

    def count_chains(self, chains, use_masks=True, planes=None):
        """(sums float64[n], unmasked lengths int64[n]) for a list of chains in one launch —
        what ``numpy.nansum(chain.get_masked_counts(ga))`` and ``chain.masked_length`` give.

        ``planes``: ``True`` sums over the count planes (building them if need be), ``False`` counts straight from
        the sorted reads without planes (``pb_chain_counts``: point rules; table-only programs never pay for
        4 bytes per genome position), ``None`` (default) takes the planes when they exist already and the
        plane-free path otherwise."""
        table = chains if isinstance(chains, ChainTable) else self.chain_table(chains, use_masks)
        need = table.__dict__.get("_strands_needed")          # geometry of the table: looked at once
        if need is None:
            need = tuple(sorted(set(_STRANDS[p] for p in np.unique(table.chain_plane)), key=_STRANDS.index)) or ("+",)
            table._strands_needed = need
            table._any_unknown = bool((~table.known).any())
        direct_ok = self._is_lowerable() and not isinstance(self.map_fn, CenterMapFactory)
        have = self._planes is not None and all(s in self._planes.planes for s in need)
        if planes is None:
            planes = have or not direct_ok
        if planes is False and not direct_ok:
            raise TypeError("plane-free counting needs a FivePrime/ThreePrime/VariableFivePrime mapping rule")
        if planes:
            sums, live = region_sums(self.count_planes(need), table)
        else:
            import torch
            stats = torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=self.device)
            sums, live = chain_counts(self._device_batch(), self.layout, self.map_fn, self._size_filter(), table,
                                      self._bin_range, stats)
            st = stats.cpu().numpy()
            if st[:3].any():
                self.map_fn._warn_dropped(int(st[:3].sum()), int(st[_lib.PB_STAT_DROPPED_LEN]))
        sums = self._allreduce(sums)
        sums, live = sums.cpu().numpy(), live.cpu().numpy()
        if table._any_unknown:
            live[~table.known] = table.unknown_live[~table.known]
        if self._normalize is True:
            sums = sums / float(self.sum()) * 1e6
        return sums, live

    # -- track export (genome_array.py:990-1111): run-length / non-zero compaction on the device ----
    def _export_records(self, chrom, strand, mode, window_size):
        """(start, end, value) arrays of one chromosome strand from ``pb_export_runs``."""
        import torch
        planes = self.count_planes((strand,))
        dev = planes.device
        n = self._chr_lengths[chrom]
        vec = self._chrom_vector(planes, strand, chrom)
        L = _lib.lib()
        ws_bytes = L.pb_export_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        n_out = torch.zeros(1, dtype=torch.int64, device=dev)
        dtype = 1 if planes.dtype == "f64" else 0

        def call(cap, st, en, va):
            _lib.check(L.pb_export_runs(_lib.ptr(vec), dtype, n, int(window_size), mode, cap, _lib.ptr(st), _lib.ptr(en),
                                        _lib.ptr(va), _lib.ptr(n_out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
            return int(n_out.item())
        total = call(0, None, None, None)            # counting pass
        st = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        en = torch.empty(max(total, 1), dtype=torch.int64, device=dev) if mode == 1 else None
        va = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
        if total:
            call(total, st, en, va)
        vals = va[:total].cpu().numpy()
        if planes.dtype == "u32":
            vals = vals.astype(np.int64)             # the reference's point rules count in int
        if self._normalize is True:
            vals = vals / float(self.sum()) * 1e6
        return st[:total].cpu().numpy(), (en[:total].cpu().numpy() if mode == 1 else None), vals

    @staticmethod
    def _write_records(fh, kind, chrom, st, en, vals, chunk=1 << 20):
        """Lines of one chromosome (``pb_format_track``: the reference's per-line ``fh.write("%s\\t%s\\n" % ...)``,
        genome_array.py:1030-1037, 1096-1111, as one native pass per 1 M records)."""
        L = _lib.lib()
        is_float = vals.dtype.kind == "f"
        vals = np.ascontiguousarray(vals, dtype=np.float64 if is_float else np.int64)
        st = np.ascontiguousarray(st, dtype=np.int64)
        en = None if en is None else np.ascontiguousarray(en, dtype=np.int64)
        name = chrom.encode()
        p = lambda a, lo: None if a is None else C.c_void_p(a.ctypes.data + lo * a.itemsize)   # noqa: E731
        cap = L.pb_format_track_bound(kind, name, min(chunk, len(st)))
        buf = np.empty(cap, dtype=np.uint8)
        for lo in range(0, len(st), chunk):
            n = min(chunk, len(st) - lo)
            got = L.pb_format_track(kind, name, p(st, lo), p(en, lo), p(vals, lo), int(is_float), n,
                                    buf.ctypes.data_as(C.c_void_p), cap, _lib.host_threads())
            if got < 0:
                raise _lib.PlastidB200Error(L.pb_last_error().decode())
            fh.write(buf[:got].tobytes().decode("utf-8"))

    @staticmethod
    def _write_track_header(fh, kind, trackname, kwargs):
        fh.write("track type=%s name=%s" % (kind, trackname))
        for k, v in sorted(kwargs.items(), key=lambda x: x[0]):
            fh.write(" %s=%s" % (k, v))
        fh.write("\n")

    def to_variable_step(self, fh, trackname, strand, window_size=100000, printer=None, **kwargs):
        """genome_array.py:990-1037: every non-zero position as ``1-based position<TAB>value``."""
        assert strand in self.strands()
        if not self._is_lowerable():
            raise TypeError("track export needs a mapping rule that lowers to whole-genome planes")
        self._write_track_header(fh, "wiggle_0", trackname, kwargs)
        for chrom in sorted(self.chroms()):
            if printer is not None:
                printer.write("Writing chromosome %s..." % chrom)
            fh.write("variableStep chrom=%s span=1\n" % chrom)
            pos, _e, vals = self._export_records(chrom, strand, 0, window_size)
            self._write_records(fh, 0, chrom, pos, None, vals)

    def to_bedgraph(self, fh, trackname, strand, window_size=100000, printer=None, **kwargs):
        """genome_array.py:1039-1111: runs of equal positive values, cut at ``window_size`` boundaries."""
        assert strand in self.strands()
        assert window_size > 0
        if not self._is_lowerable():
            raise TypeError("track export needs a mapping rule that lowers to whole-genome planes")
        self._write_track_header(fh, "bedGraph", trackname, kwargs)
        for chrom in sorted(self.chroms()):
            if printer is not None:
                printer.write("Writing chromosome %s..." % chrom)
            st, en, vals = self._export_records(chrom, strand, 1, window_size)
            self._write_records(fh, 1, chrom, st, en, vals)

    def to_genome_array(self, array_type=None):
        """genome_array.py:965-988 — including its quirk of dropping each chromosome's last base."""
        import torch
        if array_type is None:
            array_type = GenomeArray
        ga = array_type(chr_lengths=self.lengths(), strands=self.strands(), device=self.device)
        planes = self.count_planes(_STRANDS)
        for chrom in self.chroms():
            n = self._chr_lengths[chrom] - 1
            for strand in _STRANDS:
                src = self._chrom_vector(planes, strand, chrom, n)
                if planes.dtype == "u32":
                    vals = (src.to(torch.int64) & 0xFFFFFFFF).to(torch.float64)
                else:
                    vals = src.clone()
                if self._normalize is True:
                    vals = vals / float(self.sum()) * 1e6
                ga._set_device(chrom, strand, 0, vals)     # `ga[seg] = self[seg]`; both in 5'->3' order
        return ga


# ---------------------------------------------------------------------------------------------
# GenomeArray / SparseGenomeArray (mutable float64 planes)
# ---------------------------------------------------------------------------------------------
class GenomeArray(object):
    """``GenomeArray(chr_lengths=None, strands=None, min_chr_size=MIN_CHR_SIZE)`` with float64
    device planes (genome_array.py:1354-1611).  Chromosomes seen first in ``__setitem__`` are
    created on demand; reads outside the known range return zeros."""
    MIN_CHR_SIZE = 10 * 1000 * 1000

    def __init__(self, chr_lengths=None, strands=None, min_chr_size=None, device="cuda"):
        self.device = device
        self._strands = tuple(strands) if strands is not None else ("+", "-")
        self.min_chr_size = self.MIN_CHR_SIZE if min_chr_size is None else min_chr_size
        self._chr_lengths = dict(chr_lengths or {})
        self._chroms = {}
        self._sum = None
        self._normalize = False
        for chrom, n in self._chr_lengths.items():
            self._alloc(chrom, n)

    def _alloc(self, chrom, n):
        import torch
        _lib.require_cuda()
        self._chroms[chrom] = {s: torch.zeros(int(n), dtype=torch.float64, device=self.device)
                               for s in self._strands}
        self._chr_lengths[chrom] = int(n)

    def _grow(self, chrom, n):
        import torch
        for s, t in self._chroms[chrom].items():
            new = torch.zeros(int(n), dtype=torch.float64, device=self.device)
            new[:t.numel()] = t
            self._chroms[chrom][s] = new
        self._chr_lengths[chrom] = int(n)

    def chroms(self):
        return list(self._chroms.keys())

    def strands(self):
        return self._strands

    def lengths(self):
        return {c: self._chr_lengths[c] for c in self._chroms}

    def reset_sum(self):
        self._sum = None

    def set_sum(self, val):
        self._sum = val

    def sum(self):
        if self._sum is None:
            total = 0.0
            for planes in self._chroms.values():
                for t in planes.values():
                    total += float(t.sum().item())
            self._sum = total
        return self._sum

    def set_normalize(self, value=True):
        assert value in (True, False)
        self._normalize = value

    def _set_device(self, chrom, strand, start, vals):
        if chrom not in self._chroms:
            self._alloc(chrom, max(self.min_chr_size, start + vals.numel()))
        elif start + vals.numel() > self._chr_lengths[chrom]:
            self._grow(chrom, start + vals.numel() + 10000)
        self._chroms[chrom][strand][start:start + vals.numel()] = vals
        self._sum = None

    def __setitem__(self, seg, val):
        import torch
        self._sum = None
        if isinstance(seg, SegmentChain):
            if isinstance(val, np.ndarray):
                if seg.strand == "-":
                    val = val[::-1]
                x = 0
                for sub in seg:
                    n = len(sub)
                    self._set_device(sub.chrom, sub.strand, sub.start,
                                     torch.from_numpy(np.ascontiguousarray(val[x:x + n], dtype=np.float64)).to(self.device))
                    x += n
            else:
                for sub in seg:
                    self[sub] = val
            return
        if seg.strand not in self._strands:
            raise KeyError("strand %r not in array" % seg.strand)
        n = len(seg)
        if isinstance(val, np.ndarray):
            if seg.strand == "-":
                val = val[::-1]
            vals = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float64)).to(self.device)
        else:
            vals = torch.full((n,), float(val), dtype=torch.float64, device=self.device)
        self._set_device(seg.chrom, seg.strand, seg.start, vals)

    def plane(self, chrom, strand):
        return self._chroms[chrom][strand]

    def get(self, roi, roi_order=True):
        if isinstance(roi, SegmentChain):
            return roi.get_counts(self)
        n = len(roi)
        if roi.chrom not in self._chroms or roi.strand not in self._strands:
            vals = np.zeros(n, dtype=np.float64)
        else:
            t = self._chroms[roi.chrom][roi.strand]
            vals = np.zeros(n, dtype=np.float64)
            a, b = max(roi.start, 0), min(roi.end, t.numel())
            if a < b:
                vals[a - roi.start:b - roi.start] = t[a:b].cpu().numpy()
        if self._normalize is True:
            vals = 1e6 * vals / self.sum()
        if roi_order == True and roi.strand == "-":
            vals = vals[::-1]
        return vals

    def __getitem__(self, roi):
        return self.get(roi, roi_order=True)

    def count_chains(self, chains, use_masks=True):
        """Bulk masked sums over chains (same contract as ``BAMGenomeArray.count_chains``)."""
        import torch
        chroms = sorted(self._chroms)
        layout = GenomeLayout(chroms, [self._chr_lengths[c] for c in chroms])
        planes = CountPlanes(layout, "f64", self.device)
        planes.alloc(self._strands)
        for s in self._strands:
            planes.planes[s].zero_()
            for chrom in chroms:
                base = int(layout.chrom_bin_off[layout.index[chrom]])
                t = self._chroms[chrom][s]
                planes.planes[s][base:base + t.numel()] = t
        table = ChainTable.from_chains(chains, layout, use_masks=use_masks, unstranded="." in self._strands
                                       and "+" not in self._strands)
        sums, live = region_sums(planes, table)
        sums, live = sums.cpu().numpy(), live.cpu().numpy()
        if self._normalize is True:
            sums = 1e6 * sums / self.sum()
        return sums, live


class SparseGenomeArray(GenomeArray):
    """Same API as :class:`GenomeArray`; chromosome planes are allocated on first write
    (genome_array.py:2134-2298 uses scipy DOK matrices to save host RAM; 180 GB of HBM make dense
    planes per touched chromosome the cheaper representation here)."""

    def __init__(self, chr_lengths=None, strands=None, min_chr_size=None, device="cuda"):
        GenomeArray.__init__(self, None, strands, min_chr_size, device)
        self._declared = dict(chr_lengths or {})

    def _alloc(self, chrom, n):
        GenomeArray._alloc(self, chrom, max(n, self._declared.get(chrom, 0)))

    def lengths(self):
        out = dict(self._declared)
        out.update(GenomeArray.lengths(self))
        return out
