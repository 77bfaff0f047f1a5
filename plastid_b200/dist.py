"""Multi-GPU sharding of the counting path (one process per GPU, ``torch.distributed``).

The reference is single-process; counts at a position depend only on the reads
mapped there and region / window values only on their own positions, so the path shards with no
count-vector collective:

* **position-range sharding** (``position_cuts`` / ``balanced_cuts`` / ``shard_positions`` / ``clip_table``, SURVEY §8e;
  what ``BAMGenomeArray``, the programs and ``bench.py --gpus N`` use): the concatenated genome is cut into contiguous bin
  ranges of equal cost (reads streamed + plane bins written); a rank holds the planes of its range only (1/N of the
  memory), receives the reads that start inside it plus a halo of ``max_span`` before it (those go to both neighbours,
  each counts only sites inside its own range), and sums the parts of every region that fall into its range; the
  partial region tables are summed with one all-reduce.  ``snap="chromosomes"`` puts the cuts on chromosome boundaries
  (BASELINE config 5).
* **window profiles** (``window_profile``: ``metagene count``): the count matrix stays on its ranks — rows are completed
  and normalised by the rank owning their first position (``row_owners``), means are all-reduced as column sums, exact
  medians are taken per column slice after one all-to-all (``exchange_column_slices``).
* **read-range sharding** (``shard_reads``; ``bench.py --sharding reads``, round 1's weak-scaling mode): every rank maps
  a contiguous slice of the coordinate-sorted batch over the whole layout; region sums, phase sums and mean-profile
  numerators are linear in the reads, so the small result tables are summed with one all-reduce (``allreduce_sum``).
* **chromosome assignment** (``assign_chromosomes`` / ``shard_chromosomes`` / ``gather_rows`` / ``gather_matrix``):
  every rank owns whole chromosomes (longest-processing-time assignment by read count) and only the regions on them;
  the per-region tables are gathered.

NCCL is used on GPUs, gloo in the CPU tests; all messages are small (a few hundred KB for tables, 1/N of a window
matrix for exact medians).
"""
import numpy as np

from .batch import AlignmentBatch


def is_distributed():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def world():
    import torch.distributed as dist
    if not is_distributed():
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


# ------------------------------------------------------------------------------- planning (host)
def read_range(n_reads, rank, world_size):
    """Contiguous, balanced slice [a, b) of the read index space for ``rank``."""
    base, extra = divmod(int(n_reads), int(world_size))
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def shard_reads(hb, rank, world_size):
    """Read-range shard of a host batch (same chromosomes and layout; ``mapped`` stays the global
    count so RPKM normalisation is unchanged)."""
    a, b = read_range(len(hb), rank, world_size)
    off = np.clip(hb.chrom_read_off, a, b) - a
    blk_off = blk = None
    if hb.blk_off is not None:
        k0, k1 = int(hb.blk_off[a]), int(hb.blk_off[b])
        blk_off = hb.blk_off[a:b + 1].astype(np.int64) - k0
        blk = hb.blk[k0:k1]
    out = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start[a:b], hb.meta[a:b], off, blk_off, blk,
                         max_span=hb.max_span, mapped=hb.mapped)
    if hb.objects is not None:
        out.objects = hb.objects[a:b]
    return out


def position_cuts(hb, layout, world_size, snap="bins", weights=(1.0, 1.0)):
    """Global-bin cut points ``int64[world_size + 1]`` (multiples of PB_LAYOUT_ALIGN, first 0, last
    ``layout.total_bins``) splitting the genome into ``world_size`` contiguous position ranges of about
    equal cost (:func:`balanced_cuts`: reads streamed + plane bins written).  ``snap="chromosomes"``: cuts fall on chromosome boundaries only (BASELINE config 5,
    "sharded by chromosome": every rank owns a run of whole chromosomes) — the boundary whose cumulative read
    count is nearest each target."""
    from . import _lib
    A = _lib.PB_LAYOUT_ALIGN
    n = len(hb)
    if snap == "chromosomes":
        cum = np.asarray(hb.chrom_read_off, dtype=np.int64)          # reads before chromosome c
        cuts = [0]
        for r in range(1, world_size):
            c = int(np.argmin(np.abs(cum - (r * n) // world_size)))
            cuts.append(max(int(layout.chrom_bin_off[c]), cuts[-1]))
        cuts.append(int(layout.total_bins))
        return np.asarray(cuts, dtype=np.int64)
    if snap != "bins":
        raise ValueError("snap must be 'bins' or 'chromosomes'")
    starts, off = hb.ref_start, np.asarray(hb.chrom_read_off, dtype=np.int64)

    def reads_before(c, local):
        a, b = int(off[c]), int(off[c + 1])
        return a + int(np.searchsorted(starts[a:b], local, side="left"))
    return balanced_cuts(layout, n, world_size, reads_before, weights)


def balanced_cuts(layout, n_reads, world_size, reads_before, weights=(1.0, 1.0), lo=0, hi=None):
    """Cut points that give every rank the same COST, ``weights[0] * reads + weights[1] * bins``: a rank's mapping
    pass streams its reads (8 bytes each) and writes the dense planes of its bins (2 strands x 4 bytes), so equal read
    counts alone leave the rank with the sparsest stretch of genome the most plane bytes to write (N = 8, C2: 363-401 M
    bins per rank and tiles-kernel times to match, profiles/NOTES_r02.md section 5).  ``weights=(1, 0)`` is the
    equal-read-count rule.  ``reads_before(c, local)``: reads of chromosome ``c`` and before that start before
    chromosome position ``local`` (host: a search in ``ref_start``; device batches search on the device).
    ``lo`` / ``hi``: cut the global bins [lo, hi) only (default: the whole layout; ``n_reads`` is ignored then and the
    reads starting in the range are counted).  The cost is monotone in the cut position: one bisection over the
    PB_LAYOUT_ALIGN grid per cut."""
    from . import _lib
    A = _lib.PB_LAYOUT_ALIGN
    total = int(layout.total_bins)
    hi = total if hi is None else int(hi)
    lo = int(lo)
    wr, wb = float(weights[0]), float(weights[1])
    coff = np.asarray(layout.chrom_bin_off, dtype=np.int64)

    def before(g):
        c = min(int(np.searchsorted(coff, g, side="right")) - 1, len(layout.chroms) - 1)
        return reads_before(c, g - int(coff[c]))
    r_lo = before(lo) if lo > 0 else 0
    r_hi = before(hi) if hi < total else int(n_reads)

    def cost(g):
        return wr * (before(g) - r_lo) + wb * (g - lo)
    whole = wr * (r_hi - r_lo) + wb * (hi - lo)
    cuts = [lo]
    for r in range(1, world_size):
        target = whole * r / world_size
        a, b = cuts[-1] // A, hi // A                     # smallest grid point whose cost reaches the target
        while a < b:
            mid = (a + b) // 2
            if cost(mid * A) >= target:
                b = mid
            else:
                a = mid + 1
        cuts.append(min(max(a * A, lo), hi))
    cuts.append(hi)
    return np.asarray(cuts, dtype=np.int64)


def device_reads_before(dbatch, layout, off=None):
    """``reads_before(c, local)`` of :func:`balanced_cuts` for a device batch (one short device search per call)."""
    import torch
    off = dbatch.chrom_read_off.cpu().numpy() if off is None else off

    def reads_before(c, local):
        a, b = int(off[c]), int(off[c + 1])
        if b <= a or local <= 0:
            return a
        if local >= int(layout.chrom_len[c]):
            return b
        key = torch.tensor([local], dtype=dbatch.ref_start.dtype, device=dbatch.ref_start.device)
        return a + int(torch.searchsorted(dbatch.ref_start[a:b], key, right=False).item())
    return reads_before


def shard_positions(hb, layout, rank, world_size, cuts=None, snap="bins"):
    """Position-range shard: ``(sub_batch, bin_lo, bin_hi)``.  ``sub_batch`` holds, per chromosome,
    the reads starting in ``[lo - hb.max_span, hi)`` of the part of the chromosome inside the rank's
    range — every read that can put a site into ``[bin_lo, bin_hi)`` — with the same chromosomes
    and layout as ``hb`` (coordinates stay global)."""
    cuts = position_cuts(hb, layout, world_size, snap) if cuts is None else cuts
    g_lo, g_hi = int(cuts[rank]), int(cuts[rank + 1])
    keep = []
    off = [0]
    for c in range(len(hb.chroms)):
        base = int(layout.chrom_bin_off[c])
        a, b = int(hb.chrom_read_off[c]), int(hb.chrom_read_off[c + 1])
        lo, hi = g_lo - base, g_hi - base           # range in this chromosome's coordinates
        if hi <= 0 or lo >= int(hb.chrom_len[c]) or b <= a or g_hi <= g_lo:
            off.append(off[-1])
            continue
        i0 = a + int(np.searchsorted(hb.ref_start[a:b], lo - hb.max_span, side="left"))
        i1 = a + int(np.searchsorted(hb.ref_start[a:b], hi, side="left"))
        keep.append((i0, i1))
        off.append(off[-1] + max(i1 - i0, 0))
    idx = np.concatenate([np.arange(i0, i1) for i0, i1 in keep]) if keep else np.zeros(0, dtype=np.int64)
    blk_off = blk = None
    if hb.blk_off is not None and len(idx):
        rows = (hb.blk_off[idx + 1].astype(np.int64) - hb.blk_off[idx].astype(np.int64))
        blk_off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(rows, out=blk_off[1:])
        parts = [hb.blk[int(hb.blk_off[i0]):int(hb.blk_off[i1])] for i0, i1 in keep]
        blk = np.concatenate(parts) if parts else np.zeros((0, 2), dtype=np.int32)
        if len(blk) == 0:
            blk_off = blk = None
    sub = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start[idx], hb.meta[idx], off, blk_off, blk,
                         max_span=hb.max_span, mapped=hb.mapped)
    return sub, g_lo, g_hi


def shard_positions_device(dbatch, layout, rank, world_size, weights=(1.0, 1.0)):
    """:func:`shard_positions` for a batch that already lives on the device (unspliced batches): the
    same cuts (:func:`balanced_cuts`), the
    rank's reads + halo gathered per chromosome with searches on the device.
    Returns ``(sub_batch, bin_lo, bin_hi, cuts)``."""
    import torch
    from . import _lib
    from .batch import DeviceBatch
    if dbatch.blk_off is not None:
        raise ValueError("shard_positions_device handles unspliced batches; shard spliced batches on the host")
    n = dbatch.n_reads
    off = dbatch.chrom_read_off.cpu().numpy()
    reads_before = device_reads_before(dbatch, layout, off)
    cuts = [int(x) for x in balanced_cuts(layout, n, world_size, reads_before, weights)]
    g_lo, g_hi = cuts[rank], cuts[rank + 1]
    pieces, new_off = [], [0]
    for c in range(dbatch.n_chrom):
        base = int(layout.chrom_bin_off[c])
        a, b = int(off[c]), int(off[c + 1])
        lo, hi = g_lo - base, g_hi - base
        if hi <= 0 or lo >= int(layout.chrom_len[c]) or b <= a or g_hi <= g_lo:
            new_off.append(new_off[-1])
            continue
        seg = dbatch.ref_start[a:b]
        clen = int(layout.chrom_len[c])              # keep the search keys inside int32
        keys = torch.tensor([max(lo - dbatch.max_span, -clen - 1), min(hi, clen)], dtype=seg.dtype, device=seg.device)
        i0, i1 = (a + int(x) for x in torch.searchsorted(seg, keys, right=False).tolist())
        pieces.append((i0, i1))
        new_off.append(new_off[-1] + max(i1 - i0, 0))
    if pieces:
        starts = torch.cat([dbatch.ref_start[i0:i1] for i0, i1 in pieces])
        metas = torch.cat([dbatch.meta[i0:i1] for i0, i1 in pieces])
    else:
        starts, metas = dbatch.ref_start[:0].clone(), dbatch.meta[:0].clone()
    sub = DeviceBatch(int(starts.numel()), dbatch.n_chrom, dbatch.max_span, starts, metas,
                      torch.tensor(new_off, dtype=torch.int64, device=starts.device), None, None, dbatch.max_block_len)
    return sub, g_lo, g_hi, np.asarray(cuts, dtype=np.int64)


def clip_table(table, bin_lo, bin_hi):
    """The part of every chain of a :class:`~plastid_b200.regions.ChainTable` inside the global bins
    ``[bin_lo, bin_hi)``: blocks clipped (or dropped), mask bits re-based to the clipped chain
    positions.  Region sums and unmasked lengths over clipped tables of all ranks add up to those of
    the whole table, so one all-reduce finishes a position-sharded region table."""
    from .regions import ChainTable
    bstart, bend, chain_off, length = [], [], [0], []
    bits_out, mask_off = [], []
    nbits = 0
    old_bits = None
    if table.mask_bits is not None:
        old_bits = np.unpackbits(table.mask_bits, bitorder="little")
    for c in range(table.n_chains):
        pos = 0                                   # chain position of the current block's first base
        n_c = 0
        pieces = []
        for j in range(int(table.chain_off[c]), int(table.chain_off[c + 1])):
            bs, be = int(table.bstart[j]), int(table.bend[j])
            lo, hi = max(bs, bin_lo), min(be, bin_hi)
            if lo < hi:
                bstart.append(lo)
                bend.append(hi)
                if old_bits is not None:
                    m0 = int(table.mask_off[c]) + pos + (lo - bs)
                    pieces.append(old_bits[m0:m0 + (hi - lo)])
                n_c += hi - lo
            pos += be - bs
        chain_off.append(len(bstart))
        length.append(n_c)
        mask_off.append(nbits)
        if old_bits is not None:
            bits_out.append(np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.uint8))
        nbits += n_c
    mask_bits = None
    if old_bits is not None:
        flat = np.concatenate(bits_out) if bits_out else np.zeros(0, dtype=np.uint8)
        mask_bits = np.packbits(flat, bitorder="little")
        if len(mask_bits) == 0:
            mask_bits = np.zeros(1, dtype=np.uint8)
    out = ChainTable(table.layout, bstart, bend, chain_off, table.chain_plane, table.chain_reverse, length,
                     mask_bits, mask_off if old_bits is not None else None, table.known)
    return out


def assign_chromosomes(hb, world_size):
    """Longest-processing-time assignment of chromosomes to ranks by read count (ties: length).
    Returns a list of sorted chromosome-index lists, one per rank."""
    n_reads = np.diff(hb.chrom_read_off)
    order = sorted(range(len(hb.chroms)), key=lambda c: (-int(n_reads[c]), -int(hb.chrom_len[c]), c))
    load = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for c in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        owned[r].append(c)
        load[r] += int(n_reads[c]) + 1
    return [sorted(x) for x in owned]


def shard_chromosomes(hb, chrom_ids):
    """Batch restricted to ``chrom_ids`` (their reads, their chromosomes only)."""
    chrom_ids = list(chrom_ids)
    starts, metas, off = [], [], [0]
    blk_rows, blks = [], []
    for c in chrom_ids:
        a, b = int(hb.chrom_read_off[c]), int(hb.chrom_read_off[c + 1])
        starts.append(hb.ref_start[a:b])
        metas.append(hb.meta[a:b])
        off.append(off[-1] + b - a)
        if hb.blk_off is not None:
            k0, k1 = int(hb.blk_off[a]), int(hb.blk_off[b])
            blk_rows.append(np.diff(hb.blk_off[a:b + 1].astype(np.int64)))
            blks.append(hb.blk[k0:k1])
    blk_off = blk = None
    if hb.blk_off is not None and chrom_ids:
        rows = np.concatenate(blk_rows)
        blk_off = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(rows, out=blk_off[1:])
        blk = np.concatenate(blks) if blks else np.zeros((0, 2), dtype=np.int32)
        if len(blk) == 0:
            blk_off = blk = None
    cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt))
    return AlignmentBatch([hb.chroms[c] for c in chrom_ids], hb.chrom_len[chrom_ids], cat(starts, np.int32),
                          cat(metas, np.uint32), off, blk_off, blk, max_span=hb.max_span, mapped=hb.mapped)


def owner_of_chains(chains, hb, owned):
    """Rank owning each chain under a chromosome assignment (-1: chromosome unknown -> rank 0)."""
    rank_of = {}
    for r, ids in enumerate(owned):
        for c in ids:
            rank_of[hb.chroms[c]] = r
    return np.asarray([rank_of.get(ch.chrom, 0) if len(ch) else 0 for ch in chains], dtype=np.int64)


# ------------------------------------------------------------------------------- collectives
def allreduce_sum(t):
    """In-place sum over ranks of a (small) tensor: region tables, phase sums, profile numerators."""
    import torch.distributed as dist
    if is_distributed() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_rows(local_values, owner, rank=None):
    """Assemble a per-region table from per-rank pieces: rank r computed the rows with
    ``owner == r`` (in order); every rank gets the full table."""
    import torch
    import torch.distributed as dist
    owner = np.asarray(owner)
    if not is_distributed() or dist.get_world_size() == 1:
        return local_values
    rank = dist.get_rank() if rank is None else rank
    shape = (len(owner),) + tuple(local_values.shape[1:])
    full = torch.zeros(shape, dtype=local_values.dtype, device=local_values.device)
    idx = torch.from_numpy(np.nonzero(owner == rank)[0]).to(local_values.device)
    full[idx] = local_values
    dist.all_reduce(full, op=dist.ReduceOp.SUM)      # rows are disjoint across ranks: sum == gather
    return full


def gather_matrix(local_rows):
    """Concatenate row blocks of all ranks (rank order) on every rank — the exchange an exact
    multi-GPU median needs.  Row counts may differ between ranks."""
    import torch
    import torch.distributed as dist
    if not is_distributed() or dist.get_world_size() == 1:
        return local_rows
    ws = dist.get_world_size()
    n = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = tuple(local_rows.shape[1:])
    pad = torch.zeros((max(counts),) + width, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    parts = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def global_length_hist(sub, layout, bin_lo, bin_hi):
    """Center rule under position-range sharding: the aligned-length histogram of the WHOLE batch, from
    which every rank derives the same slot / fixed-point tables (``map_batch(..., length_hist=...)``) so
    that the sharded planes equal the unsharded ones bit for bit.  Every rank counts the reads of ``sub``
    (its shard incl. halo, from :func:`shard_positions`) that START in its own range — each read of the
    batch is owned by exactly one rank — and the 65536 counters are all-reduced (512 KB)."""
    import torch
    from .batch import meta_length_hist
    c_of = np.searchsorted(sub.chrom_read_off, np.arange(len(sub)), side="right") - 1
    g = layout.chrom_bin_off[c_of] + sub.ref_start.astype(np.int64)
    owned = (g >= bin_lo) & (g < bin_hi)
    hist = torch.from_numpy(meta_length_hist(sub.meta[owned]))
    return allreduce_sum(hist).numpy()


# ------------------------------------------------------------------------------- window profiles (metagene count)
def all_ranges(lo, hi, device="cpu"):
    """The bin range of every rank, ``int64[world, 2]`` (one small all-gather; contiguous in rank order for the
    partitions this module makes)."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return np.asarray([[int(lo), int(hi)]], dtype=np.int64)
    mine = torch.tensor([int(lo), int(hi)], dtype=torch.int64, device=device)
    parts = [torch.zeros_like(mine) for _ in range(ws)]
    dist.all_gather(parts, mine)
    return torch.stack(parts).cpu().numpy()


def row_owners(table, ranges):
    """Which rank completes which row of a window / chain table under position sharding: ``owner[c]`` = the rank whose
    bins hold the first position of chain ``c`` that lies on the genome (rank 0 for chains without one), and the
    indices of the chains whose positions lie on more than one rank (a few per cut).  Geometry only."""
    from .regions import VIRTUAL_BIN
    n = table.n_chains
    ranges = np.asarray(ranges, dtype=np.int64)
    world_size = len(ranges)
    chain_of = np.repeat(np.arange(n), np.diff(table.chain_off))
    real = (table.bstart < VIRTUAL_BIN) & (table.bend > table.bstart)
    his = ranges[:, 1]
    first = np.minimum(np.searchsorted(his, table.bstart[real], side="right"), world_size - 1)
    last = np.minimum(np.searchsorted(his, table.bend[real] - 1, side="right"), world_size - 1)
    lo_r = np.full(n, world_size, dtype=np.int64)
    hi_r = np.full(n, -1, dtype=np.int64)
    np.minimum.at(lo_r, chain_of[real], first)
    np.maximum.at(hi_r, chain_of[real], last)
    owner = np.where(hi_r >= 0, lo_r, 0)
    return owner, np.nonzero(hi_r > lo_r)[0]


def exchange_column_slices(rows, row_counts=None):
    """``rows`` (float64 ``[k_r, width]``, this rank's rows) -> ``(slice, bounds)``: the columns
    ``[bounds[rank], bounds[rank + 1])`` of the rows of ALL ranks (rank order), ``[sum(k_r), w_rank]``.  NCCL: one
    all-to-all (every rank ships ``1 / world`` of its columns to every peer: the exchange an exact per-column median
    needs, at ``1 / world`` of the bytes of gathering whole rows); other backends: all-gather, then cut.
    ``row_counts``: ``k_r`` of every rank when the caller knows them (no size exchange, no host synchronisation)."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    width = int(rows.shape[1])
    bounds = [(c * width) // ws for c in range(ws + 1)]
    if ws == 1:
        return rows, bounds
    if row_counts is None:
        k = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        ks = [torch.zeros_like(k) for _ in range(ws)]
        dist.all_gather(ks, k)
        row_counts = [int(x.item()) for x in ks]
    ks = [int(x) for x in row_counts]
    w_mine = bounds[rank + 1] - bounds[rank]
    if dist.get_backend() == "nccl":
        ins = [rows[:, bounds[c]:bounds[c + 1]].contiguous() for c in range(ws)]
        outs = [torch.empty((ks[r], w_mine), dtype=rows.dtype, device=rows.device) for r in range(ws)]
        dist.all_to_all(outs, ins)
        return torch.cat(outs, dim=0), bounds
    pad = torch.zeros((max(ks), width), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    parts = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat([p_[:c, bounds[rank]:bounds[rank + 1]] for p_, c in zip(parts, ks)], dim=0).contiguous(), bounds


def window_profile(mat, mmask, table, ranges, norm_lo, norm_hi, min_counts, mode="median", per_million_of=None,
                   want_rows=True):
    """``metagene count`` on a position-sharded genome without moving the count matrix (60 k x 350 float64 = 168 MB at
    BASELINE config 4): ``mat`` / ``mmask`` are this rank's ``gather_windows`` output (cells of its own positions,
    zero elsewhere, NaN where the window has no position).

    * a row is COMPLETED by one rank, the owner of its first position (:func:`row_owners`: geometry, known to every
      rank without talking); the few rows that lie on both sides of a cut are summed over ranks first (one small
      all-reduce);
    * every rank normalises and selects its own rows (``pb_window_normalize``; metagene.py:916-924);
    * mean (``--use_mean``): column sums and counts of the selected rows, all-reduced (2 x width numbers);
      median (default; not all-reducible): every rank receives ``width / world`` columns of all owned rows (one
      all-to-all, rows that are not selected travel as NaN), takes their exact medians (``pb_column_profile``) and the
      slices are gathered.  Three collectives, no host synchronisation.

    Returns ``(profile, regions_counted, denominator, row_select)`` complete on every rank (the last two ``None``
    unless ``want_rows``); profile and counts equal the single-GPU result (medians bit for bit; means up to the order
    of the float sums)."""
    import torch
    import torch.distributed as dist
    from .genome_array import window_normalize, column_profile
    rank, ws = world()
    dev = mat.device
    n, width = mat.shape
    cache = table.__dict__.setdefault("_owners", {})
    key = (tuple(int(x) for x in np.asarray(ranges).reshape(-1)), rank, str(dev))
    if key not in cache:
        owner, strad = row_owners(table, ranges)
        lo, hi = (int(x) for x in np.asarray(ranges)[rank])
        chain_of = np.repeat(np.arange(table.n_chains), np.diff(table.chain_off))
        hit = np.zeros(table.n_chains, dtype=bool)
        hit[chain_of[(table.bend > lo) & (table.bstart < hi)]] = True      # rows this rank wrote (gather_windows touched_only)
        cache[key] = (np.bincount(owner, minlength=ws)[:ws], torch.from_numpy(np.nonzero(owner == rank)[0]).to(dev),
                      torch.from_numpy(strad).to(dev), torch.from_numpy(hit[strad]).to(dev))
    counts, own_idx, strad, strad_mine = cache[key]
    if strad.numel() and ws > 1:
        # NaN cells are geometry (NaN on every rank that wrote the row); counts have one non-zero summand; a rank
        # without a position in the row never wrote it and contributes zeros
        part = torch.where(strad_mine.unsqueeze(1), mat[strad], torch.zeros((), dtype=mat.dtype, device=dev))
        allreduce_sum(part)
        mat[strad] = part
    k = int(own_idx.numel())
    x, xm = mat[own_idx], mmask[own_idx]
    if per_million_of is not None:
        x = x / float(per_million_of) * 1e6
    if k:
        denom, sel, norm, nmask = window_normalize(x, xm, norm_lo, norm_hi, min_counts)
    else:
        denom, sel = torch.zeros(0, dtype=torch.float64, device=dev), torch.zeros(0, dtype=torch.uint8, device=dev)
        norm, nmask = x, xm
    if mode == "mean":
        both = torch.zeros(2 * width, dtype=torch.float64, device=dev)
        if k:
            _p, n_regions, col_sum = column_profile(norm, nmask, sel, "mean")
            both[:width], both[width:] = col_sum, n_regions.to(torch.float64)
        allreduce_sum(both)                    # counts are far below 2^53: exact as float64
        n_regions = both[width:].to(torch.int64)
        profile = both[:width] / both[width:]
    else:
        y = torch.where((nmask != 0) | (sel == 0).unsqueeze(1), torch.full_like(norm, float("nan")), norm)
        part, bounds = exchange_column_slices(y, counts)
        w_mine = bounds[rank + 1] - bounds[rank]
        w_max = max(bounds[c + 1] - bounds[c] for c in range(ws))
        loc = torch.zeros((2, w_max), dtype=torch.float64, device=dev)
        loc[0] = float("nan")
        if part.shape[0] and w_mine:
            pmask = torch.isnan(part).to(torch.uint8)
            every = torch.ones(part.shape[0], dtype=torch.uint8, device=dev)
            p_c, n_c, _s = column_profile(torch.nan_to_num(part, nan=0.0), pmask, every, "median")
            loc[0, :w_mine], loc[1, :w_mine] = p_c, n_c.to(torch.float64)
        if ws > 1:
            parts = [torch.zeros_like(loc) for _ in range(ws)]
            dist.all_gather(parts, loc)
        else:
            parts = [loc]
        profile = torch.cat([parts[c][0, :bounds[c + 1] - bounds[c]] for c in range(ws)])
        n_regions = torch.cat([parts[c][1, :bounds[c + 1] - bounds[c]] for c in range(ws)]).to(torch.int64)
    if not want_rows:
        return profile, n_regions, None, None
    rows = torch.zeros((2, n), dtype=torch.float64, device=dev)
    rows[0, own_idx], rows[1, own_idx] = denom, sel.to(torch.float64)
    allreduce_sum(rows)
    return profile, n_regions, rows[0], rows[1].to(torch.uint8)


def mean_profile(col_sum, n_regions):
    """Multi-GPU ``--use_mean`` metagene profile: all-reduce numerators and counts, then divide."""
    allreduce_sum(col_sum)
    allreduce_sum(n_regions)
    return col_sum / n_regions.to(col_sum.dtype)
