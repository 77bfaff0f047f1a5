// pb_tiles.cu — kernels shared by the point and center mapping paths: tile candidate index,
// binning of multi-block (spliced) reads by tile, statistics, workspace and timing plumbing.
#include "pb_tiles.cuh"
#include <stdlib.h>

namespace {

// ----------------------------------------------------------------------------------------
// tile -> candidate read slice
// ----------------------------------------------------------------------------------------
__global__ void pb_tile_index_kernel(PbReads b, PbLayoutDev lay, int tile_bins, int64_t tile_begin, int64_t tile_end,
                                     int64_t read_limit, int split, PbTile *__restrict__ tiles,
                                     PbJob *__restrict__ jobs, int64_t job_capacity,
                                     unsigned long long *__restrict__ n_jobs)
{
    int64_t t = tile_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tile_end) return;
    int64_t g0 = t * tile_bins;
    int c = pb_chrom_of_bin(lay, g0);
    int64_t p0 = g0 - __ldg(lay.chrom_bin_off + c);
    int64_t clen = __ldg(lay.chrom_len + c);
    int64_t r0 = 0, r1 = 0;
    if (c < b.n_chrom) { r0 = __ldg(b.chrom_read_off + c); r1 = __ldg(b.chrom_read_off + c + 1); }
    // streaming uploads: reads at or beyond read_limit have not arrived yet (and, being sorted, start
    // beyond every tile of the range being mapped)
    if (r1 > read_limit) r1 = read_limit;
    if (r0 > r1) r0 = r1;
    // a single-block read [s, s+L) can only touch [p0, p0+T) if p0 - max_block_len < s < p0 + T;
    // multi-block reads reach their tiles through the binned records instead
    int64_t lo = pb_lower_bound(b.ref_start, r0, r1, p0 - b.max_block_len + 1);
    int64_t hi = pb_lower_bound_near(b.ref_start, lo, r1, p0 + tile_bins);
    int64_t live = clen - p0;
    live = live < 0 ? 0 : (live > tile_bins ? tile_bins : live);
    PbTile d;
    d.lo = lo; d.p0 = p0;
    d.n = (live > 0 && hi - lo < 0x7fffffff) ? (int)(hi - lo) : (live > 0 ? 0x7fffffff : 0);
    d.live = (int)live; d.chrom = c; d.pad = 0;
    if (split > 0 && d.n > split) {
        // pile-up: the tile job keeps the first `split` candidates, the rest become overflow jobs
        const long long rest = (long long)d.n - split;
        const long long nj = (rest + split - 1) / split;
        const long long at = (long long)atomicAdd(n_jobs, (unsigned long long)nj);
        for (long long j = 0; j < nj; ++j) {
            if (at + j >= job_capacity) break;      // cannot happen: capacity covers sum(n) / split
            PbJob jb;
            jb.lo = lo + split + j * split;
            jb.tile = t;
            jb.n = (int)(rest - j * split < split ? rest - j * split : split);
            jb.kind = jb.hot = jb.pad = 0;
            jobs[at + j] = jb;
        }
        d.n = split;
    }
    tiles[t] = d;
}

// ----------------------------------------------------------------------------------------
// K1: CIGAR blocks of multi-block reads -> per-tile record buckets (counting sort by tile)
// ----------------------------------------------------------------------------------------
// FILL = 0: count pass (one fire-and-forget RED per record, no record is built); FILL = 1: fill pass.  AGG: lanes that
// target the same tile share one cursor atomic (match.any) instead of one returning atomic per record.
template <bool CENTER, int OCC, bool FILL, bool AGG>
__global__ void __launch_bounds__(256, OCC)   // latency-bound gather chains: resident threads vs registers (OCC CTAs/SM)
pb_bin_kernel(PbReads b, PbRuleDev r, PbLayoutDev lay, int planes, const int16_t *__restrict__ slot_of_len,
              int tile_shift, int64_t tile_lo, int64_t tile_hi, int64_t tile_rec_lo, int64_t read_begin,
              uint32_t *__restrict__ rec_cursor,
              const uint32_t *__restrict__ rec_off, PbRec *__restrict__ recs, unsigned long long *__restrict__ stat_slots)
{
    // [tile_lo, tile_hi): the tiles this launch produces (position-sharded ranks map a bin range only);
    // records and statistics outside it belong to another rank.  Center intervals reach into later tiles, so
    // their records are kept from tile_rec_lo = tile_lo - lookback on (the tiles kernel looks that far back).
    unsigned long long drop_p = 0, drop_m = 0, drop_a = 0, map_p = 0, map_m = 0, map_a = 0;
    unsigned int drop_len = 0;
    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS,
               want_any = planes & PB_PLANE_ANY;
    // A warp takes 128 consecutive reads per round.  Only the multi-block ones (a third of the C3 batch)
    // have work to do, so the warp first compacts them into a dense shared-memory list (ballot + prefix
    // popcount) and then walks that list 32 at a time with ALL lanes busy; what is left of the list (< 32 entries)
    // waits at its front for the next round's reads.  ncu history: the uncompacted loop was issue-bound at 11 of 32
    // lanes active per instruction; the compacted loop without carry-over at 19.8 (a round leaves ~42 entries: one
    // full pass and one with 10 lanes), 690 M warp instructions per pass at 61 % issue utilisation.
    constexpr int kU = 4;   // 32-read groups per round (independent loads in flight per thread)
    __shared__ uint4 dense[8][kU * 32 + 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t i_base = (read_begin / (kU * 32)) * (kU * 32);       // list entries hold read index - i_base (32 bits)
    // a lane's read indices only grow: the chromosome of the previous read is the place to start from
    int c = 0;
    int64_t c_end = __ldg(b.chrom_read_off + 1);

    auto process = [&](const uint4 e) {
        const int64_t i = i_base + (int64_t)e.w;
        const uint32_t m = e.x;
        const uint32_t k_first = e.z;
        while (i >= c_end && c + 1 < b.n_chrom) { ++c; c_end = __ldg(b.chrom_read_off + c + 1); }
        const int64_t base = __ldg(lay.chrom_bin_off + c), clen = __ldg(lay.chrom_len + c);
        const int32_t s = (int32_t)e.y;
        const int L = PB_META_L(m);
        const bool rev = PB_META_REV(m);
        auto in_range = [&](int64_t x) {
            const int64_t t = (base + x) >> tile_shift;
            return t >= tile_lo && t < tile_hi;
        };
        const bool own = in_range(s);                            // the read's start lies in this launch's tiles
        auto emit = [&](int64_t x, int64_t y, uint32_t tag) {
            if (x < 0 || x >= clen) return;
            if (y > clen) y = clen;
            const int64_t tile = (base + x) >> tile_shift;       // tile sizes are powers of two
            if (tile < tile_rec_lo || tile >= tile_hi) return;
            if (!FILL) { atomicAdd(&rec_cursor[tile], 1u); return; }        // result unused: a RED, nothing to wait for
            uint32_t k;
            if (AGG) {
                // reads are coordinate-sorted, so the lanes that are here together mostly target the same
                // tile: one atomic per group of lanes instead of one per record
                const unsigned ln = threadIdx.x & 31;
                const unsigned act = __activemask();
                const unsigned peers = __match_any_sync(act, (unsigned long long)tile);
                const int leader = __ffs(peers) - 1;
                k = 0;
                if ((int)ln == leader) k = atomicAdd(&rec_cursor[tile], (uint32_t)__popc(peers));
                k = __shfl_sync(peers, k, leader) + (uint32_t)__popc(peers & ((1u << ln) - 1u));
            } else {
                k = atomicAdd(&rec_cursor[tile], 1u);
            }
            PbRec rec;
            rec.x = (int32_t)x; rec.y = (int32_t)y; rec.tag = tag; rec.pad = 0;
            recs[__ldg(rec_off + tile) + k] = rec;
        };
        if (CENTER) {
            const int nibble = r.param, map_len = L - 2 * nibble;
            if (map_len < 0) {                                   // map_factories.pyx:246-248
                if (own) { drop_a++; if (rev) drop_m++; else drop_p++; drop_len = L; }
                return;
            }
            if (map_len == 0) return;
            if (own) { map_a++; if (rev) map_m++; else map_p++; }   // reads_out semantics (:256)
            const int slot = (int)__ldg(slot_of_len + L);
            if (slot < 0) return;
            const uint32_t tag = (uint32_t)slot | ((uint32_t)rev << 16);
            const uint32_t k0 = k_first, k1 = k0 + (uint32_t)PB_META_NBLK(m);   // blk lists multi-block reads only
            int a = 0;  // aligned-base index of the block's first base
            for (uint32_t k = k0; k < k1; ++k) {
                const int2 bl = __ldg(b.blk + k);
                const int ia = a > nibble ? a : nibble;
                const int ib = (a + bl.y) < (L - nibble) ? (a + bl.y) : (L - nibble);
                if (ia < ib) emit((int64_t)s + bl.x + (ia - a), (int64_t)s + bl.x + (ib - a), tag);
                a += bl.y;
            }
        } else {
            const int idx_f = pb_rule_index(r, L, false);
            if (idx_f < 0) {                                     // the reference skips the read and warns
                if (own) { drop_a++; if (rev) drop_m++; else drop_p++; drop_len = L; }
                return;
            }
            int64_t p_f = -1, p_r = -1;
            uint32_t tag_f = 0, tag_r = 0;
            if (want_any || (!rev && want_plus)) {
                p_f = pb_block_position(b, i, s, idx_f);
                if (p_f >= 0 && p_f < clen && in_range(p_f)) {
                    if (want_any) { tag_f |= PB_PLANE_ANY; map_a++; }
                    if (!rev && want_plus) { tag_f |= PB_PLANE_PLUS; map_p++; }
                }
            }
            if (rev && want_minus) {
                p_r = pb_block_position(b, i, s, pb_rule_index(r, L, true));
                if (p_r >= 0 && p_r < clen && in_range(p_r)) { tag_r = PB_PLANE_MINUS; map_m++; }
            }
            if (tag_f && tag_r && p_f == p_r) { tag_f |= tag_r; tag_r = 0; }
            if (tag_f) emit(p_f, p_f + 1, tag_f);
            if (tag_r) emit(p_r, p_r + 1, tag_r);
        }
    };

    int n_dense = 0;                               // entries waiting in dense[wid][0 .. n_dense): warp-uniform, < 32 between rounds
    // reads [read_begin, b.n_reads) are looked at (a streamed upload maps a bin range from the reads that can
    // reach it: those before read_begin end before the range, those from b.n_reads on have not arrived)
    for (int64_t q = read_begin / (kU * 32) + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
         q * (kU * 32) < b.n_reads; q += n_warps) {
        const int64_t r0 = q * (kU * 32);
        uint32_t mv[kU], kv[kU];
        int32_t sv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int64_t i = r0 + u * 32 + lane;
            const bool ok = i >= read_begin && i < b.n_reads;
            mv[u] = ok ? __ldg(b.meta + i) : 0u;
            sv[u] = ok ? __ldg(b.ref_start + i) : 0;
            kv[u] = ok ? __ldg(b.blk_off + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            // n_blocks 0 is the out-of-range filler
            const bool work = PB_META_NBLK(mv[u]) > 1 && pb_passes(mv[u], r.size_min, r.size_max);
            const unsigned bal = __ballot_sync(0xffffffffu, work);
            if (work) dense[wid][n_dense + __popc(bal & ((1u << lane) - 1u))] =
                make_uint4(mv[u], (uint32_t)sv[u], kv[u], (uint32_t)(r0 - i_base) + (uint32_t)(u * 32 + lane));
            n_dense += __popc(bal);
        }
        __syncwarp();
        int j = 0;
        for (; j + 32 <= n_dense; j += 32) process(dense[wid][j + lane]);
        const int left = n_dense - j;
        uint4 keep = make_uint4(0u, 0u, 0u, 0u);
        if (lane < left) keep = dense[wid][j + lane];
        __syncwarp();                              // every lane has read its entry before the front is overwritten
        if (j > 0 && lane < left) dense[wid][lane] = keep;
        n_dense = left;
        __syncwarp();
    }
    if (lane < n_dense) process(dense[wid][lane]);
    if (!FILL) pb_flush_cta_stats(drop_p, drop_m, drop_a, drop_len, map_p, map_m, map_a, stat_slots);
}

// exclusive scan of the per-tile record counts in three small launches (4096 counts per CTA, scan
// of the CTA totals, add-back); zeroes the counts so the fill pass can reuse them as cursors.
constexpr int kScanThreads = 1024, kScanPer = 4, kScanChunk = kScanThreads * kScanPer;

__device__ __forceinline__ uint32_t pb_block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += u;
        }
        s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = warp ? s_warp[warp - 1] : 0;
    if (total) *total = s_warp[31];
    return before + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
pb_scan_chunks_kernel(uint32_t *__restrict__ cursor, uint32_t *__restrict__ off, uint32_t *__restrict__ part, int64_t n)
{
    __shared__ uint32_t s_warp[32];
    const int64_t j0 = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanPer;
    uint32_t c[kScanPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) { c[k] = j0 + k < n ? cursor[j0 + k] : 0; sum += c[k]; }
    uint32_t total = 0;
    uint32_t run = pb_block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        if (j0 + k < n) { off[j0 + k] = run; cursor[j0 + k] = 0; }
        run += c[k];
    }
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) pb_scan_top_kernel(uint32_t *__restrict__ part, int64_t nb)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += kScanThreads) {
        const int64_t j = base + threadIdx.x;
        const uint32_t v = j < nb ? part[j] : 0;
        uint32_t total = 0;
        const uint32_t ex = pb_block_exclusive_scan(v, s_warp, &total);
        const uint32_t carry = s_carry;
        if (j < nb) part[j] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) part[nb] = s_carry;   // grand total
}

__global__ void __launch_bounds__(kScanThreads)
pb_scan_add_kernel(uint32_t *__restrict__ off, const uint32_t *__restrict__ part, int64_t n, int64_t nb)
{
    const uint32_t add = part[blockIdx.x];
    const int64_t j0 = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanPer;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)
        if (j0 + k < n) off[j0 + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n] = part[nb];
}

__global__ void pb_stats_finish_kernel(const unsigned long long *__restrict__ slots,
                                       unsigned long long *__restrict__ stats)
{
    int k = threadIdx.x;
    if (k >= PB_NSTATS) return;
    unsigned long long v = 0;
    for (int s = 0; s < kStatSlots; ++s) {
        unsigned long long x = slots[s * PB_NSTATS + k];
        if (k == PB_STAT_DROPPED_LEN) v = x > v ? x : v; else v += x;
    }
    if (k == PB_STAT_DROPPED_LEN) { if (v) stats[k] = v; }
    else stats[k] += v;
}

// optional device timing of the dominant (tiles) kernel, for bench.py's roofline line: a ring of
// CUDA event pairs recorded on the launch stream, summed by pb_tiles_kernel_ms_total().
constexpr int kTimingRing = 256;
bool g_timing = false;
cudaEvent_t g_ev[kTimingRing][2];
int g_ev_created = 0, g_ev_count = 0;

}  // namespace

void pb_timing_begin(cudaStream_t stream)
{
    if (!g_timing || g_ev_count >= kTimingRing) return;
    if (g_ev_count >= g_ev_created) {
        cudaEventCreate(&g_ev[g_ev_created][0]);
        cudaEventCreate(&g_ev[g_ev_created][1]);
        g_ev_created++;
    }
    cudaEventRecord(g_ev[g_ev_count][0], stream);
}

void pb_timing_end(cudaStream_t stream)
{
    if (!g_timing || g_ev_count >= kTimingRing) return;
    cudaEventRecord(g_ev[g_ev_count][1], stream);
    g_ev_count++;
}

extern "C" void pb_enable_kernel_timing(int on) { g_timing = on != 0; g_ev_count = 0; }

extern "C" int pb_tiles_kernel_ms_total(float *ms_total, int *n_launches)
{
    if (!ms_total || !n_launches) { pb_set_error("pb_tiles_kernel_ms_total: null"); return PB_EINVAL; }
    float total = 0.f;
    for (int i = 0; i < g_ev_count; ++i) {
        float ms = 0.f;
        PB_CUDA_CHECK(cudaEventSynchronize(g_ev[i][1]));
        PB_CUDA_CHECK(cudaEventElapsedTime(&ms, g_ev[i][0], g_ev[i][1]));
        total += ms;
    }
    *ms_total = total;
    *n_launches = g_ev_count;
    return PB_OK;
}

size_t pb_ws_tile_bytes(int64_t total_bins) { return (size_t)(total_bins / 1024 + 1) * sizeof(PbTile); }
size_t pb_ws_stat_bytes() { return (size_t)kStatSlots * PB_NSTATS * sizeof(unsigned long long); }
static size_t ws_part_bytes(int64_t total_bins) { return (((size_t)(total_bins / 1024 / 4096 + 4) * sizeof(uint32_t)) + 255) & ~(size_t)255; }
static size_t ws_idx_bytes(int64_t total_bins) { return (((size_t)(total_bins / 1024 + 2) * sizeof(uint32_t)) + 255) & ~(size_t)255; }

// every candidate read appears in at most two tiles' windows, so sum(n)/split + one per tile bounds
// the overflow jobs; split is never below 4096 reads.  The Center rule also cuts the binned records of a tile
// into jobs (a record is seen by 1 + lookback tiles); a pile-up tile whose jobs do not fit is walked whole.
static int64_t ws_job_capacity(int64_t total_bins, int64_t n_reads, int64_t n_blk) { return 2 * n_reads / 4096 + 4 * n_blk / 4096 + 1024; }
// Center rule, pile-up tiles: integer partial difference arrays (16-100 KB per tile); tiles beyond the scratch are
// walked whole by their CTA
static size_t ws_hot_bytes(int64_t n_reads, int64_t n_blk)
{
    const size_t want = ((size_t)2 << 20) + (size_t)(n_reads + n_blk) * 4;
    return (want < ((size_t)64 << 20) ? want : ((size_t)64 << 20)) & ~(size_t)255;
}

extern "C" size_t pb_map_workspace_bytes(int64_t total_bins, int64_t n_blk, int64_t n_reads)
{
    if (total_bins < 0 || n_blk < 0 || n_reads < 0) return 0;
    size_t bytes = pb_ws_tile_bytes(total_bins) + 2 * pb_ws_stat_bytes() + 256;
    bytes += (size_t)ws_job_capacity(total_bins, n_reads, n_blk) * sizeof(PbJob);
    bytes += ws_hot_bytes(n_reads, n_blk) + 256;
    if (n_blk > 0) bytes += 2 * ws_idx_bytes(total_bins) + ws_part_bytes(total_bins) + (size_t)n_blk * sizeof(PbRec);
    return bytes + 256;
}

int pb_carve_workspace(void *base, size_t bytes, int64_t total_bins, int64_t n_blk, int64_t n_reads, PbWorkspace *ws)
{
    if (!base || bytes < pb_map_workspace_bytes(total_bins, n_blk, n_reads)) { pb_set_error("workspace too small"); return PB_ENOSPACE; }
    if (n_blk >= 0xffffffffll) { pb_set_error("more than 2^32-1 block rows in one batch"); return PB_EINVAL; }
    char *p = (char *)base;
    ws->tiles = (PbTile *)p;                     p += pb_ws_tile_bytes(total_bins);
    ws->slots = (unsigned long long *)p;         p += 2 * pb_ws_stat_bytes();
    ws->tile_counter = (unsigned long long *)p;  p += 256;
    ws->job_capacity = ws_job_capacity(total_bins, n_reads, n_blk);
    ws->jobs = (PbJob *)p;                       p += (size_t)ws->job_capacity * sizeof(PbJob);
    p = (char *)(((uintptr_t)p + 255) & ~(uintptr_t)255);
    ws->hot = p;                                 ws->hot_bytes = ws_hot_bytes(n_reads, n_blk);
    p += ws->hot_bytes;
    ws->rec_off = ws->rec_cursor = ws->scan_part = nullptr;
    ws->recs = nullptr;
    if (n_blk > 0) {
        ws->rec_off = (uint32_t *)p;             p += ws_idx_bytes(total_bins);
        ws->rec_cursor = (uint32_t *)p;          p += ws_idx_bytes(total_bins);
        ws->scan_part = (uint32_t *)p;           p += ws_part_bytes(total_bins);
        ws->recs = (PbRec *)p;
    }
    return PB_OK;
}

PbReads pb_to_dev(const pb_batch *b)
{
    PbReads d;
    d.ref_start = b->ref_start;
    d.meta = b->meta;
    d.blk_off = b->blk_off;
    d.blk = reinterpret_cast<const int2 *>(b->blk);
    d.chrom_read_off = b->chrom_read_off;
    d.n_reads = b->n_reads;
    d.n_blk = b->blk ? b->n_blk : 0;
    d.n_chrom = b->n_chrom;
    d.max_span = b->max_span < 1 ? 1 : b->max_span;
    d.max_block_len = b->max_block_len < 1 ? d.max_span : b->max_block_len;
    return d;
}

PbRuleDev pb_to_dev(const pb_rule *r)
{
    PbRuleDev d;
    d.kind = r->kind; d.param = r->param;
    d.lut_fw = r->lut_fw; d.lut_rc = r->lut_rc;
    d.size_min = r->size_min; d.size_max = r->size_max;
    d.strat_min = r->strat_min; d.strat_max = r->strat_max;
    return d;
}

int pb_check_common(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes)
{
    if (!batch || !layout || !rule) { pb_set_error("null batch/layout/rule"); return PB_EINVAL; }
    if (planes <= 0 || planes > 7) { pb_set_error("planes must be a non-empty mask of PB_PLANE_*"); return PB_EINVAL; }
    if (batch->n_reads < 0 || (batch->n_reads > 0 && (!batch->ref_start || !batch->meta))) {
        pb_set_error("batch arrays missing"); return PB_EINVAL;
    }
    if (batch->n_chrom != layout->n_chrom || !batch->chrom_read_off) {
        pb_set_error("batch/layout chromosome tables disagree"); return PB_EINVAL;
    }
    if (layout->total_bins <= 0 || layout->total_bins % PB_LAYOUT_ALIGN) {
        pb_set_error("layout.total_bins must be a positive multiple of PB_LAYOUT_ALIGN"); return PB_EINVAL;
    }
    if ((batch->blk_off == nullptr) != (batch->blk == nullptr)) {
        pb_set_error("blk_off and blk must both be given or both be NULL"); return PB_EINVAL;
    }
    if (batch->blk && batch->n_blk < 0) { pb_set_error("negative n_blk"); return PB_EINVAL; }
    return PB_OK;
}

int pb_sm_count(int *out)
{
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        PB_CUDA_CHECK(cudaGetDevice(&dev));
        PB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    *out = sm_count;
    return PB_OK;
}

int pb_launch_tile_index(const PbReads &b, const PbLayoutDev &lay, int tile_bins, int64_t tile_begin, int64_t tile_end,
                         int64_t read_limit, int split, const PbWorkspace &ws, cudaStream_t stream)
{
    if (tile_end <= tile_begin) return PB_OK;
    if (split > 0 && split < 4096) split = 4096;
    pb_tile_index_kernel<<<(unsigned)((tile_end - tile_begin + 255) / 256), 256, 0, stream>>>(
        b, lay, tile_bins, tile_begin, tile_end, read_limit, split, ws.tiles, ws.jobs, ws.job_capacity, ws.tile_counter + 1);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

int pb_launch_binning(const PbReads &b, const PbRuleDev &r, const PbLayoutDev &lay, int planes, int center,
                      const int16_t *slot_of_len, int tile_bins, int64_t n_tiles, int64_t tile_lo, int64_t tile_hi,
                      int64_t read_begin, int64_t read_limit, const PbWorkspace &ws, cudaStream_t stream)
{
    if (!b.blk_off || b.n_blk <= 0 || b.n_reads == 0) return PB_OK;
    if (read_begin < 0) read_begin = 0;
    if (read_limit < 0 || read_limit > b.n_reads) read_limit = b.n_reads;
    PbReads bw = b;                       // the kernel's loop bound: reads from read_limit on have not arrived
    bw.n_reads = read_limit;
    int sms = 0;
    int rc = pb_sm_count(&sms);
    if (rc) return rc;
    int64_t want = (read_limit - read_begin + 255) / 256;
    if (want < 1) want = 1;
    unsigned grid = (unsigned)(want < (int64_t)sms * 32 ? want : (int64_t)sms * 32);
    PB_CUDA_CHECK(cudaMemsetAsync(ws.rec_cursor, 0, (size_t)(n_tiles + 1) * sizeof(uint32_t), stream));
    int tile_shift = 0;
    while ((1 << tile_shift) < tile_bins) ++tile_shift;
    if ((1 << tile_shift) != tile_bins) { pb_set_error("tile size must be a power of two"); return PB_EINVAL; }
    int64_t tile_rec_lo = tile_lo;
    if (center) {
        const int64_t lookback = ((int64_t)b.max_block_len + tile_bins - 1) / tile_bins;
        tile_rec_lo = tile_lo > lookback ? tile_lo - lookback : 0;
    }
    if (b.n_reads - (read_begin / 128) * 128 > 0xffffffffll) { pb_set_error("binning: more than 2^32 reads in one launch"); return PB_EINVAL; }
    // PB_BIN_AGG=1 (A/B aid): the fill pass shares one cursor atomic among the lanes that target the same tile
    const char *env_agg = getenv("PB_BIN_AGG");
    const bool agg = env_agg && atoi(env_agg) != 0;
    // PB_BIN_OCC=8 (A/B aid): 32 registers per thread (a few spills) for eight resident CTAs per SM instead of six
    const char *env_occ = getenv("PB_BIN_OCC");
    const int occ_sel = env_occ ? atoi(env_occ) : 6;
    for (int fill = 0; fill < 2; ++fill) {
#define PB_BIN_LAUNCH_O(C_, O_, F_, A_) pb_bin_kernel<C_, O_, F_, A_><<<grid, 256, 0, stream>>>(bw, r, lay, planes, slot_of_len, tile_shift, tile_lo, tile_hi, tile_rec_lo, read_begin, ws.rec_cursor, ws.rec_off, ws.recs, ws.slots)
#define PB_BIN_LAUNCH(C_, F_, A_) do { if (occ_sel == 8) PB_BIN_LAUNCH_O(C_, 8, F_, A_); else PB_BIN_LAUNCH_O(C_, 6, F_, A_); } while (0)
        if (center) {
            if (!fill) PB_BIN_LAUNCH(true, false, false); else if (agg) PB_BIN_LAUNCH(true, true, true); else PB_BIN_LAUNCH(true, true, false);
        } else {
            if (!fill) PB_BIN_LAUNCH(false, false, false); else if (agg) PB_BIN_LAUNCH(false, true, true); else PB_BIN_LAUNCH(false, true, false);
        }
#undef PB_BIN_LAUNCH_O
#undef PB_BIN_LAUNCH
        if (!fill) {
            const int64_t nb = (n_tiles + kScanChunk - 1) / kScanChunk;
            pb_scan_chunks_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(ws.rec_cursor, ws.rec_off, ws.scan_part, n_tiles);
            pb_scan_top_kernel<<<1, kScanThreads, 0, stream>>>(ws.scan_part, nb);
            pb_scan_add_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(ws.rec_off, ws.scan_part, n_tiles, nb);
        }
    }
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

int64_t pb_scan_part_entries(int64_t n) { return (n + kScanChunk - 1) / kScanChunk + 1; }

// off[0..n] = exclusive prefix sums of counts[0..n) (off[n] = total); counts are zeroed on the way
int pb_launch_exclusive_scan_u32(uint32_t *counts, uint32_t *off, uint32_t *part, int64_t n, cudaStream_t stream)
{
    const int64_t nb = (n + kScanChunk - 1) / kScanChunk;
    if (nb == 0) { PB_CUDA_CHECK(cudaMemsetAsync(off, 0, sizeof(uint32_t), stream)); return PB_OK; }
    pb_scan_chunks_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(counts, off, part, n);
    pb_scan_top_kernel<<<1, kScanThreads, 0, stream>>>(part, nb);
    pb_scan_add_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(off, part, n, nb);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

int pb_launch_stats_finish(const unsigned long long *slots, unsigned long long *stats, cudaStream_t stream)
{
    pb_stats_finish_kernel<<<1, 32, 0, stream>>>(slots, stats);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
