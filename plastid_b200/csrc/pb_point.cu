// pb_point.cu — 5' / 3' / variable-offset mapping into dense uint32 planes (sm_100a).
//
// Design ("owner computes", no global atomics, no memset pass): the concatenated genome is cut into
// 4096-bin tiles; persistent CTAs (one resident wave) pull tiles from an atomic queue.  A tile's
// bins for every requested query strand live in shared memory; the CTA scans only the slice of the
// coordinate-sorted batch whose single-block reads can land in the tile (pb_tile_index_kernel) plus
// the tile's bucket of binned multi-block records (pb_bin_kernel), applies the mapping rule per
// read, accumulates with shared-memory integer atomics (order independent => bit-exact and
// deterministic) and hands the finished tile to the TMA engine as one bulk store per plane.
// Tiles nothing can land in are a bulk store of the (already zero) buffer: the launch is also the
// memset.
//
// Reference semantics restated (plastid/genomics/map_factories.pyx): FivePrime :308-367,
// ThreePrime :407-466, VariableFivePrime :585-650, SizeFilter :837-839; strand selection as
// genome_array.py:811-815.  Query strand '.' applies the FORWARD rule to reads of both strands, so it
// is its own plane, not '+' + '-'.
#include "pb_tiles.cuh"
#include <stdlib.h>

namespace {

constexpr unsigned kNoKey = 0xffffffffu;

// Measured (profiles/NOTES_r01.md): grouping lanes by target word with match.any before the shared
// atomic is SLOWER than the plain atomic on both C2 (4.43 -> 5.01 ms) and C5 (8.83 -> 11.33 ms):
// shared atomics cost ~2 cycles per active lane whether or not addresses collide, and MATCH.ANY
// costs more than it saves.  So: one plain shared atomic per mapped read.
__device__ __forceinline__ void pb_smem_inc(uint32_t *smem, unsigned key)
{
    if (key != kNoKey) atomicAdd(&smem[key], 1u);
}

constexpr int kPThreads = 256;     // threads per persistent CTA
constexpr int kPTileBins = 4096;   // bins per tile: 16 KB per plane in shared memory
constexpr int kPUnroll = 4;        // independent read loads in flight per thread
constexpr int kPSplit = 8192;      // candidate reads one tile job scans; the rest become overflow jobs

struct PbCounters {
    unsigned long long drop_p, drop_m, drop_a, map_p, map_m, map_a;
    unsigned int drop_len;
};

struct PbPlaneBases {
    unsigned plus, minus, any;     // word offsets of the planes inside the shared tile buffer
    bool want_plus, want_minus, want_any;
};

// Apply the rule to the single-block reads [lo, hi) of the sorted batch and add their sites that fall
// into the tile [p0, plim) to the shared tile buffer.  Trip count is CTA-uniform.
__device__ __forceinline__ void pb_scan_reads(const PbReads &b, const PbRuleDev &r, bool skip_multi, int64_t lo,
                                              int64_t hi, int64_t p0, int64_t plim, int64_t p1, uint32_t *smem,
                                              const PbPlaneBases &pl, PbCounters &c)
{
    for (int64_t base = lo; base < hi; base += (int64_t)kPUnroll * kPThreads) {
        int32_t sv[kPUnroll];
        uint32_t mv[kPUnroll];
#pragma unroll
        for (int u = 0; u < kPUnroll; ++u) {
            const int64_t i = base + (int64_t)u * kPThreads + threadIdx.x;
            const bool ok = i < hi;
            sv[u] = ok ? __ldg(b.ref_start + i) : 0;
            mv[u] = ok ? __ldg(b.meta + i) : (1u << 17);   // drop bit: skipped below
        }
#pragma unroll
        for (int u = 0; u < kPUnroll; ++u) {
            const int32_t s = sv[u];
            const uint32_t m = mv[u];
            const int L = PB_META_L(m);
            const bool rev = PB_META_REV(m);
            unsigned key_strand = kNoKey, key_any = kNoKey;   // word index into smem, or none
            // multi-block reads were mapped by pb_bin_kernel and arrive through the tile's bucket
            if (pb_passes(m, r.size_min, r.size_max) && !(skip_multi && PB_META_NBLK(m) > 1)) {
                const int idx_f = pb_rule_index(r, L, false);
                if (idx_f < 0) {
                    // the reference skips this read and warns; count it once, in the tile owning its start
                    if (s >= p0 && s < p1) {
                        c.drop_a++;
                        if (rev) c.drop_m++; else c.drop_p++;
                        c.drop_len = L;
                    }
                } else {
                    if (pl.want_any || (!rev && pl.want_plus)) {
                        const int64_t p = (int64_t)s + idx_f;
                        if (p >= p0 && p < plim) {
                            const unsigned o = (unsigned)(p - p0);
                            if (pl.want_any) { key_any = pl.any + o; c.map_a++; }
                            if (!rev && pl.want_plus) { key_strand = pl.plus + o; c.map_p++; }
                        }
                    }
                    if (rev && pl.want_minus) {
                        const int64_t p = (int64_t)s + pb_rule_index(r, L, true);
                        if (p >= p0 && p < plim) {
                            key_strand = pl.minus + (unsigned)(p - p0);
                            c.map_m++;
                        }
                    }
                }
            }
            pb_smem_inc(smem, key_strand);
            pb_smem_inc(smem, key_any);
        }
    }
}

__device__ __forceinline__ PbPlaneBases pb_plane_bases(int planes)
{
    PbPlaneBases pl;
    pl.want_plus = planes & PB_PLANE_PLUS; pl.want_minus = planes & PB_PLANE_MINUS; pl.want_any = planes & PB_PLANE_ANY;
    unsigned k = 0;
    pl.plus = pl.minus = pl.any = 0;
    if (pl.want_plus) pl.plus = (k++) * kPTileBins;
    if (pl.want_minus) pl.minus = (k++) * kPTileBins;
    if (pl.want_any) pl.any = (k++) * kPTileBins;
    return pl;
}

// Invariant: at the top of every loop iteration the shared tile buffer is all zero and visible to
// the async proxy.  Empty tiles are therefore one bulk store of the buffer as it is; tiles with
// reads accumulate into it, store it, wait until the TMA engine has READ it (not until the write
// has landed), and re-zero it.  The SM never touches the output bytes itself.
__global__ void __launch_bounds__(kPThreads)
pb_point_tiles_kernel(PbReads b, PbRuleDev r, int planes, const PbTile *__restrict__ tiles, int64_t tile_begin,
                      int64_t n_tiles, unsigned long long *__restrict__ tile_counter,
                      const uint32_t *__restrict__ rec_off, const PbRec *__restrict__ recs,
                      uint32_t *__restrict__ out_plus, uint32_t *__restrict__ out_minus,
                      uint32_t *__restrict__ out_any, unsigned long long *__restrict__ stat_slots)
{
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ long long s_q[3];

    const PbPlaneBases pl = pb_plane_bases(planes);
    const int n_planes = (int)pl.want_plus + (int)pl.want_minus + (int)pl.want_any;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 *smem4 = reinterpret_cast<uint4 *>(smem);
    for (int j = threadIdx.x; j < n_planes * kPTileBins / 4; j += kPThreads) smem4[j] = zero4;
    if (threadIdx.x == 0) {
        s_q[0] = tile_begin + (long long)atomicAdd(tile_counter, 1ull);
        s_q[1] = tile_begin + (long long)atomicAdd(tile_counter, 1ull);
    }
    pb_fence_proxy_async();
    __syncthreads();

    PbCounters c = {0, 0, 0, 0, 0, 0, 0};

    // Look-ahead tile queue: the CTA always knows its current and its next tile.  Thread 0 claims the
    // tile for iteration k+2 at the top of iteration k and only publishes it at the end, so the atomic's
    // round trip, the next descriptor's load and the L2 prefetch of the next tile's reads all overlap
    // with the current tile — the chain of dependent DRAM round trips per tile is what bounds a CTA.
    // (Claiming runs of consecutive tiles instead was measured slower on skewed data: hot tiles are
    // neighbours, and a run lands on one CTA.)
#ifdef PB_POINT_SIMPLE_QUEUE
    for (int k = 0;; ++k) {
        const long long tile = s_q[k & 1];
        if (tile >= n_tiles) break;
        if (threadIdx.x == 0) s_q[(k + 1) & 1] = tile_begin + (long long)atomicAdd(tile_counter, 1ull);
        const PbTile d = tiles[tile];
        const long long tile_nxt = n_tiles;
        const PbTile d_nxt = {0, 0, 0, 0, 0, 0};
#else
    long long tile = s_q[0], tile_nxt = s_q[1];
    PbTile d = {0, 0, 0, 0, 0, 0}, d_nxt = {0, 0, 0, 0, 0, 0};
    if (tile < n_tiles) d = tiles[tile];
    if (tile_nxt < n_tiles) d_nxt = tiles[tile_nxt];
    for (int k = 0; tile < n_tiles; ++k) {
        long long claimed = 0;
        if (threadIdx.x == 0) claimed = tile_begin + (long long)atomicAdd(tile_counter, 1ull);
#endif
        const int64_t g0 = tile * kPTileBins;
        uint32_t rec_lo = 0, rec_hi = 0;       // this tile's bucket of binned multi-block sites
        if (rec_off) { rec_lo = __ldg(rec_off + tile); rec_hi = __ldg(rec_off + tile + 1); }
        const bool has_work = d.n > 0 || rec_hi > rec_lo;

        if (has_work) {
            if (threadIdx.x == 0) pb_bulk_wait_read0();   // earlier stores of the zero buffer have read it
            __syncthreads();
            const int64_t p0 = d.p0;
            pb_scan_reads(b, r, rec_off != nullptr, d.lo, d.lo + d.n, p0, p0 + d.live, p0 + kPTileBins, smem, pl, c);
            if (tile_nxt < n_tiles) pb_prefetch_reads_l2(b, d_nxt);
            for (uint32_t j = rec_lo + threadIdx.x; j < rec_hi; j += kPThreads) {
                const PbRec rec = recs[j];      // site already bounds-checked and counted by pb_bin_kernel
                const unsigned o = (unsigned)((int64_t)rec.x - p0);
                if (rec.tag & PB_PLANE_PLUS) atomicAdd(&smem[pl.plus + o], 1u);
                if (rec.tag & PB_PLANE_MINUS) atomicAdd(&smem[pl.minus + o], 1u);
                if (rec.tag & PB_PLANE_ANY) atomicAdd(&smem[pl.any + o], 1u);
            }
            pb_fence_proxy_async();
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            int q = 0;
            if (pl.want_plus) pb_bulk_store(out_plus + g0, smem + (q++) * kPTileBins, kPTileBins * 4);
            if (pl.want_minus) pb_bulk_store(out_minus + g0, smem + (q++) * kPTileBins, kPTileBins * 4);
            if (pl.want_any) pb_bulk_store(out_any + g0, smem + (q++) * kPTileBins, kPTileBins * 4);
            pb_bulk_commit();
            if (has_work) pb_bulk_wait_read0();
#ifndef PB_POINT_SIMPLE_QUEUE
            s_q[(k + 2) % 3] = claimed;
#endif
        }
        __syncthreads();
#ifndef PB_POINT_SIMPLE_QUEUE
        const long long tile_nn = s_q[(k + 2) % 3];
        PbTile d_nn = {0, 0, 0, 0, 0, 0};
        if (tile_nn < n_tiles) d_nn = tiles[tile_nn];     // in flight while the buffer is re-zeroed
#endif
        if (has_work) {
            for (int j = threadIdx.x; j < n_planes * kPTileBins / 4; j += kPThreads) smem4[j] = zero4;
            pb_fence_proxy_async();
            __syncthreads();
        }
#ifndef PB_POINT_SIMPLE_QUEUE
        tile = tile_nxt; d = d_nxt;
        tile_nxt = tile_nn; d_nxt = d_nn;
#endif
    }

    pb_flush_cta_stats(c.drop_p, c.drop_m, c.drop_a, c.drop_len, c.map_p, c.map_m, c.map_a, stat_slots);
    if (threadIdx.x == 0) pb_bulk_wait_all();
}

// Overflow jobs: the candidate reads of a tile beyond the first kPSplit (pile-ups on highly expressed
// genes put millions of reads into one tile, and one CTA walking them alone would set the kernel's
// duration).  Runs after the tiles kernel on the same stream, so every plane already holds its tile;
// each job accumulates its slice in shared memory and ADDS it with a TMA bulk reduction
// (cp.reduce.async.bulk ... .add.u32, SASS UBLKRED) — integer adds commute, so the result stays
// bit-exact and deterministic.
__global__ void __launch_bounds__(kPThreads)
pb_point_overflow_kernel(PbReads b, PbRuleDev r, int planes, const PbTile *__restrict__ tiles, int skip_multi,
                         const PbJob *__restrict__ jobs, const unsigned long long *__restrict__ n_jobs_p,
                         unsigned long long *__restrict__ job_counter,
                         uint32_t *__restrict__ out_plus, uint32_t *__restrict__ out_minus,
                         uint32_t *__restrict__ out_any, unsigned long long *__restrict__ stat_slots)
{
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ long long s_job;
    const long long n_jobs = (long long)*n_jobs_p;
    if (n_jobs == 0) return;
    const PbPlaneBases pl = pb_plane_bases(planes);
    const int n_planes = (int)pl.want_plus + (int)pl.want_minus + (int)pl.want_any;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 *smem4 = reinterpret_cast<uint4 *>(smem);
    PbCounters c = {0, 0, 0, 0, 0, 0, 0};
    for (int j = threadIdx.x; j < n_planes * kPTileBins / 4; j += kPThreads) smem4[j] = zero4;
    // Invariant: the tile buffer is all zero at the top of every iteration.  A job's reads are a run of the sorted
    // batch, so their sites lie in [first start, last start + longest block): only that stretch of the tile is added to
    // the planes and re-zeroed.  A pile-up's jobs (hundreds on one tile) then reduce a few hundred bytes each into the
    // same lines of L2 instead of the whole 16 KB tile per plane.
    for (;;) {
        if (threadIdx.x == 0) s_job = (long long)atomicAdd(job_counter, 1ull);
        __syncthreads();
        const long long job = s_job;
        if (job >= n_jobs) break;
        const PbJob jb = jobs[job];
        const PbTile d = tiles[jb.tile];
        long long o_lo = (long long)__ldg(b.ref_start + jb.lo) - d.p0;
        long long o_hi = (long long)__ldg(b.ref_start + jb.lo + jb.n - 1) + b.max_block_len - d.p0;
        o_lo = o_lo < 0 ? 0 : (o_lo & ~3ll);                       // bulk copies move 16-byte units
        o_hi = o_hi > kPTileBins ? kPTileBins : ((o_hi + 3) & ~3ll);
        pb_scan_reads(b, r, skip_multi != 0, jb.lo, jb.lo + jb.n, d.p0, d.p0 + d.live, d.p0 + kPTileBins, smem, pl, c);
        if (o_lo < o_hi) {
            pb_fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) {
                const int64_t g0 = jb.tile * kPTileBins + o_lo;
                const uint32_t bytes = (uint32_t)(o_hi - o_lo) * 4u;
                int q = 0;
                if (pl.want_plus) pb_bulk_add_u32(out_plus + g0, smem + (q++) * kPTileBins + o_lo, bytes);
                if (pl.want_minus) pb_bulk_add_u32(out_minus + g0, smem + (q++) * kPTileBins + o_lo, bytes);
                if (pl.want_any) pb_bulk_add_u32(out_any + g0, smem + (q++) * kPTileBins + o_lo, bytes);
                pb_bulk_commit();
                pb_bulk_wait_read0();
            }
            __syncthreads();
            const int w4 = (int)((o_hi - o_lo) >> 2);
            for (int j = threadIdx.x; j < n_planes * w4; j += kPThreads) {
                const int q = j / w4, k = j - q * w4;
                smem4[(q * kPTileBins + (int)o_lo) / 4 + k] = zero4;
            }
        }
        __syncthreads();                                           // s_job is rewritten at the top
    }
    pb_flush_cta_stats(c.drop_p, c.drop_m, c.drop_a, c.drop_len, c.map_p, c.map_m, c.map_a, stat_slots);
    if (threadIdx.x == 0) pb_bulk_wait_all();
}

}  // namespace

extern "C" int pb_map_point_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                                  uint32_t *out_plus, uint32_t *out_minus, uint32_t *out_any,
                                  uint64_t *stats, void *workspace, size_t workspace_bytes,
                                  int64_t bin_begin, int64_t bin_end, int64_t read_limit, void *stream_)
{
    int rc = pb_check_common(batch, layout, rule, planes);
    if (rc) return rc;
    if (rule->kind != PB_RULE_FIVEPRIME && rule->kind != PB_RULE_THREEPRIME && rule->kind != PB_RULE_VARIABLE) {
        pb_set_error("pb_map_point: rule kind %d is not a point rule", rule->kind); return PB_EINVAL;
    }
    if (rule->kind == PB_RULE_VARIABLE && (!rule->lut_fw || !rule->lut_rc)) {
        pb_set_error("pb_map_point: variable rule needs lut_fw/lut_rc"); return PB_EINVAL;
    }
    if (rule->kind != PB_RULE_VARIABLE && rule->param < 0) {
        pb_set_error("pb_map_point: offset must be >= 0"); return PB_EINVAL;
    }
    if (((planes & PB_PLANE_PLUS) && !out_plus) || ((planes & PB_PLANE_MINUS) && !out_minus) ||
        ((planes & PB_PLANE_ANY) && !out_any) || !stats) {
        pb_set_error("pb_map_point: missing output plane or stats"); return PB_EINVAL;
    }
    if ((((uintptr_t)out_plus | (uintptr_t)out_minus | (uintptr_t)out_any) & 15) ||
        ((uintptr_t)workspace & 15)) {      // TMA bulk stores move 16-byte units
        pb_set_error("pb_map_point: planes and workspace must be 16-byte aligned"); return PB_EINVAL;
    }
    if (bin_begin < 0 || bin_end > layout->total_bins || bin_begin > bin_end || bin_begin % PB_LAYOUT_ALIGN ||
        bin_end % PB_LAYOUT_ALIGN) {
        pb_set_error("pb_map_point_range: bin range must be PB_LAYOUT_ALIGN-aligned and inside the layout"); return PB_EINVAL;
    }
    if (read_limit < 0 || read_limit > batch->n_reads) read_limit = batch->n_reads;
    if (bin_begin == bin_end) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t tile_begin = bin_begin / kPTileBins, n_tiles = bin_end / kPTileBins;   // n_tiles = end of range
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    PbWorkspace ws;
    rc = pb_carve_workspace(workspace, workspace_bytes, layout->total_bins, b.n_blk, b.n_reads, &ws);
    if (rc) return rc;
    if (b.n_blk > 0 && read_limit != batch->n_reads) {
        // binning walks every read of the batch; streamed uploads are unspliced by format (wire16 / delta8)
        pb_set_error("pb_map_point_range: batches with multi-block reads need every read resident (read_limit == n_reads)");
        return PB_EINVAL;
    }

    PB_CUDA_CHECK(cudaMemsetAsync(ws.slots, 0, 2 * pb_ws_stat_bytes() + 64, stream));
    // PB_POINT_SPLIT=<reads> (A/B aid): candidate reads a tile keeps before the rest become overflow jobs
    static const int split = [] { const char *e = getenv("PB_POINT_SPLIT"); int v = e ? atoi(e) : kPSplit; return v >= 2048 ? v : kPSplit; }();
    rc = pb_launch_tile_index(b, lay, kPTileBins, tile_begin, n_tiles, read_limit, split, ws, stream);
    if (rc) return rc;
    // multi-block (spliced) reads: map them once and bin their sites by tile
    rc = pb_launch_binning(b, r, lay, planes, 0, nullptr, kPTileBins, layout->total_bins / kPTileBins, tile_begin, n_tiles,
                           0, b.n_reads, ws, stream);
    if (rc) return rc;
    const int n_planes = __builtin_popcount(planes);
    const size_t smem = (size_t)n_planes * kPTileBins * sizeof(uint32_t);
    int sm_count = 0;
    rc = pb_sm_count(&sm_count);
    if (rc) return rc;
    PB_CUDA_CHECK(cudaFuncSetAttribute(pb_point_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pb_point_tiles_kernel, kPThreads, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sm_count * occ;     // persistent: one resident wave, tiles come from a queue
    if (grid > n_tiles - tile_begin) grid = n_tiles - tile_begin;
    pb_timing_begin(stream);
    pb_point_tiles_kernel<<<(unsigned)grid, kPThreads, smem, stream>>>(b, r, planes, ws.tiles, tile_begin, n_tiles,
                                                                      ws.tile_counter, ws.rec_off, ws.recs,
                                                                      out_plus, out_minus, out_any, ws.slots);
    // slices of pile-up tiles beyond kPSplit reads: added on top of the stored tiles (exits at once
    // when the tile index found none)
    PB_CUDA_CHECK(cudaFuncSetAttribute(pb_point_overflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pb_point_overflow_kernel, kPThreads, smem));
    if (occ < 1) occ = 1;
    pb_point_overflow_kernel<<<(unsigned)(sm_count * occ), kPThreads, smem, stream>>>(
        b, r, planes, ws.tiles, b.n_blk > 0, ws.jobs, ws.tile_counter + 1, ws.tile_counter + 2,
        out_plus, out_minus, out_any, ws.slots);
    pb_timing_end(stream);
    PB_CUDA_CHECK(cudaGetLastError());
    return pb_launch_stats_finish(ws.slots, (unsigned long long *)stats, stream);
}

extern "C" int pb_map_point(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                            uint32_t *out_plus, uint32_t *out_minus, uint32_t *out_any,
                            uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!layout || !batch) { pb_set_error("null batch/layout/rule"); return PB_EINVAL; }
    return pb_map_point_range(batch, layout, rule, planes, out_plus, out_minus, out_any, stats, workspace,
                              workspace_bytes, 0, layout->total_bins, batch->n_reads, stream_);
}

