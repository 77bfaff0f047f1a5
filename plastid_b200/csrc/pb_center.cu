// pb_center.cu — CenterMapFactory into dense float64 planes (sm_100a), deterministic.
//
// Reference: CenterMapFactory.__call__ (plastid/genomics/map_factories.pyx:200-265) adds 1.0/m to
// each of the m = L - 2*nibble trimmed aligned positions of every read, in read order.  Here, per
// 1024/2048-bin tile and per distinct map length ("slot"), reads are accumulated as an INTEGER
// difference array in shared memory (+1 at the first bin of each trimmed aligned interval, -1 one
// past its end); an exact warp-shuffle prefix scan turns it into "number of reads of that map length
// covering the bin", and the bin value is sum over slots, in ascending map length, of
// cover * (1/m).  Integer accumulation commutes, the combine order is fixed => run-to-run identical
// results, exact zeros where no read lies.
//
// Same skeleton as pb_point.cu: persistent CTAs pull tiles from an atomic queue; single-block reads
// come from the sorted candidate slice of the tile, the intervals of multi-block (spliced) reads
// from the per-tile buckets built by pb_bin_kernel (so the candidate window is the longest aligned
// block, not the longest intron); finished tiles are staged as fp64 in shared memory and leave the
// SM as TMA bulk stores (or bulk fp64 reductions when map lengths need more than one pass).
#include "pb_tiles.cuh"

namespace {

constexpr int kCThreads = 256;
constexpr int kCWarps = kCThreads / 32;
constexpr int kCUnroll = 4;
constexpr int kZeroBins = 256;   // 2 KB of fp64 zeros: tiles nothing lands in are stored from here

template <int EPT>  // bins per thread; tile = EPT * 256 bins
__global__ void __launch_bounds__(kCThreads)
pb_center_tiles_kernel(PbReads b, PbRuleDev r, int planes,
                       const int16_t *__restrict__ slot_of_len, const double *__restrict__ inv_m,
                       int slot0, int n_slots, int accumulate, int lookback,
                       const PbTile *__restrict__ tiles, int64_t n_tiles, unsigned long long *__restrict__ tile_counter,
                       const uint32_t *__restrict__ rec_off, const PbRec *__restrict__ recs,
                       double *__restrict__ out_plus, double *__restrict__ out_minus, double *__restrict__ out_any,
                       unsigned long long *__restrict__ stat_slots)
{
    constexpr int T = EPT * kCThreads;
    constexpr int seg = T / kCWarps;  // bins per warp in the scan = EPT * 32
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ PbSlot s_ring[4];

    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS,
               want_any = planes & PB_PLANE_ANY;
    const int n_planes = (int)want_plus + (int)want_minus + (int)want_any;
    const int n_arrays = n_planes * n_slots;
    double *stage = reinterpret_cast<double *>(smem_raw);           // [n_planes][T] fp64 staging
    double *zbuf = stage + (size_t)n_planes * T;                    // [kZeroBins] zeros, never written
    int *diff = reinterpret_cast<int *>(zbuf + kZeroBins);          // [n_arrays][T]
    int *warp_tot = diff + (size_t)n_arrays * T;                    // [2][n_arrays][kCWarps] (double-buffered by tile parity)
    double *outs[3];
    int *d_plus = diff, *d_minus = diff, *d_any = diff;
    {
        int k = 0;
        if (want_plus) { outs[k] = out_plus; d_plus = diff + (size_t)(k++) * n_slots * T; }
        if (want_minus) { outs[k] = out_minus; d_minus = diff + (size_t)(k++) * n_slots * T; }
        if (want_any) { outs[k] = out_any; d_any = diff + (size_t)(k++) * n_slots * T; }
    }
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4 *s4 = reinterpret_cast<uint4 *>(smem_raw);
        const int n16 = ((n_planes * T + kZeroBins) * 8 + n_arrays * T * 4 + 2 * n_arrays * kCWarps * 4) / 16;
        for (int j = threadIdx.x; j < n16; j += kCThreads) s4[j] = z;
    }
    PbQueueRegs q;
    if (threadIdx.x == 0) pb_queue_init(s_ring, q, tiles, rec_off, lookback, 0, n_tiles, T, tile_counter);
    pb_fence_proxy_async();
    __syncthreads();

    unsigned long long drop_p = 0, drop_m = 0, drop_a = 0, map_p = 0, map_m = 0, map_a = 0;
    unsigned int drop_len = 0;
    const int nibble = r.param;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int parity = 0;   // which copy of the segment sums the current tile-with-work uses
    for (int k = 0;; ++k) {
        if (threadIdx.x == 0) pb_queue_step(s_ring, q, k, tiles, rec_off, lookback, 0, n_tiles, T, tile_counter);
        const PbSlot &cur = s_ring[k & 3];
        const long long tile = cur.tile;
        if (tile >= n_tiles) break;
        const PbTile d = cur.d;
        const uint32_t rec_lo = cur.rec_lo, rec_hi = cur.rec_hi;
        const int64_t g0 = tile * T;
        const bool has_work = d.n > 0 || rec_hi > rec_lo;
        if (!has_work) {
            if (threadIdx.x == 0) {
                if (!accumulate) {
                    for (int q = 0; q < n_planes; ++q)
                        for (int z = 0; z < T; z += kZeroBins) pb_bulk_store(outs[q] + g0 + z, zbuf, kZeroBins * 8);
                    pb_bulk_commit();
                }
            }
            __syncthreads();
        } else {

        const int64_t p0 = d.p0, p1 = d.p0 + T, plim = d.p0 + d.live;
        // Difference update of one aligned reference interval [x,y) of map-length slot `sl`.  The sum
        // of each warp segment (what the scan needs as carry-in) is kept up to date with a second
        // shared atomic instead of re-reading every difference word before the scan.
        int *tot = warp_tot + parity * n_arrays * kCWarps;
        auto add_interval = [&](int64_t x, int64_t y, int sl, bool rev) {
            if (y <= p0 || x >= plim) return;
            const bool do_strand = rev ? want_minus : want_plus;
            const int a_strand = ((rev ? d_minus : d_plus) - diff) / T + sl, a_all = (d_any - diff) / T + sl;
            const unsigned ox = (unsigned)((x > p0 ? x : p0) - p0);
            if (do_strand) { atomicAdd(&diff[(size_t)a_strand * T + ox], 1); atomicAdd(&tot[a_strand * kCWarps + ox / seg], 1); }
            if (want_any) { atomicAdd(&diff[(size_t)a_all * T + ox], 1); atomicAdd(&tot[a_all * kCWarps + ox / seg], 1); }
            if (y < p1) {
                const unsigned oy = (unsigned)(y - p0);
                if (do_strand) { atomicAdd(&diff[(size_t)a_strand * T + oy], -1); atomicAdd(&tot[a_strand * kCWarps + oy / seg], -1); }
                if (want_any) { atomicAdd(&diff[(size_t)a_all * T + oy], -1); atomicAdd(&tot[a_all * kCWarps + oy / seg], -1); }
            }
        };

        // single-block reads of the candidate slice
        const int64_t hi = d.lo + d.n;
        for (int64_t base = d.lo; base < hi; base += (int64_t)kCUnroll * kCThreads) {
            int32_t sv[kCUnroll];
            uint32_t mv[kCUnroll];
#pragma unroll
            for (int u = 0; u < kCUnroll; ++u) {
                const int64_t i = base + (int64_t)u * kCThreads + threadIdx.x;
                const bool ok = i < hi;
                sv[u] = ok ? __ldg(b.ref_start + i) : 0;
                mv[u] = ok ? __ldg(b.meta + i) : (1u << 17);
            }
#pragma unroll
            for (int u = 0; u < kCUnroll; ++u) {
                const int32_t s = sv[u];
                const uint32_t m = mv[u];
                if (!pb_passes(m, r.size_min, r.size_max)) continue;
                if (rec_off && PB_META_NBLK(m) > 1) continue;      // arrives through the bucket
                const int L = PB_META_L(m);
                const bool rev = PB_META_REV(m);
                const bool own = (s >= p0 && s < p1);
                const int map_len = L - 2 * nibble;
                if (map_len < 0) {                                 // map_factories.pyx:246-248
                    if (own) { drop_a++; if (rev) drop_m++; else drop_p++; drop_len = L; }
                    continue;
                }
                if (map_len == 0) continue;
                if (own) { map_a++; if (rev) map_m++; else map_p++; }   // reads_out semantics (:256)
                const int slot = (int)__ldg(slot_of_len + L) - slot0;
                if (slot < 0 || slot >= n_slots) continue;         // another pass handles this map length
                add_interval((int64_t)s + nibble, (int64_t)s + L - nibble, slot, rev);
            }
        }
        if (s_ring[(k + 1) & 3].tile < n_tiles) pb_prefetch_tile_l2(b, recs, s_ring[(k + 1) & 3]);
        // trimmed aligned intervals of multi-block reads (already filtered and counted by pb_bin_kernel)
        for (uint32_t j = rec_lo + threadIdx.x; j < rec_hi; j += kCThreads) {
            const PbRec rec = recs[j];
            const int slot = (int)(rec.tag & 0xffffu) - slot0;
            if (slot < 0 || slot >= n_slots) continue;
            add_interval(rec.x, rec.y, slot, (rec.tag >> 16) & 1u);
        }
        __syncthreads();

        // the previous tile's bulk copies must have read the staging buffers before they are rewritten
        // (they were issued a whole read-scan ago); the barrier also publishes the difference arrays
        if (threadIdx.x == 0) pb_bulk_wait_read0();
        __syncthreads();
        // the other parity's segment sums were last used by the previous tile: clear them for the next
        for (int j = threadIdx.x; j < n_arrays * kCWarps; j += kCThreads) warp_tot[(parity ^ 1) * n_arrays * kCWarps + j] = 0;

        // pass 2: exact scan + fixed-order combine into the staging buffers; every thread zeroes the
        // difference words it consumed, so the arrays are clean for the next tile without another pass
        for (int q = 0; q < n_planes; ++q) {
            double acc[EPT];
#pragma unroll
            for (int ch = 0; ch < EPT; ++ch) acc[ch] = 0.0;
            for (int sl = 0; sl < n_slots; ++sl) {
                const int a = q * n_slots + sl;
                int *A = diff + (size_t)a * T + warp * seg;
                int carry = (lane < warp) ? tot[a * kCWarps + lane] : 0;
                carry = __reduce_add_sync(0xffffffffu, carry);
                const double w = __ldg(inv_m + slot0 + sl);
#pragma unroll
                for (int ch = 0; ch < EPT; ++ch) {
                    int v = A[ch * 32 + lane];
                    A[ch * 32 + lane] = 0;
                    // Kogge-Stone step = shuffle + add predicated on the shuffle's own in-range flag
                    // (two instructions instead of shuffle + compare + select + add)
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1)
                        asm volatile("{ .reg .s32 t; .reg .pred p;\n\t"
                                     "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
                                     "@p add.s32 %0, %0, t; }" : "+r"(v) : "r"(dd));
                    v += carry;
                    carry = __shfl_sync(0xffffffffu, v, 31);
                    acc[ch] += (double)v * w;
                }
            }
            double *buf = stage + (size_t)q * T + warp * seg;
#pragma unroll
            for (int ch = 0; ch < EPT; ++ch) buf[ch * 32 + lane] = acc[ch];
        }
        pb_fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 0; q < n_planes; ++q) {
                if (accumulate) pb_bulk_add_f64(outs[q] + g0, stage + (size_t)q * T, T * 8);
                else pb_bulk_store(outs[q] + g0, stage + (size_t)q * T, T * 8);
            }
            pb_bulk_commit();
        }
        parity ^= 1;
        }   // has_work
    }
    if (stat_slots) pb_flush_cta_stats(drop_p, drop_m, drop_a, drop_len, map_p, map_m, map_a, stat_slots);
    if (threadIdx.x == 0) pb_bulk_wait_all();
}

template <int EPT>
int launch_center(const PbReads &b, const PbRuleDev &r, int planes, const int16_t *slot_of_len, const double *inv_m,
                  int n_slots, int per_pass, int lookback, int64_t total_bins, const PbWorkspace &ws,
                  double *out_plus, double *out_minus, double *out_any, cudaStream_t stream)
{
    constexpr int T = EPT * kCThreads;
    const int n_planes = __builtin_popcount(planes);
    const int64_t n_tiles = total_bins / T;
    int sm_count = 0;
    int rc = pb_sm_count(&sm_count);
    if (rc) return rc;
    for (int s0 = 0, pass = 0; s0 < n_slots || pass == 0; s0 += per_pass, ++pass) {
        int ns = n_slots - s0 < per_pass ? n_slots - s0 : per_pass;
        if (ns < 0) ns = 0;
        const size_t smem = ((size_t)n_planes * T + kZeroBins) * 8 + (size_t)n_planes * ns * T * 4 +
                            (size_t)2 * n_planes * ns * kCWarps * 4 + 16;
        PB_CUDA_CHECK(cudaFuncSetAttribute(pb_center_tiles_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pb_center_tiles_kernel<EPT>, kCThreads, smem));
        if (occ < 1) occ = 1;
        int64_t grid = (int64_t)sm_count * occ;
        if (grid > n_tiles) grid = n_tiles;
        PB_CUDA_CHECK(cudaMemsetAsync(ws.tile_counter, 0, 64, stream));
        // statistics are accumulated by the first pass only (later passes see the same reads again)
        pb_center_tiles_kernel<EPT><<<(unsigned)grid, kCThreads, smem, stream>>>(
            b, r, planes, slot_of_len, inv_m, s0, ns, pass > 0, lookback, ws.tiles, n_tiles, ws.tile_counter,
            ws.rec_off, ws.recs, out_plus, out_minus, out_any, pass == 0 ? ws.slots : nullptr);
        if (n_slots == 0) break;
    }
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

}  // namespace

extern "C" int pb_map_center(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                             const int16_t *slot_of_len, const double *inv_m, int n_slots,
                             double *out_plus, double *out_minus, double *out_any,
                             uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream_)
{
    int rc = pb_check_common(batch, layout, rule, planes);
    if (rc) return rc;
    if (rule->kind != PB_RULE_CENTER || rule->param < 0) { pb_set_error("pb_map_center: need a center rule with nibble >= 0"); return PB_EINVAL; }
    if (!slot_of_len || (n_slots > 0 && !inv_m) || n_slots < 0 || n_slots > 32767) { pb_set_error("pb_map_center: bad slot tables"); return PB_EINVAL; }
    if (((planes & PB_PLANE_PLUS) && !out_plus) || ((planes & PB_PLANE_MINUS) && !out_minus) ||
        ((planes & PB_PLANE_ANY) && !out_any) || !stats) {
        pb_set_error("pb_map_center: missing output plane or stats"); return PB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    PbWorkspace ws;
    rc = pb_carve_workspace(workspace, workspace_bytes, layout->total_bins, b.n_blk, b.n_reads, &ws);
    if (rc) return rc;
    const int n_planes = __builtin_popcount(planes);

    // 2048-bin tiles when all difference arrays fit next to the staging buffers in ~72 KB (three or
    // four CTAs per SM); otherwise 1024-bin tiles, and several accumulating passes over groups of map lengths
    // if even those cannot hold every slot at once.
    const int ns = n_slots < 1 ? 1 : n_slots;
    int ept = 8, per_pass = ns;
    if (((size_t)n_planes * 2048 + kZeroBins) * 8 + (size_t)n_planes * ns * 2048 * 4 > 72 * 1024) {
        ept = 4;
        per_pass = (int)((200 * 1024 - ((size_t)n_planes * 1024 + kZeroBins) * 8) / ((size_t)n_planes * 1024 * 4));
        if (per_pass > ns) per_pass = ns;
        if (per_pass < 1) per_pass = 1;
    }
    const int tile_bins = ept * kCThreads;
    const int64_t n_tiles = layout->total_bins / tile_bins;
    const int lookback = (b.max_block_len + tile_bins - 1) / tile_bins;

    PB_CUDA_CHECK(cudaMemsetAsync(ws.slots, 0, 2 * pb_ws_stat_bytes() + 64, stream));
    rc = pb_launch_tile_index(b, lay, tile_bins, 0, n_tiles, batch->n_reads, 0, ws, stream);
    if (rc) return rc;
    rc = pb_launch_binning(b, r, lay, planes, 1, slot_of_len, tile_bins, n_tiles, ws, stream);
    if (rc) return rc;
    pb_timing_begin(stream);
    if (ept == 8)
        rc = launch_center<8>(b, r, planes, slot_of_len, inv_m, n_slots, per_pass, lookback, layout->total_bins, ws,
                              out_plus, out_minus, out_any, stream);
    else
        rc = launch_center<4>(b, r, planes, slot_of_len, inv_m, n_slots, per_pass, lookback, layout->total_bins, ws,
                              out_plus, out_minus, out_any, stream);
    pb_timing_end(stream);
    if (rc) return rc;
    return pb_launch_stats_finish(ws.slots, (unsigned long long *)stats, stream);
}
