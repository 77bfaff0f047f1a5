// pb_map.cu — alignments -> mapping rule -> dense per-strand count planes (sm_100a).
//
// Design ("owner computes", no global atomics, no memset pass): the concatenated genome is
// cut into position tiles; one CTA owns one tile, keeps the tile's bins for every requested
// query strand in shared memory, scans only the slice of the coordinate-sorted batch whose
// reads can land in the tile (slice bounds come from pb_tile_index_kernel), applies the
// mapping rule per read, accumulates with shared-memory integer atomics (order independent
// => bit-exact and deterministic) and writes the finished tile once with 128-bit stores.
// Tiles with no candidate reads (most of a human genome) skip shared memory and just store
// zeros, so the same launch is also the memset.
//
// Reference semantics restated (plastid/genomics/map_factories.pyx): FivePrime :308-367,
// ThreePrime :407-466, VariableFivePrime :585-650, Center :200-265, SizeFilter :837-839;
// strand selection as genome_array.py:811-815.  Query strand '.' applies the FORWARD rule to
// reads of both strands, so it is its own plane, not '+' + '-'.
#include "pb_common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kStatSlots = 64;  // stats are spread over 64 slots to keep atomics off one address

// ----------------------------------------------------------------------------------------
// tile -> candidate read slice
// ----------------------------------------------------------------------------------------
struct __align__(16) PbTile {
    long long lo;   // first candidate read
    long long p0;   // chromosome coordinate of the tile's first bin
    int n;          // number of candidate reads [lo, lo+n)
    int live;       // bins of the tile that lie inside the chromosome (0..tile_bins)
    int chrom;
    int pad;
};

__global__ void pb_tile_index_kernel(PbReads b, PbLayoutDev lay, int tile_bins, int64_t tile_begin, int64_t tile_end,
                                     int64_t read_limit, PbTile *__restrict__ tiles)
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
    // a read can only place a site in [p0, p0+T) if p0 - max_span < start < p0 + T
    int64_t lo = pb_lower_bound(b.ref_start, r0, r1, p0 - b.max_span + 1);
    int64_t hi = pb_lower_bound(b.ref_start, lo, r1, p0 + tile_bins);
    int64_t live = clen - p0;
    live = live < 0 ? 0 : (live > tile_bins ? tile_bins : live);
    PbTile d;
    d.lo = lo; d.p0 = p0;
    d.n = (live > 0 && hi - lo < 0x7fffffff) ? (int)(hi - lo) : (live > 0 ? 0x7fffffff : 0);
    d.live = (int)live; d.chrom = c; d.pad = 0;
    tiles[t] = d;
}

__device__ __forceinline__ void pb_flush_stats(const unsigned long long *local, unsigned int *s_stats,
                                               unsigned long long *stat_slots, int64_t tile)
{
    // block-level reduce in shared memory, then one global atomic per non-zero counter
#pragma unroll
    for (int k = 0; k < PB_NSTATS; ++k) {
        if (k == PB_STAT_DROPPED_LEN) continue;
        unsigned long long v = pb_warp_sum(local[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_stats[k], (unsigned int)v);
    }
    unsigned int len = (unsigned int)local[PB_STAT_DROPPED_LEN];
    len = __reduce_max_sync(0xffffffffu, len);
    if ((threadIdx.x & 31) == 0 && len) atomicMax(&s_stats[PB_STAT_DROPPED_LEN], len);
    __syncthreads();
    if (threadIdx.x < PB_NSTATS) {
        unsigned int v = s_stats[threadIdx.x];
        if (v) {
            unsigned long long *dst = stat_slots + (tile & (kStatSlots - 1)) * PB_NSTATS + threadIdx.x;
            if (threadIdx.x == PB_STAT_DROPPED_LEN) atomicMax(dst, (unsigned long long)v);
            else atomicAdd(dst, (unsigned long long)v);
        }
    }
}

// ----------------------------------------------------------------------------------------
// point rules: 5' / 3' / variable offset — persistent CTAs, dynamic tile queue, TMA bulk stores
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void pb_fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pb_bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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

// Invariant: at the top of every loop iteration the shared tile buffer is all zero and visible to
// the async proxy.  Empty tiles are therefore one bulk store of the buffer as it is; tiles with
// reads accumulate into it, store it, wait until the TMA engine has READ it (not until the write
// has landed), and re-zero it.  The SM never touches the output bytes itself.
__global__ void __launch_bounds__(kPThreads)
pb_point_tiles_kernel(PbReads b, PbRuleDev r, int planes, const PbTile *__restrict__ tiles, int64_t tile_begin,
                      int64_t n_tiles, unsigned long long *__restrict__ tile_counter,
                      uint32_t *__restrict__ out_plus, uint32_t *__restrict__ out_minus,
                      uint32_t *__restrict__ out_any, unsigned long long *__restrict__ stat_slots)
{
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ long long s_next[2];

    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS,
               want_any = planes & PB_PLANE_ANY;
    const int n_planes = (int)want_plus + (int)want_minus + (int)want_any;
    uint32_t *sm_plus = smem, *sm_minus = smem, *sm_any = smem;
    {
        int k = 0;
        if (want_plus) sm_plus = smem + (k++) * kPTileBins;
        if (want_minus) sm_minus = smem + (k++) * kPTileBins;
        if (want_any) sm_any = smem + (k++) * kPTileBins;
    }
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 *smem4 = reinterpret_cast<uint4 *>(smem);
    for (int j = threadIdx.x; j < n_planes * kPTileBins / 4; j += kPThreads) smem4[j] = zero4;
    if (threadIdx.x == 0) s_next[0] = tile_begin + (long long)atomicAdd(tile_counter, 1ull);
    pb_fence_proxy_async();
    __syncthreads();

    unsigned long long drop_p = 0, drop_m = 0, drop_a = 0, map_p = 0, map_m = 0, map_a = 0;
    unsigned int drop_len = 0;

    for (int it = 0;; ++it) {
        const long long tile = s_next[it & 1];
        if (tile >= n_tiles) break;
        if (threadIdx.x == 0) s_next[(it + 1) & 1] = tile_begin + (long long)atomicAdd(tile_counter, 1ull);
        const PbTile d = tiles[tile];
        const int64_t g0 = tile * kPTileBins;

        if (d.n > 0) {
            if (threadIdx.x == 0) pb_bulk_wait_read0();   // earlier stores of the zero buffer have read it
            __syncthreads();
            const int64_t p0 = d.p0, plim = d.p0 + d.live, p1 = d.p0 + kPTileBins;
            const int64_t hi = d.lo + d.n;
            const unsigned plus_base = (unsigned)(sm_plus - smem), minus_base = (unsigned)(sm_minus - smem),
                           any_base = (unsigned)(sm_any - smem);
            for (int64_t base = d.lo; base < hi; base += (int64_t)kPUnroll * kPThreads) {
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
                    const int64_t i = base + (int64_t)u * kPThreads + threadIdx.x;
                    const int L = PB_META_L(m);
                    const bool rev = PB_META_REV(m);
                    unsigned key_strand = kNoKey, key_any = kNoKey;   // word index into smem, or none
                    if (pb_passes(m, r.size_min, r.size_max)) {
                        const int idx_f = pb_rule_index(r, L, false);
                        if (idx_f < 0) {
                            // the reference skips this read and warns; count it once, in the tile owning its start
                            if (s >= p0 && s < p1) {
                                drop_a++;
                                if (rev) drop_m++; else drop_p++;
                                drop_len = L;
                            }
                        } else {
                            if (want_any || (!rev && want_plus)) {
                                const int64_t p = pb_position(b, i, s, m, idx_f);
                                if (p >= p0 && p < plim) {
                                    const unsigned o = (unsigned)(p - p0);
                                    if (want_any) { key_any = any_base + o; map_a++; }
                                    if (!rev && want_plus) { key_strand = plus_base + o; map_p++; }
                                }
                            }
                            if (rev && want_minus) {
                                const int idx_r = pb_rule_index(r, L, true);
                                const int64_t p = pb_position(b, i, s, m, idx_r);
                                if (p >= p0 && p < plim) {
                                    key_strand = minus_base + (unsigned)(p - p0);
                                    map_m++;
                                }
                            }
                        }
                    }
                    pb_smem_inc(smem, key_strand);
                    pb_smem_inc(smem, key_any);
                }
            }
            pb_fence_proxy_async();
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            int k = 0;
            if (want_plus) pb_bulk_store(out_plus + g0, smem + (k++) * kPTileBins, kPTileBins * 4);
            if (want_minus) pb_bulk_store(out_minus + g0, smem + (k++) * kPTileBins, kPTileBins * 4);
            if (want_any) pb_bulk_store(out_any + g0, smem + (k++) * kPTileBins, kPTileBins * 4);
            pb_bulk_commit();
            if (d.n > 0) pb_bulk_wait_read0();
        }
        __syncthreads();
        if (d.n > 0) {
            for (int j = threadIdx.x; j < n_planes * kPTileBins / 4; j += kPThreads) smem4[j] = zero4;
            pb_fence_proxy_async();
            __syncthreads();
        }
    }

    // per-CTA statistics: warp reduce -> shared -> one global atomic per counter
    {
        unsigned long long v[6] = {drop_p, drop_m, drop_a, map_p, map_m, map_a};
        const int idx[6] = {PB_STAT_DROPPED_PLUS, PB_STAT_DROPPED_MINUS, PB_STAT_DROPPED_ANY,
                            PB_STAT_MAPPED_PLUS, PB_STAT_MAPPED_MINUS, PB_STAT_MAPPED_ANY};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const unsigned long long t = pb_warp_sum(v[k]);
            if ((threadIdx.x & 31) == 0 && t)
                atomicAdd(stat_slots + (blockIdx.x & (kStatSlots - 1)) * PB_NSTATS + idx[k], t);
        }
        const unsigned int len = __reduce_max_sync(0xffffffffu, drop_len);
        if ((threadIdx.x & 31) == 0 && len)
            atomicMax(stat_slots + (blockIdx.x & (kStatSlots - 1)) * PB_NSTATS + PB_STAT_DROPPED_LEN,
                      (unsigned long long)len);
    }
    if (threadIdx.x == 0) pb_bulk_wait_all();
}

// ----------------------------------------------------------------------------------------
// center rule: integer difference arrays per map length, exact scan, fixed-order fp64 combine
// ----------------------------------------------------------------------------------------
// Shared layout: diff[plane][slot][tile_bins] int32.  Every read adds +1 at the first bin of each
// trimmed aligned interval and -1 one past its end (clipped to the tile); an inclusive scan gives
// the number of reads of that map length covering each bin; the bin value is
// sum_slot cover[slot] * (1.0 / map_length[slot]) evaluated in ascending map-length order.
template <int EPT>  // bins per thread = tile_bins / kThreads
__global__ void __launch_bounds__(kThreads)
pb_center_tiles_kernel(PbReads b, PbRuleDev r, PbLayoutDev lay, int planes,
                       const int16_t *__restrict__ slot_of_len, const double *__restrict__ inv_m,
                       int slot0, int n_slots, int accumulate,
                       const PbTile *__restrict__ tiles,
                       double *__restrict__ out_plus, double *__restrict__ out_minus,
                       double *__restrict__ out_any, unsigned long long *__restrict__ stat_slots)
{
    constexpr int tile_bins = EPT * kThreads;
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ unsigned int s_stats[PB_NSTATS];

    const int64_t tile = blockIdx.x;
    const int64_t g0 = tile * tile_bins;
    const PbTile d = tiles[tile];
    const int64_t lo = d.lo, hi = d.lo + d.n;
    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS,
               want_any = planes & PB_PLANE_ANY;
    const int n_planes = (int)want_plus + (int)want_minus + (int)want_any;
    double *outs[3];
    {
        int k = 0;
        if (want_plus) outs[k++] = out_plus;
        if (want_minus) outs[k++] = out_minus;
        if (want_any) outs[k++] = out_any;
    }

    if (lo >= hi) {
        if (!accumulate) {
            for (int q = 0; q < n_planes; ++q) {
                double2 *dst = reinterpret_cast<double2 *>(outs[q] + g0);
                for (int j = threadIdx.x; j < tile_bins / 2; j += kThreads) dst[j] = make_double2(0.0, 0.0);
            }
        }
        return;
    }

    const int64_t p0 = d.p0;
    const int64_t p1 = p0 + tile_bins;
    const int64_t plim = p0 + d.live;

    int *diff = reinterpret_cast<int *>(smem);
    int *warp_tot = diff + n_planes * n_slots * tile_bins;  // [n_planes*n_slots][kWarps]
    {
        uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4 *s4 = reinterpret_cast<uint4 *>(smem);
        for (int j = threadIdx.x; j < n_planes * n_slots * tile_bins / 4; j += kThreads) s4[j] = z;
    }
    if (threadIdx.x < PB_NSTATS) s_stats[threadIdx.x] = 0;
    __syncthreads();

    int *d_plus = diff, *d_minus = diff, *d_any = diff;
    {
        int k = 0;
        if (want_plus) d_plus = diff + (k++) * n_slots * tile_bins;
        if (want_minus) d_minus = diff + (k++) * n_slots * tile_bins;
        if (want_any) d_any = diff + (k++) * n_slots * tile_bins;
    }

    unsigned long long local[PB_NSTATS];
#pragma unroll
    for (int k = 0; k < PB_NSTATS; ++k) local[k] = 0;
    const int nibble = r.param;

    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        const int32_t s = __ldg(b.ref_start + i);
        const uint32_t m = __ldg(b.meta + i);
        if (!pb_passes(m, r.size_min, r.size_max)) continue;
        const int L = PB_META_L(m);
        const bool rev = PB_META_REV(m);
        const bool own = (s >= p0 && s < p1);
        const int map_len = L - 2 * nibble;
        if (map_len < 0) {  // map_factories.pyx:246-248
            if (own) {
                local[PB_STAT_DROPPED_ANY]++;
                local[rev ? PB_STAT_DROPPED_MINUS : PB_STAT_DROPPED_PLUS]++;
                local[PB_STAT_DROPPED_LEN] = L;
            }
            continue;
        }
        if (map_len == 0) continue;
        if (own) {  // every mapped read is counted once (reads_out semantics, :256)
            local[PB_STAT_MAPPED_ANY]++;
            local[rev ? PB_STAT_MAPPED_MINUS : PB_STAT_MAPPED_PLUS]++;
        }
        const int slot = (int)__ldg(slot_of_len + L) - slot0;
        if (slot < 0 || slot >= n_slots) continue;  // handled by another pass
        const int so = slot * tile_bins;
        const bool do_strand = rev ? want_minus : want_plus;
        int *d_strand = (rev ? d_minus : d_plus) + so;
        int *d_all = d_any + so;

        auto add_interval = [&](int64_t x, int64_t y) {  // aligned reference interval [x,y)
            if (y <= p0 || x >= plim) return;
            const unsigned ox = (unsigned)((x > p0 ? x : p0) - p0);
            if (do_strand) atomicAdd(&d_strand[ox], 1);
            if (want_any) atomicAdd(&d_all[ox], 1);
            if (y < p1) {
                const unsigned oy = (unsigned)(y - p0);
                if (do_strand) atomicAdd(&d_strand[oy], -1);
                if (want_any) atomicAdd(&d_all[oy], -1);
            }
        };

        if (PB_META_NBLK(m) <= 1 || b.blk_off == nullptr) {
            add_interval((int64_t)s + nibble, (int64_t)s + L - nibble);
        } else {
            const uint32_t k0 = __ldg(b.blk_off + i), k1 = __ldg(b.blk_off + i + 1);
            int a = 0;  // aligned-base index of the block's first base
            for (uint32_t k = k0; k < k1; ++k) {
                const int2 bl = __ldg(b.blk + k);
                const int ia = a > nibble ? a : nibble;
                const int ib = (a + bl.y) < (L - nibble) ? (a + bl.y) : (L - nibble);
                if (ia < ib) add_interval((int64_t)s + bl.x + (ia - a), (int64_t)s + bl.x + (ib - a));
                a += bl.y;
            }
        }
    }
    pb_flush_stats(local, s_stats, stat_slots, tile);  // includes __syncthreads()

    // pass 1: per-warp segment totals for every (plane, slot)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int seg = tile_bins / kWarps;  // bins per warp = EPT * 32
    const int n_arrays = n_planes * n_slots;
    for (int a = 0; a < n_arrays; ++a) {
        const int *A = diff + a * tile_bins + warp * seg;
        int t = 0;
#pragma unroll
        for (int ch = 0; ch < EPT; ++ch) t += A[ch * 32 + lane];
        t = __reduce_add_sync(0xffffffffu, t);
        if (lane == 0) warp_tot[a * kWarps + warp] = t;
    }
    __syncthreads();

    // pass 2: scan + combine, plane by plane
    for (int q = 0; q < n_planes; ++q) {
        double acc[EPT];
#pragma unroll
        for (int ch = 0; ch < EPT; ++ch) acc[ch] = 0.0;
        for (int sl = 0; sl < n_slots; ++sl) {
            const int a = q * n_slots + sl;
            const int *A = diff + a * tile_bins + warp * seg;
            int carry = (lane < warp) ? warp_tot[a * kWarps + lane] : 0;  // kWarps <= 32
            carry = __reduce_add_sync(0xffffffffu, carry);
            const double w = __ldg(inv_m + slot0 + sl);
#pragma unroll
            for (int ch = 0; ch < EPT; ++ch) {
                int v = A[ch * 32 + lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int u = __shfl_up_sync(0xffffffffu, v, d);
                    if (lane >= d) v += u;
                }
                v += carry;
                carry = __shfl_sync(0xffffffffu, v, 31);
                acc[ch] += (double)v * w;
            }
        }
        double *dst = outs[q] + g0 + warp * seg;
#pragma unroll
        for (int ch = 0; ch < EPT; ++ch) {
            if (accumulate) dst[ch * 32 + lane] += acc[ch];
            else dst[ch * 32 + lane] = acc[ch];
        }
    }
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

// ----------------------------------------------------------------------------------------
// the operator on one segment (global atomics; small inputs)
// ----------------------------------------------------------------------------------------
__global__ void pb_segment_kernel(PbReads b, PbRuleDev r, int64_t i0, int64_t i1, int strand, int flags,
                                  int64_t seg_start, int64_t seg_end,
                                  unsigned long long *counts_i, double *counts_f,
                                  uint8_t *__restrict__ kept, unsigned long long *__restrict__ stats)
{
    const int64_t n = seg_end - seg_start;
    const bool rq = (strand == PB_PLANE_MINUS);
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = __ldg(b.ref_start + i);
        const uint32_t m = __ldg(b.meta + i);
        uint8_t keep = 0;
        const bool rev = PB_META_REV(m);
        bool pass = pb_passes(m, r.size_min, r.size_max);
        if ((flags & PB_SEG_FILTER_STRAND) && strand == PB_PLANE_PLUS && rev) pass = false;    // genome_array.py:811-815
        if ((flags & PB_SEG_FILTER_STRAND) && strand == PB_PLANE_MINUS && !rev) pass = false;
        if (pass && (flags & PB_SEG_FETCH_OVERLAP)) {
            // AlignmentFile.fetch(chrom, start, end): only reads whose reference span overlaps the segment
            int64_t span = PB_META_L(m);
            if (PB_META_NBLK(m) > 1 && b.blk_off != nullptr) {
                const int2 last = __ldg(b.blk + (__ldg(b.blk_off + i + 1) - 1));
                span = (int64_t)last.x + last.y;
            }
            if (!((int64_t)s < seg_end && (int64_t)s + span > seg_start)) pass = false;
        }
        if (pass) {
            const int L = PB_META_L(m);
            const int sidx = strand == PB_PLANE_PLUS ? PB_STAT_DROPPED_PLUS
                           : strand == PB_PLANE_MINUS ? PB_STAT_DROPPED_MINUS : PB_STAT_DROPPED_ANY;
            if (r.kind == PB_RULE_CENTER) {
                const int nib = r.param, map_len = L - 2 * nib;
                if (map_len < 0) {
                    atomicAdd(&stats[sidx], 1ull);
                    stats[PB_STAT_DROPPED_LEN] = L;
                } else if (map_len > 0) {
                    const double v = 1.0 / map_len;
                    for (int k = nib; k < L - nib; ++k) {
                        const int64_t cpos = pb_position(b, i, s, m, k) - seg_start;
                        if (cpos >= 0 && cpos < n) atomicAdd(&counts_f[cpos], v);
                    }
                    keep = 1;
                }
            } else if (r.kind == PB_RULE_STRATIFIED) {
                if (L >= r.strat_min && L <= r.strat_max && L < PB_LUT_SIZE) {
                    int off = rq ? __ldg(r.lut_rc + L) : __ldg(r.lut_fw + L);
                    if (off < 0) off += L;  // map_factories.pyx:773-774: no BAD_OFFSET test, python index -1
                    const int64_t p = pb_position(b, i, s, m, off);
                    if (p >= seg_start && p < seg_end) {
                        atomicAdd(&counts_i[(int64_t)(L - r.strat_min) * n + (p - seg_start)], 1ull);
                        keep = 1;
                    }
                }
            } else {
                const int idx = pb_rule_index(r, L, rq);
                if (idx < 0) {
                    atomicAdd(&stats[sidx], 1ull);
                    stats[PB_STAT_DROPPED_LEN] = L;
                } else {
                    const int64_t p = pb_position(b, i, s, m, idx);
                    if (p >= seg_start && p < seg_end) {
                        atomicAdd(&counts_i[p - seg_start], 1ull);
                        keep = 1;
                    }
                }
            }
        }
        if (kept) kept[i - i0] = keep;
    }
}

__global__ void pb_length_hist_kernel(PbReads b, PbRuleDev r, int strand, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[1024];
    for (int j = threadIdx.x; j < 1024; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n_reads;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t m = __ldg(b.meta + i);
        if (!pb_passes(m, r.size_min, r.size_max)) continue;
        const bool rev = PB_META_REV(m);
        if (strand == PB_PLANE_PLUS && rev) continue;
        if (strand == PB_PLANE_MINUS && !rev) continue;
        const int L = PB_META_L(m);
        if (L < 1024) atomicAdd(&sh[L], 1u); else atomicAdd(&hist[L], 1ull);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 1024; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], (unsigned long long)sh[j]);
}

// ----------------------------------------------------------------------------------------
// wire16: compact host format of an unspliced batch (4 B per read) -> the SoA the kernels stream
// ----------------------------------------------------------------------------------------
// Reads are sorted, so within one 65536-position segment of a chromosome the start needs 16 bits;
// seg_off[s] is the first read of segment s, seg_base[s] the chromosome coordinate of its first
// position.  One warp expands 1024 consecutive reads: one binary search for the chunk's segment,
// then every lane walks forward (segments are crossed rarely).
__global__ void __launch_bounds__(256)
pb_unpack_wire16_kernel(const uint16_t *__restrict__ start_lo, const uint16_t *__restrict__ meta16,
                        const int64_t *__restrict__ seg_off, const int32_t *__restrict__ seg_base,
                        int64_t n_seg, int64_t read_begin, int64_t n_reads, int32_t *__restrict__ ref_start,
                        uint32_t *__restrict__ meta)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t chunk0 = read_begin + warp * 1024;
    if (chunk0 >= n_reads) return;
    int64_t lo = 0, hi = n_seg;       // last segment with seg_off[s] <= chunk0
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= chunk0) lo = mid; else hi = mid;
    }
    int64_t seg = lo;
    int64_t seg_end = __ldg(seg_off + seg + 1);
    int32_t base = __ldg(seg_base + seg);
    const int64_t chunk1 = chunk0 + 1024 < n_reads ? chunk0 + 1024 : n_reads;
    for (int64_t i = chunk0 + lane; i < chunk1; i += 32) {
        while (i >= seg_end) {        // empty segments are skipped too
            ++seg;
            seg_end = __ldg(seg_off + seg + 1);
            base = __ldg(seg_base + seg);
        }
        const uint32_t m = __ldg(meta16 + i);
        ref_start[i] = base + (int32_t)__ldg(start_lo + i);
        // L (14 bits) | reverse | drop  ->  L | reverse<<16 | drop<<17 | n_blocks(=1)<<24
        meta[i] = (m & 0x3fffu) | (((m >> 14) & 1u) << 16) | (((m >> 15) & 1u) << 17) | (1u << 24);
    }
}

PbReads to_dev(const pb_batch *b)
{
    PbReads d;
    d.ref_start = b->ref_start;
    d.meta = b->meta;
    d.blk_off = b->blk_off;
    d.blk = reinterpret_cast<const int2 *>(b->blk);
    d.chrom_read_off = b->chrom_read_off;
    d.n_reads = b->n_reads;
    d.n_chrom = b->n_chrom;
    d.max_span = b->max_span < 1 ? 1 : b->max_span;
    return d;
}

PbRuleDev to_dev(const pb_rule *r)
{
    PbRuleDev d;
    d.kind = r->kind; d.param = r->param;
    d.lut_fw = r->lut_fw; d.lut_rc = r->lut_rc;
    d.size_min = r->size_min; d.size_max = r->size_max;
    d.strat_min = r->strat_min; d.strat_max = r->strat_max;
    return d;
}

int check_common(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes)
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
    return PB_OK;
}

// optional device timing of the dominant (tiles) kernel, for bench.py's roofline line: a ring of
// CUDA event pairs recorded on the launch stream, summed by pb_tiles_kernel_ms_total().
constexpr int kTimingRing = 256;
bool g_timing = false;
cudaEvent_t g_ev[kTimingRing][2];
int g_ev_created = 0, g_ev_count = 0;

void timing_begin(cudaStream_t stream)
{
    if (!g_timing || g_ev_count >= kTimingRing) return;
    if (g_ev_count >= g_ev_created) {
        cudaEventCreate(&g_ev[g_ev_created][0]);
        cudaEventCreate(&g_ev[g_ev_created][1]);
        g_ev_created++;
    }
    cudaEventRecord(g_ev[g_ev_count][0], stream);
}
void timing_end(cudaStream_t stream)
{
    if (!g_timing || g_ev_count >= kTimingRing) return;
    cudaEventRecord(g_ev[g_ev_count][1], stream);
    g_ev_count++;
}

size_t tile_index_bytes(int64_t total_bins) { return (size_t)(total_bins / 1024 + 1) * sizeof(PbTile); }
size_t stat_slot_bytes() { return (size_t)kStatSlots * PB_NSTATS * sizeof(unsigned long long); }

}  // namespace

extern "C" size_t pb_map_workspace_bytes(int64_t total_bins)
{
    if (total_bins < 0) return 0;
    return tile_index_bytes(total_bins) + 2 * stat_slot_bytes() + 256;  // + tile counter
}

extern "C" int pb_map_point_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                                  uint32_t *out_plus, uint32_t *out_minus, uint32_t *out_any,
                                  uint64_t *stats, void *workspace, size_t workspace_bytes,
                                  int64_t bin_begin, int64_t bin_end, int64_t read_limit, void *stream_)
{
    int rc = check_common(batch, layout, rule, planes);
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
    if (workspace_bytes < pb_map_workspace_bytes(layout->total_bins) || !workspace) {
        pb_set_error("pb_map_point: workspace too small"); return PB_ENOSPACE;
    }
    if (bin_begin < 0 || bin_end > layout->total_bins || bin_begin > bin_end || bin_begin % PB_LAYOUT_ALIGN ||
        bin_end % PB_LAYOUT_ALIGN) {
        pb_set_error("pb_map_point_range: bin range must be PB_LAYOUT_ALIGN-aligned and inside the layout"); return PB_EINVAL;
    }
    if (read_limit < 0 || read_limit > batch->n_reads) read_limit = batch->n_reads;
    if (bin_begin == bin_end) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t tile_begin = bin_begin / kPTileBins, n_tiles = bin_end / kPTileBins;   // n_tiles = end of range
    PbTile *tiles = (PbTile *)workspace;
    unsigned long long *slots = (unsigned long long *)((char *)workspace + tile_index_bytes(layout->total_bins));
    unsigned long long *tile_counter = slots + 2 * kStatSlots * PB_NSTATS;
    PbReads b = to_dev(batch);
    PbRuleDev r = to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};

    PB_CUDA_CHECK(cudaMemsetAsync(slots, 0, 2 * stat_slot_bytes() + 64, stream));
    pb_tile_index_kernel<<<(unsigned)((n_tiles - tile_begin + 255) / 256), 256, 0, stream>>>(b, lay, kPTileBins, tile_begin,
                                                                                              n_tiles, read_limit, tiles);
    const int n_planes = __builtin_popcount(planes);
    const size_t smem = (size_t)n_planes * kPTileBins * sizeof(uint32_t);
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        PB_CUDA_CHECK(cudaGetDevice(&dev));
        PB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    PB_CUDA_CHECK(cudaFuncSetAttribute(pb_point_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pb_point_tiles_kernel, kPThreads, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sm_count * occ;     // persistent: one resident wave, tiles come from a queue
    if (grid > n_tiles - tile_begin) grid = n_tiles - tile_begin;
    timing_begin(stream);
    pb_point_tiles_kernel<<<(unsigned)grid, kPThreads, smem, stream>>>(b, r, planes, tiles, tile_begin, n_tiles, tile_counter,
                                                                      out_plus, out_minus, out_any, slots);
    timing_end(stream);
    pb_stats_finish_kernel<<<1, 32, 0, stream>>>(slots, (unsigned long long *)stats);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_map_point(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                            uint32_t *out_plus, uint32_t *out_minus, uint32_t *out_any,
                            uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!layout || !batch) { pb_set_error("null batch/layout/rule"); return PB_EINVAL; }
    return pb_map_point_range(batch, layout, rule, planes, out_plus, out_minus, out_any, stats, workspace,
                              workspace_bytes, 0, layout->total_bins, batch->n_reads, stream_);
}

template <int EPT>
static int launch_center(const PbReads &b, const PbRuleDev &r, const PbLayoutDev &lay, int planes,
                         const int16_t *slot_of_len, const double *inv_m, int n_slots, int slots_per_pass,
                         int64_t total_bins, const PbTile *tiles,
                         double *out_plus, double *out_minus, double *out_any,
                         unsigned long long *slots, cudaStream_t stream)
{
    constexpr int tile_bins = EPT * kThreads;
    const int n_planes = __builtin_popcount(planes);
    const int64_t n_tiles = total_bins / tile_bins;
    for (int s0 = 0, pass = 0; s0 < n_slots || pass == 0; s0 += slots_per_pass, ++pass) {
        const int ns = (n_slots - s0) < slots_per_pass ? (n_slots - s0) : slots_per_pass;
        const int ns_eff = ns < 1 ? 1 : ns;
        const size_t smem = ((size_t)n_planes * ns_eff * tile_bins + (size_t)n_planes * ns_eff * kWarps) * sizeof(int);
        PB_CUDA_CHECK(cudaFuncSetAttribute(pb_center_tiles_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // stats are only accumulated by the first pass (later passes would count reads again)
        pb_center_tiles_kernel<EPT><<<(unsigned)n_tiles, kThreads, smem, stream>>>(
            b, r, lay, planes, slot_of_len, inv_m, s0, ns < 0 ? 0 : ns, pass > 0, tiles,
            out_plus, out_minus, out_any, pass == 0 ? slots : slots + kStatSlots * PB_NSTATS);
        if (n_slots == 0) break;
    }
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_map_center(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                             const int16_t *slot_of_len, const double *inv_m, int n_slots,
                             double *out_plus, double *out_minus, double *out_any,
                             uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream_)
{
    int rc = check_common(batch, layout, rule, planes);
    if (rc) return rc;
    if (rule->kind != PB_RULE_CENTER || rule->param < 0) { pb_set_error("pb_map_center: need a center rule with nibble >= 0"); return PB_EINVAL; }
    if (!slot_of_len || (n_slots > 0 && !inv_m) || n_slots < 0 || n_slots > 32767) { pb_set_error("pb_map_center: bad slot tables"); return PB_EINVAL; }
    if (((planes & PB_PLANE_PLUS) && !out_plus) || ((planes & PB_PLANE_MINUS) && !out_minus) ||
        ((planes & PB_PLANE_ANY) && !out_any) || !stats) {
        pb_set_error("pb_map_center: missing output plane or stats"); return PB_EINVAL;
    }
    if (workspace_bytes < pb_map_workspace_bytes(layout->total_bins) || !workspace) {
        pb_set_error("pb_map_center: workspace too small"); return PB_ENOSPACE;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    PbTile *tiles = (PbTile *)workspace;
    unsigned long long *slots = (unsigned long long *)((char *)workspace + tile_index_bytes(layout->total_bins));
    PbReads b = to_dev(batch);
    PbRuleDev r = to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    const int n_planes = __builtin_popcount(planes);

    // pick the largest tile whose difference arrays fit ~96 KB (two CTAs per SM); if even the
    // smallest tile cannot hold all slots, run several passes over slot groups.
    const size_t budget = 96 * 1024;
    const int ns = n_slots < 1 ? 1 : n_slots;
    int ept = 16;
    while (ept > 2 && (size_t)n_planes * ns * ept * kThreads * 4 > budget) ept >>= 1;
    int per_pass = (int)(budget / ((size_t)n_planes * ept * kThreads * 4));
    if (per_pass < 1) per_pass = 1;
    if (per_pass > ns) per_pass = ns;
    const int tile_bins = ept * kThreads;
    const int64_t n_tiles = layout->total_bins / tile_bins;

    PB_CUDA_CHECK(cudaMemsetAsync(slots, 0, 2 * stat_slot_bytes(), stream));
    pb_tile_index_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(b, lay, tile_bins, 0, n_tiles, batch->n_reads, tiles);
    timing_begin(stream);
    switch (ept) {
    case 16: rc = launch_center<16>(b, r, lay, planes, slot_of_len, inv_m, n_slots, per_pass, layout->total_bins, tiles, out_plus, out_minus, out_any, slots, stream); break;
    case 8:  rc = launch_center<8>(b, r, lay, planes, slot_of_len, inv_m, n_slots, per_pass, layout->total_bins, tiles, out_plus, out_minus, out_any, slots, stream); break;
    case 4:  rc = launch_center<4>(b, r, lay, planes, slot_of_len, inv_m, n_slots, per_pass, layout->total_bins, tiles, out_plus, out_minus, out_any, slots, stream); break;
    default: rc = launch_center<2>(b, r, lay, planes, slot_of_len, inv_m, n_slots, per_pass, layout->total_bins, tiles, out_plus, out_minus, out_any, slots, stream); break;
    }
    timing_end(stream);
    if (rc) return rc;
    pb_stats_finish_kernel<<<1, 32, 0, stream>>>(slots, (unsigned long long *)stats);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
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

extern "C" int pb_map_segment(const pb_batch *batch, int64_t i0, int64_t i1, const pb_rule *rule, int strand,
                              int flags, int64_t seg_start, int64_t seg_end, void *counts_out, uint8_t *kept_out,
                              uint64_t *stats, void *stream_)
{
    if (!batch || !rule || !counts_out || !stats) { pb_set_error("pb_map_segment: null argument"); return PB_EINVAL; }
    if (i0 < 0 || i1 < i0 || i1 > batch->n_reads) { pb_set_error("pb_map_segment: bad read range"); return PB_EINVAL; }
    if (strand != PB_PLANE_PLUS && strand != PB_PLANE_MINUS && strand != PB_PLANE_ANY) { pb_set_error("pb_map_segment: bad strand"); return PB_EINVAL; }
    if (seg_end < seg_start) { pb_set_error("pb_map_segment: negative-length segment"); return PB_EINVAL; }
    if ((rule->kind == PB_RULE_VARIABLE || rule->kind == PB_RULE_STRATIFIED) && (!rule->lut_fw || !rule->lut_rc)) {
        pb_set_error("pb_map_segment: rule needs lut_fw/lut_rc"); return PB_EINVAL;
    }
    if (rule->kind < 0 || rule->kind > PB_RULE_STRATIFIED) { pb_set_error("pb_map_segment: unknown rule"); return PB_EINVAL; }
    if (i1 == i0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = to_dev(batch);
    PbRuleDev r = to_dev(rule);
    int64_t n = i1 - i0;
    unsigned grid = (unsigned)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    pb_segment_kernel<<<grid, 256, 0, stream>>>(b, r, i0, i1, strand, flags, seg_start, seg_end,
                                                (unsigned long long *)counts_out, (double *)counts_out, kept_out,
                                                (unsigned long long *)stats);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_length_hist(const pb_batch *batch, const pb_rule *rule, int strand, uint64_t *hist, void *stream_)
{
    if (!batch || !rule || !hist) { pb_set_error("pb_length_hist: null argument"); return PB_EINVAL; }
    if (batch->n_reads == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = to_dev(batch);
    PbRuleDev r = to_dev(rule);
    unsigned grid = (unsigned)((batch->n_reads + 511) / 512);
    if (grid > 148 * 8) grid = 148 * 8;
    pb_length_hist_kernel<<<grid, 512, 0, stream>>>(b, r, strand, (unsigned long long *)hist);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_unpack_wire16(const uint16_t *start_lo, const uint16_t *meta16, const int64_t *seg_off,
                                const int32_t *seg_base, int64_t n_seg, int64_t read_begin, int64_t read_end,
                                int32_t *ref_start_out, uint32_t *meta_out, void *stream_)
{
    if (read_begin < 0 || read_end < read_begin || n_seg < 0) { pb_set_error("pb_unpack_wire16: bad range"); return PB_EINVAL; }
    if (read_end == read_begin) return PB_OK;
    const int64_t n_reads = read_end;
    if (!start_lo || !meta16 || !seg_off || !seg_base || !ref_start_out || !meta_out || n_seg < 1) {
        pb_set_error("pb_unpack_wire16: null argument"); return PB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t warps = (read_end - read_begin + 1023) / 1024;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    pb_unpack_wire16_kernel<<<grid, 256, 0, stream>>>(start_lo, meta16, seg_off, seg_base, n_seg, read_begin, n_reads, ref_start_out, meta_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
