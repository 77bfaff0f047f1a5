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
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int kCThreads = 256;
constexpr int kCWarps = kCThreads / 32;
constexpr int kCUnroll = 4;
constexpr int kZeroBins = 256;   // 2 KB of fp64 zeros: tiles nothing lands in are stored from here

// exact int -> double for 0 <= x < 2^32 with one DADD instead of an I2F: 2^52 + x is representable
__device__ __forceinline__ double pb_u32_to_f64(int x)
{
    return __hiloint2double(0x43300000, x) - 4503599627370496.0;
}

// One barrier per tile: the difference arrays are double-buffered by iteration parity, every warp
// scans and writes out its own chunks of the tile, so a warp that is done moves on to the next tile
// while slower warps still scan.  A thread owns 4 CONSECUTIVE bins of a 128-bin chunk (one LDS.128,
// a 3-add serial scan, one 5-step warp scan per chunk); the running sum entering a chunk comes from
// per-chunk totals kept with a second shared atomic, so chunks are independent of each other.
// The first candidate read and binned record of every thread for tile k+1 are loaded into
// registers before the barrier of tile k and applied after its scan (software pipeline: their
// latency is covered by the scan instead of being waited for at the barrier).
//
// DIRECT: the finished bins leave the registers as 256-bit global stores (STG.E.256; a warp writes
// 1 KB contiguous per instruction) — no fp64 staging buffer, so the shared-memory pipe carries only
// the difference arrays.  The staged variant (TMA bulk store / bulk fp64 reduction) remains for the
// accumulating passes of batches with more map lengths than fit, and for planes not 32-byte aligned.
//
// PLANES / ONE_SLOT: compile-time copies of the plane mask and of "this pass has exactly one map
// length" for the common cases ('+' and '-' planes, all reads trimmed to one length), so the plane
// and slot loops unroll and their address arithmetic folds; PLANES = 0 / ONE_SLOT = false is the
// generic kernel.
template <int EPT, bool DIRECT, int PLANES, bool ONE_SLOT>  // EPT bins per thread (4, 8 or 16); tile = EPT * 256 bins
__global__ void __launch_bounds__(kCThreads, DIRECT ? 4 : 3)
pb_center_tiles_kernel(PbReads b, PbRuleDev r, int planes_rt,
                       const int16_t *__restrict__ slot_of_len, const double *__restrict__ inv_m,
                       int slot0, int n_slots_rt, int accumulate, int lookback,
                       const PbTile *__restrict__ tiles, int64_t tile_begin, int64_t n_tiles,
                       unsigned long long *__restrict__ tile_counter,
                       const uint32_t *__restrict__ rec_off, const PbRec *__restrict__ recs,
                       double *__restrict__ out_plus, double *__restrict__ out_minus, double *__restrict__ out_any,
                       unsigned long long *__restrict__ stat_slots)
{
    // tiles [tile_begin, n_tiles) are produced (n_tiles = end of the range; whole genome: 0 .. total tiles)
    const int planes = PLANES ? PLANES : planes_rt;
    const int n_slots = ONE_SLOT ? 1 : n_slots_rt;
    constexpr int T = EPT * kCThreads;
    constexpr int kChunk = 128;                 // bins per scan unit: 32 lanes x 4 consecutive bins
    constexpr int nChunks = T / kChunk;         // <= 32: one warp reduction yields a chunk's carry-in
    constexpr int kPerWarp = nChunks / kCWarps; // chunks scanned by each warp (contiguous)
    static_assert(EPT % 4 == 0 && nChunks <= 32, "tile = 1024, 2048 or 4096 bins");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ PbSlot s_ring[4];

    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS,
               want_any = planes & PB_PLANE_ANY;
    const int n_planes = PLANES ? ((PLANES & 1) + ((PLANES >> 1) & 1) + ((PLANES >> 2) & 1))
                                : (int)want_plus + (int)want_minus + (int)want_any;
    const int n_arrays = n_planes * n_slots;
    double *stage = reinterpret_cast<double *>(smem_raw);           // [n_planes][T] fp64 staging (staged variant)
    double *zbuf = stage + (DIRECT ? 0 : (size_t)n_planes * T);     // [kZeroBins] zeros, never written
    int *diff_all = reinterpret_cast<int *>(zbuf + kZeroBins);      // [2][n_arrays][T] by iteration parity
    int *tot_all = diff_all + (size_t)2 * n_arrays * T;             // [3][n_arrays][nChunks] by iteration mod 3
    double *outs[3];
    int a_plus = 0, a_minus = 0, a_any = 0;                         // first array (slot 0) of each plane
    {
        int k = 0;
        if (want_plus) { outs[k] = out_plus; a_plus = (k++) * n_slots; }
        if (want_minus) { outs[k] = out_minus; a_minus = (k++) * n_slots; }
        if (want_any) { outs[k] = out_any; a_any = (k++) * n_slots; }
    }
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4 *s4 = reinterpret_cast<uint4 *>(smem_raw);
        const int n16 = (((DIRECT ? 0 : n_planes * T) + kZeroBins) * 8 + 2 * n_arrays * T * 4 + 3 * n_arrays * nChunks * 4 + 15) / 16;
        for (int j = threadIdx.x; j < n16; j += kCThreads) s4[j] = z;
    }
    PbQueueRegs q;
    if (threadIdx.x == 0) pb_queue_init(s_ring, q, tiles, rec_off, lookback, tile_begin, n_tiles, T, tile_counter);
    pb_fence_proxy_async();
    __syncthreads();

    unsigned int drop_p = 0, drop_m = 0, drop_a = 0, map_p = 0, map_m = 0, map_a = 0;   // per thread: < 2^32
    unsigned int drop_len = 0;
    const int nibble = r.param;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double w_one = (ONE_SLOT && n_slots_rt > 0) ? __ldg(inv_m + slot0) : 0.0;

    // register stage of the software pipeline: this thread's first candidate read and first binned
    // record of the tile that becomes current in the next iteration (drop bit / empty interval = none)
    int32_t pre_s = 0;
    uint32_t pre_m = 1u << 17;
    PbRec pre_rec = PbRec{0, 0, 0u, 0u};
    auto preload = [&](const PbSlot &nx) {
        pre_m = 1u << 17;
        pre_rec.x = pre_rec.y = 0;
        if (nx.tile >= n_tiles) return;
        if ((int)threadIdx.x < nx.d.n) {
            pre_s = __ldg(b.ref_start + nx.d.lo + threadIdx.x);
            pre_m = __ldg(b.meta + nx.d.lo + threadIdx.x);
        }
        if (nx.rec_lo + threadIdx.x < nx.rec_hi) pre_rec = recs[nx.rec_lo + threadIdx.x];
    };
    preload(s_ring[0]);

    int k3 = 0;       // k mod 3
    for (int k = 0;; ++k, k3 = (k3 == 2 ? 0 : k3 + 1)) {
        if (threadIdx.x == 0) pb_queue_step(s_ring, q, k, tiles, rec_off, lookback, tile_begin, n_tiles, T, tile_counter);
        const PbSlot &cur = s_ring[k & 3];
        const long long tile = cur.tile;
        if (tile >= n_tiles) break;
        const PbTile d = cur.d;
        const uint32_t rec_lo = cur.rec_lo, rec_hi = cur.rec_hi;
        const int64_t g0 = tile * T;
        // pile-up tile (d.pad > 0): left to the overflow jobs and pb_center_finish_hot_kernel, bins included
        const bool has_work = d.pad == 0 && (d.n > 0 || rec_hi > rec_lo);
        int *diff = diff_all + (size_t)(k & 1) * n_arrays * T;
        int *tot = tot_all + k3 * n_arrays * nChunks;
        int *tot_next2 = tot_all + (k3 == 0 ? 2 : k3 - 1) * n_arrays * nChunks;   // (k + 2) mod 3

        if (has_work) {
            const int64_t p0 = d.p0, p1 = d.p0 + T, plim = d.p0 + d.live;
            // Difference update of one aligned reference interval [x,y) of map-length slot `sl`; the
            // total of the 128-bin chunk is kept up to date alongside (the scan's carry-in).
            auto add_interval = [&](int64_t x, int64_t y, int sl, bool rev) {
                if (y <= p0 || x >= plim) return;
                const bool do_strand = rev ? want_minus : want_plus;
                const int a_strand = (rev ? a_minus : a_plus) + sl, a_all = a_any + sl;
                const unsigned ox = (unsigned)((x > p0 ? x : p0) - p0);
                if (do_strand) { atomicAdd(&diff[(size_t)a_strand * T + ox], 1); atomicAdd(&tot[a_strand * nChunks + ox / kChunk], 1); }
                if (want_any) { atomicAdd(&diff[(size_t)a_all * T + ox], 1); atomicAdd(&tot[a_all * nChunks + ox / kChunk], 1); }
                if (y < p1) {
                    const unsigned oy = (unsigned)(y - p0);
                    if (do_strand) { atomicAdd(&diff[(size_t)a_strand * T + oy], -1); atomicAdd(&tot[a_strand * nChunks + oy / kChunk], -1); }
                    if (want_any) { atomicAdd(&diff[(size_t)a_all * T + oy], -1); atomicAdd(&tot[a_all * nChunks + oy / kChunk], -1); }
                }
            };
            auto one_read = [&](int32_t s, uint32_t m) {
                if (!pb_passes(m, r.size_min, r.size_max)) return;
                if (rec_off && PB_META_NBLK(m) > 1) return;        // arrives through the bucket
                const int L = PB_META_L(m);
                const bool rev = PB_META_REV(m);
                const bool own = (s >= p0 && s < p1);
                const int map_len = L - 2 * nibble;
                if (map_len < 0) {                                 // map_factories.pyx:246-248
                    if (own) { drop_a++; if (rev) drop_m++; else drop_p++; drop_len = L; }
                    return;
                }
                if (map_len == 0) return;
                if (own) { map_a++; if (rev) map_m++; else map_p++; }   // reads_out semantics (:256)
                const int slot = (int)__ldg(slot_of_len + L) - slot0;
                if (slot < 0 || slot >= n_slots) return;           // another pass handles this map length
                add_interval((int64_t)s + nibble, (int64_t)s + L - nibble, slot, rev);
            };
            auto one_rec = [&](const PbRec &rec) {
                const int slot = (int)(rec.tag & 0xffffu) - slot0;
                if (slot < 0 || slot >= n_slots) return;
                add_interval(rec.x, rec.y, slot, (rec.tag >> 16) & 1u);
            };

            // the first 256 candidate reads / binned records were loaded a tile ago
            one_read(pre_s, pre_m);
            if (pre_rec.y > pre_rec.x) one_rec(pre_rec);
            // dense tiles: the rest of the candidate slice, kCUnroll independent loads per thread
            const int64_t hi = d.lo + d.n;
            for (int64_t base = d.lo + kCThreads; base < hi; base += (int64_t)kCUnroll * kCThreads) {
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
                for (int u = 0; u < kCUnroll; ++u) one_read(sv[u], mv[u]);
            }
            // trimmed aligned intervals of multi-block reads (already filtered and counted by pb_bin_kernel)
            for (uint32_t j = rec_lo + kCThreads + threadIdx.x; j < rec_hi; j += kCThreads) one_rec(recs[j]);
            // this warp's previous bulk copies must have read its staging segment before it is rewritten
            if (!DIRECT && lane == 0) pb_bulk_wait_read0();
        } else if (threadIdx.x == 0 && !accumulate && d.pad == 0) {
            for (int qq = 0; qq < n_planes; ++qq)
                for (int z = 0; z < T; z += kZeroBins) pb_bulk_store(outs[qq] + g0 + z, zbuf, kZeroBins * 8);
            pb_bulk_commit();
        }
        preload(s_ring[(k + 1) & 3]);     // consumed after this tile's scan
        __syncthreads();   // the only barrier of the iteration: publishes the difference arrays
        // chunk totals of iteration k+2 (= k-1 mod 3, consumed before this barrier) are cleared here:
        // they are not touched again before the next barrier
        for (int j = threadIdx.x; j < n_arrays * nChunks; j += kCThreads) tot_next2[j] = 0;
        // the tile after next was published before the barrier: pull its reads and records into L2
        if (s_ring[(k + 2) & 3].tile < n_tiles) pb_prefetch_tile_l2(b, recs, s_ring[(k + 2) & 3]);
        if (!has_work) continue;

        // exact scan of this warp's chunks + fixed-order combine over the slots; the difference
        // words are zeroed as they are consumed
#pragma unroll 1
        for (int cc = 0; cc < kPerWarp; ++cc) {
            const int c = warp * kPerWarp + cc;
#pragma unroll
            for (int qq = 0; qq < (PLANES ? n_planes : 3); ++qq) {
                if (!PLANES && qq >= n_planes) break;
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll 1
                for (int sl = 0; sl < n_slots; ++sl) {
                    const int a = qq * n_slots + sl;
                    int4 *A = reinterpret_cast<int4 *>(diff + (size_t)a * T + c * kChunk);
                    int carry = (lane < c) ? tot[a * nChunks + lane] : 0;
                    const int4 v = A[lane];
                    A[lane] = make_int4(0, 0, 0, 0);
                    carry = __reduce_add_sync(0xffffffffu, carry);
                    const double w = ONE_SLOT ? w_one : __ldg(inv_m + slot0 + sl);
                    const int s1 = v.x, s2 = s1 + v.y, s3 = s2 + v.z, s4 = s3 + v.w;
                    int incl = s4;
                    // Kogge-Stone step = shuffle + add predicated on the shuffle's own in-range flag
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1)
                        asm volatile("{ .reg .s32 t; .reg .pred p;\n\t"
                                     "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
                                     "@p add.s32 %0, %0, t; }" : "+r"(incl) : "r"(dd));
                    const int before = incl - s4 + carry;     // coverage entering this thread's 4 bins
                    acc0 = __fma_rn(pb_u32_to_f64(before + s1), w, acc0);    // (explicit: the finish kernel of the pile-up
                    acc1 = __fma_rn(pb_u32_to_f64(before + s2), w, acc1);    // tiles must round the same way)
                    acc2 = __fma_rn(pb_u32_to_f64(before + s3), w, acc2);
                    acc3 = __fma_rn(pb_u32_to_f64(before + s4), w, acc3);
                }
                if (DIRECT) {
                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};"
                                 :: "l"(outs[qq] + g0 + c * kChunk + lane * 4), "d"(acc0), "d"(acc1), "d"(acc2), "d"(acc3)
                                 : "memory");
                } else {
                    // a lane's 32 bytes go out as two 16-byte stores; lanes 4-7 of every eight store their
                    // halves in the opposite order, which keeps each quarter-warp on 32 distinct banks
                    double2 *buf = reinterpret_cast<double2 *>(stage + (size_t)qq * T + c * kChunk);
                    const int flip = (lane >> 2) & 1;
                    const double2 lo2 = make_double2(acc0, acc1), hi2 = make_double2(acc2, acc3);
                    buf[lane * 2 + flip] = flip ? hi2 : lo2;
                    buf[lane * 2 + (flip ^ 1)] = flip ? lo2 : hi2;
                }
            }
        }
        if (!DIRECT) {
            pb_fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                constexpr int seg = kPerWarp * kChunk;
                for (int qq = 0; qq < n_planes; ++qq) {
                    double *dst = outs[qq] + g0 + warp * seg;
                    const double *src = stage + (size_t)qq * T + warp * seg;
                    if (accumulate) pb_bulk_add_f64(dst, src, seg * 8);
                    else pb_bulk_store(dst, src, seg * 8);
                }
                pb_bulk_commit();
            }
        }
    }
    if (stat_slots) pb_flush_cta_stats(drop_p, drop_m, drop_a, drop_len, map_p, map_m, map_a, stat_slots);
    if (lane == 0) pb_bulk_wait_all();   // staged tiles (every warp) and zero tiles (thread 0)
}

// ----------------------------------------------------------------------------------------
// many map lengths at once: 64-bit fixed-point weights
// ----------------------------------------------------------------------------------------
// With S distinct map lengths the exact kernel above needs S integer difference arrays per plane;
// ribo-seq reads (25-35 nt, the reference's default CenterMapFactory()) have a dozen, which does not
// fit next to a useful tile and forced many accumulating passes (13x slower than one length).  Here
// every read adds the INTEGER weight W_m = round(2^shift / m) to ONE 64-bit difference array per
// plane: integer accumulation commutes (deterministic, exact zeros), the prefix scan is exact, and
// bin = total * 2^-shift.  The host picks `shift` from the length histogram so that no total can
// overflow 63 bits and checks that m * 2^-(shift+1) — the relative error against sum(1/m) — stays far
// below the north star's 1e-6 (map_batch falls back to the exact multi-pass kernel otherwise).
// Same skeleton as the exact kernel: one barrier per tile, double-buffered arrays, 4 consecutive bins
// per thread, register-prefetched reads, 256-bit direct stores.  64-bit shared atomics are CAS loops on
// sm_100 (ATOMS.CAST.SPIN.64): fine for the ~100 reads of an ordinary tile.
// Run-aggregated 64-bit add, called by all 32 lanes: consecutive lanes with the same key (reads are
// coordinate-sorted, so equal targets come in runs) are summed with a segmented warp scan and the last
// lane of every run issues ONE atomic.  64-bit shared atomics are CAS loops; in a pile-up tile, where
// hundreds of reads share a start, this removes the same-address contention.  key < 0 = nothing to add.
template <typename W>
__device__ __forceinline__ void pb_run_add(W *arr, int key, W w)
{
    const int lane = threadIdx.x & 31;
    const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
    bool head = (lane == 0) || (key != kprev);
    const int knext = __shfl_down_sync(0xffffffffu, key, 1);
    const bool tail = (lane == 31) || (knext != key);
    W v = w;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const W up = __shfl_up_sync(0xffffffffu, v, dd);
        const bool hup = __shfl_up_sync(0xffffffffu, (int)head, dd) != 0;
        if (lane >= dd && !head) { v += up; head = hup; }
    }
    if (tail && key >= 0 && v != (W)0) atomicAdd(arr + key, v);
}
__device__ __forceinline__ void pb_run_add_u64(unsigned long long *arr, int key, unsigned long long w) { pb_run_add<unsigned long long>(arr, key, w); }

constexpr int kCDenseReads = 2048;   // candidate reads per tile from which the aggregated path is used

template <int PLANES>
__global__ void __launch_bounds__(kCThreads, 3)
pb_center_fixed_kernel(PbReads b, PbRuleDev r, int planes_rt,
                       const int16_t *__restrict__ slot_of_len, const long long *__restrict__ w_fix, double scale,
                       int lookback, const PbTile *__restrict__ tiles, int64_t tile_begin, int64_t n_tiles,
                       unsigned long long *__restrict__ tile_counter,
                       const uint32_t *__restrict__ rec_off, const PbRec *__restrict__ recs,
                       double *__restrict__ out_plus, double *__restrict__ out_minus, double *__restrict__ out_any,
                       unsigned long long *__restrict__ stat_slots)
{
    typedef unsigned long long u64;
    constexpr int EPT = 8, T = EPT * kCThreads, kChunk = 128, nChunks = T / kChunk, kPerWarp = nChunks / kCWarps;
    const int planes = PLANES ? PLANES : planes_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ PbSlot s_ring[4];
    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS,
               want_any = planes & PB_PLANE_ANY;
    const int n_planes = PLANES ? ((PLANES & 1) + ((PLANES >> 1) & 1) + ((PLANES >> 2) & 1))
                                : (int)want_plus + (int)want_minus + (int)want_any;
    double *zbuf = reinterpret_cast<double *>(smem_raw);            // [kZeroBins] zeros, never written
    u64 *diff_all = reinterpret_cast<u64 *>(zbuf + kZeroBins);      // [2][n_planes][T] by iteration parity
    u64 *tot_all = diff_all + (size_t)2 * n_planes * T;             // [3][n_planes][nChunks] by iteration mod 3
    double *outs[3];
    int a_plus = 0, a_minus = 0, a_any = 0;
    {
        int k = 0;
        if (want_plus) { outs[k] = out_plus; a_plus = k++; }
        if (want_minus) { outs[k] = out_minus; a_minus = k++; }
        if (want_any) { outs[k] = out_any; a_any = k++; }
    }
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4 *s4 = reinterpret_cast<uint4 *>(smem_raw);
        const int n16 = (kZeroBins * 8 + 2 * n_planes * T * 8 + 3 * n_planes * nChunks * 8 + 15) / 16;
        for (int j = threadIdx.x; j < n16; j += kCThreads) s4[j] = z;
    }
    PbQueueRegs q;
    if (threadIdx.x == 0) pb_queue_init(s_ring, q, tiles, rec_off, lookback, tile_begin, n_tiles, T, tile_counter);
    pb_fence_proxy_async();
    __syncthreads();

    unsigned int drop_p = 0, drop_m = 0, drop_a = 0, map_p = 0, map_m = 0, map_a = 0, drop_len = 0;
    const int nibble = r.param;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int32_t pre_s = 0;
    uint32_t pre_m = 1u << 17;
    PbRec pre_rec = PbRec{0, 0, 0u, 0u};
    auto preload = [&](const PbSlot &nx) {
        pre_m = 1u << 17;
        pre_rec.x = pre_rec.y = 0;
        if (nx.tile >= n_tiles) return;
        if ((int)threadIdx.x < nx.d.n) {
            pre_s = __ldg(b.ref_start + nx.d.lo + threadIdx.x);
            pre_m = __ldg(b.meta + nx.d.lo + threadIdx.x);
        }
        if (nx.rec_lo + threadIdx.x < nx.rec_hi) pre_rec = recs[nx.rec_lo + threadIdx.x];
    };
    preload(s_ring[0]);

    int k3 = 0;
    for (int k = 0;; ++k, k3 = (k3 == 2 ? 0 : k3 + 1)) {
        if (threadIdx.x == 0) pb_queue_step(s_ring, q, k, tiles, rec_off, lookback, tile_begin, n_tiles, T, tile_counter);
        const PbSlot &cur = s_ring[k & 3];
        const long long tile = cur.tile;
        if (tile >= n_tiles) break;
        const PbTile d = cur.d;
        const uint32_t rec_lo = cur.rec_lo, rec_hi = cur.rec_hi;
        const int64_t g0 = tile * T;
        const bool has_work = d.pad == 0 && (d.n > 0 || rec_hi > rec_lo);      // pile-up tile: see pb_center_tiles_kernel
        u64 *diff = diff_all + (size_t)(k & 1) * n_planes * T;
        u64 *tot = tot_all + k3 * n_planes * nChunks;
        u64 *tot_next2 = tot_all + (k3 == 0 ? 2 : k3 - 1) * n_planes * nChunks;

        if (has_work) {
            const int64_t p0 = d.p0, p1 = d.p0 + T, plim = d.p0 + d.live;
            auto add_interval = [&](int64_t x, int64_t y, u64 w, bool rev) {
                if (y <= p0 || x >= plim) return;
                const bool do_strand = rev ? want_minus : want_plus;
                const int a_strand = rev ? a_minus : a_plus;
                const unsigned ox = (unsigned)((x > p0 ? x : p0) - p0);
                if (do_strand) { atomicAdd(&diff[(size_t)a_strand * T + ox], w); atomicAdd(&tot[a_strand * nChunks + ox / kChunk], w); }
                if (want_any) { atomicAdd(&diff[(size_t)a_any * T + ox], w); atomicAdd(&tot[a_any * nChunks + ox / kChunk], w); }
                if (y < p1) {
                    const unsigned oy = (unsigned)(y - p0);
                    const u64 nw = 0ull - w;                       // two's complement: adds -w
                    if (do_strand) { atomicAdd(&diff[(size_t)a_strand * T + oy], nw); atomicAdd(&tot[a_strand * nChunks + oy / kChunk], nw); }
                    if (want_any) { atomicAdd(&diff[(size_t)a_any * T + oy], nw); atomicAdd(&tot[a_any * nChunks + oy / kChunk], nw); }
                }
            };
            // decode one read into its trimmed interval and weight (statistics on the way)
            auto decode = [&](int32_t s, uint32_t m, int64_t &x, int64_t &y, u64 &w, bool &rev) -> bool {
                if (!pb_passes(m, r.size_min, r.size_max)) return false;
                if (rec_off && PB_META_NBLK(m) > 1) return false;  // arrives through the bucket
                const int L = PB_META_L(m);
                rev = PB_META_REV(m);
                const bool own = (s >= p0 && s < p1);
                const int map_len = L - 2 * nibble;
                if (map_len < 0) {                                 // map_factories.pyx:246-248
                    if (own) { drop_a++; if (rev) drop_m++; else drop_p++; drop_len = L; }
                    return false;
                }
                if (map_len == 0) return false;
                if (own) { map_a++; if (rev) map_m++; else map_p++; }   // reads_out semantics (:256)
                const int slot = (int)__ldg(slot_of_len + L);
                if (slot < 0) return false;
                x = (int64_t)s + nibble;
                y = (int64_t)s + L - nibble;
                w = (u64)__ldg(w_fix + slot);
                return true;
            };
            // pile-up tiles: all lanes take part, equal targets are summed in the warp first
            auto agg_interval = [&](bool valid, int64_t x, int64_t y, u64 w, bool rev) {
                const bool in = valid && !(y <= p0 || x >= plim);
                const bool has_y = in && y < p1;
                const int kx = in ? (int)((x > p0 ? x : p0) - p0) : -1, ky = has_y ? (int)(y - p0) : -1;
                const int cx = in ? kx / kChunk : -1, cy = has_y ? ky / kChunk : -1;
                const u64 nw = 0ull - w;
                if (want_plus) {
                    const u64 wx = (in && !rev) ? w : 0ull, wy = (has_y && !rev) ? nw : 0ull;
                    pb_run_add_u64(diff + (size_t)a_plus * T, kx, wx);  pb_run_add_u64(tot + a_plus * nChunks, cx, wx);
                    pb_run_add_u64(diff + (size_t)a_plus * T, ky, wy);  pb_run_add_u64(tot + a_plus * nChunks, cy, wy);
                }
                if (want_minus) {
                    const u64 wx = (in && rev) ? w : 0ull, wy = (has_y && rev) ? nw : 0ull;
                    pb_run_add_u64(diff + (size_t)a_minus * T, kx, wx); pb_run_add_u64(tot + a_minus * nChunks, cx, wx);
                    pb_run_add_u64(diff + (size_t)a_minus * T, ky, wy); pb_run_add_u64(tot + a_minus * nChunks, cy, wy);
                }
                if (want_any) {
                    const u64 wx = in ? w : 0ull, wy = has_y ? nw : 0ull;
                    pb_run_add_u64(diff + (size_t)a_any * T, kx, wx);   pb_run_add_u64(tot + a_any * nChunks, cx, wx);
                    pb_run_add_u64(diff + (size_t)a_any * T, ky, wy);   pb_run_add_u64(tot + a_any * nChunks, cy, wy);
                }
            };
            const bool dense = d.n >= kCDenseReads || (rec_hi - rec_lo) >= (uint32_t)kCDenseReads;   // CTA-uniform
            auto one_read = [&](int32_t s, uint32_t m) {
                int64_t x = 0, y = 0;
                u64 w = 0;
                bool rev = false;
                const bool valid = decode(s, m, x, y, w, rev);
                if (dense) agg_interval(valid, x, y, w, rev);
                else if (valid) add_interval(x, y, w, rev);
            };
            auto one_rec = [&](bool have, const PbRec &rec) {
                const u64 w = have ? (u64)__ldg(w_fix + (rec.tag & 0xffffu)) : 0ull;
                const bool rev = (rec.tag >> 16) & 1u;
                if (dense) agg_interval(have, rec.x, rec.y, w, rev);
                else if (have) add_interval(rec.x, rec.y, w, rev);
            };
            one_read(pre_s, pre_m);
            one_rec(pre_rec.y > pre_rec.x, pre_rec);
            const int64_t hi = d.lo + d.n;
            for (int64_t base = d.lo + kCThreads; base < hi; base += (int64_t)kCUnroll * kCThreads) {
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
                for (int u = 0; u < kCUnroll; ++u) one_read(sv[u], mv[u]);
            }
            for (uint32_t j0 = rec_lo + kCThreads; j0 < rec_hi; j0 += kCThreads) {     // warp-uniform trip count
                const uint32_t j = j0 + threadIdx.x;
                PbRec rec = PbRec{0, 0, 0u, 0u};
                if (j < rec_hi) rec = recs[j];
                one_rec(j < rec_hi, rec);
            }
        } else if (threadIdx.x == 0 && d.pad == 0) {
            for (int qq = 0; qq < n_planes; ++qq)
                for (int z = 0; z < T; z += kZeroBins) pb_bulk_store(outs[qq] + g0 + z, zbuf, kZeroBins * 8);
            pb_bulk_commit();
        }
        preload(s_ring[(k + 1) & 3]);
        __syncthreads();
        for (int j = threadIdx.x; j < n_planes * nChunks; j += kCThreads) tot_next2[j] = 0;
        if (s_ring[(k + 2) & 3].tile < n_tiles) pb_prefetch_tile_l2(b, recs, s_ring[(k + 2) & 3]);
        if (!has_work) continue;

#pragma unroll
        for (int qq = 0; qq < (PLANES ? n_planes : 3); ++qq) {
            if (!PLANES && qq >= n_planes) break;
            // carry-in of every chunk of this plane: exclusive scan of the 16 chunk totals across lanes
            long long tsum = lane < nChunks ? (long long)tot[qq * nChunks + lane] : 0ll;
            const long long town = tsum;
#pragma unroll
            for (int dd = 1; dd < nChunks; dd <<= 1) {
                const long long up = __shfl_up_sync(0xffffffffu, tsum, dd);
                if (lane >= dd) tsum += up;
            }
            const long long tex = tsum - town;
#pragma unroll
            for (int cc = 0; cc < kPerWarp; ++cc) {
                const int c = warp * kPerWarp + cc;
                const long long carry = __shfl_sync(0xffffffffu, tex, c);
                ulonglong2 *A = reinterpret_cast<ulonglong2 *>(diff + (size_t)qq * T + c * kChunk + lane * 4);
                const ulonglong2 v01 = A[0], v23 = A[1];
                A[0] = make_ulonglong2(0ull, 0ull);
                A[1] = make_ulonglong2(0ull, 0ull);
                const long long s1 = (long long)v01.x, s2 = s1 + (long long)v01.y, s3 = s2 + (long long)v23.x,
                                s4 = s3 + (long long)v23.y;
                long long incl = s4;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const long long up = __shfl_up_sync(0xffffffffu, incl, dd);
                    if (lane >= dd) incl += up;
                }
                const long long before = incl - s4 + carry;
                const double o0 = (double)(before + s1) * scale, o1 = (double)(before + s2) * scale,
                             o2 = (double)(before + s3) * scale, o3 = (double)(before + s4) * scale;
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};"
                             :: "l"(outs[qq] + g0 + c * kChunk + lane * 4), "d"(o0), "d"(o1), "d"(o2), "d"(o3) : "memory");
            }
        }
    }
    if (stat_slots) pb_flush_cta_stats(drop_p, drop_m, drop_a, drop_len, map_p, map_m, map_a, stat_slots);
    if (threadIdx.x == 0) pb_bulk_wait_all();
}

// ----------------------------------------------------------------------------------------
// pile-up tiles (highly expressed genes put 10^5-10^6 reads into one tile; one CTA walking them alone would set
// the kernel's duration)
// ----------------------------------------------------------------------------------------
// The difference arrays are integers, so partial arrays add up exactly.  pb_center_jobs_kernel finds the tiles with
// more than `split` candidate reads or binned records, cuts ALL their candidates and records into jobs of `split`
// and gives each such tile a scratch index (PbTile.pad = 1 + index: the tiles kernel leaves the tile alone);
// pb_center_overflow_kernel (persistent CTAs) accumulates every job in shared memory — equal targets of neighbouring
// lanes summed in the warp first, the reads of a pile-up share their starts — and reduces it into the tile's scratch
// difference arrays with TMA bulk reductions (cp.reduce.async.bulk .add.u32 / .add.u64, SASS UBLKRED);
// pb_center_finish_hot_kernel scans the scratch arrays and writes the tile's bins with the arithmetic of the tiles
// kernel.  The planes are bit-identical to the unsplit run (test_center_pileup_tiles_are_split_bit_identically).
constexpr int kCSplit = 8192;          // candidate reads / records per job; more than that in a tile = pile-up
constexpr int kCHotMax = 8192;         // pile-up tiles with a scratch slot (header: their tile numbers)
constexpr size_t kCHotHeader = (size_t)kCHotMax * sizeof(long long);
constexpr int kCHotSlots = 64;         // map lengths per pass the finish kernel keeps carries for

__global__ void pb_center_jobs_kernel(PbTile *__restrict__ tiles, const uint32_t *__restrict__ rec_off, int lookback,
                                      int tile_bins, int64_t tile_begin, int64_t tile_end, int split,
                                      PbJob *__restrict__ jobs, long long job_capacity,
                                      unsigned long long *__restrict__ counters, long long hot_cap,
                                      long long *__restrict__ hot_tiles)
{
    const int64_t t = tile_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tile_end) return;
    PbTile d = tiles[t];
    long long nrec = 0;
    uint32_t rl = 0;
    if (rec_off) {      // the tile's records: its own bucket and those of up to `lookback` earlier tiles of its chromosome
        const long long first = t - d.p0 / tile_bins;
        long long from = t - lookback;
        if (from < first) from = first;
        rl = __ldg(rec_off + from);
        nrec = (long long)__ldg(rec_off + t + 1) - rl;
    }
    if (d.n <= split && nrec <= split) return;
    const long long jr = ((long long)d.n + split - 1) / split, jc = (nrec + split - 1) / split;
    const long long h = (long long)atomicAdd(counters + 3, 1ull);
    if (h >= hot_cap) return;                                                // out of scratch: the tile is walked whole
    const long long at = (long long)atomicAdd(counters + 1, (unsigned long long)(jr + jc));
    const bool fits = at + jr + jc <= job_capacity;                          // else: empty jobs, tile walked whole
    for (long long j = 0; j < jr + jc && at + j < job_capacity; ++j) {
        PbJob jb;
        const bool reads = j < jr;
        const long long skip = (long long)split * (reads ? j : j - jr);
        const long long left = (reads ? (long long)d.n : nrec) - skip;
        jb.lo = (reads ? d.lo : (long long)rl) + skip;
        jb.tile = t;
        jb.n = fits ? (int)(left < split ? left : split) : 0;
        jb.kind = reads ? 0 : 1;
        jb.hot = (int)h;
        jb.pad = 0;
        jobs[at + j] = jb;
    }
    hot_tiles[h] = fits ? (long long)t : -1ll;
    if (fits) {
        d.pad = (int)h + 1;
        tiles[t] = d;
    }
}

__global__ void pb_center_zero_hot_kernel(uint4 *__restrict__ scratch, const unsigned long long *__restrict__ counters,
                                          long long hot_cap, size_t stride16)
{
    long long nh = (long long)counters[3];
    if (nh > hot_cap) nh = hot_cap;
    const size_t n = (size_t)nh * stride16;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
        scratch[j] = make_uint4(0u, 0u, 0u, 0u);
}

template <bool FIXED> struct PbCenterWord { typedef int type; };
template <> struct PbCenterWord<true> { typedef unsigned long long type; };

template <bool FIXED>
__global__ void __launch_bounds__(kCThreads)
pb_center_overflow_kernel(PbReads b, PbRuleDev r, int planes, const int16_t *__restrict__ slot_of_len,
                          const long long *__restrict__ w_fix, int slot0, int n_slots, int T, int skip_multi,
                          const PbTile *__restrict__ tiles, const PbJob *__restrict__ jobs,
                          const unsigned long long *__restrict__ counters, long long job_capacity,
                          unsigned long long *__restrict__ job_counter, const PbRec *__restrict__ recs,
                          void *__restrict__ hot, size_t hot_stride, unsigned long long *__restrict__ stat_slots)
{
    typedef typename PbCenterWord<FIXED>::type W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ long long s_job;
    long long n_jobs = (long long)counters[1];
    if (n_jobs > job_capacity) n_jobs = job_capacity;
    if (n_jobs == 0) return;
    const bool want_plus = planes & PB_PLANE_PLUS, want_minus = planes & PB_PLANE_MINUS, want_any = planes & PB_PLANE_ANY;
    const int n_planes = (int)want_plus + (int)want_minus + (int)want_any;
    const int n_arrays = n_planes * n_slots;
    int a_plus = 0, a_minus = 0, a_any = 0;
    {
        int k = 0;
        if (want_plus) a_plus = (k++) * n_slots;
        if (want_minus) a_minus = (k++) * n_slots;
        if (want_any) a_any = (k++) * n_slots;
    }
    W *diff = reinterpret_cast<W *>(smem_raw);            // [n_arrays][T]: the layout of the tile's scratch
    const uint32_t n_bytes = (uint32_t)((size_t)n_arrays * T * sizeof(W));
    const int nibble = r.param;
    unsigned int drop_p = 0, drop_m = 0, drop_a = 0, map_p = 0, map_m = 0, map_a = 0, drop_len = 0;
    for (;;) {
        if (threadIdx.x == 0) s_job = (long long)atomicAdd(job_counter, 1ull);
        {
            uint4 *s4 = reinterpret_cast<uint4 *>(smem_raw);
            for (uint32_t j = threadIdx.x; j < n_bytes / 16; j += kCThreads) s4[j] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const long long job = s_job;
        if (job >= n_jobs) break;
        const PbJob jb = jobs[job];
        const PbTile d = tiles[jb.tile];
        const int64_t p0 = d.p0, p1 = d.p0 + T, plim = d.p0 + d.live;
        for (int base = 0; base < jb.n; base += kCThreads) {         // CTA-uniform trip count
            const int i = base + (int)threadIdx.x;
            const bool have = i < jb.n;
            bool valid = false, rev = false;
            int64_t x = 0, y = 0;
            int slot = 0;
            if (jb.kind == 0) {
                const int32_t s = have ? __ldg(b.ref_start + jb.lo + i) : 0;
                const uint32_t m = have ? __ldg(b.meta + jb.lo + i) : (1u << 17);
                if (pb_passes(m, r.size_min, r.size_max) && !(skip_multi && PB_META_NBLK(m) > 1)) {
                    const int L = PB_META_L(m);
                    rev = PB_META_REV(m);
                    const bool own = (s >= p0 && s < p1);
                    const int map_len = L - 2 * nibble;
                    if (map_len < 0) {                                 // map_factories.pyx:246-248
                        if (own) { drop_a++; if (rev) drop_m++; else drop_p++; drop_len = L; }
                    } else if (map_len > 0) {
                        if (own) { map_a++; if (rev) map_m++; else map_p++; }      // reads_out semantics (:256)
                        slot = (int)__ldg(slot_of_len + L) - slot0;
                        valid = slot >= 0 && slot < (FIXED ? 32768 : n_slots);     // exact kernel: another pass has the rest
                        x = (int64_t)s + nibble;
                        y = (int64_t)s + L - nibble;
                    }
                }
            } else if (have) {
                const PbRec rec = recs[jb.lo + i];
                slot = (int)(rec.tag & 0xffffu) - slot0;
                valid = slot >= 0 && slot < (FIXED ? 32768 : n_slots);
                rev = (rec.tag >> 16) & 1u;
                x = rec.x; y = rec.y;
            }
            const bool in = valid && !(y <= p0 || x >= plim);
            const bool has_y = in && y < p1;
            const int ox = in ? (int)((x > p0 ? x : p0) - p0) : 0, oy = has_y ? (int)(y - p0) : 0;
            const W w = FIXED ? (W)(in ? __ldg(w_fix + slot) : 0ll) : (W)1;
            const W nw = (W)0 - w;
            const int sl = FIXED ? 0 : slot;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const bool wanted = q == 0 ? want_plus : (q == 1 ? want_minus : want_any);
                if (!wanted) continue;                                  // CTA-uniform
                const bool mine = q == 0 ? !rev : (q == 1 ? rev : true);
                const int a = (q == 0 ? a_plus : (q == 1 ? a_minus : a_any)) + sl;
                pb_run_add<W>(diff, in && mine ? a * T + ox : -1, w);
                pb_run_add<W>(diff, has_y && mine ? a * T + oy : -1, nw);
            }
        }
        pb_fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(hot) + (size_t)jb.hot * hot_stride * sizeof(W);
            for (uint32_t o = 0; o < n_bytes; o += 16384u) {
                const uint32_t n = n_bytes - o < 16384u ? n_bytes - o : 16384u;
                if (FIXED) pb_bulk_add_u64(dst + o, smem_raw + o, n);
                else pb_bulk_add_u32(dst + o, smem_raw + o, n);
            }
            pb_bulk_commit();
            pb_bulk_wait_read0();
        }
        __syncthreads();
    }
    if (stat_slots) pb_flush_cta_stats(drop_p, drop_m, drop_a, drop_len, map_p, map_m, map_a, stat_slots);
    if (threadIdx.x == 0) pb_bulk_wait_all();
}

// One CTA per pile-up tile: scan the tile's scratch difference arrays and write its bins exactly as the tiles kernels
// do — exact: fma over the pass's map lengths in ascending order, stored (first pass) or added to the plane (later
// passes); fixed point: total * 2^-shift.  1024 bins per round (a thread owns 4 consecutive bins), the running sums
// carried from round to round per map length.
template <bool FIXED>
__global__ void __launch_bounds__(kCThreads)
pb_center_finish_hot_kernel(const void *__restrict__ hot, const long long *__restrict__ hot_tiles,
                            const unsigned long long *__restrict__ counters, long long hot_cap, size_t hot_stride,
                            int planes, int slot0, int n_slots, const double *__restrict__ inv_m, double scale, int T,
                            int accumulate, double *__restrict__ out_plus, double *__restrict__ out_minus,
                            double *__restrict__ out_any)
{
    typedef typename PbCenterWord<FIXED>::type W;
    __shared__ long long s_warp[kCWarps];
    __shared__ long long s_carry[kCHotSlots];
    long long nh = (long long)counters[3];
    if (nh > hot_cap) nh = hot_cap;
    double *outs[3];
    int n_planes = 0;
    if (planes & PB_PLANE_PLUS) outs[n_planes++] = out_plus;
    if (planes & PB_PLANE_MINUS) outs[n_planes++] = out_minus;
    if (planes & PB_PLANE_ANY) outs[n_planes++] = out_any;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long h = blockIdx.x; h < nh; h += gridDim.x) {
        const long long tile = hot_tiles[h];
        if (tile < 0) continue;
        const W *A = reinterpret_cast<const W *>(hot) + (size_t)h * hot_stride;
        for (int qq = 0; qq < n_planes; ++qq) {
            __syncthreads();
            if ((int)threadIdx.x < n_slots) s_carry[threadIdx.x] = 0;
            for (int seg = 0; seg < T; seg += 4 * kCThreads) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                for (int sl = 0; sl < n_slots; ++sl) {
                    const W *src = A + (size_t)(qq * n_slots + sl) * T + seg + threadIdx.x * 4;
                    const long long s1 = (long long)src[0], s2 = s1 + (long long)src[1], s3 = s2 + (long long)src[2],
                                    s4 = s3 + (long long)src[3];
                    long long incl = s4;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const long long up = __shfl_up_sync(0xffffffffu, incl, dd);
                        if (lane >= dd) incl += up;
                    }
                    __syncthreads();                         // s_warp / s_carry of the previous round have been read
                    if (lane == 31) s_warp[warp] = incl;
                    __syncthreads();
                    long long before = incl - s4 + s_carry[sl], total = 0;
#pragma unroll
                    for (int w2 = 0; w2 < kCWarps; ++w2) { const long long t2 = s_warp[w2]; total += t2; if (w2 < warp) before += t2; }
                    if (FIXED) {
                        acc[0] = (double)(before + s1) * scale; acc[1] = (double)(before + s2) * scale;
                        acc[2] = (double)(before + s3) * scale; acc[3] = (double)(before + s4) * scale;
                    } else {
                        const double w = __ldg(inv_m + slot0 + sl);
                        acc[0] = __fma_rn(pb_u32_to_f64((int)(before + s1)), w, acc[0]);
                        acc[1] = __fma_rn(pb_u32_to_f64((int)(before + s2)), w, acc[1]);
                        acc[2] = __fma_rn(pb_u32_to_f64((int)(before + s3)), w, acc[2]);
                        acc[3] = __fma_rn(pb_u32_to_f64((int)(before + s4)), w, acc[3]);
                    }
                    __syncthreads();                         // everyone has read s_carry[sl]
                    if (threadIdx.x == 0) s_carry[sl] += total;
                }
                double *dst = outs[qq] + tile * (long long)T + seg + threadIdx.x * 4;
#pragma unroll
                for (int e = 0; e < 4; ++e) dst[e] = accumulate ? dst[e] + acc[e] : acc[e];
            }
        }
    }
}

int pb_center_split()
{
    // PB_CENTER_SPLIT=<reads> (tests, A/B): candidates / records per job; 0 = never split
    if (const char *e = getenv("PB_CENTER_SPLIT")) { const int v = atoi(e); return v <= 0 ? 0x7fffffff : (v < 256 ? 256 : v); }
    return kCSplit;
}

// scratch slots that fit behind the header of tile numbers, given the bytes of one tile's arrays
long long pb_center_hot_cap(const PbWorkspace &ws, size_t tile_bytes, int slots_per_pass)
{
    if (ws.hot_bytes <= kCHotHeader || slots_per_pass > kCHotSlots || pb_center_split() == 0x7fffffff) return 0;
    const long long cap = (long long)((ws.hot_bytes - kCHotHeader) / tile_bytes);
    return cap < kCHotMax ? cap : kCHotMax;
}

// after the tile index and the binning: find the pile-up tiles, cut their candidates and records into jobs
int pb_launch_center_jobs(const PbWorkspace &ws, int lookback, int tile_bins, int64_t tile_lo, int64_t tile_hi,
                          long long hot_cap, cudaStream_t stream)
{
    if (tile_hi <= tile_lo || hot_cap < 1) return PB_OK;
    pb_center_jobs_kernel<<<(unsigned)((tile_hi - tile_lo + 255) / 256), 256, 0, stream>>>(
        ws.tiles, ws.rec_off, lookback, tile_bins, tile_lo, tile_hi, pb_center_split(), ws.jobs, (long long)ws.job_capacity,
        ws.tile_counter, hot_cap, reinterpret_cast<long long *>(ws.hot));
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// every pass, before the tiles kernel: clear the scratch of the pile-up tiles, run the overflow jobs of the pass
template <bool FIXED>
int pb_launch_center_overflow(const PbReads &b, const PbRuleDev &r, int planes, const int16_t *slot_of_len, const long long *w_fix,
                              int s0, int ns, int T, const PbWorkspace &ws, long long hot_cap, size_t hot_stride, bool stats,
                              int sm_count, cudaStream_t stream)
{
    typedef typename PbCenterWord<FIXED>::type W;
    if (hot_cap < 1) return PB_OK;
    unsigned char *arrays = reinterpret_cast<unsigned char *>(ws.hot) + kCHotHeader;
    PB_CUDA_CHECK(cudaMemsetAsync(ws.tile_counter + 2, 0, 8, stream));
    pb_center_zero_hot_kernel<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(reinterpret_cast<uint4 *>(arrays), ws.tile_counter, hot_cap,
                                                                           hot_stride * sizeof(W) / 16);
    const size_t smem = hot_stride * sizeof(W);
    auto kern = pb_center_overflow_kernel<FIXED>;
    PB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCThreads, smem));
    if (occ < 1) occ = 1;
    kern<<<(unsigned)(sm_count * occ), kCThreads, smem, stream>>>(
        b, r, planes, slot_of_len, w_fix, s0, ns, T, b.n_blk > 0, ws.tiles, ws.jobs, ws.tile_counter, (long long)ws.job_capacity,
        ws.tile_counter + 2, ws.recs, arrays, hot_stride, stats ? ws.slots : nullptr);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// every pass, after (or beside) the tiles kernel: the bins of the pile-up tiles
template <bool FIXED>
int pb_launch_center_finish(int planes, int s0, int ns, const double *inv_m, double scale, int T, int accumulate,
                            const PbWorkspace &ws, long long hot_cap, size_t hot_stride, int sm_count,
                            double *out_plus, double *out_minus, double *out_any, cudaStream_t stream)
{
    if (hot_cap < 1) return PB_OK;
    const unsigned char *arrays = reinterpret_cast<const unsigned char *>(ws.hot) + kCHotHeader;
    pb_center_finish_hot_kernel<FIXED><<<(unsigned)(sm_count * 2), kCThreads, 0, stream>>>(
        arrays, reinterpret_cast<const long long *>(ws.hot), ws.tile_counter, hot_cap, hot_stride, planes, s0, ns, inv_m, scale, T,
        accumulate, out_plus, out_minus, out_any);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// staging + zero block + double-buffered difference arrays + three rotating copies of the segment sums
size_t pb_center_smem_bytes(int n_planes, int n_slots, int tile_bins, bool direct)
{
    return ((direct ? 0 : (size_t)n_planes * tile_bins) + kZeroBins) * 8 + (size_t)2 * n_planes * n_slots * tile_bins * 4 +
           (size_t)3 * n_planes * n_slots * (tile_bins / 128) * 4 + 16;
}

template <int EPT, bool DIRECT, int PLANES, bool ONE_SLOT>
int launch_center_kernel(const PbReads &b, const PbRuleDev &r, int planes, const int16_t *slot_of_len, const double *inv_m,
                         int s0, int ns, int pass, int lookback, int64_t tile_begin, int64_t n_tiles, int sm_count,
                         const PbWorkspace &ws, double *out_plus, double *out_minus, double *out_any, cudaStream_t stream,
                         long long hot_cap)
{
    constexpr int T = EPT * kCThreads;
    const int n_planes = __builtin_popcount(planes);
    const size_t smem = pb_center_smem_bytes(n_planes, ns, T, DIRECT);
    // pile-up tiles: overflow jobs leave the difference arrays of this pass's map lengths in the scratch
    const int hns = ns < 1 ? 1 : ns;
    const size_t hot_stride = (size_t)n_planes * hns * T;
    int rc = pb_launch_center_overflow<false>(b, r, planes, slot_of_len, nullptr, s0, hns, T, ws, hot_cap, hot_stride,
                                              pass == 0, sm_count, stream);
    if (rc) return rc;
    auto kern = pb_center_tiles_kernel<EPT, DIRECT, PLANES, ONE_SLOT>;
    PB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCThreads, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sm_count * occ;
    if (grid > n_tiles - tile_begin) grid = n_tiles - tile_begin;
    PB_CUDA_CHECK(cudaMemsetAsync(ws.tile_counter, 0, 8, stream));
    // statistics are accumulated by the first pass only (later passes see the same reads again)
    kern<<<(unsigned)grid, kCThreads, smem, stream>>>(
        b, r, planes, slot_of_len, inv_m, s0, ns, pass > 0, lookback, ws.tiles, tile_begin, n_tiles, ws.tile_counter,
        ws.rec_off, ws.recs, out_plus, out_minus, out_any, pass == 0 ? ws.slots : nullptr);
    PB_CUDA_CHECK(cudaGetLastError());
    return pb_launch_center_finish<false>(planes, s0, ns, inv_m, 0.0, T, pass > 0, ws, ns < 1 ? 0 : hot_cap, hot_stride, sm_count,
                                          out_plus, out_minus, out_any, stream);
}

// specialised instantiations exist for the direct-store kernel on 2048-bin tiles with one map length
// and the plane sets the host layer asks for ('+','-' | '.' | all three); everything else is generic
template <int EPT, bool DIRECT>
int launch_center_pass(const PbReads &b, const PbRuleDev &r, int planes, const int16_t *slot_of_len, const double *inv_m,
                       int s0, int ns, int pass, int lookback, int64_t tile_begin, int64_t n_tiles, int sm_count,
                       const PbWorkspace &ws, double *out_plus, double *out_minus, double *out_any, cudaStream_t stream,
                       long long hot_cap)
{
#define PB_CENTER_ARGS b, r, planes, slot_of_len, inv_m, s0, ns, pass, lookback, tile_begin, n_tiles, sm_count, ws, out_plus, out_minus, out_any, stream, hot_cap
    if (EPT == 8 && DIRECT && ns == 1 && !getenv("PB_CENTER_GENERIC")) {
        if (planes == (PB_PLANE_PLUS | PB_PLANE_MINUS)) return launch_center_kernel<8, true, PB_PLANE_PLUS | PB_PLANE_MINUS, true>(PB_CENTER_ARGS);
        if (planes == PB_PLANE_ANY) return launch_center_kernel<8, true, PB_PLANE_ANY, true>(PB_CENTER_ARGS);
        if (planes == 7) return launch_center_kernel<8, true, 7, true>(PB_CENTER_ARGS);
    }
    return launch_center_kernel<EPT, DIRECT, 0, false>(PB_CENTER_ARGS);
#undef PB_CENTER_ARGS
}

template <int EPT>
int launch_center(const PbReads &b, const PbRuleDev &r, int planes, const int16_t *slot_of_len, const double *inv_m,
                  int n_slots, int per_pass, int lookback, int64_t bin_begin, int64_t bin_end, bool direct_ok,
                  const PbWorkspace &ws, double *out_plus, double *out_minus, double *out_any, cudaStream_t stream)
{
    constexpr int T = EPT * kCThreads;
    const int64_t tile_begin = bin_begin / T, n_tiles = bin_end / T;   // n_tiles = end
    int sm_count = 0;
    int rc = pb_sm_count(&sm_count);
    if (rc) return rc;
    // pile-up tiles: scratch arrays sized for the widest pass (no map length in the batch: nothing to split)
    const int pp = per_pass < 1 ? 1 : per_pass;
    const long long hot_cap = n_slots < 1 ? 0 : pb_center_hot_cap(ws, (size_t)__builtin_popcount(planes) * pp * T * sizeof(int), pp);
    rc = pb_launch_center_jobs(ws, lookback, T, tile_begin, n_tiles, hot_cap, stream);
    if (rc) return rc;
    for (int s0 = 0, pass = 0; s0 < n_slots || pass == 0; s0 += per_pass, ++pass) {
        int ns = n_slots - s0 < per_pass ? n_slots - s0 : per_pass;
        if (ns < 0) ns = 0;
        // the first pass stores every bin (direct 256-bit stores when allowed); later passes add to them
        if (pass == 0 && direct_ok)
            rc = launch_center_pass<EPT, true>(b, r, planes, slot_of_len, inv_m, s0, ns, pass, lookback, tile_begin, n_tiles,
                                               sm_count, ws, out_plus, out_minus, out_any, stream, hot_cap);
        else
            rc = launch_center_pass<EPT, false>(b, r, planes, slot_of_len, inv_m, s0, ns, pass, lookback, tile_begin, n_tiles,
                                                sm_count, ws, out_plus, out_minus, out_any, stream, hot_cap);
        if (rc) return rc;
        if (n_slots == 0) break;
    }
    return PB_OK;
}

}  // namespace

namespace {
// [bin_begin, bin_end) must be multiples of PB_LAYOUT_ALIGN inside the layout (0 .. total_bins = everything)
int pb_check_bin_range(const pb_layout *layout, int64_t bin_begin, int64_t bin_end, const char *who)
{
    if (bin_begin < 0 || bin_end > layout->total_bins || bin_begin > bin_end || bin_begin % PB_LAYOUT_ALIGN ||
        bin_end % PB_LAYOUT_ALIGN) {
        pb_set_error("%s: bin range must lie in the layout on multiples of PB_LAYOUT_ALIGN", who);
        return PB_EINVAL;
    }
    return PB_OK;
}
}  // namespace

extern "C" int pb_map_center_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                                   const int16_t *slot_of_len, const double *inv_m, int n_slots,
                                   double *out_plus, double *out_minus, double *out_any,
                                   uint64_t *stats, void *workspace, size_t workspace_bytes,
                                   int64_t bin_begin, int64_t bin_end, int64_t read_begin, int64_t read_limit,
                                   void *stream_)
{
    int rc = pb_check_common(batch, layout, rule, planes);
    if (rc) return rc;
    rc = pb_check_bin_range(layout, bin_begin, bin_end, "pb_map_center_range");
    if (rc) return rc;
    if (bin_begin == bin_end) return PB_OK;
    if (read_begin < 0) read_begin = 0;
    if (read_limit < 0 || read_limit > batch->n_reads) read_limit = batch->n_reads;
    if (rule->kind != PB_RULE_CENTER || rule->param < 0) { pb_set_error("pb_map_center: need a center rule with nibble >= 0"); return PB_EINVAL; }
    if (!slot_of_len || (n_slots > 0 && !inv_m) || n_slots < 0 || n_slots > 32767) { pb_set_error("pb_map_center: bad slot tables"); return PB_EINVAL; }
    if (((planes & PB_PLANE_PLUS) && !out_plus) || ((planes & PB_PLANE_MINUS) && !out_minus) ||
        ((planes & PB_PLANE_ANY) && !out_any) || !stats) {
        pb_set_error("pb_map_center: missing output plane or stats"); return PB_EINVAL;
    }
    if ((((uintptr_t)out_plus | (uintptr_t)out_minus | (uintptr_t)out_any) & 15) ||
        ((uintptr_t)workspace & 15)) {      // TMA bulk stores / reductions move 16-byte units
        pb_set_error("pb_map_center: planes and workspace must be 16-byte aligned"); return PB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    PbWorkspace ws;
    rc = pb_carve_workspace(workspace, workspace_bytes, layout->total_bins, b.n_blk, b.n_reads, &ws);
    if (rc) return rc;
    const int n_planes = __builtin_popcount(planes);

    // 2048-bin tiles when everything fits in ~75 KB (three CTAs per SM); otherwise 1024-bin tiles, and
    // several accumulating passes over groups of map lengths if even those cannot hold every slot.
    const int ns = n_slots < 1 ? 1 : n_slots;
    int ept = 8, per_pass = ns;
    if (pb_center_smem_bytes(n_planes, ns, 2048, false) > 75 * 1024) {
        ept = 4;
        while (per_pass > 1 && pb_center_smem_bytes(n_planes, per_pass, 1024, false) > 200 * 1024) --per_pass;
    }
    // 256-bit global stores need 32-byte aligned planes (chromosome offsets are multiples of 16384 bins)
    bool direct_ok = true;
    for (double *p : {out_plus, out_minus, out_any})
        if (p && ((uintptr_t)p & 31)) direct_ok = false;
    if (const char *e = getenv("PB_CENTER_EPT")) {       // measurement overrides (profiles/NOTES)
        const int v = atoi(e);
        if ((v == 4 || v == 8 || v == 16) && pb_center_smem_bytes(n_planes, per_pass, v * kCThreads, false) <= 200 * 1024) ept = v;
    }
    if (const char *e = getenv("PB_CENTER_DIRECT")) direct_ok = direct_ok && atoi(e) != 0;
    const int tile_bins = ept * kCThreads;
    const int64_t n_tiles = layout->total_bins / tile_bins;
    const int64_t tile_lo = bin_begin / tile_bins, tile_hi = bin_end / tile_bins;   // PB_LAYOUT_ALIGN is a multiple of every tile size
    const int lookback = (b.max_block_len + tile_bins - 1) / tile_bins;

    PB_CUDA_CHECK(cudaMemsetAsync(ws.slots, 0, 2 * pb_ws_stat_bytes() + 64, stream));
    rc = pb_launch_tile_index(b, lay, tile_bins, tile_lo, tile_hi, read_limit, 0, ws, stream);
    if (rc) return rc;
    rc = pb_launch_binning(b, r, lay, planes, 1, slot_of_len, tile_bins, n_tiles, tile_lo, tile_hi, read_begin, read_limit, ws, stream);
    if (rc) return rc;
    pb_timing_begin(stream);
    if (ept == 16)
        rc = launch_center<16>(b, r, planes, slot_of_len, inv_m, n_slots, per_pass, lookback, bin_begin, bin_end, direct_ok, ws,
                               out_plus, out_minus, out_any, stream);
    else if (ept == 8)
        rc = launch_center<8>(b, r, planes, slot_of_len, inv_m, n_slots, per_pass, lookback, bin_begin, bin_end, direct_ok, ws,
                              out_plus, out_minus, out_any, stream);
    else
        rc = launch_center<4>(b, r, planes, slot_of_len, inv_m, n_slots, per_pass, lookback, bin_begin, bin_end, direct_ok, ws,
                              out_plus, out_minus, out_any, stream);
    pb_timing_end(stream);
    if (rc) return rc;
    return pb_launch_stats_finish(ws.slots, (unsigned long long *)stats, stream);
}

extern "C" int pb_map_center(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                             const int16_t *slot_of_len, const double *inv_m, int n_slots,
                             double *out_plus, double *out_minus, double *out_any,
                             uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!layout) { pb_set_error("pb_map_center: null argument"); return PB_EINVAL; }
    return pb_map_center_range(batch, layout, rule, planes, slot_of_len, inv_m, n_slots, out_plus, out_minus, out_any,
                               stats, workspace, workspace_bytes, 0, layout->total_bins, 0, -1, stream_);
}

namespace {
size_t pb_center_fixed_smem(int n_planes)
{
    constexpr int T = 8 * kCThreads;
    return (size_t)kZeroBins * 8 + (size_t)2 * n_planes * T * 8 + (size_t)3 * n_planes * (T / 128) * 8 + 16;
}

template <int PLANES>
int launch_center_fixed(const PbReads &b, const PbRuleDev &r, int planes, const int16_t *slot_of_len, const long long *w_fix,
                        double scale, int lookback, int64_t tile_begin, int64_t n_tiles, const PbWorkspace &ws,
                        double *out_plus, double *out_minus, double *out_any, cudaStream_t stream)
{
    const size_t smem = pb_center_fixed_smem(__builtin_popcount(planes));
    auto kern = pb_center_fixed_kernel<PLANES>;
    PB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sm_count = 0;
    PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCThreads, smem));
    if (occ < 1) occ = 1;
    int rc = pb_sm_count(&sm_count);
    if (rc) return rc;
    int64_t grid = (int64_t)sm_count * occ;
    if (grid > n_tiles - tile_begin) grid = n_tiles - tile_begin;
    // pile-up tiles: jobs, then their partial 64-bit difference arrays, then the tiles
    constexpr int T = 8 * kCThreads;
    const size_t hot_stride = (size_t)__builtin_popcount(planes) * T;
    const long long hot_cap = pb_center_hot_cap(ws, hot_stride * sizeof(unsigned long long), 1);
    rc = pb_launch_center_jobs(ws, lookback, T, tile_begin, n_tiles, hot_cap, stream);
    if (rc) return rc;
    rc = pb_launch_center_overflow<true>(b, r, planes, slot_of_len, w_fix, 0, 1, T, ws, hot_cap, hot_stride, true, sm_count, stream);
    if (rc) return rc;
    PB_CUDA_CHECK(cudaMemsetAsync(ws.tile_counter, 0, 8, stream));
    kern<<<(unsigned)grid, kCThreads, smem, stream>>>(b, r, planes, slot_of_len, w_fix, scale, lookback, ws.tiles, tile_begin, n_tiles,
                                                      ws.tile_counter, ws.rec_off, ws.recs, out_plus, out_minus, out_any, ws.slots);
    PB_CUDA_CHECK(cudaGetLastError());
    return pb_launch_center_finish<true>(planes, 0, 1, nullptr, scale, T, 0, ws, hot_cap, hot_stride, sm_count, out_plus, out_minus,
                                         out_any, stream);
}
}  // namespace

extern "C" int pb_map_center_fixed_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                                   const int16_t *slot_of_len, const int64_t *w_fix, int n_slots, int shift,
                                   double *out_plus, double *out_minus, double *out_any,
                                   uint64_t *stats, void *workspace, size_t workspace_bytes,
                                   int64_t bin_begin, int64_t bin_end, int64_t read_begin, int64_t read_limit,
                                   void *stream_)
{
    int rc = pb_check_common(batch, layout, rule, planes);
    if (rc) return rc;
    rc = pb_check_bin_range(layout, bin_begin, bin_end, "pb_map_center_fixed_range");
    if (rc) return rc;
    if (bin_begin == bin_end) return PB_OK;
    if (read_begin < 0) read_begin = 0;
    if (read_limit < 0 || read_limit > batch->n_reads) read_limit = batch->n_reads;
    if (rule->kind != PB_RULE_CENTER || rule->param < 0) { pb_set_error("pb_map_center_fixed: need a center rule with nibble >= 0"); return PB_EINVAL; }
    if (!slot_of_len || !w_fix || n_slots < 1 || n_slots > 32767 || shift < 1 || shift > 62) {
        pb_set_error("pb_map_center_fixed: bad weight tables"); return PB_EINVAL;
    }
    if (((planes & PB_PLANE_PLUS) && !out_plus) || ((planes & PB_PLANE_MINUS) && !out_minus) ||
        ((planes & PB_PLANE_ANY) && !out_any) || !stats) {
        pb_set_error("pb_map_center_fixed: missing output plane or stats"); return PB_EINVAL;
    }
    if ((((uintptr_t)out_plus | (uintptr_t)out_minus | (uintptr_t)out_any) & 31) || ((uintptr_t)workspace & 15)) {
        pb_set_error("pb_map_center_fixed: planes must be 32-byte aligned (256-bit stores), workspace 16-byte"); return PB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    PbWorkspace ws;
    rc = pb_carve_workspace(workspace, workspace_bytes, layout->total_bins, b.n_blk, b.n_reads, &ws);
    if (rc) return rc;
    const int tile_bins = 8 * kCThreads;
    const int64_t n_tiles = layout->total_bins / tile_bins;
    const int64_t tile_lo = bin_begin / tile_bins, tile_hi = bin_end / tile_bins;
    const int lookback = (b.max_block_len + tile_bins - 1) / tile_bins;
    const double scale = ldexp(1.0, -shift);

    PB_CUDA_CHECK(cudaMemsetAsync(ws.slots, 0, 2 * pb_ws_stat_bytes() + 64, stream));
    rc = pb_launch_tile_index(b, lay, tile_bins, tile_lo, tile_hi, read_limit, 0, ws, stream);
    if (rc) return rc;
    rc = pb_launch_binning(b, r, lay, planes, 1, slot_of_len, tile_bins, n_tiles, tile_lo, tile_hi, read_begin, read_limit, ws, stream);
    if (rc) return rc;
    pb_timing_begin(stream);
    const long long *w = reinterpret_cast<const long long *>(w_fix);
    if (planes == (PB_PLANE_PLUS | PB_PLANE_MINUS))
        rc = launch_center_fixed<PB_PLANE_PLUS | PB_PLANE_MINUS>(b, r, planes, slot_of_len, w, scale, lookback, tile_lo, tile_hi, ws, out_plus, out_minus, out_any, stream);
    else if (planes == 7)
        rc = launch_center_fixed<7>(b, r, planes, slot_of_len, w, scale, lookback, tile_lo, tile_hi, ws, out_plus, out_minus, out_any, stream);
    else
        rc = launch_center_fixed<0>(b, r, planes, slot_of_len, w, scale, lookback, tile_lo, tile_hi, ws, out_plus, out_minus, out_any, stream);
    pb_timing_end(stream);
    if (rc) return rc;
    return pb_launch_stats_finish(ws.slots, (unsigned long long *)stats, stream);
}

extern "C" int pb_map_center_fixed(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                                   const int16_t *slot_of_len, const int64_t *w_fix, int n_slots, int shift,
                                   double *out_plus, double *out_minus, double *out_any,
                                   uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!layout) { pb_set_error("pb_map_center_fixed: null argument"); return PB_EINVAL; }
    return pb_map_center_fixed_range(batch, layout, rule, planes, slot_of_len, w_fix, n_slots, shift, out_plus, out_minus,
                                     out_any, stats, workspace, workspace_bytes, 0, layout->total_bins, 0, -1, stream_);
}
