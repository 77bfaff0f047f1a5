// pb_regions.cu — region tables over SegmentChain exon blocks (sm_100a): masked sums, window matrices and the
// plane-free count of point-rule sites per chain.
//
// Reference semantics restated: SegmentChain.get_counts / get_masked_counts (plastid/genomics/roitools.pyx:3221-3315),
// the counts_in_region / cs inner loops (plastid/bin/counts_in_region.py:113-125, plastid/bin/cs.py:705-711) and the
// window-matrix fill of metagene / psite (plastid/bin/metagene.py:895-914).
//
// Shape of the work: tens of thousands of chains, each a few exon blocks of a few hundred positions scattered over a
// 12-25 GB plane.  Round 1 walked a chain block by block, 32 positions per dependent step, and ran at a quarter of
// the HBM rate with every warp waiting on one 128-byte row at a time (ncu: 11-15 warps stalled on the long scoreboard
// per issue, profiles/ncu_gather_r02a.txt).  Here the positions of a chain are FLATTENED over the lanes of its warp:
// the warp loads the bounds of up to 32 blocks at once (one round trip), a shuffle scan gives every block its first
// unit, and every lane then owns units u = lane, lane + 32, ... of the whole chain — four independent loads per lane
// are in flight before the first value is used, whatever the block structure.  Sums read 16-byte aligned groups
// (uint4 / double2), trimming the group's head and tail against the block; window rows are written column by column
// from the same flattened index.
//
// Every kernel takes the global-bin range [lo, hi) the calling rank owns (position sharding): positions
// outside it count zero but keep their place in the chain, so partial tables of all ranks add up to the whole table
// and geometry-only outputs (unmasked lengths, mask matrices, NaN cells) are identical on every rank.
#include "pb_tiles.cuh"
#include <math.h>

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct PbPlanes { const void *p[3]; };

template <typename T> struct VecOf;
template <> struct VecOf<uint32_t> { static constexpr int V = 4; typedef uint4 L; typedef unsigned long long Acc; };
template <> struct VecOf<double> { static constexpr int V = 2; typedef double2 L; typedef double Acc; };

__device__ __forceinline__ long long shfl_ll(long long v, int src)
{
    return __shfl_sync(kFull, v, src);
}

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

// bits [b0, b0 + n) (1 <= n <= 32) of a little-endian bit array held as 32-bit words
__device__ __forceinline__ uint32_t mask_bits_at(const uint32_t *__restrict__ words, long long b0, int n)
{
    const long long w = b0 >> 5;
    const int sh = (int)(b0 & 31);
    uint32_t bits = __ldg(words + w) >> sh;
    if (sh + n > 32) bits |= __ldg(words + w + 1) << (32 - sh);
    return n == 32 ? bits : (bits & ((1u << n) - 1u));
}

// number of set bits in [b0, b0 + n) by a whole warp
__device__ __forceinline__ long long warp_popcount_bits(const uint32_t *__restrict__ words, long long b0, long long n, int lane)
{
    if (n <= 0) return 0;
    const long long wa = b0 >> 5, wb = (b0 + n - 1) >> 5;
    long long cnt = 0;
    for (long long w = wa + lane; w <= wb; w += 32) {
        uint32_t x = __ldg(words + w);
        if (w == wa) x &= kFull << (b0 & 31);
        if (w == wb) x &= kFull >> (31 - ((b0 + n - 1) & 31));
        cnt += __popc(x);
    }
    return (long long)pb_warp_sum((unsigned long long)cnt);
}

template <typename T> __device__ __forceinline__ void add_group(typename VecOf<T>::Acc &acc, const typename VecOf<T>::L &v, uint32_t m);
template <> __device__ __forceinline__ void add_group<uint32_t>(unsigned long long &acc, const uint4 &v, uint32_t m)
{
    acc += (unsigned long long)((m & 1u) ? v.x : 0u) + ((m & 2u) ? v.y : 0u);
    acc += (unsigned long long)((m & 4u) ? v.z : 0u) + ((m & 8u) ? v.w : 0u);
}
template <> __device__ __forceinline__ void add_group<double>(double &acc, const double2 &v, uint32_t m)
{
    if (m & 1u) acc += v.x;
    if (m & 2u) acc += v.y;
}

// ------------------------------------------------------------------------------------------------------------
// masked sums: one warp per chain, 16-byte groups flattened over the lanes
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool MASK>
__global__ void __launch_bounds__(256)
pb_region_sums_kernel(PbPlanes planes, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                      const int64_t *__restrict__ chain_off, const uint8_t *__restrict__ chain_plane, int64_t n_chains,
                      const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                      long long lo, long long hi, double *__restrict__ sums, int64_t *__restrict__ live_len)
{
    constexpr int V = VecOf<T>::V, U = 4;
    typedef typename VecOf<T>::L L;
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_chains) return;
    const T *__restrict__ vec = static_cast<const T *>(planes.p[__ldg(chain_plane + c)]);
    const int64_t k0 = __ldg(chain_off + c), k1 = __ldg(chain_off + c + 1);
    const long long moff = MASK ? __ldg(mask_off + c) : 0;
    typename VecOf<T>::Acc acc = 0;
    long long jbase = 0;                     // chain position of the batch's first block
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int nb = (int)(k1 - kb < 32 ? k1 - kb : 32);
        long long bs = 0, be = 0;
        if (lane < nb) { bs = __ldg(bstart + kb + lane); be = __ldg(bend + kb + lane); }
        const long long cs = bs > lo ? bs : lo, ce = be < hi ? be : hi;      // the part this rank owns
        const long long g0 = cs & ~(long long)(V - 1);
        const int units = cs < ce ? (int)((ce - g0 + V - 1) / V) : 0;
        int incl_u = units;
        long long incl_len = be - bs;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up_u = __shfl_up_sync(kFull, incl_u, d);
            const long long up_l = __shfl_up_sync(kFull, incl_len, d);
            if (lane >= d) { incl_u += up_u; incl_len += up_l; }
        }
        const int pre_u = incl_u - units;
        const int total_u = __shfl_sync(kFull, incl_u, 31);
        const int head = (int)(cs - g0);                                   // elements of the first group before cs
        const int span = (int)(ce - g0);                                   // end of the owned part, relative to g0
        const long long mbase = moff + jbase + (incl_len - (be - bs)) + (g0 - bs);   // mask bit of element g0
        jbase += shfl_ll(incl_len, 31);
        for (int u0 = 0; u0 < total_u; u0 += 32 * U) {
            int uu[U], blk[U];
#pragma unroll
            for (int x = 0; x < U; ++x) { uu[x] = u0 + x * 32 + lane; blk[x] = 0; }
            for (int i = 1; i < nb; ++i) {
                const int s = __shfl_sync(kFull, pre_u, i);
#pragma unroll
                for (int x = 0; x < U; ++x) blk[x] += (uu[x] >= s);
            }
            L v[U];
            int rel[U], hd[U], sp[U];
            long long mb[U];
            bool ok[U];
#pragma unroll
            for (int x = 0; x < U; ++x) {
                const long long g = shfl_ll(g0, blk[x]);
                const int pre = __shfl_sync(kFull, pre_u, blk[x]);
                hd[x] = __shfl_sync(kFull, head, blk[x]);
                sp[x] = __shfl_sync(kFull, span, blk[x]);
                if (MASK) mb[x] = shfl_ll(mbase, blk[x]);
                ok[x] = uu[x] < total_u;
                rel[x] = (uu[x] - pre) * V;                                // first element of the group, relative to g0
                if (ok[x]) v[x] = __ldg(reinterpret_cast<const L *>(vec + g + rel[x]));
            }
#pragma unroll
            for (int x = 0; x < U; ++x) {
                if (!ok[x]) continue;
                const int e_lo = hd[x] > rel[x] ? hd[x] - rel[x] : 0;
                const int e_hi = sp[x] - rel[x] < V ? sp[x] - rel[x] : V;
                uint32_t m = ((1u << e_hi) - 1u) & ~((1u << e_lo) - 1u);
                if (MASK) m &= ~(mask_bits_at(mask_words, mb[x] + rel[x] + e_lo, e_hi - e_lo) << e_lo);
                add_group<T>(acc, v[x], m);
            }
        }
    }
    long long live = jbase;
    if (MASK) live -= warp_popcount_bits(mask_words, moff, jbase, lane);
    double total;
    if (sizeof(T) == 4) total = (double)pb_warp_sum((unsigned long long)acc);
    else total = warp_sum_f64((double)acc);
    if (lane == 0) { sums[c] = total; live_len[c] = live; }
}

// ------------------------------------------------------------------------------------------------------------
// window matrices: one warp per chain, chain positions flattened over the lanes, one store per cell
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool MASK>
__global__ void __launch_bounds__(256)
pb_gather_windows_kernel(PbPlanes planes, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                         const int64_t *__restrict__ chain_off, const uint8_t *__restrict__ chain_plane,
                         const uint8_t *__restrict__ chain_reverse, const int32_t *__restrict__ row_col,
                         const int64_t *__restrict__ row_off, int64_t n_chains, int32_t width_,
                         const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                         long long lo, long long hi, double *__restrict__ matrix, uint8_t *__restrict__ maskmat)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_chains) return;
    const T *__restrict__ vec = static_cast<const T *>(planes.p[__ldg(chain_plane + c)]);
    const int64_t k0 = __ldg(chain_off + c), k1 = __ldg(chain_off + c + 1);
    const bool rev = __ldg(chain_reverse + c);
    const long long moff = MASK ? __ldg(mask_off + c) : 0;
    long long len = 0;
    for (int64_t k = k0 + lane; k < k1; k += 32) len += __ldg(bend + k) - __ldg(bstart + k);
    len = (long long)pb_warp_sum((unsigned long long)len);
    // window rows (row_off == NULL): row c of a width-wide matrix, the chain laid from column row_col[c];
    // ragged rows: chain c owns cells [row_off[c], row_off[c] + its length) of a flat vector
    const long long col0 = row_off ? 0 : __ldg(row_col + c);
    const long long width = row_off ? len : width_;
    const int64_t cell0 = row_off ? __ldg(row_off + c) : c * (int64_t)width_;
    double *__restrict__ row = matrix + cell0;
    uint8_t *__restrict__ mrow = maskmat + cell0;
    // columns no chain position reaches stay "masked NaN" (metagene.py:895-898)
    for (long long col = lane; col < width; col += 32)
        if (col < col0 || col >= col0 + len) { row[col] = nan(""); mrow[col] = 1; }
    long long jbase = 0;
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int nb = (int)(k1 - kb < 32 ? k1 - kb : 32);
        long long bs = 0, be = 0;
        if (lane < nb) { bs = __ldg(bstart + kb + lane); be = __ldg(bend + kb + lane); }
        const int n_i = (int)(be - bs);
        int incl = n_i;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += up;
        }
        const int pre = incl - n_i;
        const int total = __shfl_sync(kFull, incl, 31);
        for (int t0 = 0; t0 < total; t0 += 32 * U) {
            int tt[U], blk[U];
#pragma unroll
            for (int x = 0; x < U; ++x) { tt[x] = t0 + x * 32 + lane; blk[x] = 0; }
            for (int i = 1; i < nb; ++i) {
                const int s = __shfl_sync(kFull, pre, i);
#pragma unroll
                for (int x = 0; x < U; ++x) blk[x] += (tt[x] >= s);
            }
            T v[U];
            bool ok[U];
#pragma unroll
            for (int x = 0; x < U; ++x) {
                const long long b0 = shfl_ll(bs, blk[x]);
                const int p0 = __shfl_sync(kFull, pre, blk[x]);
                ok[x] = tt[x] < total;
                const long long p = b0 + (tt[x] - p0);
                v[x] = T(0);
                if (ok[x] && p >= lo && p < hi) v[x] = __ldg(vec + p);
            }
#pragma unroll
            for (int x = 0; x < U; ++x) {
                if (!ok[x]) continue;
                const long long jj = jbase + tt[x];
                const long long col = col0 + (rev ? (len - 1 - jj) : jj);
                if (col >= 0 && col < width) {
                    row[col] = (double)v[x];
                    mrow[col] = MASK ? (uint8_t)mask_bits_at(mask_words, moff + jj, 1) : (uint8_t)0;
                }
            }
        }
        jbase += total;
    }
}

// ------------------------------------------------------------------------------------------------------------
// plane-free region counts of a point rule: one CTA per chain walks the reads that can map into each of its
// blocks (a contiguous slice of the coordinate-sorted batch) — no count vectors are materialised.
// Equals pb_region_sums over the planes pb_map_point would write (same strand pre-filter per query strand,
// genome_array.py:811-815; same rule direction; same masks).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pb_chain_counts_kernel(PbReads b, PbRuleDev r, PbLayoutDev lay,
                       const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                       const int64_t *__restrict__ chain_off, const uint8_t *__restrict__ chain_plane, int64_t n_chains,
                       const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                       long long lo, long long hi, long long total_bins,
                       double *__restrict__ sums, int64_t *__restrict__ live_len, unsigned long long *__restrict__ stats)
{
    __shared__ unsigned long long s_count;
    __shared__ unsigned int s_drop, s_drop_len;
    const int64_t c = blockIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_count = 0; s_drop = 0; s_drop_len = 0; }
    __syncthreads();
    const int64_t k0 = __ldg(chain_off + c), k1 = __ldg(chain_off + c + 1);
    const int plane = __ldg(chain_plane + c);            // 0 '+', 1 '-', 2 '.'
    const bool rq = plane == 1;                          // rule direction follows the chain's strand
    const long long moff = mask_words ? __ldg(mask_off + c) : 0;
    unsigned long long count = 0;
    unsigned int drop = 0, drop_len = 0;
    long long j0 = 0;
    for (int64_t k = k0; k < k1; ++k) {
        const long long gs = __ldg(bstart + k), ge = __ldg(bend + k);
        const long long cs = gs > lo ? gs : lo, ce = ge < hi ? ge : hi;      // owned part, global bins
        if (cs < ce && gs >= 0 && gs < total_bins) {
            const int ch = pb_chrom_of_bin(lay, gs);
            const long long base = __ldg(lay.chrom_bin_off + ch);
            const long long ps = cs - base, pe = ce - base;                  // chromosome coordinates
            int64_t r0 = 0, r1 = 0;
            if (ch < b.n_chrom) { r0 = __ldg(b.chrom_read_off + ch); r1 = __ldg(b.chrom_read_off + ch + 1); }
            const int64_t first = pb_lower_bound_warp(b.ref_start, r0, r1, ps - b.max_span + 1);
            const int64_t last = pb_lower_bound_warp(b.ref_start, first, r1, pe);
            constexpr int kU = 4;
            for (int64_t i0 = first + threadIdx.x; i0 < last; i0 += (int64_t)kU * blockDim.x) {
                uint32_t mv[kU];
                int32_t sv[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int64_t i = i0 + (int64_t)u * blockDim.x;
                    mv[u] = i < last ? __ldg(b.meta + i) : (1u << 17);
                    sv[u] = i < last ? __ldg(b.ref_start + i) : 0;
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t m = mv[u];
                    if (!pb_passes(m, r.size_min, r.size_max)) continue;
                    const bool rev = PB_META_REV(m);
                    if ((plane == 0 && rev) || (plane == 1 && !rev)) continue;
                    const int L = PB_META_L(m);
                    const int idx = pb_rule_index(r, L, rq);
                    if (idx < 0) { drop = 1; drop_len = L; continue; }
                    const long long p = pb_position(b, i0 + (int64_t)u * blockDim.x, sv[u], m, idx);
                    if (p < ps || p >= pe) continue;
                    if (mask_words && mask_bits_at(mask_words, moff + j0 + (base + p - gs), 1)) continue;
                    count++;
                }
            }
        }
        j0 += ge - gs;
    }
    count = pb_warp_sum(count);
    drop = __reduce_or_sync(kFull, drop);
    drop_len = __reduce_max_sync(kFull, drop_len);
    if (lane == 0) {
        if (count) atomicAdd(&s_count, count);
        if (drop) { atomicOr(&s_drop, 1u); atomicMax(&s_drop_len, drop_len); }
    }
    __syncthreads();
    long long live = j0;
    if (threadIdx.x < 32 && mask_words) live -= warp_popcount_bits(mask_words, moff, j0, lane);
    if (threadIdx.x == 0) {
        sums[c] = (double)s_count;
        live_len[c] = live;
        if (s_drop && stats) {
            const int which = plane == 0 ? PB_STAT_DROPPED_PLUS : (plane == 1 ? PB_STAT_DROPPED_MINUS : PB_STAT_DROPPED_ANY);
            atomicAdd(stats + which, 1ull);
            atomicMax(stats + PB_STAT_DROPPED_LEN, (unsigned long long)s_drop_len);
        }
    }
}

int check_chains(const void *const *planes, const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                 const uint8_t *chain_plane, int64_t n_chains, const uint8_t *mask_bits, const int64_t *mask_off, int vec_dtype,
                 int64_t bin_begin, int64_t bin_end)
{
    if (!planes || !bstart || !bend || !chain_off || !chain_plane) { pb_set_error("gather: null chain tables"); return PB_EINVAL; }
    if (n_chains < 0) { pb_set_error("gather: negative chain count"); return PB_EINVAL; }
    if (mask_bits && !mask_off) { pb_set_error("gather: mask_bits without mask_off"); return PB_EINVAL; }
    if (mask_bits && ((uintptr_t)mask_bits & 3)) { pb_set_error("gather: mask_bits must be 4-byte aligned (and padded to whole words)"); return PB_EINVAL; }
    if (vec_dtype != 0 && vec_dtype != 1) { pb_set_error("gather: vec_dtype must be 0 (uint32) or 1 (float64)"); return PB_EINVAL; }
    if (bin_begin > bin_end) { pb_set_error("gather: empty or inverted bin range"); return PB_EINVAL; }
    return PB_OK;
}

}  // namespace

extern "C" int pb_region_sums_range(const void *const *planes, int vec_dtype,
                                    const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                    const uint8_t *chain_plane, int64_t n_chains,
                                    const uint8_t *mask_bits, const int64_t *mask_off,
                                    int64_t bin_begin, int64_t bin_end,
                                    double *sums, int64_t *live_len, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off, vec_dtype, bin_begin, bin_end);
    if (rc) return rc;
    if (!sums || !live_len) { pb_set_error("pb_region_sums: null outputs"); return PB_EINVAL; }
    for (int i = 0; i < 3; ++i)
        if ((uintptr_t)planes[i] & 15) { pb_set_error("pb_region_sums: planes must be 16-byte aligned"); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbPlanes pl{{planes[0], planes[1], planes[2]}};
    const unsigned grid = (unsigned)((n_chains * 32 + 255) / 256);
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask_bits);
#define PB_LAUNCH_SUMS(T, M) pb_region_sums_kernel<T, M><<<grid, 256, 0, stream>>>(pl, bstart, bend, chain_off, chain_plane, n_chains, mw, mask_off, bin_begin, bin_end, sums, live_len)
    if (vec_dtype == 0) { if (mw) PB_LAUNCH_SUMS(uint32_t, true); else PB_LAUNCH_SUMS(uint32_t, false); }
    else { if (mw) PB_LAUNCH_SUMS(double, true); else PB_LAUNCH_SUMS(double, false); }
#undef PB_LAUNCH_SUMS
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_region_sums(const void *const *planes, int vec_dtype,
                              const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                              const uint8_t *chain_plane, int64_t n_chains,
                              const uint8_t *mask_bits, const int64_t *mask_off,
                              double *sums, int64_t *live_len, void *stream_)
{
    return pb_region_sums_range(planes, vec_dtype, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off,
                                0, INT64_MAX, sums, live_len, stream_);
}

extern "C" int pb_gather_windows_range(const void *const *planes, int vec_dtype,
                                       const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                       const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                       const int32_t *row_col, int64_t n_chains, int32_t width,
                                       const uint8_t *mask_bits, const int64_t *mask_off,
                                       int64_t bin_begin, int64_t bin_end,
                                       double *matrix, uint8_t *maskmat, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off, vec_dtype, bin_begin, bin_end);
    if (rc) return rc;
    if (!chain_reverse || !row_col || !matrix || !maskmat || width <= 0) { pb_set_error("pb_gather_windows: bad arguments"); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbPlanes pl{{planes[0], planes[1], planes[2]}};
    const unsigned grid = (unsigned)((n_chains * 32 + 255) / 256);
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask_bits);
#define PB_LAUNCH_WIN(T, M) pb_gather_windows_kernel<T, M><<<grid, 256, 0, stream>>>(pl, bstart, bend, chain_off, chain_plane, chain_reverse, row_col, nullptr, n_chains, width, mw, mask_off, bin_begin, bin_end, matrix, maskmat)
    if (vec_dtype == 0) { if (mw) PB_LAUNCH_WIN(uint32_t, true); else PB_LAUNCH_WIN(uint32_t, false); }
    else { if (mw) PB_LAUNCH_WIN(double, true); else PB_LAUNCH_WIN(double, false); }
#undef PB_LAUNCH_WIN
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_gather_chains_range(const void *const *planes, int vec_dtype,
                                      const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                      const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                      const int64_t *row_off, int64_t n_chains,
                                      const uint8_t *mask_bits, const int64_t *mask_off,
                                      int64_t bin_begin, int64_t bin_end,
                                      double *values, uint8_t *masked, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off, vec_dtype, bin_begin, bin_end);
    if (rc) return rc;
    if (!chain_reverse || !row_off || !values || !masked) { pb_set_error("pb_gather_chains: bad arguments"); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbPlanes pl{{planes[0], planes[1], planes[2]}};
    const unsigned grid = (unsigned)((n_chains * 32 + 255) / 256);
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask_bits);
#define PB_LAUNCH_FLAT(T, M) pb_gather_windows_kernel<T, M><<<grid, 256, 0, stream>>>(pl, bstart, bend, chain_off, chain_plane, chain_reverse, nullptr, row_off, n_chains, 0, mw, mask_off, bin_begin, bin_end, values, masked)
    if (vec_dtype == 0) { if (mw) PB_LAUNCH_FLAT(uint32_t, true); else PB_LAUNCH_FLAT(uint32_t, false); }
    else { if (mw) PB_LAUNCH_FLAT(double, true); else PB_LAUNCH_FLAT(double, false); }
#undef PB_LAUNCH_FLAT
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_gather_windows(const void *const *planes, int vec_dtype,
                                 const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                 const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                 const int32_t *row_col, int64_t n_chains, int32_t width,
                                 const uint8_t *mask_bits, const int64_t *mask_off,
                                 double *matrix, uint8_t *maskmat, void *stream_)
{
    return pb_gather_windows_range(planes, vec_dtype, bstart, bend, chain_off, chain_plane, chain_reverse, row_col, n_chains,
                                   width, mask_bits, mask_off, 0, INT64_MAX, matrix, maskmat, stream_);
}

extern "C" int pb_chain_counts(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                               const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                               const uint8_t *chain_plane, int64_t n_chains,
                               const uint8_t *mask_bits, const int64_t *mask_off,
                               int64_t bin_begin, int64_t bin_end,
                               double *sums, int64_t *live_len, uint64_t *stats, void *stream_)
{
    if (!batch || !layout || !rule || !bstart || !bend || !chain_off || !chain_plane || !sums || !live_len || n_chains < 0) {
        pb_set_error("pb_chain_counts: null argument"); return PB_EINVAL;
    }
    if (rule->kind != PB_RULE_FIVEPRIME && rule->kind != PB_RULE_THREEPRIME && rule->kind != PB_RULE_VARIABLE) {
        pb_set_error("pb_chain_counts: needs a point rule (5' / 3' / variable)"); return PB_EINVAL;
    }
    if (rule->kind == PB_RULE_VARIABLE && (!rule->lut_fw || !rule->lut_rc)) { pb_set_error("pb_chain_counts: variable rule needs LUTs"); return PB_EINVAL; }
    if (mask_bits && (!mask_off || ((uintptr_t)mask_bits & 3))) { pb_set_error("pb_chain_counts: mask_bits need mask_off and 4-byte alignment"); return PB_EINVAL; }
    if (bin_begin > bin_end) { pb_set_error("pb_chain_counts: inverted bin range"); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    if (n_chains > 0x7fffffffll) { pb_set_error("pb_chain_counts: too many chains for one launch"); return PB_EINVAL; }
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    pb_chain_counts_kernel<<<(unsigned)n_chains, 128, 0, (cudaStream_t)stream_>>>(
        b, r, lay, bstart, bend, chain_off, chain_plane, n_chains, reinterpret_cast<const uint32_t *>(mask_bits), mask_off,
        bin_begin, bin_end, layout->total_bins, sums, live_len, reinterpret_cast<unsigned long long *>(stats));
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
