// pb_gather.cu — gather / segmented reductions over SegmentChain exon blocks (sm_100a).
//
// Reference semantics restated: SegmentChain.get_counts / get_masked_counts
// (plastid/genomics/roitools.pyx:3221-3315), the counts_in_region / cs inner loops
// (plastid/bin/counts_in_region.py:113-125, plastid/bin/cs.py:705-711) and the metagene / psite
// window matrices and profiles (plastid/bin/metagene.py:895-960, plastid/bin/psite.py:204-234).
// All of this is HBM/L2-bound gather work: one warp walks one chain with coalesced 128-byte
// reads of the dense count plane.
#include "pb_tiles.cuh"
#include <stdlib.h>
#include <math.h>

namespace {

struct PbPlanes { const void *p[3]; };

template <typename T> struct Acc;
template <> struct Acc<uint32_t> { typedef unsigned long long type; };
template <> struct Acc<double> { typedef double type; };

__device__ __forceinline__ double pb_warp_sum_f64(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ bool pb_mask_bit(const uint8_t *__restrict__ bits, int64_t bit)
{
    return (__ldg(bits + (bit >> 3)) >> (bit & 7)) & 1;
}

// phase_by_size.py:197-214: counts laid 5'->3', cut into codons, `[front:back]` codon slice, summed
// per sub-codon phase.  One warp per chain; out[c*3 + phase].
template <typename T>
__global__ void __launch_bounds__(256)
pb_phase_sums_kernel(PbPlanes planes, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                     const int64_t *__restrict__ chain_off, const uint8_t *__restrict__ chain_plane,
                     const uint8_t *__restrict__ chain_reverse, int64_t n_chains, int32_t front, int32_t back,
                     long long lo, long long hi, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_chains) return;
    const T *__restrict__ vec = static_cast<const T *>(planes.p[__ldg(chain_plane + c)]);
    const int64_t k0 = __ldg(chain_off + c), k1 = __ldg(chain_off + c + 1);
    int64_t len = 0;
    for (int64_t k = k0; k < k1; ++k) len += __ldg(bend + k) - __ldg(bstart + k);
    const int64_t ncod = len / 3;                       // a trailing partial codon is ignored (:203-212)
    int64_t cod_lo = front < 0 ? ncod + front : front;  // python slice semantics for [front:back]
    int64_t cod_hi = back < 0 ? ncod + back : back;
    cod_lo = cod_lo < 0 ? 0 : (cod_lo > ncod ? ncod : cod_lo);
    cod_hi = cod_hi < 0 ? 0 : (cod_hi > ncod ? ncod : cod_hi);
    const bool rev = __ldg(chain_reverse + c);
    typename Acc<T>::type acc0 = 0, acc1 = 0, acc2 = 0;
    int64_t j = 0;
    for (int64_t k = k0; k < k1; ++k) {
        const int64_t bs = __ldg(bstart + k), be = __ldg(bend + k);
        for (int64_t p = bs + lane; p < be; p += 32) {
            const int64_t jj = j + (p - bs);
            const int64_t t = rev ? (len - 1 - jj) : jj;  // 5'->3' index
            const int64_t cod = t / 3;
            if (cod >= cod_lo && cod < cod_hi && p >= lo && p < hi) {     // positions of other ranks count zero
                const int ph = (int)(t - cod * 3);
                const T v = vec[p];
                if (ph == 0) acc0 += v; else if (ph == 1) acc1 += v; else acc2 += v;
            }
        }
        j += be - bs;
    }
    double r0, r1, r2;
    if (sizeof(T) == 4) {
        r0 = (double)pb_warp_sum((unsigned long long)acc0);
        r1 = (double)pb_warp_sum((unsigned long long)acc1);
        r2 = (double)pb_warp_sum((unsigned long long)acc2);
    } else {
        r0 = pb_warp_sum_f64((double)acc0); r1 = pb_warp_sum_f64((double)acc1); r2 = pb_warp_sum_f64((double)acc2);
    }
    if (lane == 0) { out[c * 3] = r0; out[c * 3 + 1] = r1; out[c * 3 + 2] = r2; }
}

// one warp per row: denominator, selection, normalisation (metagene.py:918-924)
__global__ void __launch_bounds__(256)
pb_window_normalize_kernel(const double *__restrict__ matrix, const uint8_t *__restrict__ maskmat,
                           int64_t n_rows, int32_t width, int32_t norm_lo, int32_t norm_hi, double min_counts,
                           double *__restrict__ denom, uint8_t *__restrict__ row_select,
                           double *__restrict__ norm, uint8_t *__restrict__ normmask)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const double *row = matrix + r * (int64_t)width;
    const uint8_t *mrow = maskmat + r * (int64_t)width;
    double acc = 0.0;
    unsigned long long live = 0;
    for (int col = norm_lo + lane; col < norm_hi; col += 32) {
        if (col >= 0 && col < width && !mrow[col]) { acc += row[col]; live++; }
    }
    acc = pb_warp_sum_f64(acc);
    live = pb_warp_sum(live);
    const bool den_masked = (live == 0);  // nansum of an all-masked slice is the masked constant
    const double d = den_masked ? nan("") : acc;
    if (lane == 0) {
        denom[r] = d;
        row_select[r] = (!den_masked && acc >= min_counts) ? 1 : 0;
    }
    if (norm) {
        for (int col = lane; col < width; col += 32) {
            const double v = row[col] / d;
            norm[r * (int64_t)width + col] = v;
            normmask[r * (int64_t)width + col] = (mrow[col] || den_masked || isnan(v) || isinf(v)) ? 1 : 0;
        }
    }
}

// order-preserving map double -> uint64
__device__ __forceinline__ unsigned long long pb_key_of(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double pb_value_of(unsigned long long k)
{
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// transpose selected, unmasked cells into per-column key lists; invalid cells become the max key
__global__ void pb_column_keys_kernel(const double *__restrict__ norm, const uint8_t *__restrict__ normmask,
                                      const uint8_t *__restrict__ row_select, int64_t n_rows, int32_t width,
                                      unsigned long long *__restrict__ keys)
{
    __shared__ unsigned long long tile[32][33];
    // blockIdx.z = matrix of a batch (psite: one per read length); each has its own rows, select flags, keys
    norm += (int64_t)blockIdx.z * n_rows * width;
    normmask += (int64_t)blockIdx.z * n_rows * width;
    row_select += (int64_t)blockIdx.z * n_rows;
    keys += (int64_t)blockIdx.z * n_rows * width;
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int64_t r = r0 + dy;
        const int col = c0 + threadIdx.x;
        unsigned long long k = ~0ull;
        if (r < n_rows && col < width && row_select[r] && !normmask[r * (int64_t)width + col])
            k = pb_key_of(norm[r * (int64_t)width + col]);
        tile[dy][threadIdx.x] = k;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int col = c0 + dy;
        const int64_t r = r0 + threadIdx.x;
        if (col < width && r < n_rows) keys[(int64_t)col * n_rows + r] = tile[threadIdx.x][dy];
    }
}

// psite.py:200-234 for integer count matrices in one pass over the counts: per row the denominator over
// the normalisation window (unmasked cells only; all masked => masked row), the row's selection flag,
// and — instead of materialising float64, normalised and mask matrices — the order-preserving keys of
// the normalised cells written straight into the per-column key lists the median select reads.
// Integer counts sum exactly in fp64, so the denominator does not depend on summation order; the
// quotient is the same IEEE division `matrix / denom` the separate kernels perform.
// grid = (ceil(n_rows / 32), n_batch); the position mask may be shared by all matrices of the batch.
__global__ void __launch_bounds__(256)
pb_norm_keys_u32_kernel(const uint32_t *__restrict__ counts, const uint8_t *__restrict__ maskmat, int mask_shared,
                        int64_t n_rows, int32_t width, int32_t norm_lo, int32_t norm_hi, double min_counts,
                        uint8_t *__restrict__ row_select, unsigned long long *__restrict__ keys)
{
    __shared__ double s_den[32];
    __shared__ unsigned char s_sel[32];
    __shared__ unsigned long long tile[32][33];
    const int64_t mat = (int64_t)blockIdx.y * n_rows * width;
    counts += mat;
    keys += mat;
    if (!mask_shared) maskmat += mat;
    row_select += (int64_t)blockIdx.y * n_rows;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp * 4; rr < warp * 4 + 4; ++rr) {
        const int64_t r = r0 + rr;
        double acc = 0.0;
        int live = 0;
        if (r < n_rows)
            for (int col = norm_lo + lane; col < norm_hi && col < width; col += 32)
                if (col >= 0 && !maskmat[r * width + col]) { acc += (double)counts[r * width + col]; live++; }
        acc = pb_warp_sum_f64(acc);
        live = __reduce_add_sync(0xffffffffu, live);
        if (lane == 0) {
            const bool den_masked = (live == 0);       // nansum of an all-masked slice is the masked constant
            s_den[rr] = den_masked ? nan("") : acc;
            s_sel[rr] = (!den_masked && acc >= min_counts) ? 1 : 0;
            if (r < n_rows) row_select[r] = s_sel[rr];
        }
    }
    __syncthreads();
    for (int c0 = 0; c0 < width; c0 += 32) {
        for (int dy = warp; dy < 32; dy += 8) {
            const int64_t r = r0 + dy;
            const int col = c0 + lane;
            unsigned long long k = ~0ull;
            if (r < n_rows && col < width && s_sel[dy] && !maskmat[r * width + col]) {
                const double v = (double)counts[r * width + col] / s_den[dy];
                if (!(isnan(v) || isinf(v))) k = pb_key_of(v);                 // metagene.py:923-924
            }
            tile[dy][lane] = k;
        }
        __syncthreads();
        for (int dy = warp; dy < 32; dy += 8) {
            const int col = c0 + dy;
            const int64_t r = r0 + lane;
            if (col < width && r < n_rows) keys[(int64_t)col * n_rows + r] = tile[lane][dy];
        }
        __syncthreads();
    }
}

// One CTA per column: count / sum of valid cells, and the two middle order statistics by
// 8-bit-digit radix select (exact; numpy.ma.median averages the two middle values).
// Three passes read the whole column: (1) count, sum and the histogram of the top digit, (2) the
// second digit, (3) the third digit while the keys that match the 16 selected bits are COMPACTED
// into `cand`; the remaining five digits and the successor search only walk the candidates (a few
// percent of the column for normalised counts, whose top 16 bits — sign, exponent, 4 mantissa bits —
// already separate most values).
__device__ __forceinline__ void pb_select_digit(const unsigned int *hist, unsigned long long rank, int shift,
                                                unsigned long long prefix, unsigned long long *s_prefix,
                                                unsigned long long *s_rank)
{
    unsigned long long run = 0;
    int d = 0;
    for (; d < 256; ++d) {
        if (run + hist[d] > rank) break;
        run += hist[d];
    }
    *s_prefix = prefix | ((unsigned long long)d << shift);
    *s_rank = rank - run;
}

__global__ void __launch_bounds__(512)
pb_column_stats_kernel(const unsigned long long *__restrict__ keys, unsigned long long *__restrict__ cand_all,
                       int64_t n_rows, int32_t width, int mode,
                       double *__restrict__ profile, int64_t *__restrict__ n_regions, double *__restrict__ col_sum)
{
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_cnt, s_prefix, s_rank, s_succ;
    __shared__ unsigned int s_m;
    __shared__ double s_sum[16];
    const int col = blockIdx.x;
    const unsigned long long *__restrict__ K = keys + (int64_t)col * n_rows;
    unsigned long long *__restrict__ cand = cand_all + (int64_t)col * n_rows;

    // pass 1: valid count + deterministic sum + histogram of the top digit
    for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
    if (threadIdx.x == 0) { s_cnt = 0; s_m = 0; s_succ = ~0ull; }
    __syncthreads();
    unsigned long long cnt = 0;
    double sum = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const unsigned long long k = K[r];
        if (k != ~0ull) {
            cnt++;
            sum += pb_value_of(k);
            if (mode == 0) atomicAdd(&hist[k >> 56], 1u);
        }
    }
    cnt = pb_warp_sum(cnt);
    sum = pb_warp_sum_f64(sum);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt, cnt); s_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    const unsigned long long n_valid = s_cnt;
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_sum[w];
        n_regions[col] = (int64_t)n_valid;
        col_sum[col] = t;
        if (mode == 1) profile[col] = n_valid ? t / (double)n_valid : nan("");
        if (mode == 2) profile[col] = t;
    }
    if (mode != 0) return;
    if (n_valid == 0) { if (threadIdx.x == 0) profile[col] = nan(""); return; }

    // lower middle element: 0-based rank (n-1)/2
    const unsigned long long rank0 = (n_valid - 1) / 2;
    if (threadIdx.x == 0) pb_select_digit(hist, rank0, 56, 0ull, &s_prefix, &s_rank);
    __syncthreads();
    unsigned long long prefix = s_prefix, rank = s_rank;
    __syncthreads();
    // pass 2: second digit among the keys that share the top digit
    for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
    __syncthreads();
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const unsigned long long k = K[r];
        if (k != ~0ull && (k >> 56) == (prefix >> 56)) atomicAdd(&hist[(k >> 48) & 0xff], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) pb_select_digit(hist, rank, 48, prefix, &s_prefix, &s_rank);
    __syncthreads();
    prefix = s_prefix; rank = s_rank;
    __syncthreads();
    // keys below the selected 16-bit bucket: what the residual rank no longer counts
    const unsigned long long below = rank0 - rank;
    // pass 3: third digit + compaction of the bucket + smallest key above the bucket
    for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
    __syncthreads();
    unsigned long long succ_out = ~0ull;
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const unsigned long long k = K[r];
        if (k == ~0ull) continue;
        if ((k >> 48) == (prefix >> 48)) {
            atomicAdd(&hist[(k >> 40) & 0xff], 1u);
            cand[atomicAdd(&s_m, 1u)] = k;
        } else if ((k >> 48) > (prefix >> 48) && k < succ_out) succ_out = k;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, succ_out, d);
        succ_out = o < succ_out ? o : succ_out;
    }
    if ((threadIdx.x & 31) == 0 && succ_out != ~0ull) atomicMin(&s_succ, succ_out);
    __syncthreads();       // also publishes cand[] (written and read by this CTA only)
    const unsigned int m = s_m;
    if (threadIdx.x == 0) pb_select_digit(hist, rank, 40, prefix, &s_prefix, &s_rank);
    __syncthreads();
    prefix = s_prefix; rank = s_rank;
    __syncthreads();
    // remaining five digits over the candidates only
    for (int shift = 32; shift >= 0; shift -= 8) {
        for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
        __syncthreads();
        const unsigned long long himask = ~0ull << (shift + 8);
        for (unsigned int r = threadIdx.x; r < m; r += blockDim.x) {
            const unsigned long long k = cand[r];
            if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 0xff], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) pb_select_digit(hist, rank, shift, prefix, &s_prefix, &s_rank);
        __syncthreads();
        prefix = s_prefix; rank = s_rank;
        __syncthreads();
    }
    const unsigned long long key0 = prefix;
    // upper middle element (rank n/2): the same key when n is odd or key0 repeats far enough, else the
    // smallest key above it (among the candidates, or the smallest key above the bucket)
    unsigned long long key1 = key0;
    if ((n_valid & 1ull) == 0) {
        unsigned long long le = 0, succ = ~0ull;
        for (unsigned int r = threadIdx.x; r < m; r += blockDim.x) {
            const unsigned long long k = cand[r];
            if (k <= key0) le++; else if (k < succ) succ = k;
        }
        le = pb_warp_sum(le);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, succ, d);
            succ = o < succ ? o : succ;
        }
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt, le); if (succ != ~0ull) atomicMin(&s_succ, succ); }
        __syncthreads();
        if (below + s_cnt <= n_valid / 2) key1 = s_succ;      // fewer than n/2 + 1 keys are <= key0
    }
    if (threadIdx.x == 0) profile[col] = (pb_value_of(key0) + pb_value_of(key1)) / 2.0;
}

// ----------------------------------------------------------------------------------------
// mask pipeline: per-chain mask bits from a sorted interval set, all chains in one launch
// ----------------------------------------------------------------------------------------
// Replaces, for a whole region list at once, GenomeHash.get_overlapping_features + SegmentChain.add_masks
// (genome_hash.py:259-436, roitools.pyx:2213-2301): the masked positions of a chain are (union of the
// mask features on its chromosome and strand) ∩ (chain positions).  Mask intervals arrive merged and
// sorted per strand class in global-bin coordinates, so interval ends are sorted too: one binary
// search per exon block finds the first interval that can overlap it.  One warp per chain, lanes over
// blocks; bits are OR-ed in (a chain's bit range is not word aligned, neighbours share words).
__global__ void pb_mask_chains_kernel(const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                                      const int64_t *__restrict__ chain_off, const uint8_t *__restrict__ chain_plane,
                                      int64_t n_chains, const int64_t *__restrict__ mstart, const int64_t *__restrict__ mend,
                                      const int64_t *__restrict__ mclass_off, const int64_t *__restrict__ mask_off,
                                      unsigned int *__restrict__ mask_words)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_chains) return;
    const int cls = chain_plane[c];
    const int64_t m0 = __ldg(mclass_off + cls), m1 = __ldg(mclass_off + cls + 1);
    if (m0 == m1) return;
    const int64_t b0 = __ldg(chain_off + c), b1 = __ldg(chain_off + c + 1);
    int64_t chain_pos = __ldg(mask_off + c);       // bit index of the chain position the batch starts at
    for (int64_t j0 = b0; j0 < b1; j0 += 32) {
        const int64_t j = j0 + lane;
        const int64_t bs = j < b1 ? __ldg(bstart + j) : 0, be = j < b1 ? __ldg(bend + j) : 0;
        long long incl = be - bs;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int64_t bit0 = chain_pos + (incl - (be - bs));
        chain_pos += __shfl_sync(0xffffffffu, incl, 31);
        if (be <= bs) continue;
        int64_t lo = m0, hi = m1;                   // first interval with mend > bs
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(mend + mid) <= bs) lo = mid + 1; else hi = mid;
        }
        for (int64_t k = lo; k < m1; ++k) {
            const int64_t ms = __ldg(mstart + k);
            if (ms >= be) break;
            const int64_t me = __ldg(mend + k);
            const int64_t a = bit0 + ((ms > bs ? ms : bs) - bs), b = bit0 + ((me < be ? me : be) - bs);   // bits [a, b)
            const int64_t w0 = a >> 5, w1 = (b - 1) >> 5;
            for (int64_t w = w0; w <= w1; ++w) {
                unsigned int m = 0xffffffffu;
                if (w == w0) m &= 0xffffffffu << (a & 31);
                if (w == w1) m &= 0xffffffffu >> (31 - ((b - 1) & 31));
                atomicOr(mask_words + w, m);
            }
        }
    }
}

}  // namespace

static int check_chains(const void *const *planes, const int64_t *bstart, const int64_t *bend,
                        const int64_t *chain_off, const uint8_t *chain_plane, int64_t n_chains,
                        const uint8_t *mask_bits, const int64_t *mask_off, int vec_dtype)
{
    if (!planes || !bstart || !bend || !chain_off || !chain_plane) { pb_set_error("gather: null chain tables"); return PB_EINVAL; }
    if (n_chains < 0) { pb_set_error("gather: negative chain count"); return PB_EINVAL; }
    if (mask_bits && !mask_off) { pb_set_error("gather: mask_bits without mask_off"); return PB_EINVAL; }
    if (vec_dtype != 0 && vec_dtype != 1) { pb_set_error("gather: vec_dtype must be 0 (uint32) or 1 (float64)"); return PB_EINVAL; }
    return PB_OK;
}

extern "C" int pb_window_normalize(const double *matrix, const uint8_t *maskmat, int64_t n_rows, int32_t width,
                                   int32_t norm_lo, int32_t norm_hi, double min_counts,
                                   double *denom, uint8_t *row_select, double *norm_out, uint8_t *normmask_out,
                                   void *stream_)
{
    if (!matrix || !maskmat || !denom || !row_select || n_rows < 0 || width <= 0) { pb_set_error("pb_window_normalize: bad arguments"); return PB_EINVAL; }
    if ((norm_out == nullptr) != (normmask_out == nullptr)) { pb_set_error("pb_window_normalize: norm_out and normmask_out go together"); return PB_EINVAL; }
    if (n_rows == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const unsigned grid = (unsigned)((n_rows * 32 + 255) / 256);
    pb_window_normalize_kernel<<<grid, 256, 0, stream>>>(matrix, maskmat, n_rows, width, norm_lo, norm_hi, min_counts, denom, row_select, norm_out, normmask_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" size_t pb_column_profile_workspace_bytes(int64_t n_rows, int32_t width)
{
    if (n_rows < 0 || width < 0) return 0;
    return 2 * (size_t)n_rows * (size_t)width * sizeof(unsigned long long) + 256;   // keys + select candidates
}

extern "C" int pb_column_profile_batched(const double *values, const uint8_t *valmask, const uint8_t *row_select,
                                         int32_t n_batch, int64_t n_rows, int32_t width, int mode,
                                         double *profile, int64_t *n_regions, double *col_sum,
                                         void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!values || !valmask || !row_select || !profile || !n_regions || !col_sum || n_rows < 0 || width <= 0 ||
        mode < 0 || mode > 2 || n_batch < 1 || n_batch > 65535) {
        pb_set_error("pb_column_profile: bad arguments"); return PB_EINVAL;
    }
    if (!workspace || workspace_bytes < (size_t)n_batch * pb_column_profile_workspace_bytes(n_rows, width)) {
        pb_set_error("pb_column_profile: workspace too small"); return PB_ENOSPACE;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long *keys = (unsigned long long *)workspace;
    if (n_rows > 0) {
        dim3 grid((unsigned)((width + 31) / 32), (unsigned)((n_rows + 31) / 32), (unsigned)n_batch);
        if (grid.y > 65535) { pb_set_error("pb_column_profile: more than 2,097,120 rows"); return PB_EINVAL; }
        pb_column_keys_kernel<<<grid, dim3(32, 8), 0, stream>>>(values, valmask, row_select, n_rows, width, keys);
    }
    // one CTA per column of every matrix: column b*width + c reads keys[(b*width + c) * n_rows ...]
    unsigned long long *cand = keys + (size_t)n_batch * (size_t)n_rows * (size_t)width;
    pb_column_stats_kernel<<<(unsigned)(n_batch * width), 512, 0, stream>>>(keys, cand, n_rows, width, mode, profile, n_regions, col_sum);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_count_profiles_u32(const uint32_t *counts, const uint8_t *maskmat, int mask_shared,
                                     int32_t n_batch, int64_t n_rows, int32_t width,
                                     int32_t norm_lo, int32_t norm_hi, double min_counts, int mode,
                                     uint8_t *row_select, double *profile, int64_t *n_regions, double *col_sum,
                                     void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!counts || !maskmat || !row_select || !profile || !n_regions || !col_sum || n_rows < 0 || width <= 0 ||
        (mode != 0 && mode != 1) || n_batch < 1 || n_batch > 65535) {
        pb_set_error("pb_count_profiles_u32: bad arguments"); return PB_EINVAL;
    }
    if (!workspace || workspace_bytes < (size_t)n_batch * pb_column_profile_workspace_bytes(n_rows, width)) {
        pb_set_error("pb_count_profiles_u32: workspace too small"); return PB_ENOSPACE;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long *keys = (unsigned long long *)workspace;
    unsigned long long *cand = keys + (size_t)n_batch * (size_t)n_rows * (size_t)width;
    if (n_rows > 0) {
        dim3 grid((unsigned)((n_rows + 31) / 32), (unsigned)n_batch);
        pb_norm_keys_u32_kernel<<<grid, 256, 0, stream>>>(counts, maskmat, mask_shared, n_rows, width, norm_lo, norm_hi,
                                                          min_counts, row_select, keys);
    }
    pb_column_stats_kernel<<<(unsigned)(n_batch * width), 512, 0, stream>>>(keys, cand, n_rows, width, mode, profile, n_regions, col_sum);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_column_profile(const double *values, const uint8_t *valmask, const uint8_t *row_select,
                                 int64_t n_rows, int32_t width, int mode,
                                 double *profile, int64_t *n_regions, double *col_sum,
                                 void *workspace, size_t workspace_bytes, void *stream_)
{
    return pb_column_profile_batched(values, valmask, row_select, 1, n_rows, width, mode, profile, n_regions, col_sum,
                                     workspace, workspace_bytes, stream_);
}

extern "C" int pb_phase_sums_range(const void *const *planes, int vec_dtype,
                                   const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                   const uint8_t *chain_plane, const uint8_t *chain_reverse, int64_t n_chains,
                                   int32_t codon_front, int32_t codon_back, int64_t bin_begin, int64_t bin_end,
                                   double *out, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, nullptr, nullptr, vec_dtype);
    if (rc) return rc;
    if (!chain_reverse || !out) { pb_set_error("pb_phase_sums: bad arguments"); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbPlanes pl{{planes[0], planes[1], planes[2]}};
    const unsigned grid = (unsigned)((n_chains * 32 + 255) / 256);
    if (vec_dtype == 0)
        pb_phase_sums_kernel<uint32_t><<<grid, 256, 0, stream>>>(pl, bstart, bend, chain_off, chain_plane, chain_reverse, n_chains, codon_front, codon_back, bin_begin, bin_end, out);
    else
        pb_phase_sums_kernel<double><<<grid, 256, 0, stream>>>(pl, bstart, bend, chain_off, chain_plane, chain_reverse, n_chains, codon_front, codon_back, bin_begin, bin_end, out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_phase_sums(const void *const *planes, int vec_dtype,
                             const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                             const uint8_t *chain_plane, const uint8_t *chain_reverse, int64_t n_chains,
                             int32_t codon_front, int32_t codon_back, double *out, void *stream_)
{
    return pb_phase_sums_range(planes, vec_dtype, bstart, bend, chain_off, chain_plane, chain_reverse, n_chains,
                               codon_front, codon_back, 0, INT64_MAX, out, stream_);
}

// ----------------------------------------------------------------------------------------
// psite / phase_by_size inner loops in one launch: per-read-length window matrices
// ----------------------------------------------------------------------------------------
// Reference (plastid/bin/psite.py:176-199, plastid/bin/phase_by_size.py:186-194): per window and
// exon, fetch the reads the mapping rule keeps there, bucket them by aligned length, map every
// bucket, lay the vectors 5'->3'.  Here one CTA owns one window: the reads that can map into each
// of its blocks are a contiguous slice of the coordinate-sorted batch (two binary searches), the rule
// is applied per read and the (length, column) cell is incremented in a shared-memory histogram.
namespace {

// one warp per exon block: the slice [lo, hi) of the coordinate-sorted batch whose reads can map into it
__global__ void __launch_bounds__(256)
pb_stratified_slices_kernel(PbReads b, PbLayoutDev lay, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                            int64_t n_blocks, long long *__restrict__ slices)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_blocks) return;
    const int64_t gs = __ldg(bstart + k), ge = __ldg(bend + k);
    const int ch = pb_chrom_of_bin(lay, gs);
    const int64_t base = __ldg(lay.chrom_bin_off + ch);
    int64_t r0 = 0, r1 = 0;
    if (ch < b.n_chrom) { r0 = __ldg(b.chrom_read_off + ch); r1 = __ldg(b.chrom_read_off + ch + 1); }
    const int64_t lo = pb_lower_bound_warp(b.ref_start, r0, r1, gs - base - b.max_span + 1);
    const int64_t hi = pb_lower_bound_warp(b.ref_start, lo, r1, ge - base);
    if ((threadIdx.x & 31) == 0) { slices[2 * k] = lo; slices[2 * k + 1] = hi; }
}

__global__ void __launch_bounds__(128, 10)
pb_stratified_windows_kernel(PbReads b, PbRuleDev r, PbLayoutDev lay, int min_len, int n_len,
                             const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                             const int64_t *__restrict__ chain_off, const uint8_t *__restrict__ chain_plane,
                             const uint8_t *__restrict__ chain_reverse, const int32_t *__restrict__ row_col,
                             int64_t n_chains, int32_t width, int phase_mode, int32_t codon_front, int32_t codon_back,
                             const uint8_t *__restrict__ mask_bits, const int64_t *__restrict__ mask_off,
                             long long lo_bin, long long hi_bin,
                             uint32_t *__restrict__ out, uint8_t *__restrict__ maskmat,
                             int site_tab, const long long *__restrict__ slices)
{
    extern __shared__ __align__(16) uint32_t hist[];   // [n_len][width]  (phase mode: width == 3 sub-codon phases)
    __shared__ __align__(16) int16_t s_tab[kSiteKeys];   // point rules: the site table of this window's strand class
    const int64_t c = blockIdx.x;
    {
        const int cells = n_len * width, quads = cells >> 2;
        for (int j = threadIdx.x; j < quads; j += blockDim.x) reinterpret_cast<uint4 *>(hist)[j] = make_uint4(0u, 0u, 0u, 0u);
        for (int j = (quads << 2) + threadIdx.x; j < cells; j += blockDim.x) hist[j] = 0;
    }
    const int plane = __ldg(chain_plane + c);            // 0 '+', 1 '-', 2 '.'
    if (site_tab) {
        // only the lengths of the stratification range count (psite.py:187): everything else is "skip", and the
        // 2 x n_len live entries (forward / reverse reads) cost one rule look-up each
        const uint4 skip4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);     // kSiteSkip in every entry
        for (int j = threadIdx.x; j < kSiteKeys * 2 / 16; j += blockDim.x) reinterpret_cast<uint4 *>(s_tab)[j] = skip4;
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * n_len; j += blockDim.x) {
            const int L = min_len + (j >> 1);
            const uint32_t rev = (uint32_t)(j & 1);
            if (L < 256) {
                const int v = pb_site_entry(r, plane, (uint32_t)L | (rev << 16));
                s_tab[L | (rev << 8)] = (int16_t)(v < 0 ? kSiteSkip : v);
            }
        }
    }
    __syncthreads();
    const int64_t k0 = __ldg(chain_off + c), k1 = __ldg(chain_off + c + 1);
    int64_t len = 0;
    for (int64_t k = k0; k < k1; ++k) len += __ldg(bend + k) - __ldg(bstart + k);
    const bool rev_out = __ldg(chain_reverse + c);
    const bool rq = plane == 1;                          // rule direction follows the window's strand
    const int64_t col0 = phase_mode ? 0 : __ldg(row_col + c);
    // phase mode (phase_by_size.py:197-214): codons [codon_front:codon_back] (python slice) of the chain
    const int64_t ncod = len / 3;
    int64_t cod_lo = codon_front < 0 ? ncod + codon_front : codon_front, cod_hi = codon_back < 0 ? ncod + codon_back : codon_back;
    cod_lo = cod_lo < 0 ? 0 : (cod_lo > ncod ? ncod : cod_lo);
    cod_hi = cod_hi < 0 ? 0 : (cod_hi > ncod ? ncod : cod_hi);
    int64_t j0 = 0;
    for (int64_t k = k0; k < k1; ++k) {
        const int64_t gs = __ldg(bstart + k), ge = __ldg(bend + k);
        const int ch = pb_chrom_of_bin(lay, gs);
        const int64_t base = __ldg(lay.chrom_bin_off + ch);
        const int64_t bs = gs - base, be = ge - base;    // chromosome coordinates of the block
        int64_t r0 = 0, r1 = 0;
        if (ch < b.n_chrom) { r0 = __ldg(b.chrom_read_off + ch); r1 = __ldg(b.chrom_read_off + ch + 1); }
        // the block's read slice: looked up when pb_stratified_slices_kernel ran first (all blocks of the table searched
        // side by side; ncu: a third of this kernel's stall samples sat in the two dependent searches per exon), else
        // every warp runs the same two 32-ary searches (same addresses: L1 hits after the first warp)
        int64_t lo, hi;
        if (slices) { lo = __ldg(slices + 2 * k); hi = __ldg(slices + 2 * k + 1); }
        else {
            lo = pb_lower_bound_warp(b.ref_start, r0, r1, bs - b.max_span + 1);
            hi = pb_lower_bound_warp(b.ref_start, lo, r1, be);
        }
        constexpr int kU = 4;       // independent reads in flight per thread
        constexpr int kT = 128;     // threads per CTA (the launch below); 32-bit indices relative to the slice: ncu showed
                                    // ~105 instructions of 64-bit index arithmetic per 4 loads with `i0 + u * blockDim.x`
        const int n_slice = (int)(hi - lo < 0x7fffffff ? hi - lo : 0x7fffffff);
        const uint32_t *__restrict__ mp = b.meta + lo;
        const int32_t *__restrict__ sp = b.ref_start + lo;
        for (int t0 = threadIdx.x; t0 < n_slice; t0 += kU * kT) {
            const int64_t i0 = lo + t0;
            uint32_t mv[kU];
            int32_t sv[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int t = t0 + u * kT;
                mv[u] = t < n_slice ? __ldg(mp + t) : (1u << 17);
                sv[u] = t < n_slice ? __ldg(sp + t) : 0;
            }
            if (site_tab) {
                // point rules: drop bit, size window, strand class, stratification range and rule offset are ONE look-up
                // (built per CTA above); the per-read path is straight-line up to the histogram atomic.
                // ncu on the generic path below: 190 thread instructions per read, issue slots 71 % busy.
                const int own_lo = (int)max((long long)bs, lo_bin - base), own_w = (int)min((long long)be, hi_bin - base) - own_lo;
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t m = mv[u];
                    int idx;
                    if (__builtin_expect((m & 0xFF00u) != 0u, 0)) {
                        const int L = PB_META_L(m);
                        idx = (L < min_len || L >= min_len + n_len) ? kSiteSkip : pb_site_entry(r, plane, m);
                    } else idx = s_tab[(m & 0xFFu) | ((m >> 8) & 0x300u)];
                    int p = sv[u] + idx;
                    if (idx >= 0 && PB_META_NBLK(m) > 1 && b.blk_off != nullptr) {
                        const int64_t pp = pb_block_position(b, i0 + u * kT, sv[u], idx);
                        p = pp < 0 ? own_lo - 1 : (int)pp;
                    }
                    if (idx >= 0 && own_w > 0 && (unsigned)(p - own_lo) < (unsigned)own_w) {      // the site belongs to this block and rank
                        const int jj = (int)j0 + (p - (int)bs);
                        int col = (int)col0 + (rev_out ? ((int)len - 1 - jj) : jj);
                        if (phase_mode) {
                            const int cod = col / 3;
                            col = (cod >= cod_lo && cod < cod_hi) ? col - cod * 3 : -1;
                        }
                        if ((unsigned)col < (unsigned)width) atomicAdd(&hist[((int)(m & 0xFFFFu) - min_len) * width + col], 1u);
                    }
                }
                continue;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int64_t i = i0 + u * kT;
                const uint32_t m = mv[u];
                if (!pb_passes(m, r.size_min, r.size_max)) continue;
                const bool rev = PB_META_REV(m);
                if ((plane == 0 && rev) || (plane == 1 && !rev)) continue;    // genome_array.py:811-815
                const int L = PB_META_L(m);
                if (L < min_len || L >= min_len + n_len) continue;           // psite.py:187: len(positions) in read_dict
                // phase mode: the reference does not reset its per-length read lists between the exons of a coding
                // region (phase_by_size.py:186-194); the lists hold what `get_reads` returned, i.e. the rule's
                // `reads_out`.  CenterMapFactory returns every read it was given (map_factories.pyx:256), so a read
                // that the fetch of an EARLIER exon of this chain returned too — its span reaches back over that exon's
                // end — is mapped once more against this exon per such exon.  Point rules return only the reads whose
                // site lies in the exon (:351-353): a read kept for an earlier exon has no site here, multiplicity 1.
                uint32_t mult = 1;
                if (phase_mode && r.kind == PB_RULE_CENTER)
                    for (int64_t j = k - 1; j >= k0 && __ldg(bend + j) - base > (int64_t)sv[u]; --j) ++mult;
                auto add_site = [&](int64_t p) {
                    if (base + p < lo_bin || base + p >= hi_bin) return;         // the site belongs to another rank
                    const int64_t jj = j0 + (p - bs);
                    int64_t col = col0 + (rev_out ? (len - 1 - jj) : jj);
                    if (phase_mode) {
                        const int64_t cod = col / 3;
                        col = (cod >= cod_lo && cod < cod_hi) ? col - cod * 3 : -1;
                    }
                    if (col >= 0 && col < width) atomicAdd(&hist[(L - min_len) * width + (int)col], mult);
                };
                if (r.kind == PB_RULE_CENTER) {
                    // phase mode only: every trimmed aligned position of the read inside this exon counts `mult`
                    // times 1/(L - 2 nibble); the cell holds the integer, the host applies the weight of its length
                    const int nib = r.param;
                    if (L - 2 * nib <= 0) continue;                              // map_factories.pyx:246-248
                    auto add_range = [&](int64_t x, int64_t y) {
                        if (x < bs) x = bs;
                        if (y > be) y = be;
                        for (int64_t p = x; p < y; ++p) add_site(p);
                    };
                    if (PB_META_NBLK(m) <= 1 || b.blk_off == nullptr) {
                        add_range((int64_t)sv[u] + nib, (int64_t)sv[u] + L - nib);
                    } else {
                        int a = 0;
                        for (uint32_t q = __ldg(b.blk_off + i), q1 = __ldg(b.blk_off + i + 1); q < q1; ++q) {
                            const int2 bl = __ldg(b.blk + q);
                            const int ia = a > nib ? a : nib, ib = (a + bl.y) < (L - nib) ? (a + bl.y) : (L - nib);
                            if (ia < ib) add_range((int64_t)sv[u] + bl.x + (ia - a), (int64_t)sv[u] + bl.x + (ib - a));
                            a += bl.y;
                        }
                    }
                    continue;
                }
                const int idx = pb_rule_index(r, L, rq);
                if (idx < 0) continue;
                const int64_t p = pb_position(b, i, sv[u], m, idx);
                if (p < bs || p >= be) continue;
                add_site(p);
            }
        }
        j0 += be - bs;
    }
    __syncthreads();
    // row by row (ncu: the flat loop's `j / width` made this write-out 37 instructions per cell, a third of the kernel)
    for (int row = 0; row < n_len; ++row) {
        uint32_t *__restrict__ dst = out + ((int64_t)row * n_chains + c) * width;
        const uint32_t *src = hist + row * width;
        for (int col = threadIdx.x; col < width; col += blockDim.x) dst[col] = src[col];
    }
    // the validity mask every length shares (get_masked_counts(ga).mask, psite.py:166)
    if (phase_mode || !maskmat) return;
    const int64_t moff = mask_bits ? __ldg(mask_off + c) : 0;
    for (int col = threadIdx.x; col < width; col += blockDim.x) {
        uint8_t mk = 1;
        const int64_t t = col - col0;                    // 5'->3' index along the chain
        if (t >= 0 && t < len) {
            const int64_t jj = rev_out ? (len - 1 - t) : t;
            mk = mask_bits ? (uint8_t)pb_mask_bit(mask_bits, moff + jj) : (uint8_t)0;
        }
        maskmat[c * (int64_t)width + col] = mk;
    }
}

}  // namespace

extern "C" size_t pb_stratified_windows_workspace_bytes(int64_t n_blocks)
{
    return n_blocks < 0 ? 0 : (size_t)n_blocks * 16 + 16;
}

static int stratified_windows_impl(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                                   int min_len, int max_len,
                                   const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                   const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                   const int32_t *row_col, int64_t n_chains, int64_t n_blocks, int32_t width,
                                   int phase_mode, int32_t codon_front, int32_t codon_back,
                                   const uint8_t *mask_bits, const int64_t *mask_off,
                                   int64_t bin_begin, int64_t bin_end,
                                   uint32_t *out, uint8_t *maskmat, void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!batch || !layout || !rule || !bstart || !bend || !chain_off || !chain_plane || !chain_reverse || !out ||
        (!phase_mode && (!row_col || !maskmat))) { pb_set_error("pb_stratified_windows: null argument"); return PB_EINVAL; }
    if (phase_mode) width = 3;
    if (rule->kind != PB_RULE_FIVEPRIME && rule->kind != PB_RULE_THREEPRIME && rule->kind != PB_RULE_VARIABLE &&
        !(rule->kind == PB_RULE_CENTER && phase_mode && rule->param >= 0)) {
        pb_set_error("pb_stratified_windows: needs a point rule (phase mode: or the center rule)"); return PB_EINVAL;
    }
    if (rule->kind == PB_RULE_VARIABLE && (!rule->lut_fw || !rule->lut_rc)) { pb_set_error("pb_stratified_windows: variable rule needs LUTs"); return PB_EINVAL; }
    if (max_len < min_len || min_len < 0 || width <= 0 || n_chains < 0) { pb_set_error("pb_stratified_windows: bad sizes"); return PB_EINVAL; }
    if (mask_bits && !mask_off) { pb_set_error("pb_stratified_windows: mask_bits without mask_off"); return PB_EINVAL; }
    const int n_len = max_len - min_len + 1;
    const size_t smem = (size_t)n_len * width * sizeof(uint32_t);
    if (smem > 200 * 1024) { pb_set_error("pb_stratified_windows: %d lengths x %d columns do not fit in shared memory", n_len, width); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    PB_CUDA_CHECK(cudaFuncSetAttribute(pb_stratified_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // point rules take the site-table path (the Center form of phase mode walks trimmed intervals); PB_STRAT_GENERIC=1 (A/B aid)
    const int site_tab = rule->kind != PB_RULE_CENTER && !getenv("PB_STRAT_GENERIC");
    long long *slices = nullptr;
    if (workspace && n_blocks > 0) {
        if (workspace_bytes < pb_stratified_windows_workspace_bytes(n_blocks) || ((uintptr_t)workspace & 7)) {
            pb_set_error("pb_stratified_windows_ws: workspace too small or misaligned"); return PB_ENOSPACE;
        }
        slices = (long long *)workspace;
        pb_stratified_slices_kernel<<<(unsigned)((n_blocks * 32 + 255) / 256), 256, 0, stream>>>(b, lay, bstart, bend, n_blocks, slices);
    }
    pb_stratified_windows_kernel<<<(unsigned)n_chains, 128, smem, stream>>>(b, r, lay, min_len, n_len, bstart, bend, chain_off,
                                                                           chain_plane, chain_reverse, row_col, n_chains, width,
                                                                           phase_mode, codon_front, codon_back,
                                                                           mask_bits, mask_off, bin_begin, bin_end, out, maskmat,
                                                                           site_tab, slices);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_stratified_windows_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                                           int min_len, int max_len,
                                           const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                           const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                           const int32_t *row_col, int64_t n_chains, int32_t width,
                                           int phase_mode, int32_t codon_front, int32_t codon_back,
                                           const uint8_t *mask_bits, const int64_t *mask_off,
                                           int64_t bin_begin, int64_t bin_end,
                                           uint32_t *out, uint8_t *maskmat, void *stream_)
{
    return stratified_windows_impl(batch, layout, rule, min_len, max_len, bstart, bend, chain_off, chain_plane, chain_reverse,
                                   row_col, n_chains, 0, width, phase_mode, codon_front, codon_back, mask_bits, mask_off,
                                   bin_begin, bin_end, out, maskmat, nullptr, 0, stream_);
}

extern "C" int pb_stratified_windows_ws(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                                        int min_len, int max_len,
                                        const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                        const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                        const int32_t *row_col, int64_t n_chains, int64_t n_blocks, int32_t width,
                                        int phase_mode, int32_t codon_front, int32_t codon_back,
                                        const uint8_t *mask_bits, const int64_t *mask_off,
                                        int64_t bin_begin, int64_t bin_end,
                                        uint32_t *out, uint8_t *maskmat, void *workspace, size_t workspace_bytes, void *stream_)
{
    return stratified_windows_impl(batch, layout, rule, min_len, max_len, bstart, bend, chain_off, chain_plane, chain_reverse,
                                   row_col, n_chains, n_blocks, width, phase_mode, codon_front, codon_back, mask_bits, mask_off,
                                   bin_begin, bin_end, out, maskmat, workspace, workspace_bytes, stream_);
}

extern "C" int pb_stratified_windows(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                                     int min_len, int max_len,
                                     const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                     const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                     const int32_t *row_col, int64_t n_chains, int32_t width,
                                     int phase_mode, int32_t codon_front, int32_t codon_back,
                                     const uint8_t *mask_bits, const int64_t *mask_off,
                                     uint32_t *out, uint8_t *maskmat, void *stream_)
{
    return pb_stratified_windows_range(batch, layout, rule, min_len, max_len, bstart, bend, chain_off, chain_plane,
                                       chain_reverse, row_col, n_chains, width, phase_mode, codon_front, codon_back,
                                       mask_bits, mask_off, 0, INT64_MAX, out, maskmat, stream_);
}

extern "C" int pb_mask_chains(const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                              const uint8_t *chain_plane, int64_t n_chains,
                              const int64_t *mask_start, const int64_t *mask_end, const int64_t *mask_class_off,
                              const int64_t *mask_off, uint8_t *mask_bits, void *stream_)
{
    if (!bstart || !bend || !chain_off || !chain_plane || !mask_start || !mask_end || !mask_class_off || !mask_off ||
        !mask_bits || n_chains < 0) {
        pb_set_error("pb_mask_chains: null argument"); return PB_EINVAL;
    }
    if ((uintptr_t)mask_bits & 3) { pb_set_error("pb_mask_chains: mask_bits must be 4-byte aligned"); return PB_EINVAL; }
    if (n_chains == 0) return PB_OK;
    const int64_t grid = (n_chains * 32 + 255) / 256;
    pb_mask_chains_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream_>>>(
        bstart, bend, chain_off, chain_plane, n_chains, mask_start, mask_end, mask_class_off, mask_off,
        reinterpret_cast<unsigned int *>(mask_bits));
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
