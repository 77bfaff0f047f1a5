// pb_chains.cu — sorted-interval set algebra over many chains at once: the position-set arithmetic of
// `cs generate`.
//
// Reference (plastid v0.6.1, plastid/bin/cs.py:242-496 process_partial_group): per merged gene, python
// `set`s of genomic positions are pooled (`|=`), intersected (`&`) and subtracted (`-=`) and turned back into
// segments with positions_to_segments — O(gene length) set elements per operation.  A chain here is a sorted
// list of disjoint, non-touching blocks [start, end); unions, intersections and differences of chains are
// merges of their block lists, done for every gene / transcript of the annotation in one launch per
// operation (count pass, exclusive scan, fill pass).
#include "pb_common.cuh"

namespace {

struct Chains {
    const int64_t *__restrict__ bstart;
    const int64_t *__restrict__ bend;
    const int64_t *__restrict__ off;      // chain c owns blocks [off[c], off[c+1])
};

// first block k in [lo, hi) with bend[k] > x  (blocks sorted, disjoint)
__device__ __forceinline__ int64_t first_end_above(const Chains &c, int64_t lo, int64_t hi, int64_t x)
{
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(c.bend + mid) > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ int64_t warp_min(int64_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const int64_t o = __shfl_xor_sync(0xffffffffu, v, d); v = o < v ? o : v; }
    return v;
}

__device__ __forceinline__ int64_t warp_max(int64_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const int64_t o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
    return v;
}

#define PB_POS_INF ((int64_t)0x7fffffffffffffffLL)

// Union of the member chains of every group, one warp per group, lanes over members.  The sweep keeps a
// position x: the next block of the union starts at the smallest covered position >= x and grows while some
// member block starts at or before its end (touching blocks merge, like positions_to_segments).
template <bool FILL>
__global__ void __launch_bounds__(128)
pb_chain_union_kernel(Chains in, const int64_t *__restrict__ grp_off, const int64_t *__restrict__ members,
                      int64_t n_grp, int32_t *__restrict__ n_blk, const int64_t *__restrict__ out_off,
                      int64_t *__restrict__ out_bstart, int64_t *__restrict__ out_bend)
{
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); g < n_grp; g += n_warps) {
        const int64_t m0 = __ldg(grp_off + g), m1 = __ldg(grp_off + g + 1);
        const int64_t obase = FILL ? __ldg(out_off + g) : 0;
        const int32_t cap = FILL ? __ldg(n_blk + g) : 0;
        int32_t n = 0;
        int64_t x = -PB_POS_INF;
        for (;;) {
            // earliest covered position >= x
            int64_t s = PB_POS_INF;
            for (int64_t i = m0 + lane; i < m1; i += 32) {
                const int64_t c = __ldg(members + i);
                const int64_t lo = __ldg(in.off + c), hi = __ldg(in.off + c + 1);
                const int64_t k = first_end_above(in, lo, hi, x);
                if (k < hi) { const int64_t b = __ldg(in.bstart + k); const int64_t v = b > x ? b : x; s = v < s ? v : s; }
            }
            s = warp_min(s);
            if (s == PB_POS_INF) break;
            // grow: any member block with start <= e and end > e extends the run
            int64_t e = s;
            for (;;) {
                int64_t best = e;
                for (int64_t i = m0 + lane; i < m1; i += 32) {
                    const int64_t c = __ldg(members + i);
                    const int64_t lo = __ldg(in.off + c), hi = __ldg(in.off + c + 1);
                    const int64_t k = first_end_above(in, lo, hi, e);
                    if (k < hi && __ldg(in.bstart + k) <= e) { const int64_t v = __ldg(in.bend + k); best = v > best ? v : best; }
                }
                best = warp_max(best);
                if (best == e) break;
                e = best;
            }
            if (FILL && lane == 0 && n < cap) { out_bstart[obase + n] = s; out_bend[obase + n] = e; }
            ++n;
            x = e;
        }
        if (!FILL && lane == 0) n_blk[g] = n;
    }
}

// out chain i = A[a_idx[i]] AND / SUB B[b_idx[i]] (b_idx < 0: empty B); one thread per output chain,
// two-pointer merge after a binary-search skip to the first B block that can matter.
template <bool FILL>
__global__ void pb_chain_binary_kernel(Chains A, const int64_t *__restrict__ a_idx, Chains B,
                                       const int64_t *__restrict__ b_idx, int64_t n_out, int op,
                                       int32_t *__restrict__ n_blk, const int64_t *__restrict__ out_off,
                                       int64_t *__restrict__ out_bstart, int64_t *__restrict__ out_bend)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int64_t a = __ldg(a_idx + i), b = __ldg(b_idx + i);
    const int64_t a0 = __ldg(A.off + a), a1 = __ldg(A.off + a + 1);
    int64_t b0 = 0, b1 = 0;
    if (b >= 0) { b0 = __ldg(B.off + b); b1 = __ldg(B.off + b + 1); }
    const int64_t obase = FILL ? __ldg(out_off + i) : 0;
    const int32_t cap = FILL ? __ldg(n_blk + i) : 0;
    int32_t n = 0;
    int64_t kb = (a1 > a0 && b1 > b0) ? first_end_above(B, b0, b1, __ldg(A.bstart + a0)) : b1;
    for (int64_t ka = a0; ka < a1; ++ka) {
        const int64_t as = __ldg(A.bstart + ka), ae = __ldg(A.bend + ka);
        while (kb < b1 && __ldg(B.bend + kb) <= as) ++kb;
        int64_t cur = as;                                     // SUB: first position of A's block not yet emitted
        for (int64_t k = kb; k < b1; ++k) {
            const int64_t bs = __ldg(B.bstart + k), be = __ldg(B.bend + k);
            if (bs >= ae) break;
            if (op == PB_CHAIN_AND) {
                const int64_t lo = bs > as ? bs : as, hi = be < ae ? be : ae;
                if (lo < hi) {
                    if (FILL && n < cap) { out_bstart[obase + n] = lo; out_bend[obase + n] = hi; }
                    ++n;
                }
            } else {
                if (bs > cur) {
                    if (FILL && n < cap) { out_bstart[obase + n] = cur; out_bend[obase + n] = bs; }
                    ++n;
                }
                if (be > cur) cur = be;
            }
        }
        if (op == PB_CHAIN_SUB && cur < ae) {
            if (FILL && n < cap) { out_bstart[obase + n] = cur; out_bend[obase + n] = ae; }
            ++n;
        }
    }
    if (!FILL) n_blk[i] = n;
}

}  // namespace

extern "C" int pb_chain_union(const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                              const int64_t *grp_off, const int64_t *members, int64_t n_grp,
                              int32_t *n_blk, const int64_t *out_off, int64_t *out_bstart, int64_t *out_bend,
                              void *stream)
{
    if (n_grp < 0) { pb_set_error("pb_chain_union: negative size"); return PB_EINVAL; }
    if (n_grp == 0) return PB_OK;
    if (!bstart || !bend || !chain_off || !grp_off || !members || !n_blk) { pb_set_error("pb_chain_union: NULL argument"); return PB_EINVAL; }
    const bool fill = out_off != nullptr;
    if (fill && (!out_bstart || !out_bend)) { pb_set_error("pb_chain_union: out_off without block buffers"); return PB_EINVAL; }
    Chains in{bstart, bend, chain_off};
    int dev = 0, sms = 148;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    PB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 128;
    int64_t blocks = (n_grp + 3) / 4;
    const int64_t cap = (int64_t)sms * 16;
    if (blocks > cap) blocks = cap;
    if (fill)
        pb_chain_union_kernel<true><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(in, grp_off, members, n_grp, n_blk,
                                                                                            out_off, out_bstart, out_bend);
    else
        pb_chain_union_kernel<false><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(in, grp_off, members, n_grp, n_blk,
                                                                                             out_off, out_bstart, out_bend);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_chain_binary(int op,
                               const int64_t *a_bstart, const int64_t *a_bend, const int64_t *a_off, const int64_t *a_idx,
                               const int64_t *b_bstart, const int64_t *b_bend, const int64_t *b_off, const int64_t *b_idx,
                               int64_t n_out, int32_t *n_blk, const int64_t *out_off, int64_t *out_bstart,
                               int64_t *out_bend, void *stream)
{
    if (n_out < 0 || (op != PB_CHAIN_AND && op != PB_CHAIN_SUB)) { pb_set_error("pb_chain_binary: bad op or size"); return PB_EINVAL; }
    if (n_out == 0) return PB_OK;
    if (!a_bstart || !a_bend || !a_off || !a_idx || !b_bstart || !b_bend || !b_off || !b_idx || !n_blk) {
        pb_set_error("pb_chain_binary: NULL argument");
        return PB_EINVAL;
    }
    const bool fill = out_off != nullptr;
    if (fill && (!out_bstart || !out_bend)) { pb_set_error("pb_chain_binary: out_off without block buffers"); return PB_EINVAL; }
    Chains A{a_bstart, a_bend, a_off}, B{b_bstart, b_bend, b_off};
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_out + threads - 1) / threads);
    if (fill)
        pb_chain_binary_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(A, a_idx, B, b_idx, n_out, op, n_blk, out_off,
                                                                                   out_bstart, out_bend);
    else
        pb_chain_binary_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(A, a_idx, B, b_idx, n_out, op, n_blk, out_off,
                                                                                    out_bstart, out_bend);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
