// pb_common.cuh — shared device helpers for libplastid_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "plastid_b200.h"

#define PB_META_L(m)    ((int)((m) & 0xFFFFu))
#define PB_META_REV(m)  ((int)(((m) >> 16) & 1u))
#define PB_META_DROP(m) ((int)(((m) >> 17) & 1u))
#define PB_META_NBLK(m) ((int)((m) >> 24))

void pb_set_error(const char *fmt, ...);

#define PB_CUDA_CHECK(expr)                                                          \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            pb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                         __FILE__, __LINE__);                                        \
            return PB_ECUDA;                                                         \
        }                                                                            \
    } while (0)

// Device-side view of a batch + rule, passed by value as a kernel parameter.
struct PbReads {
    const int32_t  *__restrict__ ref_start;
    const uint32_t *__restrict__ meta;
    const uint32_t *__restrict__ blk_off;
    const int2     *__restrict__ blk;
    const int64_t  *__restrict__ chrom_read_off;
    int64_t n_reads;
    int64_t n_blk;          // entries of blk (0 when every read is one block)
    int32_t n_chrom;
    int32_t max_span;
    int32_t max_block_len;  // longest aligned block in the batch: halo of the tile candidate window
};

struct PbRuleDev {
    int32_t kind, param;
    const int32_t *__restrict__ lut_fw;
    const int32_t *__restrict__ lut_rc;
    int32_t size_min, size_max, strat_min, strat_max;
};

struct PbLayoutDev {
    const int64_t *__restrict__ chrom_len;
    const int64_t *__restrict__ chrom_bin_off;
    int32_t n_chrom;
};

// SizeFilterFactory (map_factories.pyx:837-839) + host keep-mask.
__device__ __forceinline__ bool pb_passes(uint32_t m, int size_min, int size_max)
{
    if (PB_META_DROP(m)) return false;
    int L = PB_META_L(m);
    if (size_min > 0 && !(L >= size_min && (L <= size_max || size_max == -1))) return false;
    return true;
}

// Index into read.positions the rule picks when applied left-to-right ("forward": query strand
// '+' or '.') or right-to-left ("reverse": query strand '-').  Returns -1 when the reference
// skips the read and raises its DataWarning (map_factories.pyx:351-353, 450-452, 633-636).
__device__ __forceinline__ int pb_rule_index(const PbRuleDev &r, int L, bool reverse_query)
{
    if (r.kind == PB_RULE_VARIABLE) {
        if (L >= PB_LUT_SIZE) return -1;  // reference: out-of-bounds read; we drop
        return reverse_query ? __ldg(r.lut_rc + L) : __ldg(r.lut_fw + L);
    }
    if (r.param >= L) return -1;
    bool from_left = (r.kind == PB_RULE_FIVEPRIME) ? !reverse_query : reverse_query;
    return from_left ? r.param : L - 1 - r.param;
}

// read.positions[idx] for a multi-block read (CIGAR expansion, pysam get_reference_positions).
__device__ __forceinline__ int64_t pb_block_position(const PbReads &b, int64_t i, int32_t start, int idx)
{
    uint32_t k0 = __ldg(b.blk_off + i), k1 = __ldg(b.blk_off + i + 1);
    for (uint32_t k = k0; k < k1; ++k) {
        int2 bl = __ldg(b.blk + k);
        if (idx < bl.y) return (int64_t)start + bl.x + idx;
        idx -= bl.y;
    }
    return -1;
}

__device__ __forceinline__ int64_t pb_position(const PbReads &b, int64_t i, int32_t start, uint32_t m, int idx)
{
    if (PB_META_NBLK(m) <= 1 || b.blk_off == nullptr) return (int64_t)start + idx;
    return pb_block_position(b, i, start, idx);
}

// Site table: everything a read's meta word decides — drop bit, size window, strand class of the query strand, rule
// direction, rule offset — folded into ONE 16-bit look-up per read.  key = aligned length (< 256) | reverse << 8 |
// drop << 9; entry = index into read.positions of the mapped site, kSiteSkip (the read does not count on this strand
// class) or kSiteDropped (the rule has no site for this length: the reference skips the read and warns).  One table per
// strand class (plane 0 '+', 1 '-', 2 '.'), built by the read-index launch, copied to shared memory by every CTA.
// ncu on the form that evaluated the tests per read (profiles/ncu_regions_r02.txt): ~64 thread instructions per read,
// a third of them branches and reconvergence barriers around the per-read `continue`s; with the table the per-read
// path is straight-line and predicated.
constexpr int kSiteKeys = 1024;
constexpr int kSiteSkip = -1, kSiteDropped = -2;

__device__ __forceinline__ int pb_site_entry(const PbRuleDev &r, int plane, uint32_t m)
{
    const bool rev = PB_META_REV(m);
    if (!pb_passes(m, r.size_min, r.size_max) || (plane == 0 && rev) || (plane == 1 && !rev)) return kSiteSkip;   // genome_array.py:811-815
    const int idx = pb_rule_index(r, PB_META_L(m), plane == 1);          // rule direction follows the chain's strand
    return idx < 0 ? kSiteDropped : idx;
}

// first index in [lo,hi) with a[i] >= key
__device__ __forceinline__ int64_t pb_lower_bound(const int32_t *__restrict__ a, int64_t lo, int64_t hi, int64_t key)
{
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if ((int64_t)__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// lower_bound when the answer is expected a short way past `lo` (candidate slices are a few hundred
// reads): gallop 1,2,4,... from lo — probes that stay within a couple of cache lines — then bisect
__device__ __forceinline__ int64_t pb_lower_bound_near(const int32_t *__restrict__ a, int64_t lo, int64_t hi, int64_t key)
{
    if (lo >= hi || (int64_t)__ldg(a + lo) >= key) return lo;
    int64_t step = 1, prev = lo;      // invariant: a[prev] < key
    while (prev + step < hi && (int64_t)__ldg(a + prev + step) < key) { prev += step; step <<= 1; }
    const int64_t top = prev + step < hi ? prev + step : hi;
    return pb_lower_bound(a, prev + 1, top, key);
}

// lower_bound by a whole warp (all 32 lanes call with the same arguments): every step probes 32
// evenly spaced elements at once, so a range of n needs log32(n) dependent round trips instead of log2(n)
__device__ __forceinline__ int64_t pb_lower_bound_warp(const int32_t *__restrict__ a, int64_t lo, int64_t hi, int64_t key)
{
    const int lane = threadIdx.x & 31;
    while (hi - lo > 32) {
        const int64_t step = (hi - lo) / 32;
        const int64_t p = lo + (int64_t)(lane + 1) * step - 1;           // <= hi - 1
        const int n = __popc(__ballot_sync(0xffffffffu, (int64_t)__ldg(a + p) < key));   // sorted: a prefix of lanes
        if (n < 32) hi = lo + (int64_t)(n + 1) * step - 1;               // a[that probe] >= key
        lo += (int64_t)n * step;                                         // a[lo - 1] < key
    }
    const int64_t p = lo + lane;
    return lo + __popc(__ballot_sync(0xffffffffu, p < hi && (int64_t)__ldg(a + p) < key));
}

// chromosome owning global bin g: last c with chrom_bin_off[c] <= g
__device__ __forceinline__ int pb_chrom_of_bin(const PbLayoutDev &lay, int64_t g)
{
    int lo = 0, hi = lay.n_chrom;  // invariant: off[lo] <= g < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(lay.chrom_bin_off + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ unsigned long long pb_warp_sum(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
