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
// per issue).  The unit of work is now the exon BLOCK: one warp per block reads 16-byte aligned groups (uint4 /
// double2, four loads in flight per lane), trims the group's head and tail against the block and its mask bits, and
// writes one partial sum (region tables) or the block's cells of its window row (window matrices); a second small
// launch adds the partial sums per chain in block order.  Measured alternatives that were slower — a chain's positions
// flattened over one warp, several blocks per warp, deeper unrolling — are in profiles/NOTES_r02.md sections 1 and 7.3.
// The plane-free counts (pb_chain_counts) never touch a plane: reads are counted straight into the chain blocks.
//
// Every kernel takes the global-bin range [lo, hi) the calling rank owns (position sharding): positions
// outside it count zero but keep their place in the chain, so partial tables of all ranks add up to the whole table
// and geometry-only outputs (unmasked lengths, mask matrices, NaN cells) are identical on every rank.
#include "pb_tiles.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct PbPlanes { const void *p[3]; };

template <typename T> struct VecOf;
template <> struct VecOf<uint32_t> { static constexpr int V = 4; typedef uint4 L; typedef unsigned long long Acc; };
template <> struct VecOf<double> { static constexpr int V = 2; typedef double2 L; typedef double Acc; };

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
// masked sums: one warp per exon BLOCK (16-byte loads, four in flight per lane), then one warp per chain
// ------------------------------------------------------------------------------------------------------------
// History (profiles/NOTES_r02.md): one warp per chain walking block after block ran at 0.23 of the HBM rate (every
// warp waiting on one 128-byte row at a time); flattening a chain's — then four chains' — units over the lanes of one
// warp kept more bytes in flight but spent ~120 warp instructions per 32 loads on block look-ups and 64-bit shuffles
// (ncu: issue slots 49 % busy, 28 % of the warps resident at 76 registers): 0.31-0.34.  The probe pb_gather_probe shows
// what the memory system gives for this access pattern when nothing else is in the way: 4.4-6.3 TB/s for scattered
// 1-6 KB segments.  So the block is the unit of work: its warp does what the probe does (plus head / tail trimming
// and mask bits), writes ONE partial sum, and a second tiny launch adds the partial sums of every chain in a fixed
// order (deterministic for float64 planes too).  block_chain / block_pos come from the host table.
template <typename T, bool MASK>
__global__ void __launch_bounds__(256, 8)
pb_block_sums_kernel(PbPlanes planes, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                     const int32_t *__restrict__ block_chain, const int64_t *__restrict__ block_pos,
                     const uint8_t *__restrict__ block_plane, int64_t n_blocks,
                     const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                     long long lo, long long hi, double *__restrict__ block_sum)
{
    constexpr int V = VecOf<T>::V, U = 4;
    typedef typename VecOf<T>::L L;
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_blocks) return;
    const long long bs = __ldg(bstart + k), be = __ldg(bend + k);
    const long long cs = bs > lo ? bs : lo, ce = be < hi ? be : hi;          // the part this rank owns
    if (cs >= ce) { if (lane == 0) block_sum[k] = 0.0; return; }
    // the plane of a block comes from the host table: the data loads depend on ONE round trip (bounds + plane), not
    // on block -> chain -> plane
    const T *__restrict__ vec = static_cast<const T *>(planes.p[__ldg(block_plane + k)]);
    const long long g0 = cs & ~(long long)(V - 1);
    const int head = (int)(cs - g0), span = (int)(ce - g0);                  // owned elements: [head, span) relative to g0
    const int n_groups = (span + V - 1) / V;
    const L *__restrict__ src = reinterpret_cast<const L *>(vec + g0);
    const long long mbase = MASK ? __ldg(mask_off + __ldg(block_chain + k)) + __ldg(block_pos + k) + (g0 - bs) : 0;   // mask bit of element g0
    typename VecOf<T>::Acc acc = 0;
    for (int u0 = lane; u0 < n_groups; u0 += 32 * U) {
        L v[U];
#pragma unroll
        for (int x = 0; x < U; ++x) if (u0 + 32 * x < n_groups) v[x] = __ldg(src + u0 + 32 * x);
#pragma unroll
        for (int x = 0; x < U; ++x) {
            const int u = u0 + 32 * x;
            if (u >= n_groups) break;
            const int rel = u * V;
            const int e_lo = head > rel ? head - rel : 0;
            const int e_hi = span - rel < V ? span - rel : V;
            uint32_t m = ((1u << e_hi) - 1u) & ~((1u << e_lo) - 1u);
            if (MASK) m &= ~(mask_bits_at(mask_words, mbase + rel + e_lo, e_hi - e_lo) << e_lo);
            add_group<T>(acc, v[x], m);
        }
    }
    double total;
    if (sizeof(T) == 4) total = (double)pb_warp_sum((unsigned long long)acc);   // exact: counts stay far below 2^53
    else total = warp_sum_f64((double)acc);
    if (lane == 0) block_sum[k] = total;
}

// ------------------------------------------------------------------------------------------------------------
// the same partial sums with the blocks STAGED by the copy engine (cp.async.bulk global -> shared, completion on an
// mbarrier): a warp takes kTmaSlots consecutive blocks, one lane per block issues ONE bulk copy of the block's
// 16-byte-aligned bytes (up to kTmaSlotBytes; longer blocks read the rest with plain loads), and the warp then sums
// slot after slot out of shared memory.  Bytes in flight cost no registers here (8 x ~1.5 KB per warp).  A/B aid
// (PB_REGION_TMA=1): see profiles/NOTES_r02.md section 7.14 for what it measured.
// ------------------------------------------------------------------------------------------------------------
constexpr int kTmaSlots = 8, kTmaSlotBytes = 2048, kTmaWarps = 4;

__device__ __forceinline__ uint32_t pb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pb_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" :: "r"(pb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pb_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" :: "r"(pb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_bulk_load(void *sdst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(pb_smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(pb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pb_mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(pb_smem_u32(bar)), "r"(parity) : "memory");
}

template <typename T, bool MASK>
__global__ void __launch_bounds__(kTmaWarps * 32)
pb_block_sums_tma_kernel(PbPlanes planes, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                         const int32_t *__restrict__ block_chain, const int64_t *__restrict__ block_pos,
                         const uint8_t *__restrict__ block_plane, int64_t n_blocks,
                         const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                         long long lo, long long hi, double *__restrict__ block_sum)
{
    constexpr int V = VecOf<T>::V;
    typedef typename VecOf<T>::L L;
    extern __shared__ __align__(128) unsigned char stage[];          // [warp][slot][kTmaSlotBytes]
    __shared__ __align__(8) uint64_t bars[kTmaWarps][kTmaSlots];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane < kTmaSlots) pb_mbar_init(&bars[wid][lane], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pb_fence_proxy_async();
    __syncthreads();
    const int64_t k0 = (((int64_t)blockIdx.x * kTmaWarps) + wid) * kTmaSlots;
    if (k0 >= n_blocks) return;
    // lane j < kTmaSlots owns block k0 + j: its row, its geometry, its bulk copy
    int head = 0, span = 0, n_groups = 0;
    long long mbase = 0;
    const L *src = nullptr;
    if (lane < kTmaSlots && k0 + lane < n_blocks) {
        const int64_t k = k0 + lane;
        const long long bs = __ldg(bstart + k), be = __ldg(bend + k);
        const long long cs = bs > lo ? bs : lo, ce = be < hi ? be : hi;      // the part this rank owns
        if (cs < ce) {
            const T *vec = static_cast<const T *>(planes.p[__ldg(block_plane + k)]);
            const long long g0 = cs & ~(long long)(V - 1);
            head = (int)(cs - g0); span = (int)(ce - g0);
            n_groups = (span + V - 1) / V;
            src = reinterpret_cast<const L *>(vec + g0);
            if (MASK) mbase = __ldg(mask_off + __ldg(block_chain + k)) + __ldg(block_pos + k) + (g0 - bs);
            const unsigned bytes = (unsigned)(n_groups * 16 < kTmaSlotBytes ? n_groups * 16 : kTmaSlotBytes);
            pb_mbar_expect_tx(&bars[wid][lane], bytes);
            pb_bulk_load(stage + ((size_t)wid * kTmaSlots + lane) * kTmaSlotBytes, src, bytes, &bars[wid][lane]);
        }
    }
#pragma unroll 1
    for (int j = 0; j < kTmaSlots; ++j) {
        if (k0 + j >= n_blocks) break;
        const int ng = __shfl_sync(kFull, n_groups, j);
        if (ng == 0) { if (lane == 0) block_sum[k0 + j] = 0.0; continue; }
        const int hd = __shfl_sync(kFull, head, j), sp = __shfl_sync(kFull, span, j);
        const long long mb = MASK ? __shfl_sync(kFull, mbase, j) : 0;
        const L *gsrc = reinterpret_cast<const L *>(__shfl_sync(kFull, (unsigned long long)src, j));
        const L *ssrc = reinterpret_cast<const L *>(stage + ((size_t)wid * kTmaSlots + j) * kTmaSlotBytes);
        const int staged = ng < kTmaSlotBytes / 16 ? ng : kTmaSlotBytes / 16;
        pb_mbar_wait(&bars[wid][j], 0);
        typename VecOf<T>::Acc acc = 0;
        for (int u = lane; u < ng; u += 32) {
            const L v = u < staged ? ssrc[u] : __ldg(gsrc + u);
            const int rel = u * V;
            const int e_lo = hd > rel ? hd - rel : 0;
            const int e_hi = sp - rel < V ? sp - rel : V;
            uint32_t m = ((1u << e_hi) - 1u) & ~((1u << e_lo) - 1u);
            if (MASK) m &= ~(mask_bits_at(mask_words, mb + rel + e_lo, e_hi - e_lo) << e_lo);
            add_group<T>(acc, v, m);
        }
        double total;
        if (sizeof(T) == 4) total = (double)pb_warp_sum((unsigned long long)acc);
        else total = warp_sum_f64((double)acc);
        if (lane == 0) block_sum[k0 + j] = total;
    }
}

// one THREAD per chain (chains have a handful of blocks): partial sums of its blocks in block order, its length, its
// masked positions
__global__ void __launch_bounds__(256)
pb_chain_totals_kernel(const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                       const int64_t *__restrict__ chain_off, int64_t n_chains,
                       const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                       const double *__restrict__ block_sum, const unsigned long long *__restrict__ counts,
                       double *__restrict__ sums, int64_t *__restrict__ live_len)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    const int64_t k0 = __ldg(chain_off + c), k1 = __ldg(chain_off + c + 1);
    long long len = 0;
    double acc = 0.0;
    for (int64_t k = k0; k < k1; ++k) {
        len += __ldg(bend + k) - __ldg(bstart + k);
        if (block_sum) acc += block_sum[k];
    }
    long long live = len;
    if (mask_words && len > 0) {
        const long long b0 = __ldg(mask_off + c);
        const long long wa = b0 >> 5, wb = (b0 + len - 1) >> 5;
        long long cnt = 0;
        for (long long w = wa; w <= wb; ++w) {
            uint32_t x = __ldg(mask_words + w);
            if (w == wa) x &= kFull << (b0 & 31);
            if (w == wb) x &= kFull >> (31 - ((b0 + len - 1) & 31));
            cnt += __popc(x);
        }
        live -= cnt;
    }
    sums[c] = counts ? (double)counts[c] : acc;
    live_len[c] = live;
}

// ------------------------------------------------------------------------------------------------------------
// window matrices: one warp per exon block writes the cells of its positions; one warp per chain fills the rest
// ------------------------------------------------------------------------------------------------------------
// Window rows (row_off == NULL): row c of a width-wide matrix, chain c laid 5'->3' from column row_col[c].  Ragged rows
// (pb_gather_chains): chain c owns cells [row_off[c], row_off[c] + its length) of a flat vector.
template <typename T, bool MASK>
__global__ void __launch_bounds__(256)
pb_window_blocks_kernel(PbPlanes planes, const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                        const int32_t *__restrict__ block_chain, const int64_t *__restrict__ block_pos,
                        const uint8_t *__restrict__ block_plane, const uint8_t *__restrict__ chain_reverse,
                        const int64_t *__restrict__ chain_len, const int32_t *__restrict__ row_col,
                        const int64_t *__restrict__ row_off, int64_t n_blocks, int32_t width_,
                        const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                        long long lo, long long hi, double *__restrict__ matrix, uint8_t *__restrict__ maskmat)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_blocks) return;
    const long long bs = __ldg(bstart + k), be = __ldg(bend + k);
    const int n = (int)(be - bs);
    const int c = __ldg(block_chain + k);
    const T *__restrict__ vec = static_cast<const T *>(planes.p[__ldg(block_plane + k)]);
    const long long len = __ldg(chain_len + c), pos0 = __ldg(block_pos + k);
    const bool rev = __ldg(chain_reverse + c);
    const long long col0 = row_off ? 0 : __ldg(row_col + c);
    const long long width = row_off ? len : width_;
    const int64_t cell0 = row_off ? __ldg(row_off + c) : c * (int64_t)width_;
    double *__restrict__ row = matrix + cell0;
    uint8_t *__restrict__ mrow = maskmat + cell0;
    const long long mbase = MASK ? __ldg(mask_off + c) + pos0 : 0;
    // column of the block's first position and the direction columns run in
    const long long cfirst = col0 + (rev ? len - 1 - pos0 : pos0);
    const int dir = rev ? -1 : 1;
    for (int i0 = lane; i0 < n; i0 += 32 * U) {
        T v[U];
#pragma unroll
        for (int x = 0; x < U; ++x) {
            const int i = i0 + 32 * x;
            const long long p = bs + i;
            v[x] = T(0);
            if (i < n && p >= lo && p < hi) v[x] = __ldg(vec + p);          // positions of other ranks count zero
        }
#pragma unroll
        for (int x = 0; x < U; ++x) {
            const int i = i0 + 32 * x;
            if (i >= n) break;
            const long long col = cfirst + dir * (long long)i;
            if (col >= 0 && col < width) {
                row[col] = (double)v[x];
                mrow[col] = MASK ? (uint8_t)mask_bits_at(mask_words, mbase + i, 1) : (uint8_t)0;
            }
        }
    }
}

// columns no chain position reaches stay "masked NaN" (metagene.py:895-898)
__global__ void __launch_bounds__(256)
pb_window_fill_kernel(const int64_t *__restrict__ chain_len, const int32_t *__restrict__ row_col, int64_t n_chains,
                      int32_t width, double *__restrict__ matrix, uint8_t *__restrict__ maskmat)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_chains) return;
    const long long col0 = __ldg(row_col + c), len = __ldg(chain_len + c);
    double *__restrict__ row = matrix + c * (int64_t)width;
    uint8_t *__restrict__ mrow = maskmat + c * (int64_t)width;
    const long long a = col0 < 0 ? 0 : (col0 < width ? col0 : width);                 // covered columns: [a, b)
    const long long b = col0 + len < 0 ? 0 : (col0 + len < width ? col0 + len : width);
    for (long long col = lane; col < a; col += 32) { row[col] = nan(""); mrow[col] = 1; }
    for (long long col = (b > a ? b : a) + lane; col < width; col += 32) { row[col] = nan(""); mrow[col] = 1; }
}

// ------------------------------------------------------------------------------------------------------------
// plane-free region counts of a point rule: no count vectors are materialised.
// Equals pb_region_sums over the planes pb_map_point would write (same strand pre-filter per query strand,
// genome_array.py:811-815; same rule direction; same masks).
//
// The reads that can map into a block are a contiguous slice of the coordinate-sorted batch.  Expression is skewed —
// a few chains hold millions of reads — so the work is cut by READS, not by chains:
//   1. pb_read_index_kernel   first read at or beyond every 16384-bin boundary of the layout (a few hundred KB, L2)
//   2. pb_chain_first_items_kernel  one warp per block: its read slice (two short searches inside the index cell), its
//                             FIRST 2048-read work item counted right away (most blocks need no more), and the number
//                             of further items
//   3. exclusive scan of the further-item counts
//   4. pb_chain_items_kernel  persistent warps take the further items round-robin (the hot blocks): look the block up
//                             (32-ary search of the offsets), count the sites of the item's reads that land on unmasked
//                             positions of the block inside this rank's bins, one 64-bit atomic per item
//   5. pb_chain_totals_kernel one warp per chain: the count as float64, the unmasked length
// ------------------------------------------------------------------------------------------------------------
constexpr int kItemReads = 2048;
constexpr int kIndexShift = 14;          // PB_LAYOUT_ALIGN = 1 << 14: an index cell never spans two chromosomes

__global__ void pb_read_index_kernel(PbReads b, PbLayoutDev lay, long long n_cells, long long *__restrict__ index,
                                     PbRuleDev r, int16_t *__restrict__ site_tab)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (site_tab && g < 3 * 1024) {          // the site tables of pb_chain_counts ride along (kSiteKeys entries per plane)
        const uint32_t key = (uint32_t)g & 1023u;
        // key bits 8 / 9 are meta bits 16 (reverse) / 17 (drop)
        site_tab[g] = (int16_t)pb_site_entry(r, (int)(g >> 10), (key & 0xFFu) | ((key & 0x300u) << 8));
    }
    if (g > n_cells) return;
    if (g == n_cells) { index[g] = b.n_reads; return; }
    const long long bin = g << kIndexShift;
    const int ch = pb_chrom_of_bin(lay, bin);
    long long r0 = b.n_reads, r1 = b.n_reads;
    if (ch < b.n_chrom) { r0 = __ldg(b.chrom_read_off + ch); r1 = __ldg(b.chrom_read_off + ch + 1); }
    index[g] = pb_lower_bound(b.ref_start, r0, r1, bin - __ldg(lay.chrom_bin_off + ch));
}

// first read of chromosome ch (reads [r0, r1)) starting at or beyond chromosome position p, through the index
__device__ __forceinline__ long long pb_indexed_lower_bound(const PbReads &b, const long long *__restrict__ index,
                                                            long long base, long long padded_end, long long r0, long long r1,
                                                            long long p)
{
    if (p <= 0) return r0;
    if (base + p >= padded_end) return r1;
    const long long g = (base + p) >> kIndexShift;
    long long a = __ldg(index + g), e = __ldg(index + g + 1);
    a = a < r0 ? r0 : a;
    e = e > r1 ? r1 : e;
    return pb_lower_bound_warp(b.ref_start, a, e, p);
}

// last k with off[k] <= item (off ascending, off[0] = 0, off[n] = total > item), by a whole warp
__device__ __forceinline__ long long pb_block_of_item(const uint32_t *__restrict__ off, long long n, uint32_t item)
{
    const int lane = threadIdx.x & 31;
    long long lo = 0, hi = n;                 // invariant: off[lo] <= item < off[hi]
    while (hi - lo > 32) {
        const long long step = (hi - lo) / 32;
        const long long p = lo + (long long)(lane + 1) * step;           // <= hi
        const bool le = p < hi && __ldg(off + p) <= item;
        const int cnt = __popc(__ballot_sync(kFull, le));                 // probes are ascending: a prefix of lanes
        if (cnt < 32) hi = lo + (long long)(cnt + 1) * step < hi ? lo + (long long)(cnt + 1) * step : hi;
        lo += (long long)cnt * step;
    }
    const long long p = lo + 1 + lane;
    return lo + __popc(__ballot_sync(kFull, p < hi && __ldg(off + p) <= item));
}

// Sites of the reads [first, first + n) that land on unmasked positions [ps, ps + width) of one block.  Positions are
// 32-bit chromosome coordinates and the range test is one unsigned compare.  A lane takes FOUR consecutive reads per
// 16-byte load of each array (the item is walked from the 4-aligned read at or before `first`), two such loads of
// each array in flight.  `tab`: the plane's site table in shared memory.
template <bool BLOCKS, bool MASK>
__device__ __forceinline__ unsigned int pb_count_item(const PbReads &b, const PbRuleDev &r, int plane, const int16_t *tab,
                                                      long long first, int n, int lane,
                                                      int ps, unsigned width, const uint32_t *__restrict__ mask_words, long long mbit,
                                                      unsigned int &drop_len)
{
    const long long a0 = first & ~3ll;                         // 16-byte aligned start
    const int skip = (int)(first - a0), end = skip + n;        // reads [skip, end) of the aligned run are the item's
    const uint4 *__restrict__ meta4 = reinterpret_cast<const uint4 *>(b.meta + a0);
    const int4 *__restrict__ start4 = reinterpret_cast<const int4 *>(b.ref_start + a0);
    const int n_quads = (end + 3) >> 2;
    unsigned int count = 0;
    constexpr int kU = 2;
    for (int q0 = lane; q0 < n_quads; q0 += 32 * kU) {
        uint4 mq[kU];
        int4 sq[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int q = q0 + 32 * u;
            // the last quad may reach past the batch's last read by up to 3 elements: the arrays of a batch are
            // allocated in whole 16-byte units by every producer (torch allocations are 512-byte granular)
            if (q < n_quads) { mq[u] = __ldg(meta4 + q); sq[u] = __ldg(start4 + q); }
            else { mq[u] = make_uint4(1u << 17, 1u << 17, 1u << 17, 1u << 17); sq[u] = make_int4(0, 0, 0, 0); }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const uint32_t mm[4] = {mq[u].x, mq[u].y, mq[u].z, mq[u].w};
            const int32_t ss[4] = {sq[u].x, sq[u].y, sq[u].z, sq[u].w};
            const int j0 = (q0 + 32 * u) * 4 - skip;             // index of the quad's first read within the item
            int idx[4];
            if (__builtin_expect(((mm[0] | mm[1] | mm[2] | mm[3]) & 0xFF00u) != 0u, 0)) {     // a read of 256 aligned bases or more
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    idx[e] = (mm[e] & 0xFF00u) ? pb_site_entry(r, plane, mm[e]) : tab[(mm[e] & 0xFFu) | ((mm[e] >> 8) & 0x300u)];
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) idx[e] = tab[(mm[e] & 0xFFu) | ((mm[e] >> 8) & 0x300u)];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool mine = (unsigned)(j0 + e) < (unsigned)n;                           // the read belongs to the item
                int p = ss[e] + idx[e];
                if (BLOCKS && (mm[e] >> 24) > 1 && idx[e] >= 0 && mine) {
                    const long long pp = pb_block_position(b, a0 + skip + j0 + e, ss[e], idx[e]);
                    p = pp < 0 ? ps - 1 : (int)pp;
                }
                bool ok = mine && idx[e] >= 0 && (unsigned)(p - ps) < width;
                if (MASK) { if (ok) ok = !mask_bits_at(mask_words, mbit + p, 1); }
                count += ok ? 1u : 0u;
            }
            const int low = min(min(idx[0], idx[1]), min(idx[2], idx[3]));
            if (__builtin_expect(low == kSiteDropped, 0)) {                                  // the rule has no site for a length
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (idx[e] == kSiteDropped && (unsigned)(j0 + e) < (unsigned)n) drop_len = mm[e] & 0xFFFFu;
            }
        }
    }
    return count;
}

// What one work item fixes: the block's owned positions, its strand class and rule direction, its mask bits.
struct PbItemCtx {
    int ps;                      // first owned position (chromosome coordinate)
    unsigned width;              // owned positions
    long long mbit;              // + chromosome position = mask bit index
    int plane;                   // 0 '+', 1 '-', 2 '.'
    int chain;
};

__device__ __forceinline__ PbItemCtx pb_item_ctx(const PbLayoutDev &lay, long long gs, long long ge, long long lo, long long hi,
                                                 int c, int plane, const uint32_t *__restrict__ mask_words,
                                                 const int64_t *__restrict__ mask_off, long long bpos)
{
    const long long cs = gs > lo ? gs : lo, ce = ge < hi ? ge : hi;
    const int ch = pb_chrom_of_bin(lay, gs);
    const long long base = __ldg(lay.chrom_bin_off + ch);
    PbItemCtx x;
    x.ps = (int)(cs - base);
    x.width = ce > cs ? (unsigned)(ce - cs) : 0u;
    x.mbit = mask_words ? __ldg(mask_off + c) + bpos - (gs - base) : 0;
    x.plane = plane;
    x.chain = c;
    return x;
}

// count the sites of reads [first, first + n) on the item's block; `tabs`: the three site tables in shared memory
__device__ __forceinline__ unsigned int pb_count_reads(const PbReads &b, const PbRuleDev &r, const PbItemCtx &x, const int16_t *tabs,
                                                       long long first, int n, int lane, const uint32_t *__restrict__ mask_words,
                                                       unsigned int &drop_len)
{
    if (n <= 0 || x.width == 0) return 0;
    const int16_t *tab = tabs + x.plane * kSiteKeys;
    if (mask_words)
        return b.blk_off != nullptr ? pb_count_item<true, true>(b, r, x.plane, tab, first, n, lane, x.ps, x.width, mask_words, x.mbit, drop_len)
                                    : pb_count_item<false, true>(b, r, x.plane, tab, first, n, lane, x.ps, x.width, mask_words, x.mbit, drop_len);
    return b.blk_off != nullptr ? pb_count_item<true, false>(b, r, x.plane, tab, first, n, lane, x.ps, x.width, mask_words, x.mbit, drop_len)
                                : pb_count_item<false, false>(b, r, x.plane, tab, first, n, lane, x.ps, x.width, mask_words, x.mbit, drop_len);
}

// the three site tables: global (built once per call) -> shared memory of the CTA
__device__ __forceinline__ void pb_load_site_tables(int16_t *s_tab, const int16_t *__restrict__ g_tab)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(g_tab);
    uint4 *dst = reinterpret_cast<uint4 *>(s_tab);
    for (int j = threadIdx.x; j < 3 * kSiteKeys * 2 / 16; j += blockDim.x) dst[j] = __ldg(src + j);
    __syncthreads();
}

__device__ __forceinline__ void pb_flush_drops(unsigned long long *__restrict__ stats, unsigned int drop_len, int drop_plane)
{
    if (stats && __any_sync(kFull, drop_len != 0)) {
        const unsigned int len = __reduce_max_sync(kFull, drop_len);
        if (drop_len == len) {                                // flags, not counts: which strand class saw a drop
            const int which = drop_plane == 0 ? PB_STAT_DROPPED_PLUS : (drop_plane == 1 ? PB_STAT_DROPPED_MINUS : PB_STAT_DROPPED_ANY);
            atomicAdd(stats + which, 1ull);
            atomicMax(stats + PB_STAT_DROPPED_LEN, (unsigned long long)len);
        }
    }
}

// one warp per block: its read slice through the index, its FIRST work item right away (most blocks have no more),
// and how many further items it needs
__global__ void __launch_bounds__(128)
pb_chain_first_items_kernel(PbReads b, PbRuleDev r, PbLayoutDev lay, const long long *__restrict__ index,
                            const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                            const int32_t *__restrict__ block_chain, const int64_t *__restrict__ block_pos,
                            const uint8_t *__restrict__ block_plane, int64_t n_blocks,
                            const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                            long long lo, long long hi, long long total_bins,
                            long long *__restrict__ slice_first, uint32_t *__restrict__ extra_items,
                            unsigned long long *__restrict__ counts, unsigned long long *__restrict__ stats,
                            const int16_t *__restrict__ site_tab, int first_reads, int item_reads)
{
    __shared__ __align__(16) int16_t s_tab[3 * kSiteKeys];
    pb_load_site_tables(s_tab, site_tab);
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_blocks) return;
    const long long gs = __ldg(bstart + k), ge = __ldg(bend + k);
    const long long cs = gs > lo ? gs : lo, ce = ge < hi ? ge : hi;
    long long first = 0, last = 0;
    unsigned int drop_len = 0;
    int plane = 0;
    if (cs < ce && gs >= 0 && gs < total_bins) {
        const int ch = pb_chrom_of_bin(lay, gs);
        const long long base = __ldg(lay.chrom_bin_off + ch), padded_end = __ldg(lay.chrom_bin_off + ch + 1);
        long long r0 = 0, r1 = 0;
        if (ch < b.n_chrom) { r0 = __ldg(b.chrom_read_off + ch); r1 = __ldg(b.chrom_read_off + ch + 1); }
        first = pb_indexed_lower_bound(b, index, base, padded_end, r0, r1, cs - base - b.max_span + 1);
        last = pb_indexed_lower_bound(b, index, base, padded_end, r0, r1, ce - base);
        if (last < first) last = first;
        if (last > first) {
            const int c = __ldg(block_chain + k);
            plane = __ldg(block_plane + k);
            const PbItemCtx x = pb_item_ctx(lay, gs, ge, lo, hi, c, plane, mask_words, mask_off, __ldg(block_pos + k));
            const int n = (int)(last - first < first_reads ? last - first : first_reads);
            unsigned int count = pb_count_reads(b, r, x, s_tab, first, n, lane, mask_words, drop_len);
            count = __reduce_add_sync(kFull, count);
            if (lane == 0 && count) atomicAdd(counts + c, (unsigned long long)count);
        }
    }
    if (lane == 0) {
        slice_first[2 * k] = first;
        slice_first[2 * k + 1] = last;
        const long long rest = last - first - first_reads;            // reads beyond the first item
        extra_items[k] = rest > 0 ? (uint32_t)((rest + item_reads - 1) / item_reads) : 0u;
    }
    pb_flush_drops(stats, drop_len, plane);
}

// persistent warps over the items beyond the first of every block (the hot blocks)
__global__ void __launch_bounds__(256)
pb_chain_items_kernel(PbReads b, PbRuleDev r, PbLayoutDev lay,
                      const int64_t *__restrict__ bstart, const int64_t *__restrict__ bend,
                      const int32_t *__restrict__ block_chain, const int64_t *__restrict__ block_pos,
                      const uint8_t *__restrict__ block_plane, int64_t n_blocks,
                      const uint32_t *__restrict__ mask_words, const int64_t *__restrict__ mask_off,
                      long long lo, long long hi, const long long *__restrict__ slice_first,
                      const uint32_t *__restrict__ item_off, unsigned long long *__restrict__ counts,
                      unsigned long long *__restrict__ stats, const int16_t *__restrict__ site_tab, int first_reads, int item_reads)
{
    __shared__ __align__(16) int16_t s_tab[3 * kSiteKeys];
    pb_load_site_tables(s_tab, site_tab);
    const int lane = threadIdx.x & 31;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t total = __ldg(item_off + n_blocks);
    unsigned int drop_len = 0;
    int drop_plane = 0;
    for (long long item = warp; item < (long long)total; item += n_warps) {
        const long long k = pb_block_of_item(item_off, n_blocks, (uint32_t)item);
        const int c = __ldg(block_chain + k);
        const int plane = __ldg(block_plane + k);
        const PbItemCtx x = pb_item_ctx(lay, __ldg(bstart + k), __ldg(bend + k), lo, hi, c, plane, mask_words, mask_off, __ldg(block_pos + k));
        // item e of block k is its e-th run of kItemReads reads after the first `first_reads`, which were counted with the slice
        const long long first = __ldg(slice_first + 2 * k) + first_reads + (item - (long long)__ldg(item_off + k)) * item_reads;
        const long long slice_end = __ldg(slice_first + 2 * k + 1);
        const int n = (int)(slice_end - first < item_reads ? slice_end - first : item_reads);
        unsigned int dl = 0;
        unsigned int count = pb_count_reads(b, r, x, s_tab, first, n, lane, mask_words, dl);
        if (dl) { drop_len = dl; drop_plane = plane; }
        count = __reduce_add_sync(kFull, count);
        if (lane == 0 && count) atomicAdd(counts + c, (unsigned long long)count);
    }
    pb_flush_drops(stats, drop_len, drop_plane);
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

extern "C" size_t pb_region_sums_workspace_bytes(int64_t n_blocks)
{
    return n_blocks < 0 ? 0 : (size_t)n_blocks * sizeof(double) + 16;
}

extern "C" int pb_region_sums(const void *const *planes, int vec_dtype,
                              const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                              const uint8_t *chain_plane, const int32_t *block_chain, const int64_t *block_pos,
                              const uint8_t *block_plane,
                              int64_t n_chains, int64_t n_blocks, const uint8_t *mask_bits, const int64_t *mask_off,
                              int64_t bin_begin, int64_t bin_end,
                              double *sums, int64_t *live_len, void *workspace, size_t workspace_bytes, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off, vec_dtype, bin_begin, bin_end);
    if (rc) return rc;
    if (!sums || !live_len || !block_chain || !block_pos || !block_plane || n_blocks < 0) { pb_set_error("pb_region_sums: null outputs or block tables"); return PB_EINVAL; }
    for (int i = 0; i < 3; ++i)
        if ((uintptr_t)planes[i] & 15) { pb_set_error("pb_region_sums: planes must be 16-byte aligned"); return PB_EINVAL; }
    if (!workspace || workspace_bytes < pb_region_sums_workspace_bytes(n_blocks)) { pb_set_error("pb_region_sums: workspace too small"); return PB_ENOSPACE; }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbPlanes pl{{planes[0], planes[1], planes[2]}};
    double *block_sum = (double *)workspace;
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask_bits);
    if (n_blocks > 0 && getenv("PB_REGION_TMA")) {
        const unsigned grid = (unsigned)((n_blocks + kTmaWarps * kTmaSlots - 1) / (kTmaWarps * kTmaSlots));
        const size_t smem = (size_t)kTmaWarps * kTmaSlots * kTmaSlotBytes;
#define PB_LAUNCH_TMA(T, M) do { PB_CUDA_CHECK(cudaFuncSetAttribute(pb_block_sums_tma_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pb_block_sums_tma_kernel<T, M><<<grid, kTmaWarps * 32, smem, stream>>>(pl, bstart, bend, block_chain, block_pos, block_plane, n_blocks, mw, mask_off, bin_begin, bin_end, block_sum); } while (0)
        if (vec_dtype == 0) { if (mw) PB_LAUNCH_TMA(uint32_t, true); else PB_LAUNCH_TMA(uint32_t, false); }
        else { if (mw) PB_LAUNCH_TMA(double, true); else PB_LAUNCH_TMA(double, false); }
#undef PB_LAUNCH_TMA
    } else if (n_blocks > 0) {
        const unsigned grid = (unsigned)((n_blocks * 32 + 255) / 256);
#define PB_LAUNCH_SUMS(T, M) pb_block_sums_kernel<T, M><<<grid, 256, 0, stream>>>(pl, bstart, bend, block_chain, block_pos, block_plane, n_blocks, mw, mask_off, bin_begin, bin_end, block_sum)
        if (vec_dtype == 0) { if (mw) PB_LAUNCH_SUMS(uint32_t, true); else PB_LAUNCH_SUMS(uint32_t, false); }
        else { if (mw) PB_LAUNCH_SUMS(double, true); else PB_LAUNCH_SUMS(double, false); }
#undef PB_LAUNCH_SUMS
    }
    pb_chain_totals_kernel<<<(unsigned)((n_chains + 255) / 256), 256, 0, stream>>>(
        bstart, bend, chain_off, n_chains, mw, mask_off, block_sum, nullptr, sums, live_len);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

static int launch_windows(const void *const *planes, int vec_dtype,
                          const int64_t *bstart, const int64_t *bend, const int32_t *block_chain, const int64_t *block_pos,
                          const uint8_t *block_plane, const uint8_t *chain_reverse, const int64_t *chain_len,
                          const int32_t *row_col, const int64_t *row_off, int64_t n_chains, int64_t n_blocks, int32_t width,
                          const uint8_t *mask_bits, const int64_t *mask_off, int64_t bin_begin, int64_t bin_end,
                          double *matrix, uint8_t *maskmat, cudaStream_t stream)
{
    PbPlanes pl{{planes[0], planes[1], planes[2]}};
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask_bits);
    if (!row_off)
        pb_window_fill_kernel<<<(unsigned)((n_chains * 32 + 255) / 256), 256, 0, stream>>>(chain_len, row_col, n_chains, width, matrix, maskmat);
    if (n_blocks > 0) {
        const unsigned grid = (unsigned)((n_blocks * 32 + 255) / 256);
#define PB_LAUNCH_WIN(T, M) pb_window_blocks_kernel<T, M><<<grid, 256, 0, stream>>>(pl, bstart, bend, block_chain, block_pos, block_plane, chain_reverse, chain_len, row_col, row_off, n_blocks, width, mw, mask_off, bin_begin, bin_end, matrix, maskmat)
        if (vec_dtype == 0) { if (mw) PB_LAUNCH_WIN(uint32_t, true); else PB_LAUNCH_WIN(uint32_t, false); }
        else { if (mw) PB_LAUNCH_WIN(double, true); else PB_LAUNCH_WIN(double, false); }
#undef PB_LAUNCH_WIN
    }
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_gather_windows(const void *const *planes, int vec_dtype,
                                 const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                 const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                 const int32_t *block_chain, const int64_t *block_pos, const uint8_t *block_plane,
                                 const int64_t *chain_len,
                                 const int32_t *row_col, int64_t n_chains, int64_t n_blocks, int32_t width,
                                 const uint8_t *mask_bits, const int64_t *mask_off,
                                 int64_t bin_begin, int64_t bin_end,
                                 double *matrix, uint8_t *maskmat, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off, vec_dtype, bin_begin, bin_end);
    if (rc) return rc;
    if (!chain_reverse || !row_col || !matrix || !maskmat || !block_chain || !block_pos || !block_plane || !chain_len || width <= 0 || n_blocks < 0) {
        pb_set_error("pb_gather_windows: bad arguments"); return PB_EINVAL;
    }
    if (n_chains == 0) return PB_OK;
    return launch_windows(planes, vec_dtype, bstart, bend, block_chain, block_pos, block_plane, chain_reverse, chain_len, row_col, nullptr,
                          n_chains, n_blocks, width, mask_bits, mask_off, bin_begin, bin_end, matrix, maskmat, (cudaStream_t)stream_);
}

extern "C" int pb_gather_chains(const void *const *planes, int vec_dtype,
                                const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                const int32_t *block_chain, const int64_t *block_pos, const uint8_t *block_plane,
                                const int64_t *chain_len,
                                const int64_t *row_off, int64_t n_chains, int64_t n_blocks,
                                const uint8_t *mask_bits, const int64_t *mask_off,
                                int64_t bin_begin, int64_t bin_end,
                                double *values, uint8_t *masked, void *stream_)
{
    int rc = check_chains(planes, bstart, bend, chain_off, chain_plane, n_chains, mask_bits, mask_off, vec_dtype, bin_begin, bin_end);
    if (rc) return rc;
    if (!chain_reverse || !row_off || !values || !masked || !block_chain || !block_pos || !block_plane || !chain_len || n_blocks < 0) {
        pb_set_error("pb_gather_chains: bad arguments"); return PB_EINVAL;
    }
    if (n_chains == 0) return PB_OK;
    return launch_windows(planes, vec_dtype, bstart, bend, block_chain, block_pos, block_plane, chain_reverse, chain_len, nullptr, row_off,
                          n_chains, n_blocks, 0, mask_bits, mask_off, bin_begin, bin_end, values, masked, (cudaStream_t)stream_);
}

extern "C" size_t pb_chain_counts_workspace_bytes(int64_t total_bins, int64_t n_blocks, int64_t n_chains)
{
    if (total_bins < 0 || n_blocks < 0 || n_chains < 0) return 0;
    const size_t cells = (size_t)(total_bins >> kIndexShift) + 2;
    return cells * 8 + (size_t)n_blocks * 16 + ((size_t)n_blocks + 1) * 8 + (size_t)pb_scan_part_entries(n_blocks + 1) * 4
        + (size_t)n_chains * 8 + 256 + 3 * 1024 * sizeof(int16_t) + 16;
}

extern "C" int pb_chain_counts(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                               const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                               const uint8_t *chain_plane, const int32_t *block_chain, const int64_t *block_pos,
                               const uint8_t *block_plane, int64_t n_chains, int64_t n_blocks,
                               const uint8_t *mask_bits, const int64_t *mask_off,
                               int64_t bin_begin, int64_t bin_end,
                               double *sums, int64_t *live_len, uint64_t *stats,
                               void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!batch || !layout || !rule || !bstart || !bend || !chain_off || !chain_plane || !block_chain || !block_pos || !block_plane || !sums ||
        !live_len || n_chains < 0 || n_blocks < 0) {
        pb_set_error("pb_chain_counts: null argument"); return PB_EINVAL;
    }
    if (rule->kind != PB_RULE_FIVEPRIME && rule->kind != PB_RULE_THREEPRIME && rule->kind != PB_RULE_VARIABLE) {
        pb_set_error("pb_chain_counts: needs a point rule (5' / 3' / variable)"); return PB_EINVAL;
    }
    if (rule->kind == PB_RULE_VARIABLE && (!rule->lut_fw || !rule->lut_rc)) { pb_set_error("pb_chain_counts: variable rule needs LUTs"); return PB_EINVAL; }
    if (mask_bits && (!mask_off || ((uintptr_t)mask_bits & 3))) { pb_set_error("pb_chain_counts: mask_bits need mask_off and 4-byte alignment"); return PB_EINVAL; }
    if (bin_begin > bin_end) { pb_set_error("pb_chain_counts: inverted bin range"); return PB_EINVAL; }
    if (((uintptr_t)batch->ref_start & 15) || ((uintptr_t)batch->meta & 15)) {
        pb_set_error("pb_chain_counts: ref_start and meta must be 16-byte aligned (the reads are loaded four at a time)"); return PB_EINVAL;
    }
    if (n_blocks >= ((int64_t)1 << 31) || batch->n_reads / kItemReads + n_blocks >= ((int64_t)1 << 32)) {
        pb_set_error("pb_chain_counts: too many blocks / work items for one call"); return PB_EINVAL;
    }
    if (!workspace || workspace_bytes < pb_chain_counts_workspace_bytes(layout->total_bins, n_blocks, n_chains)) {
        pb_set_error("pb_chain_counts: workspace too small"); return PB_ENOSPACE;
    }
    if (n_chains == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    PbLayoutDev lay{layout->chrom_len, layout->chrom_bin_off, layout->n_chrom};
    const long long n_cells = layout->total_bins >> kIndexShift;
    char *w = (char *)workspace;
    long long *index = (long long *)w;                          w += (size_t)(n_cells + 2) * 8;
    long long *slices = (long long *)w;                         w += (size_t)n_blocks * 16;
    uint32_t *items = (uint32_t *)w;                            w += ((size_t)n_blocks + 1) * 4;
    uint32_t *item_off = (uint32_t *)w;                         w += ((size_t)n_blocks + 1) * 4;
    uint32_t *part = (uint32_t *)w;                             w += (size_t)pb_scan_part_entries(n_blocks + 1) * 4;
    w = (char *)(((uintptr_t)w + 15) & ~(uintptr_t)15);
    unsigned long long *counts = (unsigned long long *)w;                     w += (size_t)n_chains * 8;
    w = (char *)(((uintptr_t)w + 15) & ~(uintptr_t)15);
    int16_t *site_tab = (int16_t *)w;
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask_bits);
    PB_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)n_chains * 8, stream));
    const long long index_threads = n_cells + 1 > 3 * kSiteKeys ? n_cells + 1 : 3 * kSiteKeys;
    pb_read_index_kernel<<<(unsigned)((index_threads + 255) / 256), 256, 0, stream>>>(b, lay, n_cells, index, r, site_tab);
    if (n_blocks > 0) {
        // 4 warps per CTA: a block's first item is anything from 0 to 2048 reads, and a CTA holds its slot until its
        // slowest warp is done
        // PB_FIRST_READS (A/B aid): reads of a block counted by the warp that finds its slice; the rest are 2048-read items
        const char *env_f = getenv("PB_FIRST_READS");
        const char *env_i = getenv("PB_ITEM_READS");
        int item_reads = env_i ? atoi(env_i) : kItemReads;
        if (item_reads < 128 || item_reads > (1 << 20)) item_reads = kItemReads;
        int first_reads = env_f ? atoi(env_f) : item_reads;
        if (first_reads < 0 || first_reads > (1 << 20)) first_reads = item_reads;
        pb_chain_first_items_kernel<<<(unsigned)((n_blocks * 32 + 127) / 128), 128, 0, stream>>>(
            b, r, lay, index, bstart, bend, block_chain, block_pos, block_plane, n_blocks, mw, mask_off, bin_begin, bin_end,
            layout->total_bins, slices, items, counts, reinterpret_cast<unsigned long long *>(stats), site_tab, first_reads, item_reads);
        int rc = pb_launch_exclusive_scan_u32(items, item_off, part, n_blocks, stream);
        if (rc) return rc;
        int sms = 148;
        pb_sm_count(&sms);
        pb_chain_items_kernel<<<(unsigned)(sms * 6), 256, 0, stream>>>(
            b, r, lay, bstart, bend, block_chain, block_pos, block_plane, n_blocks, mw, mask_off, bin_begin, bin_end,
            slices, item_off, counts, reinterpret_cast<unsigned long long *>(stats), site_tab, first_reads, item_reads);
    }
    pb_chain_totals_kernel<<<(unsigned)((n_chains + 255) / 256), 256, 0, stream>>>(
        bstart, bend, chain_off, n_chains, mw, mask_off, nullptr, counts, sums, live_len);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
