// pb_tiles.cuh — shared machinery of the "owner computes" mapping kernels (sm_100a):
// tile descriptors, deferred-block records for multi-block (spliced) reads, TMA bulk-store helpers,
// the workspace carve-up and the launch helpers implemented in pb_tiles.cu.
#pragma once
#include "pb_common.cuh"

constexpr int kStatSlots = 64;  // stats are spread over 64 slots to keep atomics off one address

struct __align__(16) PbTile {
    long long lo;   // first candidate read
    long long p0;   // chromosome coordinate of the tile's first bin
    int n;          // number of candidate reads [lo, lo+n)
    int live;       // bins of the tile that lie inside the chromosome (0..tile_bins)
    int chrom;
    int pad;
};

// One deferred contribution of a multi-block read, binned by the tile holding x (K1 of SURVEY §7:
// "CIGAR -> aligned blocks").  Point rules: x = mapped site, tag = PB_PLANE_* mask of the planes it
// counts in.  Center rule: [x,y) = one trimmed aligned interval, tag = slot | reverse << 16.
struct __align__(16) PbRec {
    int32_t x, y;
    uint32_t tag;
    uint32_t pad;
};

// Overflow job of the point kernel: candidate reads [lo, lo+n) of `tile` beyond the tile job's share.
struct __align__(16) PbJob {
    long long lo;
    long long tile;
    int n;
    int pad[3];
};

struct PbWorkspace {
    PbTile *tiles;                  // [total_bins/1024 + 1]
    unsigned long long *slots;      // [2][kStatSlots][PB_NSTATS] (second copy: scratch for repeat passes)
    unsigned long long *tile_counter;  // [0] tile queue, [1] number of overflow jobs, [2] overflow queue
    PbJob *jobs;                    // [job_capacity]
    int64_t job_capacity;
    uint32_t *rec_off;              // [n_tiles + 1] exclusive offsets of the per-tile record buckets
    uint32_t *rec_cursor;           // [n_tiles + 1] counts, then fill cursors
    uint32_t *scan_part;            // per-4096-tile partial sums of the offset scan
    PbRec *recs;                    // [n_blk]
};

size_t pb_ws_tile_bytes(int64_t total_bins);
size_t pb_ws_stat_bytes();
int pb_carve_workspace(void *base, size_t bytes, int64_t total_bins, int64_t n_blk, int64_t n_reads, PbWorkspace *ws);

PbReads pb_to_dev(const pb_batch *b);
PbRuleDev pb_to_dev(const pb_rule *r);
int pb_check_common(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes);
int pb_sm_count(int *out);

// split > 0: a tile keeps at most `split` candidate reads; the rest are appended to ws.jobs in slices
// of `split` reads (ws.tile_counter[1] counts them).
int pb_launch_tile_index(const PbReads &b, const PbLayoutDev &lay, int tile_bins, int64_t tile_begin, int64_t tile_end,
                         int64_t read_limit, int split, const PbWorkspace &ws, cudaStream_t stream);
// Bin the contributions of multi-block reads by tile (no-op when the batch has none): count,
// exclusive scan, fill.  center = 0: point-rule sites; center = 1: trimmed aligned intervals.
int pb_launch_binning(const PbReads &b, const PbRuleDev &r, const PbLayoutDev &lay, int planes, int center,
                      const int16_t *slot_of_len, int tile_bins, int64_t n_tiles, const PbWorkspace &ws,
                      cudaStream_t stream);
int pb_launch_stats_finish(const unsigned long long *slots, unsigned long long *stats, cudaStream_t stream);
void pb_timing_begin(cudaStream_t stream);
void pb_timing_end(cudaStream_t stream);

// ---- TMA bulk copies shared memory -> global (SASS: UBLKCP) -------------------------------------
__device__ __forceinline__ void pb_fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pb_bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
// global[dst] += shared[src] elementwise in fp64, performed by the copy engine at L2
__device__ __forceinline__ void pb_bulk_add_f64(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
// global[dst] += shared[src] elementwise in uint32 (SASS UBLKRED.ADD)
__device__ __forceinline__ void pb_bulk_add_u32(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


// Pull the cache lines holding the candidate reads of a tile into L2 (no register destination).
__device__ __forceinline__ void pb_prefetch_reads_l2(const PbReads &b, const PbTile &d)
{
    const long long end = d.lo + d.n;
    for (long long j = (d.lo & ~31ll) + (long long)threadIdx.x * 32; j < end; j += (long long)blockDim.x * 32) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(b.ref_start + j));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(b.meta + j));
    }
}

// per-CTA statistics: warp reduce -> one global atomic per non-zero counter and warp
__device__ __forceinline__ void pb_flush_cta_stats(unsigned long long drop_p, unsigned long long drop_m,
                                                   unsigned long long drop_a, unsigned int drop_len,
                                                   unsigned long long map_p, unsigned long long map_m,
                                                   unsigned long long map_a, unsigned long long *stat_slots)
{
    unsigned long long v[6] = {drop_p, drop_m, drop_a, map_p, map_m, map_a};
    const int idx[6] = {PB_STAT_DROPPED_PLUS, PB_STAT_DROPPED_MINUS, PB_STAT_DROPPED_ANY,
                        PB_STAT_MAPPED_PLUS, PB_STAT_MAPPED_MINUS, PB_STAT_MAPPED_ANY};
    unsigned long long *dst = stat_slots + (blockIdx.x & (kStatSlots - 1)) * PB_NSTATS;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const unsigned long long t = pb_warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(dst + idx[k], t);
    }
    const unsigned int len = __reduce_max_sync(0xffffffffu, drop_len);
    if ((threadIdx.x & 31) == 0 && len) atomicMax(dst + PB_STAT_DROPPED_LEN, (unsigned long long)len);
}
