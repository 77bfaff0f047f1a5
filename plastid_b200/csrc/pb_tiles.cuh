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

// Overflow job: candidate reads [lo, lo+n) of `tile` beyond the tile job's share (point kernel: kind 0 only).
// Center kernels: kind 0 = candidate reads, kind 1 = binned records [lo, lo+n); hot = index of the tile's scratch
// arrays (partial difference arrays the jobs reduce into, read back by the tile's own CTA).
struct __align__(16) PbJob {
    long long lo;
    long long tile;
    int n;
    int kind;
    int hot;
    int pad;
};

struct PbWorkspace {
    PbTile *tiles;                  // [total_bins/1024 + 1]
    unsigned long long *slots;      // [2][kStatSlots][PB_NSTATS] (second copy: scratch for repeat passes)
    unsigned long long *tile_counter;  // [0] tile queue, [1] number of overflow jobs, [2] overflow queue, [3] pile-up tiles
    PbJob *jobs;                    // [job_capacity]
    int64_t job_capacity;
    uint32_t *rec_off;              // [n_tiles + 1] exclusive offsets of the per-tile record buckets
    uint32_t *rec_cursor;           // [n_tiles + 1] counts, then fill cursors
    uint32_t *scan_part;            // per-4096-tile partial sums of the offset scan
    PbRec *recs;                    // [n_blk]
    void *hot;                      // scratch of the Center rule's pile-up tiles (integer partial difference arrays)
    size_t hot_bytes;
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
// Only records (and statistics) of tiles [tile_lo, tile_hi) are produced.
int pb_launch_binning(const PbReads &b, const PbRuleDev &r, const PbLayoutDev &lay, int planes, int center,
                      const int16_t *slot_of_len, int tile_bins, int64_t n_tiles, int64_t tile_lo, int64_t tile_hi,
                      int64_t read_begin, int64_t read_limit, const PbWorkspace &ws, cudaStream_t stream);
int pb_launch_stats_finish(const unsigned long long *slots, unsigned long long *stats, cudaStream_t stream);
// exclusive prefix sums of uint32 counts (three small launches); part needs pb_scan_part_entries(n) words
int64_t pb_scan_part_entries(int64_t n);
int pb_launch_exclusive_scan_u32(uint32_t *counts, uint32_t *off, uint32_t *part, int64_t n, cudaStream_t stream);
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
// global[dst] += shared[src] elementwise in 64-bit integers (two's complement: serves signed sums too)
__device__ __forceinline__ void pb_bulk_add_u64(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void pb_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


// ---- look-ahead tile queue of the persistent mapping kernels -----------------------------------
// A CTA always has its current and its next tile (descriptor + record-bucket bounds) in shared
// memory.  Thread 0 runs a three-stage pipeline, one stage per loop iteration, so that nothing it
// consumes was requested less than a whole tile ago: claim a tile from the global counter -> load its
// descriptor -> publish it in the 4-entry ring two iterations before it becomes current.  The chain
// of dependent round trips (atomic, descriptor, reads) is what would otherwise bound a CTA.
// (Claiming runs of consecutive tiles instead was measured slower on skewed data: hot tiles are
// neighbours, and a run lands on one CTA.)  Used by the center kernel; the point kernel keeps its
// register-held two-deep look-ahead, which measured faster there (profiles/NOTES_r01.md).
struct __align__(16) PbSlot {
    PbTile d;
    long long tile;
    uint32_t rec_lo, rec_hi;
    long long pad;
};

struct PbQueueRegs {      // meaningful in thread 0 only
    PbTile d;
    long long tile_loaded, tile_claimed;
    uint32_t rl, rh;
};

__device__ __forceinline__ void pb_queue_load(const PbTile *__restrict__ tiles, const uint32_t *__restrict__ rec_off,
                                              long long lookback, long long n_tiles, long long t, PbTile &d,
                                              uint32_t &rl, uint32_t &rh)
{
    d = PbTile{0, 0, 0, 0, 0, 0};
    rl = rh = 0;
    if (t >= n_tiles) return;
    d = tiles[t];
    if (rec_off) {
        rl = __ldg(rec_off + (t > lookback ? t - lookback : 0));
        rh = __ldg(rec_off + t + 1);
    }
}

__device__ __forceinline__ void pb_queue_publish(PbSlot &slot, const PbQueueRegs &q, const uint32_t *__restrict__ rec_off,
                                                 long long lookback, int tile_bins)
{
    slot.d = q.d;
    slot.tile = q.tile_loaded;
    uint32_t rl = q.rl;
    if (rec_off && lookback > 0) {
        // records may only be taken from earlier tiles of the same chromosome (rare: its first tiles)
        const long long first = q.tile_loaded - q.d.p0 / tile_bins;
        if (q.tile_loaded - lookback < first) rl = __ldg(rec_off + first);
    }
    slot.rec_lo = rl;
    slot.rec_hi = q.rh;
}

// thread 0, before the first barrier: claim four tiles, publish the first two, keep two in flight
__device__ __forceinline__ void pb_queue_init(PbSlot *ring, PbQueueRegs &q, const PbTile *__restrict__ tiles,
                                              const uint32_t *__restrict__ rec_off, long long lookback,
                                              long long tile_begin, long long n_tiles, int tile_bins,
                                              unsigned long long *__restrict__ counter)
{
    for (int i = 0; i < 3; ++i) {
        q.tile_loaded = tile_begin + (long long)atomicAdd(counter, 1ull);
        pb_queue_load(tiles, rec_off, lookback, n_tiles, q.tile_loaded, q.d, q.rl, q.rh);
        if (i < 2) pb_queue_publish(ring[i], q, rec_off, lookback, tile_bins);
    }
    q.tile_claimed = tile_begin + (long long)atomicAdd(counter, 1ull);
}

// thread 0, top of iteration k: publish the tile for iteration k+2, load the next descriptor, claim
__device__ __forceinline__ void pb_queue_step(PbSlot *ring, PbQueueRegs &q, int k, const PbTile *__restrict__ tiles,
                                              const uint32_t *__restrict__ rec_off, long long lookback,
                                              long long tile_begin, long long n_tiles, int tile_bins,
                                              unsigned long long *__restrict__ counter)
{
    pb_queue_publish(ring[(k + 2) & 3], q, rec_off, lookback, tile_bins);
    q.tile_loaded = q.tile_claimed;
    pb_queue_load(tiles, rec_off, lookback, n_tiles, q.tile_loaded, q.d, q.rl, q.rh);
    q.tile_claimed = tile_begin + (long long)atomicAdd(counter, 1ull);
}

// Pull the cache lines holding the candidate reads of a tile into L2 (no register destination).
__device__ __forceinline__ void pb_prefetch_reads_l2(const PbReads &b, const PbTile &d)
{
    const long long end = d.lo + d.n;
    for (long long j = (d.lo & ~31ll) + (long long)threadIdx.x * 32; j < end; j += (long long)blockDim.x * 32) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(b.ref_start + j));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(b.meta + j));
    }
}

// Pull the cache lines holding the candidate reads (and binned records) of a tile into L2.
__device__ __forceinline__ void pb_prefetch_tile_l2(const PbReads &b, const PbRec *__restrict__ recs, const PbSlot &s)
{
    const long long end = s.d.lo + s.d.n;
    for (long long j = (s.d.lo & ~31ll) + (long long)threadIdx.x * 32; j < end; j += (long long)blockDim.x * 32) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(b.ref_start + j));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(b.meta + j));
    }
    if (recs) {   // 8 records per 128-byte line
        for (long long j = (long long)(s.rec_lo & ~7u) + (long long)threadIdx.x * 8; j < (long long)s.rec_hi;
             j += (long long)blockDim.x * 8)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(recs + j));
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
