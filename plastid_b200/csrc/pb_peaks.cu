// pb_peaks.cu — measurement probes for the roofline denominators SURVEY §8(d) asks for next to
// the HBM copy bandwidth: the L2 atomic (red.global.add.u32) update rate a scatter-add design of
// the mapping path would be bound by.  Not on the product path (the tiles kernels use no global
// atomics on the planes); bench.py --workload peaks times these launches with CUDA events.
#include "pb_common.cuh"

namespace {

__device__ __forceinline__ uint64_t pb_mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// mode 0: update i hits a uniformly random bin of [0, n_bins) — the unsorted-reads worst case.
// mode 1: update i hits bin floor(i * n_bins / n_updates) + jitter in [-j, +j] — coordinate-sorted
//         reads whose mapped sites wander a read length around the sort key.
__global__ void pb_atomic_probe_kernel(uint32_t *__restrict__ bins, int64_t n_bins, int64_t n_updates,
                                       int mode, int jitter)
{
    const double scale = (double)n_bins / (double)n_updates;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_updates;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = pb_mix64((uint64_t)i + 0x9e3779b97f4a7c15ull);
        int64_t t;
        if (mode == 0) {
            t = (int64_t)(h % (uint64_t)n_bins);
        } else {
            t = (int64_t)((double)i * scale) + (int64_t)(h % (uint64_t)(2 * jitter + 1)) - jitter;
            t = t < 0 ? 0 : (t >= n_bins ? n_bins - 1 : t);
        }
        atomicAdd(bins + t, 1u);      // result unused: SASS RED.E.ADD
    }
}

}  // namespace

extern "C" int pb_atomic_probe(uint32_t *bins, int64_t n_bins, int64_t n_updates, int mode, int jitter,
                               void *stream)
{
    if (!bins || n_bins <= 0 || n_updates <= 0 || (mode != 0 && mode != 1) || jitter < 0) {
        pb_set_error("pb_atomic_probe: bad argument");
        return PB_EINVAL;
    }
    int dev = 0, sms = 148;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    PB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pb_atomic_probe_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(bins, n_bins, n_updates, mode, jitter);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// ---- gather probe: the access pattern of the region kernels with nothing else in the way -------------------------
// n_chunks warps-worth of work: chunk j is `chunk_bins` consecutive uint32 bins starting at a pseudo-random, 4-aligned
// bin of [0, n_bins); a warp reads its chunk with 16-byte loads (four in flight per lane), folds it into one word and
// writes that word.  What this launch achieves in bytes per second is the rate HBM delivers for scattered
// kilobyte-sized segments — the denominator the region sums and window gathers can be held to
// (bench.py --workload peaks; no reference counterpart, not on the product path).
namespace {

__global__ void __launch_bounds__(256)
pb_gather_probe_kernel(const uint4 *__restrict__ vec, int64_t n_groups, int chunk_groups, int64_t n_chunks,
                       uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_chunks) return;
    const int64_t g0 = (int64_t)(pb_mix64((uint64_t)j + 0x51ed270b0f4cull) % (uint64_t)(n_groups - chunk_groups));
    uint32_t acc = 0;
    for (int u0 = lane; u0 < chunk_groups; u0 += 128) {
        uint4 v[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) v[x] = u0 + 32 * x < chunk_groups ? __ldg(vec + g0 + u0 + 32 * x) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int x = 0; x < 4; ++x) acc += v[x].x + v[x].y + v[x].z + v[x].w;
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    if (lane == 0) out[j] = acc;
}

}  // namespace

extern "C" int pb_gather_probe(const uint32_t *vec, int64_t n_bins, int chunk_bins, int64_t n_chunks, uint32_t *out, void *stream)
{
    if (!vec || !out || chunk_bins < 4 || (chunk_bins & 3) || n_bins <= chunk_bins || n_chunks <= 0 || ((uintptr_t)vec & 15)) {
        pb_set_error("pb_gather_probe: bad argument");
        return PB_EINVAL;
    }
    pb_gather_probe_kernel<<<(unsigned)((n_chunks * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4 *>(vec), n_bins / 4, chunk_bins / 4, n_chunks, out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
