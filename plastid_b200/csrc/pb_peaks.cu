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
