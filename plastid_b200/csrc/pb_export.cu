// pb_export.cu — count vectors -> browser-track records.
//
// Reference: BAMGenomeArray.to_variable_step / to_bedgraph (plastid/genomics/genome_array.py:990-1111)
// walk every chromosome in windows of `window_size` positions and write, per window with a positive
// sum, either every non-zero position (variableStep) or every run of equal positive values (bedGraph;
// runs are cut at window boundaries).  Here both are a stream compaction over the device-resident
// vector: flag run starts / ends per bin, count per 2048-bin block, exclusive scan, fill.  A positive
// run has exactly one start and one end, so the k-th start pairs with the k-th end, and the number of
// ends before a block is the number of starts before it minus "a positive run is open across its
// first bin" — one scan serves both.
#include "pb_tiles.cuh"

namespace {

constexpr int kEThreads = 256, kEPer = 8, kEBlock = kEThreads * kEPer;   // 2048 bins per CTA step

template <typename T> __device__ __forceinline__ double pb_as_f64(T v) { return (double)v; }

// flags of bin i within a chromosome vector of n bins: bit 0 = a record starts here, bit 1 = ends here
template <typename T>
__device__ __forceinline__ unsigned pb_run_flags(const T *__restrict__ v, int64_t i, int64_t n, int64_t window, int mode, T x)
{
    if (!(x > (T)0)) return 0u;
    if (mode == 0) return 1u;                                             // variableStep: every non-zero bin
    const bool starts = (i % window == 0) || __ldg(v + i - 1) != x;       // genome_array.py:1094-1105
    const bool ends = (i + 1 == n) || ((i + 1) % window == 0) || __ldg(v + i + 1) != x;
    return (starts ? 1u : 0u) | (ends ? 2u : 0u);
}

template <typename T>
__global__ void __launch_bounds__(kEThreads)
pb_export_count_kernel(const T *__restrict__ v, int64_t n, int64_t window, int mode, uint32_t *__restrict__ counts)
{
    __shared__ uint32_t s_warp[kEThreads / 32];
    const int64_t i0 = (int64_t)blockIdx.x * kEBlock + (int64_t)threadIdx.x * kEPer;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kEPer; ++k)
        if (i0 + k < n) c += pb_run_flags(v, i0 + k, n, window, mode, __ldg(v + i0 + k)) & 1u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kEThreads / 32; ++w) t += s_warp[w];
        counts[blockIdx.x] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(kEThreads)
pb_export_fill_kernel(const T *__restrict__ v, int64_t n, int64_t window, int mode, const uint32_t *__restrict__ off,
                      int64_t n_blocks, int64_t capacity, int64_t *__restrict__ out_start,
                      int64_t *__restrict__ out_end, double *__restrict__ out_val, int64_t *__restrict__ n_out)
{
    __shared__ uint32_t s_warp[kEThreads / 32];
    const uint32_t total = off[n_blocks];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = (int64_t)total;
    if ((int64_t)total > capacity) return;                                // counting call (or buffers too small)
    const int64_t b0 = (int64_t)blockIdx.x * kEBlock, i0 = b0 + (int64_t)threadIdx.x * kEPer;
    T x[kEPer];
    unsigned f[kEPer];
    uint32_t ns = 0, ne = 0;
#pragma unroll
    for (int k = 0; k < kEPer; ++k) {
        x[k] = i0 + k < n ? __ldg(v + i0 + k) : (T)0;
        f[k] = i0 + k < n ? pb_run_flags(v, i0 + k, n, window, mode, x[k]) : 0u;
        ns += f[k] & 1u;
        ne += (f[k] >> 1) & 1u;
    }
    // block-exclusive scan of (starts | ends << 16): at most 2048 of each per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = ns | (ne << 16);
    const uint32_t own = inc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const uint32_t ex = before + inc - own;
    // a positive run open across the block's first bin has its start in an earlier block
    uint32_t open = 0;
    if (mode != 0 && b0 > 0 && b0 < n && b0 % window != 0) {
        const T a = __ldg(v + b0 - 1);
        open = (a > (T)0 && a == __ldg(v + b0)) ? 1u : 0u;
    }
    int64_t ks = (int64_t)off[blockIdx.x] + (ex & 0xffffu);
    int64_t ke = (int64_t)off[blockIdx.x] - open + (ex >> 16);
#pragma unroll
    for (int k = 0; k < kEPer; ++k) {
        if (f[k] & 1u) {
            out_start[ks] = i0 + k;
            if (mode == 0) out_val[ks] = pb_as_f64(x[k]);
            ++ks;
        }
        if (f[k] & 2u) {
            out_end[ke] = i0 + k + 1;
            out_val[ke] = pb_as_f64(x[k]);
            ++ke;
        }
    }
}

template <typename T>
int export_runs(const T *v, int64_t n, int64_t window, int mode, int64_t capacity, int64_t *out_start,
                int64_t *out_end, double *out_val, int64_t *n_out, uint32_t *counts, uint32_t *off, uint32_t *part,
                cudaStream_t stream)
{
    const int64_t n_blocks = (n + kEBlock - 1) / kEBlock;
    pb_export_count_kernel<T><<<(unsigned)n_blocks, kEThreads, 0, stream>>>(v, n, window, mode, counts);
    int rc = pb_launch_exclusive_scan_u32(counts, off, part, n_blocks, stream);
    if (rc) return rc;
    pb_export_fill_kernel<T><<<(unsigned)n_blocks, kEThreads, 0, stream>>>(v, n, window, mode, off, n_blocks, capacity,
                                                                           out_start, out_end, out_val, n_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

}  // namespace

extern "C" size_t pb_export_workspace_bytes(int64_t n_bins)
{
    const int64_t n_blocks = (n_bins + kEBlock - 1) / kEBlock;
    return (size_t)(2 * (n_blocks + 1) + pb_scan_part_entries(n_blocks) + 16) * sizeof(uint32_t);
}

extern "C" int pb_export_runs(const void *vec, int vec_dtype, int64_t n_bins, int64_t window, int mode,
                              int64_t capacity, int64_t *out_start, int64_t *out_end, double *out_val,
                              int64_t *n_out, void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!vec || !n_out || !workspace || n_bins < 0 || window <= 0 || (mode != 0 && mode != 1) ||
        (vec_dtype != 0 && vec_dtype != 1) || capacity < 0) {
        pb_set_error("pb_export_runs: bad argument"); return PB_EINVAL;
    }
    if (n_bins >= ((int64_t)1 << 32) * 1) { pb_set_error("pb_export_runs: one chromosome at a time (n_bins < 2^32)"); return PB_EINVAL; }
    if (capacity > 0 && (!out_start || !out_val || (mode == 1 && !out_end))) {
        pb_set_error("pb_export_runs: output buffers missing"); return PB_EINVAL;
    }
    if (workspace_bytes < pb_export_workspace_bytes(n_bins)) { pb_set_error("pb_export_runs: workspace too small"); return PB_ENOSPACE; }
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_bins == 0) { PB_CUDA_CHECK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), stream)); return PB_OK; }
    const int64_t n_blocks = (n_bins + kEBlock - 1) / kEBlock;
    uint32_t *counts = (uint32_t *)workspace, *off = counts + n_blocks + 1, *part = off + n_blocks + 1;
    if (vec_dtype == 0)
        return export_runs<uint32_t>((const uint32_t *)vec, n_bins, window, mode, capacity, out_start, out_end, out_val,
                                     n_out, counts, off, part, stream);
    return export_runs<double>((const double *)vec, n_bins, window, mode, capacity, out_start, out_end, out_val,
                               n_out, counts, off, part, stream);
}
