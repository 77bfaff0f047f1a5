// pb_misc.cu — the per-segment operator, the aligned-length histogram and the wire16 expansion.
#include "pb_tiles.cuh"

namespace {

// ----------------------------------------------------------------------------------------
// the operator on one segment (global atomics; small inputs)
// ----------------------------------------------------------------------------------------
__global__ void pb_segment_kernel(PbReads b, PbRuleDev r, int64_t i0, int64_t i1, int strand, int flags,
                                  int64_t seg_start, int64_t seg_end,
                                  unsigned long long *counts_i, double *counts_f,
                                  uint8_t *__restrict__ kept, unsigned long long *__restrict__ stats)
{
    const int64_t n = seg_end - seg_start;
    const bool rq = (strand == PB_PLANE_MINUS);
    for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = __ldg(b.ref_start + i);
        const uint32_t m = __ldg(b.meta + i);
        uint8_t keep = 0;
        const bool rev = PB_META_REV(m);
        bool pass = pb_passes(m, r.size_min, r.size_max);
        if ((flags & PB_SEG_FILTER_STRAND) && strand == PB_PLANE_PLUS && rev) pass = false;    // genome_array.py:811-815
        if ((flags & PB_SEG_FILTER_STRAND) && strand == PB_PLANE_MINUS && !rev) pass = false;
        if (pass && (flags & PB_SEG_FETCH_OVERLAP)) {
            // AlignmentFile.fetch(chrom, start, end): only reads whose reference span overlaps the segment
            int64_t span = PB_META_L(m);
            if (PB_META_NBLK(m) > 1 && b.blk_off != nullptr) {
                const int2 last = __ldg(b.blk + (__ldg(b.blk_off + i + 1) - 1));
                span = (int64_t)last.x + last.y;
            }
            if (!((int64_t)s < seg_end && (int64_t)s + span > seg_start)) pass = false;
        }
        if (pass) {
            const int L = PB_META_L(m);
            const int sidx = strand == PB_PLANE_PLUS ? PB_STAT_DROPPED_PLUS
                           : strand == PB_PLANE_MINUS ? PB_STAT_DROPPED_MINUS : PB_STAT_DROPPED_ANY;
            if (r.kind == PB_RULE_CENTER) {
                const int nib = r.param, map_len = L - 2 * nib;
                if (map_len < 0) {
                    atomicAdd(&stats[sidx], 1ull);
                    stats[PB_STAT_DROPPED_LEN] = L;
                } else if (map_len > 0) {
                    const double v = 1.0 / map_len;
                    for (int k = nib; k < L - nib; ++k) {
                        const int64_t cpos = pb_position(b, i, s, m, k) - seg_start;
                        if (cpos >= 0 && cpos < n) atomicAdd(&counts_f[cpos], v);
                    }
                    keep = 1;
                }
            } else if (r.kind == PB_RULE_STRATIFIED) {
                if (L >= r.strat_min && L <= r.strat_max && L < PB_LUT_SIZE) {
                    int off = rq ? __ldg(r.lut_rc + L) : __ldg(r.lut_fw + L);
                    if (off < 0) off += L;  // map_factories.pyx:773-774: no BAD_OFFSET test, python index -1
                    const int64_t p = pb_position(b, i, s, m, off);
                    if (p >= seg_start && p < seg_end) {
                        atomicAdd(&counts_i[(int64_t)(L - r.strat_min) * n + (p - seg_start)], 1ull);
                        keep = 1;
                    }
                }
            } else {
                const int idx = pb_rule_index(r, L, rq);
                if (idx < 0) {
                    atomicAdd(&stats[sidx], 1ull);
                    stats[PB_STAT_DROPPED_LEN] = L;
                } else {
                    const int64_t p = pb_position(b, i, s, m, idx);
                    if (p >= seg_start && p < seg_end) {
                        atomicAdd(&counts_i[p - seg_start], 1ull);
                        keep = 1;
                    }
                }
            }
        }
        if (kept) kept[i - i0] = keep;
    }
}

__global__ void pb_length_hist_kernel(PbReads b, PbRuleDev r, int strand, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[1024];
    for (int j = threadIdx.x; j < 1024; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.n_reads;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t m = __ldg(b.meta + i);
        if (!pb_passes(m, r.size_min, r.size_max)) continue;
        const bool rev = PB_META_REV(m);
        if (strand == PB_PLANE_PLUS && rev) continue;
        if (strand == PB_PLANE_MINUS && !rev) continue;
        const int L = PB_META_L(m);
        if (L < 1024) atomicAdd(&sh[L], 1u); else atomicAdd(&hist[L], 1ull);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 1024; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], (unsigned long long)sh[j]);
}

// ----------------------------------------------------------------------------------------
// wire16: compact host format of an unspliced batch (4 B per read) -> the SoA the kernels stream
// ----------------------------------------------------------------------------------------
// Reads are sorted, so within one 65536-position segment of a chromosome the start needs 16 bits;
// seg_off[s] is the first read of segment s, seg_base[s] the chromosome coordinate of its first
// position.  One warp expands 1024 consecutive reads: one binary search for the chunk's segment,
// then every lane walks forward (segments are crossed rarely).
__global__ void __launch_bounds__(256)
pb_unpack_wire16_kernel(const uint16_t *__restrict__ start_lo, const uint16_t *__restrict__ meta16,
                        const int64_t *__restrict__ seg_off, const int32_t *__restrict__ seg_base,
                        int64_t n_seg, int64_t read_begin, int64_t n_reads, int32_t *__restrict__ ref_start,
                        uint32_t *__restrict__ meta)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t chunk0 = read_begin + warp * 1024;
    if (chunk0 >= n_reads) return;
    int64_t lo = 0, hi = n_seg;       // last segment with seg_off[s] <= chunk0
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= chunk0) lo = mid; else hi = mid;
    }
    int64_t seg = lo;
    int64_t seg_end = __ldg(seg_off + seg + 1);
    int32_t base = __ldg(seg_base + seg);
    const int64_t chunk1 = chunk0 + 1024 < n_reads ? chunk0 + 1024 : n_reads;
    for (int64_t i = chunk0 + lane; i < chunk1; i += 32) {
        while (i >= seg_end) {        // empty segments are skipped too
            ++seg;
            seg_end = __ldg(seg_off + seg + 1);
            base = __ldg(seg_base + seg);
        }
        const uint32_t m = __ldg(meta16 + i);
        ref_start[i] = base + (int32_t)__ldg(start_lo + i);
        // L (14 bits) | reverse | drop  ->  L | reverse<<16 | drop<<17 | n_blocks(=1)<<24
        meta[i] = (m & 0x3fffu) | (((m >> 14) & 1u) << 16) | (((m >> 15) & 1u) << 17) | (1u << 24);
    }
}

// ----------------------------------------------------------------------------------------
// delta8: 2-byte-per-read host format of a sorted unspliced batch -> the SoA
// ----------------------------------------------------------------------------------------
// Reads are coordinate-sorted, so consecutive starts differ by little: per read one byte of start
// delta and one byte indexing a 255-entry dictionary of meta words.  Reads are grouped in blocks of
// 128; blk_base[B] is the start of the block's first read.  dstart == 255 marks an exception (delta
// >= 255, first read of a chromosome, meta word not in the dictionary): start and meta come from
// exc_start / exc_meta at ordinal blk_exc_off[B] + (exceptions before it in the block).
// One warp expands one block, 4 consecutive reads per lane: ballots give the exception ordinals, a
// segmented warp scan (exceptions reset the running start) the absolute starts; 16-byte stores.
__global__ void __launch_bounds__(256)
pb_unpack_delta8_kernel(const uint32_t *__restrict__ dstart4, const uint32_t *__restrict__ code4,
                        const int32_t *__restrict__ blk_base, const uint32_t *__restrict__ blk_exc_off,
                        const int32_t *__restrict__ exc_start, const uint32_t *__restrict__ exc_meta,
                        const uint32_t *__restrict__ dict, int64_t n_reads, int64_t blk_begin, int64_t blk_end,
                        int32_t *__restrict__ ref_start, uint32_t *__restrict__ meta)
{
    __shared__ uint32_t s_dict[256];
    s_dict[threadIdx.x] = __ldg(dict + threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t B = blk_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (B >= blk_end) return;
    const int64_t i0 = B * 128 + lane * 4;
    const uint32_t d4 = __ldg(dstart4 + B * 32 + lane), c4 = __ldg(code4 + B * 32 + lane);
    uint32_t d[4], ord[4];
    bool ex[4];
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t before = __ldg(blk_exc_off + B), own = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        d[j] = (d4 >> (8 * j)) & 0xffu;
        ex[j] = d[j] == 255u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) before += __popc(__ballot_sync(0xffffffffu, ex[j]) & lt);
    int32_t es[4];
    uint32_t m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ord[j] = before + own;
        own += ex[j];
        es[j] = ex[j] ? __ldg(exc_start + ord[j]) : 0;
        m[j] = ex[j] ? __ldg(exc_meta + ord[j]) : s_dict[(c4 >> (8 * j)) & 0xffu];
    }
    // lane summary (f = saw an absolute start, v = running start or sum of deltas)
    bool f = false;
    int32_t v = 0;
    if (lane == 0) { f = true; v = __ldg(blk_base + B); d[0] = ex[0] ? d[0] : 0u; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ex[j]) { f = true; v = es[j]; } else v += (int32_t)d[j];
    }
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const int32_t pv = __shfl_up_sync(0xffffffffu, v, dd);
        const bool pf = __shfl_up_sync(0xffffffffu, (int)f, dd) != 0;
        if (lane >= dd && !f) { v += pv; f = pf; }
    }
    int32_t cur = __shfl_up_sync(0xffffffffu, v, 1);   // inclusive result of the lane before = carry-in
    if (lane == 0) cur = __ldg(blk_base + B);
    int32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ex[j]) cur = es[j]; else cur += (int32_t)d[j];
        out[j] = cur;
    }
    if (i0 + 4 <= n_reads) {
        *reinterpret_cast<int4 *>(ref_start + i0) = make_int4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<uint4 *>(meta + i0) = make_uint4(m[0], m[1], m[2], m[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j < n_reads) { ref_start[i0 + j] = out[j]; meta[i0 + j] = m[j]; }
    }
}

// ----------------------------------------------------------------------------------------
// delta3: 1-byte-per-read host format (3-bit start delta + 5-bit meta-dictionary index) -> the SoA
// ----------------------------------------------------------------------------------------
// The transfer is PCIe-bound, so bytes are time.  Per read one byte: bits 0-2 = start delta 0..6, 7 = the
// delta is 7 + the next byte of the `wide` stream (wide byte 255 = exception: absolute start and meta
// word from exc_start / exc_meta); bits 3-7 = index into a 31-entry dictionary of meta words (31 is only
// written for exceptions).  Blocks of 128 reads carry an absolute base and the ordinals of their first
// wide byte and first exception.  One warp per block, 4 consecutive reads per lane: two rounds of
// ballots (wide ordinals, then exception ordinals), a segmented warp scan, 16-byte stores.
__global__ void __launch_bounds__(256)
pb_unpack_delta3_kernel(const uint32_t *__restrict__ packed4, const uint8_t *__restrict__ wide,
                        const int32_t *__restrict__ blk_base, const uint32_t *__restrict__ blk_wide_off,
                        const uint32_t *__restrict__ blk_exc_off, const int32_t *__restrict__ exc_start,
                        const uint32_t *__restrict__ exc_meta, const uint32_t *__restrict__ dict,
                        int64_t n_reads, int64_t blk_begin, int64_t blk_end,
                        int32_t *__restrict__ ref_start, uint32_t *__restrict__ meta)
{
    __shared__ uint32_t s_dict[32];
    if (threadIdx.x < 32) s_dict[threadIdx.x] = __ldg(dict + threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t B = blk_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (B >= blk_end) return;
    const int64_t i0 = B * 128 + lane * 4;
    const uint32_t p4 = __ldg(packed4 + B * 32 + lane);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t d[4], code[4];
    bool wd[4], ex[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t byte = (p4 >> (8 * j)) & 0xffu;
        d[j] = byte & 7u;
        code[j] = byte >> 3;
        wd[j] = d[j] == 7u;
    }
    uint32_t wbefore = __ldg(blk_wide_off + B), own = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) wbefore += __popc(__ballot_sync(0xffffffffu, wd[j]) & lt);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t wv = wd[j] ? (uint32_t)__ldg(wide + wbefore + own) : 0u;
        own += wd[j];
        ex[j] = wd[j] && wv == 255u;
        if (wd[j]) d[j] = 7u + wv;
    }
    uint32_t ebefore = __ldg(blk_exc_off + B);
    own = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) ebefore += __popc(__ballot_sync(0xffffffffu, ex[j]) & lt);
    int32_t es[4];
    uint32_t m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t e = ebefore + own;
        own += ex[j];
        es[j] = ex[j] ? __ldg(exc_start + e) : 0;
        m[j] = ex[j] ? __ldg(exc_meta + e) : s_dict[code[j]];
    }
    bool f = false;
    int32_t v = 0;
    if (lane == 0) { f = true; v = __ldg(blk_base + B); d[0] = ex[0] ? d[0] : 0u; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ex[j]) { f = true; v = es[j]; } else v += (int32_t)d[j];
    }
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const int32_t pv = __shfl_up_sync(0xffffffffu, v, dd);
        const bool pf = __shfl_up_sync(0xffffffffu, (int)f, dd) != 0;
        if (lane >= dd && !f) { v += pv; f = pf; }
    }
    int32_t cur = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) cur = __ldg(blk_base + B);
    int32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ex[j]) cur = es[j]; else cur += (int32_t)d[j];
        out[j] = cur;
    }
    if (i0 + 4 <= n_reads) {
        *reinterpret_cast<int4 *>(ref_start + i0) = make_int4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<uint4 *>(meta + i0) = make_uint4(m[0], m[1], m[2], m[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j < n_reads) { ref_start[i0 + j] = out[j]; meta[i0 + j] = m[j]; }
    }
}

}  // namespace

extern "C" int pb_map_segment(const pb_batch *batch, int64_t i0, int64_t i1, const pb_rule *rule, int strand,
                              int flags, int64_t seg_start, int64_t seg_end, void *counts_out, uint8_t *kept_out,
                              uint64_t *stats, void *stream_)
{
    if (!batch || !rule || !counts_out || !stats) { pb_set_error("pb_map_segment: null argument"); return PB_EINVAL; }
    if (i0 < 0 || i1 < i0 || i1 > batch->n_reads) { pb_set_error("pb_map_segment: bad read range"); return PB_EINVAL; }
    if (strand != PB_PLANE_PLUS && strand != PB_PLANE_MINUS && strand != PB_PLANE_ANY) { pb_set_error("pb_map_segment: bad strand"); return PB_EINVAL; }
    if (seg_end < seg_start) { pb_set_error("pb_map_segment: negative-length segment"); return PB_EINVAL; }
    if ((rule->kind == PB_RULE_VARIABLE || rule->kind == PB_RULE_STRATIFIED) && (!rule->lut_fw || !rule->lut_rc)) {
        pb_set_error("pb_map_segment: rule needs lut_fw/lut_rc"); return PB_EINVAL;
    }
    if (rule->kind < 0 || rule->kind > PB_RULE_STRATIFIED) { pb_set_error("pb_map_segment: unknown rule"); return PB_EINVAL; }
    if (i1 == i0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    int64_t n = i1 - i0;
    unsigned grid = (unsigned)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    pb_segment_kernel<<<grid, 256, 0, stream>>>(b, r, i0, i1, strand, flags, seg_start, seg_end,
                                                (unsigned long long *)counts_out, (double *)counts_out, kept_out,
                                                (unsigned long long *)stats);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_length_hist(const pb_batch *batch, const pb_rule *rule, int strand, uint64_t *hist, void *stream_)
{
    if (!batch || !rule || !hist) { pb_set_error("pb_length_hist: null argument"); return PB_EINVAL; }
    if (batch->n_reads == 0) return PB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    PbReads b = pb_to_dev(batch);
    PbRuleDev r = pb_to_dev(rule);
    unsigned grid = (unsigned)((batch->n_reads + 511) / 512);
    if (grid > 148 * 8) grid = 148 * 8;
    pb_length_hist_kernel<<<grid, 512, 0, stream>>>(b, r, strand, (unsigned long long *)hist);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_unpack_wire16(const uint16_t *start_lo, const uint16_t *meta16, const int64_t *seg_off,
                                const int32_t *seg_base, int64_t n_seg, int64_t read_begin, int64_t read_end,
                                int32_t *ref_start_out, uint32_t *meta_out, void *stream_)
{
    if (read_begin < 0 || read_end < read_begin || n_seg < 0) { pb_set_error("pb_unpack_wire16: bad range"); return PB_EINVAL; }
    if (read_end == read_begin) return PB_OK;
    const int64_t n_reads = read_end;
    if (!start_lo || !meta16 || !seg_off || !seg_base || !ref_start_out || !meta_out || n_seg < 1) {
        pb_set_error("pb_unpack_wire16: null argument"); return PB_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t warps = (read_end - read_begin + 1023) / 1024;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    pb_unpack_wire16_kernel<<<grid, 256, 0, stream>>>(start_lo, meta16, seg_off, seg_base, n_seg, read_begin, n_reads, ref_start_out, meta_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_unpack_delta8(const uint8_t *dstart, const uint8_t *code, const int32_t *blk_base,
                                const uint32_t *blk_exc_off, const int32_t *exc_start, const uint32_t *exc_meta,
                                const uint32_t *dict, int64_t n_reads, int64_t read_begin, int64_t read_end,
                                int32_t *ref_start_out, uint32_t *meta_out, void *stream)
{
    if (!dstart || !code || !blk_base || !blk_exc_off || !dict || !ref_start_out || !meta_out) {
        pb_set_error("pb_unpack_delta8: null argument"); return PB_EINVAL;
    }
    if (read_begin < 0 || read_end < read_begin || read_end > n_reads || (read_begin & 127)) {
        pb_set_error("pb_unpack_delta8: bad read range (begin must be a multiple of 128)"); return PB_EINVAL;
    }
    if ((((uintptr_t)dstart | (uintptr_t)code) & 3) || (((uintptr_t)ref_start_out | (uintptr_t)meta_out) & 15)) {
        pb_set_error("pb_unpack_delta8: streams must be 4-byte aligned, outputs 16-byte aligned"); return PB_EINVAL;
    }
    if (read_end == read_begin) return PB_OK;
    const int64_t b0 = read_begin >> 7, b1 = (read_end + 127) >> 7;
    const int64_t grid = (b1 - b0 + 7) / 8;   // 8 warps = 8 blocks of 128 reads per CTA
    pb_unpack_delta8_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        (const uint32_t *)dstart, (const uint32_t *)code, blk_base, blk_exc_off, exc_start, exc_meta, dict,
        read_end, b0, b1, ref_start_out, meta_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_unpack_delta3(const uint8_t *packed, const uint8_t *wide, const int32_t *blk_base,
                                const uint32_t *blk_wide_off, const uint32_t *blk_exc_off,
                                const int32_t *exc_start, const uint32_t *exc_meta, const uint32_t *dict,
                                int64_t n_reads, int64_t read_begin, int64_t read_end,
                                int32_t *ref_start_out, uint32_t *meta_out, void *stream)
{
    if (!packed || !wide || !blk_base || !blk_wide_off || !blk_exc_off || !dict || !ref_start_out || !meta_out) {
        pb_set_error("pb_unpack_delta3: null argument"); return PB_EINVAL;
    }
    if (read_begin < 0 || read_end < read_begin || read_end > n_reads || (read_begin & 127)) {
        pb_set_error("pb_unpack_delta3: bad read range (begin must be a multiple of 128)"); return PB_EINVAL;
    }
    if (((uintptr_t)packed & 3) || (((uintptr_t)ref_start_out | (uintptr_t)meta_out) & 15)) {
        pb_set_error("pb_unpack_delta3: packed stream must be 4-byte aligned, outputs 16-byte aligned"); return PB_EINVAL;
    }
    if (read_end == read_begin) return PB_OK;
    const int64_t b0 = read_begin >> 7, b1 = (read_end + 127) >> 7;
    const int64_t grid = (b1 - b0 + 7) / 8;
    pb_unpack_delta3_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        (const uint32_t *)packed, wide, blk_base, blk_wide_off, blk_exc_off, exc_start, exc_meta, dict,
        read_end, b0, b1, ref_start_out, meta_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

// ----------------------------------------------------------------------------------------
// block words: the aligned-block table of a spliced batch in 4 bytes per block (PCIe transfer format)
// ----------------------------------------------------------------------------------------
namespace {

__global__ void pb_block_counts_kernel(const uint32_t *__restrict__ meta, int64_t n, uint32_t *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nb = PB_META_NBLK(__ldg(meta + i));
    counts[i] = nb > 1 ? (uint32_t)nb : 0u;           // the block table lists multi-block reads only
}

__global__ void pb_add_u32_kernel(uint32_t *__restrict__ v, int64_t n, uint32_t add)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += add;
}

__global__ void pb_unpack_blocks_kernel(const uint32_t *__restrict__ meta, int64_t n,
                                        const uint32_t *__restrict__ bwords, int64_t n_rows,
                                        const uint32_t *__restrict__ bexc_row, const int32_t *__restrict__ bexc,
                                        int64_t n_exc, const uint32_t *__restrict__ blk_off, int2 *__restrict__ blk)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nb = PB_META_NBLK(__ldg(meta + i));
    if (nb <= 1) return;
    const uint32_t k0 = __ldg(blk_off + i);
    int32_t pos = 0;                                  // end of the previous block, relative to ref_start
    for (int j = 0; j < nb; ++j) {
        const int64_t k = (int64_t)k0 + j;
        if (k >= n_rows) return;                      // inconsistent input: never write past the table
        const uint32_t w = __ldg(bwords + k);
        int32_t gap, len;
        if (w == 0xFFFFFFFFu) {                       // does not fit 20 + 12 bits: listed by row
            int64_t lo = 0, hi = n_exc;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)__ldg(bexc_row + mid) < k) lo = mid + 1; else hi = mid; }
            if (lo >= n_exc) return;
            gap = __ldg(bexc + 2 * lo);
            len = __ldg(bexc + 2 * lo + 1);
        } else {
            len = (int32_t)(w & 0xFFFu);
            gap = (int32_t)(w >> 12);
        }
        blk[k] = make_int2(pos + gap, len);
        pos += gap + len;
    }
}

}  // namespace

extern "C" size_t pb_unpack_blocks_workspace_bytes(int64_t n_reads)
{
    if (n_reads < 0) return 0;
    return (size_t)(n_reads + pb_scan_part_entries(n_reads) + 16) * sizeof(uint32_t);
}

extern "C" int pb_unpack_blocks_range(const uint32_t *meta, int64_t n_reads, int64_t read_begin, int64_t read_end,
                                      int64_t row_base, const uint32_t *bwords, int64_t n_rows,
                                      const uint32_t *bexc_row, const int32_t *bexc, int64_t n_exc,
                                      uint32_t *blk_off_out, int32_t *blk_out, void *workspace, size_t workspace_bytes,
                                      void *stream)
{
    if (!meta || !blk_off_out || !workspace || n_reads < 0 || n_rows < 0 || n_exc < 0 ||
        (n_rows > 0 && (!bwords || !blk_out)) || (n_exc > 0 && (!bexc_row || !bexc))) {
        pb_set_error("pb_unpack_blocks: null argument or negative size"); return PB_EINVAL;
    }
    if (read_begin < 0 || read_end < read_begin || read_end > n_reads || row_base < 0 || row_base > n_rows) {
        pb_set_error("pb_unpack_blocks: bad read range or row base"); return PB_EINVAL;
    }
    const int64_t n = read_end - read_begin;
    if (workspace_bytes < pb_unpack_blocks_workspace_bytes(n)) {
        pb_set_error("pb_unpack_blocks: workspace too small"); return PB_ENOSPACE;
    }
    if (n_reads >= ((int64_t)1 << 32) || n_rows >= ((int64_t)1 << 32)) {
        pb_set_error("pb_unpack_blocks: block offsets are 32-bit"); return PB_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (read_begin == 0) PB_CUDA_CHECK(cudaMemsetAsync(blk_off_out, 0, sizeof(uint32_t), st));
        return PB_OK;
    }
    // blk_off[read_begin .. read_end] = row_base + exclusive scan of the block counts of these reads
    uint32_t *counts = (uint32_t *)workspace;
    uint32_t *part = counts + n;
    const unsigned grid = (unsigned)((n + 255) / 256);
    pb_block_counts_kernel<<<grid, 256, 0, st>>>(meta + read_begin, n, counts);
    int rc = pb_launch_exclusive_scan_u32(counts, blk_off_out + read_begin, part, n, st);
    if (rc != PB_OK) return rc;
    if (row_base > 0)
        pb_add_u32_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(blk_off_out + read_begin, n + 1, (uint32_t)row_base);
    if (n_rows > 0)
        pb_unpack_blocks_kernel<<<grid, 256, 0, st>>>(meta + read_begin, n, bwords, n_rows, bexc_row, bexc, n_exc,
                                                      blk_off_out + read_begin, (int2 *)blk_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_unpack_blocks(const uint32_t *meta, int64_t n_reads, const uint32_t *bwords, int64_t n_rows,
                                const uint32_t *bexc_row, const int32_t *bexc, int64_t n_exc,
                                uint32_t *blk_off_out, int32_t *blk_out, void *workspace, size_t workspace_bytes,
                                void *stream)
{
    return pb_unpack_blocks_range(meta, n_reads, 0, n_reads, 0, bwords, n_rows, bexc_row, bexc, n_exc, blk_off_out,
                                  blk_out, workspace, workspace_bytes, stream);
}
