// pb_windows.cu — geometry of `metagene generate`: landmark windows of all transcripts and
// the maximal spanning window of every gene in two launches.
//
// Reference (plastid v0.6.1): window_landmark plastid/bin/metagene.py:180-239 and
// maximal_spanning_window :343-502, which per gene builds a (transcripts x window) matrix of genomic
// positions with python loops (`position_matrix`, :438-455), keeps the columns on which every row agrees
// (:466-471) and re-segments them (`positions_to_segments`, :477-479).  Here one warp owns a gene: lanes are
// window columns, the column test walks the gene's transcripts, and runs of adjacent shared positions
// become blocks through warp ballots (k-th run start pairs with the k-th run end).
#include "pb_common.cuh"

namespace {

struct TxTable {
    const int64_t *__restrict__ bstart;   // blocks, ascending per transcript (any common coordinate system)
    const int64_t *__restrict__ bend;
    const int64_t *__restrict__ bcum;     // chain coordinate of each block's first base (genomic order)
    const int64_t *__restrict__ tx_off;   // blocks of transcript t: [tx_off[t], tx_off[t+1])
    const uint8_t *__restrict__ reverse;  // 0 '+', 1 '-', 2 '.': coordinates run like '+', window columns are laid like '-'
};

__device__ __forceinline__ int64_t tx_length(const TxTable &tx, int64_t t)
{
    const int64_t b0 = __ldg(tx.tx_off + t), b1 = __ldg(tx.tx_off + t + 1);
    if (b1 <= b0) return 0;
    return __ldg(tx.bcum + b1 - 1) + (__ldg(tx.bend + b1 - 1) - __ldg(tx.bstart + b1 - 1));
}

// SegmentChain.c_get_genomic_coordinate (roitools.pyx:3055-3119) for 0 <= x < length, stranded
__device__ __forceinline__ int64_t tx_genomic(const TxTable &tx, int64_t t, int64_t x, int64_t length)
{
    if (__ldg(tx.reverse + t) == 1) x = length - 1 - x;
    int64_t lo = __ldg(tx.tx_off + t), hi = __ldg(tx.tx_off + t + 1);   // invariant: bcum[lo] <= x < bcum[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(tx.bcum + mid) <= x) lo = mid; else hi = mid;
    }
    return __ldg(tx.bstart + lo) + (x - __ldg(tx.bcum + lo));
}

// window_landmark with ref_delta = 0 (metagene.py:219-239), one thread per transcript
__global__ void pb_landmark_windows_kernel(TxTable tx, const int64_t *__restrict__ landmark, int64_t n_tx,
                                           int64_t up, int64_t down, int64_t *__restrict__ win,
                                           uint8_t *__restrict__ flags)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tx) return;
    const int64_t lm = __ldg(landmark + t);
    int64_t w_start = 0, w_end = 0, w_off = 0, ref = 0;
    uint8_t f = 0;
    if (lm >= 0) {                                         // < 0: no landmark (window_cds_start :277-278)
        const int64_t len = tx_length(tx, t);
        if (lm >= up) { w_off = 0; w_start = lm - up; } else { w_off = up - lm; w_start = 0; }
        w_end = min(len, lm + down);
        if (w_start > len) w_start = len;                  // get_subchain slices the position hash (:3213)
        if (w_end < w_start) w_end = w_start;
        if (lm == len) {                                   // landmark just past the 3' end (:232-236)
            const int64_t b0 = __ldg(tx.tx_off + t), b1 = __ldg(tx.tx_off + t + 1);
            if (b1 > b0) { ref = __ldg(tx.reverse + t) ? __ldg(tx.bstart + b0) - 1 : __ldg(tx.bend + b1 - 1); f = PB_WIN_HAS_REF; }
            else f = PB_WIN_INDEX_ERROR;
        } else if (lm > len) {
            f = PB_WIN_INDEX_ERROR;                        // get_genomic_coordinate raises IndexError (:3097)
        } else {
            ref = tx_genomic(tx, t, lm, len);
            f = PB_WIN_HAS_REF;
        }
    }
    win[4 * t + 0] = w_start;
    win[4 * t + 1] = w_end;
    win[4 * t + 2] = w_off;
    win[4 * t + 3] = ref;
    flags[t] = f;
}

template <bool FILL>
__global__ void __launch_bounds__(128)
pb_spanning_windows_kernel(TxTable tx, const int64_t *__restrict__ win, const uint8_t *__restrict__ flags,
                           const int64_t *__restrict__ grp_off, const int64_t *__restrict__ grp_tx, int64_t n_grp,
                           int32_t up, int32_t down,
                           uint8_t *__restrict__ status, int32_t *__restrict__ offset, int32_t *__restrict__ n_pos_out,
                           int32_t *__restrict__ n_blk_out, int64_t *__restrict__ refpos_out,
                           const int64_t *__restrict__ out_off, int64_t *__restrict__ out_bstart,
                           int64_t *__restrict__ out_bend)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int32_t W = up + down;
    for (int64_t g = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); g < n_grp; g += n_warps) {
        const int64_t i0 = __ldg(grp_off + g), i1 = __ldg(grp_off + g + 1);
        // every region needs a landmark, the same reference point (:465), and a non-empty window: an empty
        // or failed window leaves a nan row in position_matrix, and nan never equals nan (:469)
        bool ok = i1 > i0;
        int64_t ref0 = 0;
        uint8_t rev0 = 0;
        if (ok) {
            const int64_t tf = __ldg(grp_tx + i0);
            ref0 = __ldg(win + 4 * tf + 3);
            rev0 = __ldg(tx.reverse + tf);
            for (int64_t i = i0 + lane; i < i1; i += 32) {
                const int64_t t = __ldg(grp_tx + i);
                const uint8_t f = __ldg(flags + t);
                if (!(f & PB_WIN_HAS_REF) || (f & PB_WIN_INDEX_ERROR) || __ldg(win + 4 * t + 3) != ref0 ||
                    __ldg(tx.reverse + t) != rev0 || __ldg(win + 4 * t + 1) <= __ldg(win + 4 * t + 0))
                    ok = false;
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        int32_t n_pos = 0, n_start = 0, n_end = 0, zero_before = 0;
        bool zero_shared = false;
        int32_t n_blk_known = 0;
        int64_t obase = 0;
        if (FILL && ok) { n_blk_known = __ldg(n_blk_out + g); obase = __ldg(out_off + g); }
        if (ok) {
            bool carry_shared = false;
            int64_t carry_pos = 0;
            for (int32_t c0 = 0; c0 <= W; c0 += 32) {          // column W is a virtual, unshared column
                const int32_t c = c0 + lane;
                bool shared = c < W;
                int64_t pos = 0;
                for (int64_t i = i0; shared && i < i1; ++i) {
                    const int64_t t = __ldg(grp_tx + i);
                    const int64_t w_start = __ldg(win + 4 * t + 0), w_end = __ldg(win + 4 * t + 1);
                    const int64_t w_off = __ldg(win + 4 * t + 2);
                    // unstranded rows hold their window's positions back to front (metagene.py:450-455)
                    const int64_t x = rev0 == 2 ? w_end - 1 - ((int64_t)c - w_off) : w_start + ((int64_t)c - w_off);
                    if ((int64_t)c < w_off || (int64_t)c - w_off >= w_end - w_start) { shared = false; break; }
                    const int64_t p = tx_genomic(tx, t, x, tx_length(tx, t));
                    if (i == i0) pos = p; else if (p != pos) shared = false;
                }
                bool prev_shared = __shfl_up_sync(0xffffffffu, (int)shared, 1) != 0;
                int64_t prev_pos = __shfl_up_sync(0xffffffffu, pos, 1);
                if (lane == 0) { prev_shared = carry_shared; prev_pos = carry_pos; }
                const int64_t d = pos - prev_pos;
                const bool adj = shared && prev_shared && (d == 1 || d == -1);
                const bool is_start = shared && !adj;           // column c opens a run of adjacent positions
                const bool is_end = prev_shared && !adj;        // column c-1 closed one
                const unsigned b_shared = __ballot_sync(0xffffffffu, shared);
                const unsigned b_start = __ballot_sync(0xffffffffu, is_start);
                const unsigned b_end = __ballot_sync(0xffffffffu, is_end);
                n_pos += __popc(b_shared);
                // get_segmentchain_coordinate of the landmark in the new window (:498): columns before the zero point
                // — for an unstranded window its coordinates run left to right while the columns run right to left,
                // so it is the shared positions LEFT of the landmark that count
                zero_before += __popc(__ballot_sync(0xffffffffu, shared && (rev0 == 2 ? pos < ref0 : c < up)));
                zero_shared |= __any_sync(0xffffffffu, shared && (rev0 == 2 ? pos == ref0 : c == up));
                if (FILL) {
                    // runs come out in column order = 5'->3'; blocks are stored in genomic order
                    if (is_start) {
                        const int32_t k = n_start + __popc(b_start & lt);
                        if (k < n_blk_known) {
                            if (rev0) out_bend[obase + (n_blk_known - 1 - k)] = pos + 1;
                            else out_bstart[obase + k] = pos;
                        }
                    }
                    if (is_end) {
                        const int32_t k = n_end + __popc(b_end & lt);
                        if (k < n_blk_known) {
                            if (rev0) out_bstart[obase + (n_blk_known - 1 - k)] = prev_pos;
                            else out_bend[obase + k] = prev_pos + 1;
                        }
                    }
                }
                n_start += __popc(b_start);
                n_end += __popc(b_end);
                carry_shared = __shfl_sync(0xffffffffu, (int)shared, 31) != 0;
                carry_pos = __shfl_sync(0xffffffffu, pos, 31);
            }
        }
        if (lane == 0 && !FILL) {
            uint8_t st = PB_SPAN_NONE;
            int32_t off = 0;
            if (ok && n_pos > 0) {
                // metagene.py:495-499: the LAST region's window decides between its own offset and the
                // landmark's coordinate in the new window
                const int64_t tl = __ldg(grp_tx + i1 - 1);
                const int64_t l_start = __ldg(win + 4 * tl + 0), l_end = __ldg(win + 4 * tl + 1);
                const int64_t l_off = __ldg(win + 4 * tl + 2);
                st = PB_SPAN_WINDOW;
                if ((int64_t)up - l_off == l_end - l_start) off = (int32_t)l_off;
                else if (zero_shared) off = up - zero_before;
                else st = PB_SPAN_REF_OUTSIDE;                  // get_segmentchain_coordinate raises KeyError
            }
            status[g] = st;
            offset[g] = off;
            n_pos_out[g] = st == PB_SPAN_NONE ? 0 : n_pos;
            n_blk_out[g] = st == PB_SPAN_NONE ? 0 : n_start;
            refpos_out[g] = ref0;
        }
    }
}

}  // namespace

extern "C" int pb_landmark_windows(const int64_t *tx_bstart, const int64_t *tx_bend, const int64_t *tx_bcum,
                                   const int64_t *tx_off, const uint8_t *tx_reverse, const int64_t *tx_landmark,
                                   int64_t n_tx, int32_t flank_up, int32_t flank_down,
                                   int64_t *win_out, uint8_t *flags_out, void *stream)
{
    if (n_tx < 0 || flank_up < 0 || flank_down < 0) { pb_set_error("pb_landmark_windows: negative size"); return PB_EINVAL; }
    if (n_tx == 0) return PB_OK;
    if (!tx_bstart || !tx_bend || !tx_bcum || !tx_off || !tx_reverse || !tx_landmark || !win_out || !flags_out) {
        pb_set_error("pb_landmark_windows: NULL argument");
        return PB_EINVAL;
    }
    TxTable tx{tx_bstart, tx_bend, tx_bcum, tx_off, tx_reverse};
    const int threads = 256;
    const unsigned blocks = (unsigned)((n_tx + threads - 1) / threads);
    pb_landmark_windows_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(tx, tx_landmark, n_tx, flank_up, flank_down,
                                                                             win_out, flags_out);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_spanning_windows(const int64_t *tx_bstart, const int64_t *tx_bend, const int64_t *tx_bcum,
                                   const int64_t *tx_off, const uint8_t *tx_reverse,
                                   const int64_t *win, const uint8_t *flags,
                                   const int64_t *grp_off, const int64_t *grp_tx, int64_t n_grp,
                                   int32_t flank_up, int32_t flank_down,
                                   uint8_t *status, int32_t *offset, int32_t *n_pos, int32_t *n_blk, int64_t *refpos,
                                   const int64_t *out_off, int64_t *out_bstart, int64_t *out_bend, void *stream)
{
    if (n_grp < 0 || flank_up < 0 || flank_down < 0) { pb_set_error("pb_spanning_windows: negative size"); return PB_EINVAL; }
    if (n_grp == 0) return PB_OK;
    if (!tx_bstart || !tx_bend || !tx_bcum || !tx_off || !tx_reverse || !win || !flags || !grp_off || !grp_tx ||
        !status || !offset || !n_pos || !n_blk || !refpos) {
        pb_set_error("pb_spanning_windows: NULL argument");
        return PB_EINVAL;
    }
    const bool fill = out_off != nullptr;
    if (fill && (!out_bstart || !out_bend)) { pb_set_error("pb_spanning_windows: out_off without block buffers"); return PB_EINVAL; }
    TxTable tx{tx_bstart, tx_bend, tx_bcum, tx_off, tx_reverse};
    int dev = 0, sms = 148;
    PB_CUDA_CHECK(cudaGetDevice(&dev));
    PB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 128;                                   // 4 genes per CTA
    int64_t blocks = (n_grp + 3) / 4;
    const int64_t cap = (int64_t)sms * 16;                     // one resident wave; warps loop over genes
    if (blocks > cap) blocks = cap;
    if (fill)
        pb_spanning_windows_kernel<true><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            tx, win, flags, grp_off, grp_tx, n_grp, flank_up, flank_down, status, offset, n_pos, n_blk, refpos,
            out_off, out_bstart, out_bend);
    else
        pb_spanning_windows_kernel<false><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            tx, win, flags, grp_off, grp_tx, n_grp, flank_up, flank_down, status, offset, n_pos, n_blk, refpos,
            out_off, out_bstart, out_bend);
    PB_CUDA_CHECK(cudaGetLastError());
    return PB_OK;
}
