/*
 * oracle.c — plain-C CPU restatement of plastid's read-to-coverage path over the
 * packed SoA alignment batch.  TEST INFRASTRUCTURE ONLY: linked/loaded solely by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product (plastid_b200/) never calls into this file.
 *
 * Each function is the reference's per-read loop, statement for statement, with
 * the Python list `read.positions` replaced by an index->position walk over the
 * read's aligned blocks (M/=/X runs; pysam 0.19 get_reference_positions
 * [3rd-party], consume table kent/src/htslib/htslib/sam.h:79-104).
 *
 * Parity: pinned for "<L>M" reads by the reference's known-answer tests
 * (plastid/test/unit/genomics/test_map_factories.py:17-200, transcribed in
 * tests/test_oracle_kat.py, which also cross-checks this file against
 * oracle/pyoracle.py).  Non-M CIGAR ops: PARITY UNPINNED against pysam (absent).
 *
 * SoA schema (shared with include/plastid_b200.h):
 *   ref_start int32[N]; meta uint32[N] = L | reverse<<16 | drop<<17 | n_blocks<<24;
 *   blk_off uint32[N+1] (NULL => every read is one block [start,start+L));
 *   blk int32[B][2] = {start relative to ref_start, length}, listed only for
 *   reads with n_blocks > 1.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define META_L(m)    ((int)((m) & 0xFFFFu))
#define META_REV(m)  ((int)(((m) >> 16) & 1u))
#define META_DROP(m) ((int)(((m) >> 17) & 1u))
#define META_NBLK(m) ((int)((m) >> 24))

enum { STRAND_PLUS = 1, STRAND_MINUS = 2, STRAND_ANY = 3 }; /* c_common.pxd:1-6 */
enum { RULE_FIVEPRIME = 0, RULE_THREEPRIME = 1, RULE_VARIABLE = 2 };

typedef struct {
    const int32_t  *ref_start;
    const uint32_t *meta;
    const uint32_t *blk_off;
    const int32_t  *blk;
} or_batch;

/* positions[idx] for read i (0 <= idx < L) */
static inline int64_t or_position(const or_batch *b, int64_t i, int idx)
{
    uint32_t m = b->meta[i];
    int64_t s = b->ref_start[i];
    if (META_NBLK(m) <= 1 || !b->blk_off) return s + idx;
    uint32_t k0 = b->blk_off[i], k1 = b->blk_off[i + 1];
    for (uint32_t k = k0; k < k1; ++k) {
        int len = b->blk[2 * k + 1];
        if (idx < len) return s + b->blk[2 * k] + idx;
        idx -= len;
    }
    return -1; /* unreachable for consistent batches */
}

/* genome_array.py:811-820: strand filter on is_reverse, then user filters; only
 * the size filter (map_factories.pyx:837-839) and a host keep-mask are lowered. */
static inline int or_passes(uint32_t m, int strand, int size_min, int size_max)
{
    if (META_DROP(m)) return 0;
    if (strand == STRAND_PLUS && META_REV(m)) return 0;
    if (strand == STRAND_MINUS && !META_REV(m)) return 0;
    int L = META_L(m);
    if (size_min > 0 && !(L >= size_min && (L <= size_max || size_max == -1))) return 0;
    return 1;
}

/* FivePrime / ThreePrime / VariableFivePrime  (map_factories.pyx:308-367, 407-466, 585-650).
 * counts: int64[seg_end-seg_start] (caller zeroes); kept: uint8[i1-i0] or NULL;
 * dropped[0] += reads skipped for length (the do_warn paths), dropped[1] = last such length. */
int or_map_point(const int32_t *ref_start, const uint32_t *meta, const uint32_t *blk_off, const int32_t *blk,
                 int64_t i0, int64_t i1, int rule, int offset,
                 const int32_t *lut_fw, const int32_t *lut_rc,
                 int size_min, int size_max, int strand,
                 int64_t seg_start, int64_t seg_end,
                 int64_t *counts, uint8_t *kept, int64_t *dropped)
{
    or_batch b = { ref_start, meta, blk_off, blk };
    const int32_t *lut = (strand == STRAND_MINUS) ? lut_rc : lut_fw;       /* :625-626 */
    for (int64_t i = i0; i < i1; ++i) {
        uint32_t m = meta[i];
        if (kept) kept[i - i0] = 0;
        if (!or_passes(m, strand, size_min, size_max)) continue;
        int L = META_L(m), idx;
        if (rule == RULE_VARIABLE) {
            int off = lut[L];                                              /* :631 */
            if (off == -1) { dropped[0]++; dropped[1] = L; continue; }      /* :633-636 */
            idx = off;
        } else {
            if (offset >= L) { dropped[0]++; dropped[1] = L; continue; }    /* :351-353 / :450-452 */
            int from_left = (rule == RULE_FIVEPRIME) ? (strand != STRAND_MINUS)   /* :345-346 */
                                                     : (strand == STRAND_MINUS);  /* :444-445 */
            idx = from_left ? offset : L - 1 - offset;
        }
        int64_t p = or_position(&b, i, idx);
        if (p >= seg_start && p < seg_end) {
            counts[p - seg_start] += 1;
            if (kept) kept[i - i0] = 1;
        }
    }
    return 0;
}

/* CenterMapFactory.__call__ (map_factories.pyx:200-265): fp64, read-order accumulation. */
int or_map_center(const int32_t *ref_start, const uint32_t *meta, const uint32_t *blk_off, const int32_t *blk,
                  int64_t i0, int64_t i1, int nibble,
                  int size_min, int size_max, int strand,
                  int64_t seg_start, int64_t seg_end,
                  double *counts, uint8_t *kept, int64_t *dropped)
{
    or_batch b = { ref_start, meta, blk_off, blk };
    int64_t n = seg_end - seg_start;
    for (int64_t i = i0; i < i1; ++i) {
        uint32_t m = meta[i];
        if (kept) kept[i - i0] = 0;
        if (!or_passes(m, strand, size_min, size_max)) continue;
        int L = META_L(m);
        int map_length = L - 2 * nibble;                                   /* :245 */
        if (map_length < 0) { dropped[0]++; dropped[1] = L; continue; }    /* :246-248 */
        if (map_length > 0) {
            double val = 1.0 / map_length;                                 /* :250 */
            if (META_NBLK(m) <= 1 || !blk_off) {
                int64_t s = ref_start[i];
                for (int k = nibble; k < L - nibble; ++k) {                /* :251-254 */
                    int64_t c = s + k - seg_start;
                    if (c >= 0 && c < n) counts[c] += val;
                }
            } else {
                for (int k = nibble; k < L - nibble; ++k) {
                    int64_t c = or_position(&b, i, k) - seg_start;
                    if (c >= 0 && c < n) counts[c] += val;
                }
            }
            if (kept) kept[i - i0] = 1;                                    /* :256 */
        }
    }
    return 0;
}

/* StratifiedVariableFivePrimeMapFactory.__call__ (map_factories.pyx:724-780):
 * counts int64[(max-min+1) * n]; quirk kept: LUT value -1 indexes positions[-1]. */
int or_map_stratified(const int32_t *ref_start, const uint32_t *meta, const uint32_t *blk_off, const int32_t *blk,
                      int64_t i0, int64_t i1, const int32_t *lut_fw, const int32_t *lut_rc,
                      int min_len, int max_len,
                      int size_min, int size_max, int strand,
                      int64_t seg_start, int64_t seg_end,
                      int64_t *counts, uint8_t *kept)
{
    or_batch b = { ref_start, meta, blk_off, blk };
    const int32_t *lut = (strand == STRAND_MINUS) ? lut_rc : lut_fw;       /* :765-766 */
    int64_t n = seg_end - seg_start;
    for (int64_t i = i0; i < i1; ++i) {
        uint32_t m = meta[i];
        if (kept) kept[i - i0] = 0;
        if (!or_passes(m, strand, size_min, size_max)) continue;
        int L = META_L(m);
        if (L >= min_len && L <= max_len) {                                /* :771 */
            int off = lut[L];
            if (off < 0) off += L;                                         /* python negative index */
            int64_t p = or_position(&b, i, off);
            if (p >= seg_start && p < seg_end) {
                counts[(int64_t)(L - min_len) * n + (p - seg_start)] += 1;
                if (kept) kept[i - i0] = 1;
            }
        }
    }
    return 0;
}

/* SegmentChain.get_masked_counts + nansum / masked_length as used by
 * counts_in_region.py:113-125 and cs.py:705-711, over a dense per-strand vector.
 * chain c covers blocks [chain_off[c], chain_off[c+1]); block = [bstart,bend) in the
 * vector's own coordinates; mask bit j of chain c (genomic order along the chain) at
 * mask_bits[(mask_off[c] + j) >> 3] bit ((mask_off[c]+j)&7).  Sums are sequential in
 * genomic order (python sum / nansum over float64).  */
int or_region_sums_u32(const uint32_t *vec, const int64_t *bstart, const int64_t *bend,
                       const int64_t *chain_off, int64_t n_chains,
                       const uint8_t *mask_bits, const int64_t *mask_off,
                       double *sums, int64_t *masked_len)
{
    for (int64_t c = 0; c < n_chains; ++c) {
        double acc = 0.0;
        int64_t j = 0, keep = 0;
        for (int64_t k = chain_off[c]; k < chain_off[c + 1]; ++k) {
            for (int64_t p = bstart[k]; p < bend[k]; ++p, ++j) {
                int masked = 0;
                if (mask_bits) {
                    int64_t bit = mask_off[c] + j;
                    masked = (mask_bits[bit >> 3] >> (bit & 7)) & 1;
                }
                if (!masked) { acc += (double)vec[p]; keep++; }
            }
        }
        sums[c] = acc;
        masked_len[c] = keep;
    }
    return 0;
}

int or_region_sums_f64(const double *vec, const int64_t *bstart, const int64_t *bend,
                       const int64_t *chain_off, int64_t n_chains,
                       const uint8_t *mask_bits, const int64_t *mask_off,
                       double *sums, int64_t *masked_len)
{
    for (int64_t c = 0; c < n_chains; ++c) {
        double acc = 0.0;
        int64_t j = 0, keep = 0;
        for (int64_t k = chain_off[c]; k < chain_off[c + 1]; ++k) {
            for (int64_t p = bstart[k]; p < bend[k]; ++p, ++j) {
                int masked = 0;
                if (mask_bits) {
                    int64_t bit = mask_off[c] + j;
                    masked = (mask_bits[bit >> 3] >> (bit & 7)) & 1;
                }
                if (!masked) { acc += vec[p]; keep++; }
            }
        }
        sums[c] = acc;
        masked_len[c] = keep;
    }
    return 0;
}

/* The same over the int64 vector the reference's point rules return (map_factories.pyx:334: numpy.zeros(..., dtype=int)),
 * so that a caller holding such a vector does not have to narrow it first. */
int or_region_sums_i64(const int64_t *vec, const int64_t *bstart, const int64_t *bend,
                       const int64_t *chain_off, int64_t n_chains,
                       const uint8_t *mask_bits, const int64_t *mask_off,
                       double *sums, int64_t *masked_len)
{
    for (int64_t c = 0; c < n_chains; ++c) {
        double acc = 0.0;
        int64_t j = 0, keep = 0;
        for (int64_t k = chain_off[c]; k < chain_off[c + 1]; ++k) {
            for (int64_t p = bstart[k]; p < bend[k]; ++p, ++j) {
                int masked = 0;
                if (mask_bits) {
                    int64_t bit = mask_off[c] + j;
                    masked = (mask_bits[bit >> 3] >> (bit & 7)) & 1;
                }
                if (!masked) { acc += (double)vec[p]; keep++; }
            }
        }
        sums[c] = acc;
        masked_len[c] = keep;
    }
    return 0;
}
