/*
 * plastid_b200.h — C-ABI of libplastid_b200.so: the B200 (sm_100a) implementation of
 * plastid's read-to-coverage hot path.
 *
 * Every entry point is extern "C", takes plain pointers and sizes (device pointers
 * unless a parameter says "host"), returns PB_OK (0) or a negative PB_E* code, and
 * enqueues its work on `stream` (a cudaStream_t passed as void*; NULL = default
 * stream) without synchronising unless stated.  pb_last_error() gives the text of the
 * last failure on the calling thread.  There is no CPU fallback anywhere: without a
 * CUDA device every compute entry point returns PB_ECUDA.
 *
 * Reference interfaces replaced (paths relative to the plastid source tree):
 *   pb_map_point     FivePrimeMapFactory.__call__          plastid/genomics/map_factories.pyx:308-367
 *                    ThreePrimeMapFactory.__call__         plastid/genomics/map_factories.pyx:407-466
 *                    VariableFivePrimeMapFactory.__call__  plastid/genomics/map_factories.pyx:585-650
 *                    SizeFilterFactory.__call__            plastid/genomics/map_factories.pyx:837-839
 *                    as driven per chromosome x strand by BAMGenomeArray.get_reads_and_counts
 *                                                          plastid/genomics/genome_array.py:760-832
 *   pb_map_center    CenterMapFactory.__call__             plastid/genomics/map_factories.pyx:200-265
 *   pb_map_segment   the operator call `map_fn(reads, seg)` plastid/genomics/genome_array.py:823,
 *                    incl. StratifiedVariableFivePrimeMapFactory.__call__ map_factories.pyx:724-780
 *   pb_length_hist   len(read.positions) bucketing         plastid/bin/psite.py:187-188,
 *                                                          plastid/bin/phase_by_size.py:188-189
 *   pb_region_sums   SegmentChain.get_counts / get_masked_counts + nansum / masked_length
 *                                                          plastid/genomics/roitools.pyx:3221-3315,
 *                                                          plastid/bin/counts_in_region.py:113-125,
 *                                                          plastid/bin/cs.py:705-711
 *   pb_gather_windows  the window-matrix fill loops        plastid/bin/metagene.py:895-914,
 *                                                          plastid/bin/psite.py:153-199
 *   pb_window_normalize  denominators / row selection / normalisation  plastid/bin/metagene.py:918-924
 *   pb_column_profile  median | mean | sum per column      plastid/bin/metagene.py:934-953,
 *                                                          plastid/bin/psite.py:204-234
 *   pb_mask_chains   GenomeHash.get_overlapping_features + SegmentChain.add_masks for all regions
 *                                                          plastid/genomics/genome_hash.py:259-436,
 *                                                          plastid/genomics/roitools.pyx:2213-2301
 *   pb_export_runs   BAMGenomeArray.to_variable_step / to_bedgraph  plastid/genomics/genome_array.py:990-1111
 *   pb_format_track  the per-line text writes of the same two methods (host)  :1030-1037, 1096-1111
 *   pb_phase_sums    sub-codon phase accumulation          plastid/bin/phase_by_size.py:165-235
 *   pb_stratified_windows  per-read-length window matrices plastid/bin/psite.py:176-199,
 *                                                          plastid/bin/phase_by_size.py:186-194
 *   pb_landmark_windows  window_landmark / window_cds_start / window_cds_stop
 *                                                          plastid/bin/metagene.py:180-340
 *   pb_spanning_windows  maximal_spanning_window per gene  plastid/bin/metagene.py:343-502, 702-735
 *   pb_chain_union / pb_chain_binary  position-set pooling / masking of `cs generate`
 *                                                          plastid/bin/cs.py:242-496
 *   pb_bam_*         (host) pysam.AlignmentFile open / references / lengths / mapped / fetch and
 *                    AlignedSegment.positions / is_reverse  plastid/genomics/genome_array.py:660-690, 800-815
 *   pb_inflate_raw   (host) zlib's inflate() per BGZF member underneath pysam's htslib
 *                                                          kent/src/htslib/bgzf.c:292-316
 *
 * Alignment batch (SoA, sorted by (chromosome, ref_start); what pysam hands the reference
 * as AlignedSegment.reference_start / .positions / .is_reverse):
 *   ref_start int32[N]       0-based leftmost aligned reference position
 *   meta      uint32[N]      bits 0-15  L = number of reference-aligned bases (CIGAR M/=/X),
 *                                       i.e. len(read.positions) — NOT the query length
 *                            bit  16    is_reverse
 *                            bit  17    drop (host-evaluated filter said no)
 *                            bits 24-31 n_blocks = number of maximal M/=/X runs (1..255)
 *   blk_off   uint32[N+1]    offsets into blk for every read; NULL when all reads have one block
 *   blk       int32[B][2]    {start relative to ref_start, length} of each aligned block, listed
 *                            only for reads with n_blocks > 1
 *   chrom_read_off int64[C+1]  reads of chromosome c are [off[c], off[c+1])
 *
 * Dense count planes: one vector per query strand ('+', '-', '.'), all chromosomes
 * concatenated; chromosome c starts at bin chrom_bin_off[c] (a multiple of
 * PB_LAYOUT_ALIGN) and has chrom_len[c] live bins; padding bins are written as 0.
 */
#ifndef PLASTID_B200_H
#define PLASTID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_OK        0
#define PB_EINVAL   -1   /* bad argument */
#define PB_ECUDA    -2   /* CUDA runtime error (see pb_last_error) */
#define PB_ENOSPACE -3   /* workspace too small */

#define PB_LAYOUT_ALIGN 16384   /* chrom_bin_off[] granularity, bins */
#define PB_LUT_SIZE     10000   /* VariableFivePrimeMapFactory LUT length, map_factories.pxd:11-12 */
#define PB_BAD_OFFSET   (-1)
#define PB_MAX_BLOCKS   255     /* aligned blocks per read (meta bits 24-31) */

/* query-strand planes (bit mask) — c_common.pxd:1-6 strand enum */
#define PB_PLANE_PLUS  1
#define PB_PLANE_MINUS 2
#define PB_PLANE_ANY   4

/* mapping rules */
#define PB_RULE_FIVEPRIME  0
#define PB_RULE_THREEPRIME 1
#define PB_RULE_VARIABLE   2
#define PB_RULE_CENTER     3
#define PB_RULE_STRATIFIED 4

/* indices into the uint64 stats[PB_NSTATS] vector every mapping call accumulates into */
#define PB_STAT_DROPPED_PLUS  0   /* reads the rule could not place (do_warn paths), '+' query */
#define PB_STAT_DROPPED_MINUS 1
#define PB_STAT_DROPPED_ANY   2
#define PB_STAT_DROPPED_LEN   3   /* one aligned length that was dropped (for the warning text) */
#define PB_STAT_MAPPED_PLUS   4   /* reads whose site landed in a live bin, per plane */
#define PB_STAT_MAPPED_MINUS  5
#define PB_STAT_MAPPED_ANY    6
#define PB_NSTATS             8

typedef struct pb_batch {
    int64_t         n_reads;
    const int32_t  *ref_start;       /* device int32[n_reads] */
    const uint32_t *meta;            /* device uint32[n_reads] */
    const uint32_t *blk_off;         /* device uint32[n_reads+1] or NULL */
    const int32_t  *blk;             /* device int32[B][2] or NULL */
    const int64_t  *chrom_read_off;  /* device int64[n_chrom+1] */
    int32_t         n_chrom;
    int32_t         max_span;        /* >= max over reads of (reference_end - ref_start) */
    int64_t         n_blk;           /* rows of blk (0 when blk is NULL) */
    int32_t         max_block_len;   /* >= longest aligned block (= longest L when blk is NULL) */
    int32_t         reserved;
} pb_batch;

typedef struct pb_layout {
    int32_t         n_chrom;
    int32_t         reserved;
    const int64_t  *chrom_len;       /* device int64[n_chrom] */
    const int64_t  *chrom_bin_off;   /* device int64[n_chrom+1], multiples of PB_LAYOUT_ALIGN */
    int64_t         total_bins;      /* == host copy of chrom_bin_off[n_chrom] */
} pb_layout;

typedef struct pb_rule {
    int32_t         kind;            /* PB_RULE_* */
    int32_t         param;           /* offset (5'/3') or nibble (center) */
    const int32_t  *lut_fw;          /* device int32[PB_LUT_SIZE]  (variable / stratified) */
    const int32_t  *lut_rc;          /* device int32[PB_LUT_SIZE] */
    int32_t         size_min;        /* SizeFilterFactory; 0 = no size filter */
    int32_t         size_max;        /* -1 = no upper bound */
    int32_t         strat_min;       /* stratified: first / last length row */
    int32_t         strat_max;
} pb_rule;

/* ---- host side: BAM -> SoA batch (no CUDA involved) -----------------------------------------------
 * Streams a coordinate-sorted BAM once (BGZF inflate on `n_threads` host threads, 0 = all), keeps
 * every mapped record as one batch row (CIGAR M/=/X runs merged into aligned blocks) and skips
 * records without reference / with the unmapped flag.  Replaces pysam's AlignmentFile / fetch /
 * AlignedSegment.positions on the way into the path (plastid/genomics/genome_array.py:660-690,
 * 800-809).  Every member's CRC32 is verified (PB_BAM_NOCRC=1 skips it; the reference's vendored htslib 1.3 never
 * checks it on read, kent/src/htslib/bgzf.c:292-316).  pb_bam_n_mapped is what `bamfile.mapped` reports (genome_array.py:690).  pb_bam_copy
 * writes the decoded arrays into caller-owned HOST buffers (e.g. pinned): ref_start int32[n_reads],
 * meta uint32[n_reads], chrom_read_off int64[n_ref+1], and — when pb_bam_n_blk > 0 — blk_off
 * uint32[n_reads+1] and blk int32[n_blk][2]. */
typedef struct pb_bam pb_bam;
int pb_bam_open(const char *path, pb_bam **out);
int pb_bam_decode(pb_bam *h, int n_threads);
int pb_bam_n_ref(const pb_bam *h);
const char *pb_bam_ref_name(const pb_bam *h, int i);
int64_t pb_bam_ref_len(const pb_bam *h, int i);
int64_t pb_bam_n_reads(const pb_bam *h);
int64_t pb_bam_n_blk(const pb_bam *h);
int64_t pb_bam_n_mapped(const pb_bam *h);
int64_t pb_bam_n_skipped(const pb_bam *h);
int32_t pb_bam_max_span(const pb_bam *h);
int pb_bam_copy(const pb_bam *h, int32_t *ref_start, uint32_t *meta, uint32_t *blk_off, int32_t *blk,
                int64_t *chrom_read_off);
void pb_bam_close(pb_bam *h);

/* Indexed region access (a .bai beside the file): replaces pysam's AlignmentFile.fetch(reference, start, end) as
 * BAMGenomeArray.get_reads_and_counts calls it (plastid/genomics/genome_array.py:800-809) and the index statistic
 * behind `bamfile.mapped` (:690).  pb_bai_open parses the index (SAM/BAM specification 5.2); pb_bai_mapped gives the
 * mapped-record count of one reference (ref >= 0) or of the file (ref < 0), -1 when the index holds no statistics.
 * pb_bam_read_header fills the reference names / lengths without touching a record.  pb_bam_fetch makes the records
 * of reference `tid` whose span overlaps [beg, end) the handle's batch (pb_bam_n_reads / pb_bam_copy as after
 * pb_bam_decode): only the BGZF members the index points at are read and inflated. */
typedef struct pb_bai pb_bai;
int pb_bai_open(const char *path, pb_bai **out);
void pb_bai_close(pb_bai *idx);
int pb_bai_n_ref(const pb_bai *idx);
int64_t pb_bai_mapped(const pb_bai *idx, int ref);
int pb_bam_read_header(pb_bam *h);
int pb_bam_fetch(pb_bam *h, const pb_bai *idx, int tid, int64_t beg, int64_t end);
/* `samtools index` (what pysam needs before the reference can fetch): writes the .bai of a coordinate-sorted BAM. */
int pb_bam_build_index(const char *bam_path, const char *bai_path);

/* The decoder's own raw-DEFLATE (RFC 1951) inflater, one BGZF member at a time: `in_len` compressed bytes
 * -> exactly `out_len` bytes (the member's ISIZE).  0 on success, -1 on malformed / truncated data or a
 * size mismatch; never writes outside [out, out + out_len).  Stands where htslib's bgzf.c calls zlib's
 * inflate() underneath pysam (kent/src/htslib/bgzf.c:292-316, `inflate_block`).  PB_BAM_ZLIB=1 in the environment
 * makes pb_bam_decode use zlib instead (A/B and cross-checks). */
int pb_inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

const char *pb_version(void);
const char *pb_last_error(void);
int pb_device_count(void);

/* Bytes of device workspace pb_map_point / pb_map_center need for this layout and a batch of
 * n_reads reads with n_blk block rows (0 for unspliced batches). */
size_t pb_map_workspace_bytes(int64_t total_bins, int64_t n_blk, int64_t n_reads);

/* wire16: compact 4-byte-per-read transfer format of an unspliced batch (every read one block,
 * L < 16384), what the host decoder hands over PCIe.  start_lo uint16[N] = ref_start & 0xFFFF;
 * meta16 uint16[N] = L | is_reverse<<14 | drop<<15; the chromosome axis is cut into 65536-position
 * segments, seg_off int64[n_seg+1] = first read of each segment, seg_base int32[n_seg] = chromosome
 * coordinate of its first position (all device pointers).  Expands into ref_start / meta of the
 * SoA batch above, for reads [read_begin, read_end) (a streamed upload expands chunk by chunk). */
int pb_unpack_wire16(const uint16_t *start_lo, const uint16_t *meta16, const int64_t *seg_off,
                     const int32_t *seg_base, int64_t n_seg, int64_t read_begin, int64_t read_end,
                     int32_t *ref_start_out, uint32_t *meta_out, void *stream);

/* delta8: 2-byte-per-read transfer format of a sorted unspliced batch.  Reads are grouped in blocks of
 * 128 consecutive reads.  dstart uint8[N128] = start - start of the previous read (0 for the first read of
 * a block, whose start is blk_base int32[n_blk]); code uint8[N128] = index into dict uint32[256] of
 * full meta words (N128 = N rounded up to 128).  dstart == 255 marks an exception (delta >= 255, a
 * chromosome change, a meta word outside the dictionary): its start and meta word are exc_start /
 * exc_meta at ordinal blk_exc_off[B] + (exceptions before it within block B); blk_exc_off
 * uint32[n_blk+1].  Expands reads [read_begin, read_end) (read_begin a multiple of 128) into
 * ref_start / meta of the SoA batch; every read before read_end must be < n_reads. */
int pb_unpack_delta8(const uint8_t *dstart, const uint8_t *code, const int32_t *blk_base,
                     const uint32_t *blk_exc_off, const int32_t *exc_start, const uint32_t *exc_meta,
                     const uint32_t *dict, int64_t n_reads, int64_t read_begin, int64_t read_end,
                     int32_t *ref_start_out, uint32_t *meta_out, void *stream);

/* delta3: 1-byte-per-read transfer format of a sorted unspliced batch (the upload is PCIe-bound: bytes
 * are time).  packed uint8[N128]: bits 0-2 = start delta 0..6 from the previous read (0 for the first read
 * of a 128-read block, whose start is blk_base[B]), 7 = the delta is 7 + the next byte of the `wide`
 * stream, whose value 255 marks an exception (chromosome change, delta > 261, rare meta word) resolved
 * from exc_start / exc_meta; bits 3-7 = index into dict uint32[32] of meta words (31 only with
 * exceptions).  blk_wide_off / blk_exc_off uint32[n_blk+1]: ordinals of each block's first wide byte /
 * exception.  Same contract as pb_unpack_delta8 otherwise. */
int pb_unpack_delta3(const uint8_t *packed, const uint8_t *wide, const int32_t *blk_base,
                     const uint32_t *blk_wide_off, const uint32_t *blk_exc_off,
                     const int32_t *exc_start, const uint32_t *exc_meta, const uint32_t *dict,
                     int64_t n_reads, int64_t read_begin, int64_t read_end,
                     int32_t *ref_start_out, uint32_t *meta_out, void *stream);

/* Block words: the aligned-block table of a spliced batch in 4 bytes per block, the companion of delta3 for
 * batches with multi-block reads (delta3 carries ref_start / meta of every read; meta bits 24-31 say how many
 * blocks a read has).  bwords uint32[n_rows], one word per block of every multi-block read in read order:
 * bits 0-11 = block length (1..4095), bits 12-31 = gap between the end of the read's previous block (0 for
 * its first block) and this block's start (0..1048574); 0xFFFFFFFF = the block does not fit and is listed as
 * {gap, len} in bexc int32[n_exc][2] under its row number bexc_row uint32[n_exc] (ascending).  Rebuilds
 * blk_off uint32[n_reads+1] (exclusive scan of the block counts of multi-block reads) and
 * blk int32[n_rows][2] = {start relative to ref_start, length} of the SoA batch. */
size_t pb_unpack_blocks_workspace_bytes(int64_t n_reads);
int pb_unpack_blocks(const uint32_t *meta, int64_t n_reads, const uint32_t *bwords, int64_t n_rows,
                     const uint32_t *bexc_row, const int32_t *bexc, int64_t n_exc,
                     uint32_t *blk_off_out, int32_t *blk_out, void *workspace, size_t workspace_bytes, void *stream);
/* The same for the reads [read_begin, read_end) of a batch that is still being uploaded: row_base = number of
 * block rows of the reads before read_begin (the sender knows it); writes blk_off[read_begin .. read_end] and
 * the rows of these reads.  Workspace: pb_unpack_blocks_workspace_bytes(read_end - read_begin). */
int pb_unpack_blocks_range(const uint32_t *meta, int64_t n_reads, int64_t read_begin, int64_t read_end,
                           int64_t row_base, const uint32_t *bwords, int64_t n_rows,
                           const uint32_t *bexc_row, const int32_t *bexc, int64_t n_exc,
                           uint32_t *blk_off_out, int32_t *blk_out, void *workspace, size_t workspace_bytes,
                           void *stream);

/* Host side (no CUDA): pack a sorted unspliced SoA batch (HOST arrays) into the delta3 streams above, on
 * n_threads host threads (0 = all).  Caller-owned HOST buffers sized for the worst case: packed
 * uint8[n_blk*128], wide uint8[n_reads], blk_base int32[n_blk], blk_wide_off / blk_exc_off uint32[n_blk+1],
 * exc_start int32[n_reads], exc_meta uint32[n_reads], dict32 uint32[32] (n_blk = ceil(n_reads / 128));
 * *n_wide_out / *n_exc_out receive how many wide bytes / exceptions were written. */
int pb_pack_delta3(const int32_t *ref_start, const uint32_t *meta, const int64_t *chrom_read_off,
                   int32_t n_chrom, int64_t n_reads, int n_threads,
                   uint8_t *packed, uint8_t *wide, int32_t *blk_base, uint32_t *blk_wide_off,
                   uint32_t *blk_exc_off, int32_t *exc_start, uint32_t *exc_meta, uint32_t *dict32,
                   int64_t *n_wide_out, int64_t *n_exc_out);
/* Host side (no CUDA): reads per aligned length (int64[65536], meta bits 0-15; reads with the drop bit left out) — the
 * batch metadata the Center rule derives its map-length tables from (len(read.positions) bucketing, psite.py:187-188;
 * what pb_length_hist measures on the device).  n_threads < 1: all hardware threads. */
int pb_meta_length_hist(const uint32_t *meta, int64_t n_reads, int n_threads, int64_t *hist);

/* 5' / 3' / variable-offset mapping of a whole batch into dense uint32 planes.
 * `planes` selects which of out_plus/out_minus/out_any are produced; every bin of a selected
 * plane is written (no prior memset needed).  stats: device uint64[PB_NSTATS], accumulated. */
int pb_map_point(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                 uint32_t *out_plus, uint32_t *out_minus, uint32_t *out_any,
                 uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream);

/* The same over the bins [bin_begin, bin_end) only (both multiples of PB_LAYOUT_ALIGN), reading no
 * read at or beyond read_limit: lets a host->device upload of a sorted batch be overlapped with
 * mapping — once the reads up to a position have landed, the planes up to that position can be
 * produced — and lets a position-sharded rank produce its own bin range from the reads
 * that start in it plus a halo.  Only bins of the range are written, so a plane pointer may be the
 * address bin 0 WOULD have for a buffer holding just [bin_begin, bin_end).  Statistics count the
 * reads / sites of the range only (they add up over disjoint ranges).  Batches with multi-block reads
 * need read_limit == n_reads. */
int pb_map_point_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                       uint32_t *out_plus, uint32_t *out_minus, uint32_t *out_any,
                       uint64_t *stats, void *workspace, size_t workspace_bytes,
                       int64_t bin_begin, int64_t bin_end, int64_t read_limit, void *stream);

/* Center mapping into dense float64 planes.  slot_of_len: device int16[65536] mapping aligned
 * length L to a slot (index into inv_m) or -1 (L-2*nibble <= 0); inv_m: device double[n_slots]
 * = 1.0/(L-2*nibble) in ascending order of map length.  Deterministic: bin = sum over slots, in
 * slot order, of (integer coverage of that slot) * inv_m[slot]. */
int pb_map_center(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                  const int16_t *slot_of_len, const double *inv_m, int n_slots,
                  double *out_plus, double *out_minus, double *out_any,
                  uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream);

/* Center mapping of a batch with MANY distinct map lengths (the reference's default CenterMapFactory() on
 * 25-35 nt ribo-seq reads has a dozen) in one pass: every read adds the integer weight
 * w_fix[slot] = round(2^shift / m) to one 64-bit difference array per plane; bin = total * 2^-shift.
 * Deterministic (integer accumulation), exact zeros; relative error against sum(1/m) is at most
 * m * 2^-(shift+1) per map length — the caller picks shift <= 62 - log2(sum over reads of 1/m) so that no
 * total overflows and keeps that error far below the 1e-6 tolerance (plastid_b200.genome_array.map_batch
 * requires <= 1e-9, else it uses pb_map_center).  Planes must be 32-byte aligned. */
int pb_map_center_fixed(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                        const int16_t *slot_of_len, const int64_t *w_fix, int n_slots, int shift,
                        double *out_plus, double *out_minus, double *out_any,
                        uint64_t *stats, void *workspace, size_t workspace_bytes, void *stream);

/* pb_map_center / pb_map_center_fixed over the bins [bin_begin, bin_end) only (multiples of PB_LAYOUT_ALIGN):
 * the Center rule for one rank of a position-sharded genome.  Same contract as
 * pb_map_point_range for plane pointers (the address bin 0 WOULD have) and statistics (reads are counted by
 * the range holding their start, so ranges add up).  A bin's value depends only on the reads covering it, so
 * ranges reproduce the whole-genome planes bit for bit.  Only the reads [read_begin, read_limit) are looked at
 * (0 and -1 = all): a streamed upload of a sorted batch maps the bins below the start of read `read_limit`
 * from the reads that have landed, and may skip the reads before read_begin when it knows that they end
 * before bin_begin (their reference span, introns included, lies below it). */
int pb_map_center_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                        const int16_t *slot_of_len, const double *inv_m, int n_slots,
                        double *out_plus, double *out_minus, double *out_any,
                        uint64_t *stats, void *workspace, size_t workspace_bytes,
                        int64_t bin_begin, int64_t bin_end, int64_t read_begin, int64_t read_limit, void *stream);
int pb_map_center_fixed_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule, int planes,
                              const int16_t *slot_of_len, const int64_t *w_fix, int n_slots, int shift,
                              double *out_plus, double *out_minus, double *out_any,
                              uint64_t *stats, void *workspace, size_t workspace_bytes,
                              int64_t bin_begin, int64_t bin_end, int64_t read_begin, int64_t read_limit, void *stream);

/* Measurement hook (bench.py roofline): while enabled, pb_map_point / pb_map_center bracket their
 * tiles kernel with CUDA events on the launch stream (up to 256 launches since the last enable);
 * pb_tiles_kernel_ms_total waits for them and returns the summed device time and launch count. */
void pb_enable_kernel_timing(int on);
int pb_tiles_kernel_ms_total(float *ms_total, int *n_launches);

/* The operator itself: map reads [i0,i1) of the batch onto one segment [seg_start,seg_end) of
 * their chromosome for query strand `strand` (PB_PLANE_*), which sets the direction the rule is
 * applied in (map_factories.pyx:345-346, 444-445, 625-626).  flags:
 *   PB_SEG_FILTER_STRAND  drop reads of the other strand first, as BAMGenomeArray.get_reads_and_counts
 *                         does before calling the rule (genome_array.py:811-815);
 *   PB_SEG_FETCH_OVERLAP  drop reads whose reference span does not overlap the segment, i.e. those
 *                         pysam's fetch(chrom, start, end) would not have returned (genome_array.py:800-809).
 * A bare `map_fn(reads, seg)` call passes 0.  counts_out (caller-zeroed): int64[n] (5'/3'/variable),
 * int64[(strat_max-strat_min+1)*n] (stratified), double[n] (center).  kept_out: uint8[i1-i0] or NULL
 * — 1 where the reference appends the read to reads_out. */
#define PB_SEG_FILTER_STRAND 1
#define PB_SEG_FETCH_OVERLAP 2
int pb_map_segment(const pb_batch *batch, int64_t i0, int64_t i1, const pb_rule *rule, int strand,
                   int flags, int64_t seg_start, int64_t seg_end, void *counts_out,
                   uint8_t *kept_out, uint64_t *stats, void *stream);

/* Histogram of aligned length L over reads passing drop/size/strand filters.
 * hist: device uint64[65536], accumulated. */
int pb_length_hist(const pb_batch *batch, const pb_rule *rule, int strand, uint64_t *hist, void *stream);

/* Masked sums over exon-block chains.  vec_dtype: 0 = uint32 planes, 1 = float64 planes.
 * planes[3]: host array of device plane pointers ('+','-','.'), each the address bin 0 WOULD have (see
 * pb_map_point_range), 16-byte aligned; chain_plane: device uint8[n_chains] in {0,1,2}.  Blocks are in global-bin
 * coordinates [bstart,bend), chain c owns blocks [chain_off[c], chain_off[c+1]); block_chain int32[B] = chain of
 * every block, block_pos int64[B] = chain position (genomic order) of its first base, block_plane uint8[B] =
 * chain_plane of its chain (so that a block's data loads depend on one round trip only).  mask_bits (or NULL): bit
 * (mask_off[c] + j) set = j-th position of chain c (genomic order) is masked; 4-byte aligned, padded to whole
 * 32-bit words.  sums: double[n_chains] over unmasked positions; live_len: int64[n_chains] unmasked length.
 *
 * [bin_begin, bin_end): the global bins the calling rank owns (position sharding; 0 .. total_bins on one
 * GPU).  Only positions inside are read and summed — the others count zero but keep their place in the chain — so
 * the sums of all ranks add up to the whole table (one all-reduce); live_len is geometry and comes out whole on
 * every rank.  Blocks lying wholly outside the layout (the part of a region beyond its chromosome's end, lowered
 * to coordinates >= total_bins) count zero the same way: the reference returns zeros there (fetch yields no reads).
 * Two launches: one warp per BLOCK writes a partial sum into the workspace (n_blocks doubles), one warp per chain adds
 * them in a fixed order (deterministic for float64 planes too). */
size_t pb_region_sums_workspace_bytes(int64_t n_blocks);
int pb_region_sums(const void *const *planes, int vec_dtype,
                   const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                   const uint8_t *chain_plane, const int32_t *block_chain, const int64_t *block_pos,
                   const uint8_t *block_plane,
                   int64_t n_chains, int64_t n_blocks, const uint8_t *mask_bits, const int64_t *mask_off,
                   int64_t bin_begin, int64_t bin_end,
                   double *sums, int64_t *live_len, void *workspace, size_t workspace_bytes, void *stream);

/* Plane-free region counts of a point rule (5' / 3' / variable): sums[c] = number of reads whose mapped site
 * lies on an unmasked position of chain c (strand-matched like BAMGenomeArray.get_reads_and_counts,
 * plastid/genomics/genome_array.py:811-815; rule direction from the chain's strand), live_len[c] = unmasked
 * length — exactly what pb_region_sums returns over the planes pb_map_point would write, without writing them
 * (table-only programs: plastid/bin/counts_in_region.py:107-125, plastid/bin/cs.py:688-714).  The reads that can
 * reach a block are a contiguous slice of the sorted batch; the work is cut into 2048-read items over all blocks
 * (expression is skewed: a few chains hold millions of reads) and taken by persistent warps.  Sites outside
 * [bin_begin, bin_end) count zero (position sharding).  n_blocks = rows of bstart / bend / block_chain / block_pos.
 * stats (or NULL): PB_STAT_DROPPED_* / _LEN are raised when a read near a chain could not be placed by the rule
 * (the reference's DataWarning paths).  Workspace: pb_chain_counts_workspace_bytes.  batch->ref_start and batch->meta
 * must be 16-byte aligned and allocated in whole 16-byte units (reads are loaded four at a time). */
size_t pb_chain_counts_workspace_bytes(int64_t total_bins, int64_t n_blocks, int64_t n_chains);
int pb_chain_counts(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                    const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                    const uint8_t *chain_plane, const int32_t *block_chain, const int64_t *block_pos,
                    const uint8_t *block_plane, int64_t n_chains, int64_t n_blocks,
                    const uint8_t *mask_bits, const int64_t *mask_off,
                    int64_t bin_begin, int64_t bin_end,
                    double *sums, int64_t *live_len, uint64_t *stats,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Mask pipeline: mask bits of every chain from one interval set, replacing the per-region
 * GenomeHash.get_overlapping_features + SegmentChain.add_masks of counts_in_region.py:114-115
 * (plastid/genomics/genome_hash.py:259-436, plastid/genomics/roitools.pyx:2213-2301).
 * mask_start/mask_end int64[M]: mask intervals in global-bin coordinates, merged and sorted within each
 * strand class; class k ('+','-','.' = 0,1,2, the chain_plane numbering) owns
 * [mask_class_off[k], mask_class_off[k+1]) (int64[4]).  For every chain c, bit (mask_off[c] + j) of
 * mask_bits is OR-ed with "the j-th chain position (genomic order) lies in a mask interval of the chain's
 * class".  mask_bits: caller-initialised (zeros, or bits from add_masks), 4-byte aligned, padded to a
 * multiple of 4 bytes; the layout pb_region_sums / pb_gather_windows / pb_stratified_windows read. */
int pb_mask_chains(const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                   const uint8_t *chain_plane, int64_t n_chains,
                   const int64_t *mask_start, const int64_t *mask_end, const int64_t *mask_class_off,
                   const int64_t *mask_off, uint8_t *mask_bits, void *stream);

/* Window matrices (metagene / psite): row r = chain r laid 5'->3' (reversed when
 * chain_reverse[r]) starting at column row_col[r] of a width-W row.  matrix: double[n*W]
 * (NaN where no chain position), maskmat: uint8[n*W] (1 = masked or uncovered).  block_chain / block_pos as for
 * pb_region_sums, chain_len int64[n_chains] = positions per chain.  Cells whose position lies outside
 * [bin_begin, bin_end) are written as 0 (not NaN), so that the matrices of all ranks of a position-sharded genome add
 * up; NaN cells (no chain position) and maskmat are identical on every rank.  One warp per exon block. */
int pb_gather_windows(const void *const *planes, int vec_dtype,
                      const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                      const uint8_t *chain_plane, const uint8_t *chain_reverse,
                      const int32_t *block_chain, const int64_t *block_pos, const uint8_t *block_plane,
                      const int64_t *chain_len,
                      const int32_t *row_col, int64_t n_chains, int64_t n_blocks, int32_t width,
                      const uint8_t *mask_bits, const int64_t *mask_off,
                      int64_t bin_begin, int64_t bin_end,
                      double *matrix, uint8_t *maskmat, void *stream);

/* Count vectors of many chains at once, ragged: chain c laid 5'->3' (reversed when chain_reverse[c]) into cells
 * [row_off[c], row_off[c] + length of c) of `values` (double) and `masked` (uint8, 1 = masked position) —
 * SegmentChain.get_masked_counts for every region of plastid/bin/get_count_vectors.py:92-104 in one launch.
 * Same range semantics as pb_gather_windows (cells of other ranks' positions are 0). */
int pb_gather_chains(const void *const *planes, int vec_dtype,
                     const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                     const uint8_t *chain_plane, const uint8_t *chain_reverse,
                     const int32_t *block_chain, const int64_t *block_pos, const uint8_t *block_plane,
                     const int64_t *chain_len,
                     const int64_t *row_off, int64_t n_chains, int64_t n_blocks,
                     const uint8_t *mask_bits, const int64_t *mask_off,
                     int64_t bin_begin, int64_t bin_end,
                     double *values, uint8_t *masked, void *stream);

/* metagene.py:918-924 on a window matrix (one warp per row): denom[r] = sum of unmasked cells in
 * columns [norm_lo,norm_hi) (NaN when every cell there is masked), row_select[r] = denom >=
 * min_counts, and (when norm_out != NULL) norm_out = matrix / denom with normmask_out = maskmat |
 * isnan | isinf. */
int pb_window_normalize(const double *matrix, const uint8_t *maskmat, int64_t n_rows, int32_t width,
                        int32_t norm_lo, int32_t norm_hi, double min_counts,
                        double *denom, uint8_t *row_select, double *norm_out, uint8_t *normmask_out,
                        void *stream);

/* metagene.py:934-953 / psite.py:213-234: per-column statistic over the unmasked cells of the
 * selected rows.  mode 0 = median (numpy.ma.median: mean of the two middle order statistics,
 * exact radix select), 1 = mean, 2 = sum (psite --aggregate).  n_regions[col] = number of cells
 * used, col_sum[col] = their sum (what a multi-GPU mean all-reduces).  Columns with no usable cell
 * give NaN (mode 0/1). */
size_t pb_column_profile_workspace_bytes(int64_t n_rows, int32_t width);
int pb_column_profile(const double *values, const uint8_t *valmask, const uint8_t *row_select,
                      int64_t n_rows, int32_t width, int mode,
                      double *profile, int64_t *n_regions, double *col_sum,
                      void *workspace, size_t workspace_bytes, void *stream);
/* The same for n_batch equally shaped matrices stacked row-wise (psite.py:176-234: one matrix per read
 * length): values/valmask [n_batch][n_rows][width], row_select [n_batch][n_rows]; profile, n_regions,
 * col_sum [n_batch][width]; workspace n_batch x pb_column_profile_workspace_bytes(n_rows, width). */
int pb_column_profile_batched(const double *values, const uint8_t *valmask, const uint8_t *row_select,
                              int32_t n_batch, int64_t n_rows, int32_t width, int mode,
                              double *profile, int64_t *n_regions, double *col_sum,
                              void *workspace, size_t workspace_bytes, void *stream);

/* psite.py:200-234 on integer count matrices (the output of pb_stratified_windows) in two launches:
 * normalisation (denominator over [norm_lo, norm_hi), unmasked cells only; rows with denominator >=
 * min_counts are selected; nan / inf quotients are masked) fused with the key extraction, then the
 * per-(matrix, column) median (mode 0) or mean (mode 1) of the normalised cells.  counts
 * uint32[n_batch][n_rows][width]; maskmat uint8[n_rows][width] when mask_shared (one position mask for
 * all matrices) else [n_batch][n_rows][width]; row_select uint8[n_batch][n_rows] (output); profile,
 * n_regions, col_sum [n_batch][width]; workspace n_batch x pb_column_profile_workspace_bytes. */
int pb_count_profiles_u32(const uint32_t *counts, const uint8_t *maskmat, int mask_shared,
                          int32_t n_batch, int64_t n_rows, int32_t width,
                          int32_t norm_lo, int32_t norm_hi, double min_counts, int mode,
                          uint8_t *row_select, double *profile, int64_t *n_regions, double *col_sum,
                          void *workspace, size_t workspace_bytes, void *stream);

/* psite.py:176-199 / phase_by_size.py:186-194 in one launch: for every window chain and every aligned
 * length in [min_len, max_len], the counts of the reads the point rule maps into the window, laid
 * 5'->3' from column row_col[c] of a width-W row (strand-matched like get_reads_and_counts; the rule
 * runs right-to-left for '-' windows).  out: uint32[(max_len-min_len+1)][n_chains][width];
 * maskmat: uint8[n_chains][width], 1 = masked or outside the chain (shared by all lengths).
 * phase_mode != 0 (phase_by_size.py:197-214): width is 3, columns are sub-codon phases of the codons
 * in the python slice [codon_front:codon_back] of the chain; row_col and maskmat are unused. */
int pb_stratified_windows(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                          int min_len, int max_len,
                          const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                          const uint8_t *chain_plane, const uint8_t *chain_reverse,
                          const int32_t *row_col, int64_t n_chains, int32_t width,
                          int phase_mode, int32_t codon_front, int32_t codon_back,
                          const uint8_t *mask_bits, const int64_t *mask_off,
                          uint32_t *out, uint8_t *maskmat, void *stream);
/* Position-sharded form: only sites whose global bin lies in [bin_begin, bin_end) are counted (a rank's batch
 * holds its own reads plus a halo; each site is owned by exactly one rank), so the matrices of all ranks add up. */
int pb_stratified_windows_range(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                                int min_len, int max_len,
                                const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                                const uint8_t *chain_plane, const uint8_t *chain_reverse,
                                const int32_t *row_col, int64_t n_chains, int32_t width,
                                int phase_mode, int32_t codon_front, int32_t codon_back,
                                const uint8_t *mask_bits, const int64_t *mask_off,
                                int64_t bin_begin, int64_t bin_end,
                                uint32_t *out, uint8_t *maskmat, void *stream);
/* The same with a caller-provided workspace (pb_stratified_windows_workspace_bytes(n_blocks), 8-byte aligned; n_blocks =
 * chain_off[n_chains]): the read slice of every exon block of the table is found by one launch over all blocks before
 * the windows are counted, instead of two dependent searches per exon inside every window's CTA. */
size_t pb_stratified_windows_workspace_bytes(int64_t n_blocks);
int pb_stratified_windows_ws(const pb_batch *batch, const pb_layout *layout, const pb_rule *rule,
                             int min_len, int max_len,
                             const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                             const uint8_t *chain_plane, const uint8_t *chain_reverse,
                             const int32_t *row_col, int64_t n_chains, int64_t n_blocks, int32_t width,
                             int phase_mode, int32_t codon_front, int32_t codon_back,
                             const uint8_t *mask_bits, const int64_t *mask_off,
                             int64_t bin_begin, int64_t bin_end,
                             uint32_t *out, uint8_t *maskmat, void *workspace, size_t workspace_bytes, void *stream);

/* phase_by_size.py:197-214: per chain, counts laid 5'->3' are cut into codons (a trailing partial
 * codon is ignored), the python slice [codon_front:codon_back] of codons is kept, and counts are
 * summed per sub-codon phase.  out: double[n_chains*3]. */
int pb_phase_sums(const void *const *planes, int vec_dtype,
                  const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                  const uint8_t *chain_plane, const uint8_t *chain_reverse, int64_t n_chains,
                  int32_t codon_front, int32_t codon_back, double *out, void *stream);
int pb_phase_sums_range(const void *const *planes, int vec_dtype,
                        const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                        const uint8_t *chain_plane, const uint8_t *chain_reverse, int64_t n_chains,
                        int32_t codon_front, int32_t codon_back, int64_t bin_begin, int64_t bin_end,
                        double *out, void *stream);

/* Track export: BAMGenomeArray.to_variable_step / to_bedgraph (plastid/genomics/genome_array.py:990-1111)
 * as a stream compaction of one chromosome's count vector vec[0..n_bins) (vec_dtype 0 = uint32, 1 =
 * float64).  mode 0 (variableStep): one record per non-zero bin — out_start = 0-based position, out_val
 * its value.  mode 1 (bedGraph): one record per run of equal positive values, runs cut at multiples of
 * `window` like the reference's window loop — out_start/out_end = [start, end), out_val the value.
 * Records come out in ascending position.  *n_out (device) receives the number of records; they are
 * written only when it is <= capacity, so a call with capacity 0 just counts. */
size_t pb_export_workspace_bytes(int64_t n_bins);
int pb_export_runs(const void *vec, int vec_dtype, int64_t n_bins, int64_t window, int mode,
                   int64_t capacity, int64_t *out_start, int64_t *out_end, double *out_val,
                   int64_t *n_out, void *workspace, size_t workspace_bytes, void *stream);

/* The text of those records (host side, no CUDA), formatted like the reference's per-line Python writes:
 * kind 0 (variableStep) `"%s\t%s\n" % (start + 1, value)` (plastid/genomics/genome_array.py:1030-1037), kind 1
 * (bedGraph) `"%s\t%s\t%s\t%s\n" % (chrom, start, end, value)` (:1096-1111).  `values` is int64[n] or — when
 * values_are_float — float64[n], printed as str() of a numpy.float64 prints them (shortest round-trip digits,
 * fixed notation for 1e-4 <= |x| < 1e16 with ".0" on integers, else exponent form with two exponent digits at least).
 * HOST pointers.  `out` has to hold pb_format_track_bound(kind, chrom, n) bytes; returns the bytes written
 * (no terminator), -1 on bad arguments.  n_threads 0 = all. */
int64_t pb_format_track_bound(int kind, const char *chrom, int64_t n);
int64_t pb_format_track(int kind, const char *chrom, const int64_t *start, const int64_t *end, const void *values,
                        int values_are_float, int64_t n, char *out, int64_t cap, int n_threads);

/* `metagene generate` geometry.  Transcript table: blocks [tx_bstart, tx_bend) ascending per
 * transcript, in ONE coordinate system shared by all chromosomes (global bins: chrom_bin_off[c] + position),
 * so that equal coordinates mean equal genomic positions; tx_bcum = chain coordinate of each block's first
 * base in genomic order (exclusive running sum of block lengths within the transcript); transcript t owns
 * blocks [tx_off[t], tx_off[t+1]); tx_reverse[t] = 1 for '-' strand transcripts.
 *
 * pb_landmark_windows: window_landmark(region, flank_up, flank_down, landmark=tx_landmark[t]) with
 * ref_delta = 0 (plastid/bin/metagene.py:180-239; window_cds_start / window_cds_stop :241-340 pass
 * cds_start / cds_end - 3), one thread per transcript.  tx_landmark[t] < 0 = no landmark (non-coding).
 * win_out int64[n_tx][4] = {w_start, w_end, w_off, ref_pos}: the window covers transcript coordinates
 * [w_start, w_end) and is placed at column w_off of a (flank_up + flank_down)-wide row; ref_pos = genomic
 * coordinate of the landmark.  flags_out uint8[n_tx]: PB_WIN_HAS_REF, PB_WIN_INDEX_ERROR (landmark beyond
 * the transcript: the reference catches IndexError and ignores the region, :457-461). */
#define PB_WIN_HAS_REF     1
#define PB_WIN_INDEX_ERROR 2
int pb_landmark_windows(const int64_t *tx_bstart, const int64_t *tx_bend, const int64_t *tx_bcum,
                        const int64_t *tx_off, const uint8_t *tx_reverse, const int64_t *tx_landmark,
                        int64_t n_tx, int32_t flank_up, int32_t flank_down,
                        int64_t *win_out, uint8_t *flags_out, void *stream);

/* pb_spanning_windows: maximal_spanning_window (plastid/bin/metagene.py:343-502) for every group of
 * transcripts (group_regions_make_windows :702-735 calls it once per gene) in one launch, one warp per
 * group.  Group g = transcripts grp_tx[grp_off[g] .. grp_off[g+1]) (indices into the transcript table, in
 * the order the reference would iterate them: the LAST one decides the offset quirk of :495-499);
 * win / flags as written by pb_landmark_windows (or filled by the host from a custom window function:
 * blocks = the window's own blocks, w_start = 0, w_end = its length).  A column of the
 * (flank_up + flank_down)-wide alignment is kept when every transcript of the group has the same genomic
 * position there.  Outputs per group: status (PB_SPAN_NONE: landmarks differ / a transcript has none / no
 * shared column; PB_SPAN_WINDOW; PB_SPAN_REF_OUTSIDE: the reference would raise KeyError at :498),
 * offset = alignment_offset, n_pos = window length, n_blk = its number of blocks, refpos = the shared
 * landmark coordinate.  Call once with out_off == NULL (count), exclusive-scan n_blk into out_off, and call
 * again with out_off / out_bstart / out_bend to write the blocks of group g, ascending, at out_off[g]
 * (the second call reads n_blk and leaves the per-group outputs untouched). */
#define PB_SPAN_NONE        0
#define PB_SPAN_WINDOW      1
#define PB_SPAN_REF_OUTSIDE 2
int pb_spanning_windows(const int64_t *tx_bstart, const int64_t *tx_bend, const int64_t *tx_bcum,
                        const int64_t *tx_off, const uint8_t *tx_reverse,
                        const int64_t *win, const uint8_t *flags,
                        const int64_t *grp_off, const int64_t *grp_tx, int64_t n_grp,
                        int32_t flank_up, int32_t flank_down,
                        uint8_t *status, int32_t *offset, int32_t *n_pos, int32_t *n_blk, int64_t *refpos,
                        const int64_t *out_off, int64_t *out_bstart, int64_t *out_bend, void *stream);

/* `cs generate` position-set arithmetic (plastid/bin/cs.py:242-496 process_partial_group pools,
 * intersects and subtracts python sets of genomic positions per merged gene).  A chain table is
 * {bstart, bend, chain_off}: chain c owns blocks [chain_off[c], chain_off[c+1]), sorted, disjoint and
 * non-touching, in one coordinate system for all chromosomes (global bins).  Both operations are
 * count / fill pairs like pb_spanning_windows: call with out_off == NULL to get n_blk[i], exclusive-scan it
 * into out_off, call again to write the blocks of output chain i at out_off[i].  Outputs are again sorted,
 * disjoint and non-touching.
 *
 * pb_chain_union: output chain g = union of the chains members[grp_off[g] .. grp_off[g+1]) (touching
 * blocks merge, as positions_to_segments does); one warp per group, lanes over members.
 * pb_chain_binary: output chain i = A[a_idx[i]] AND B[b_idx[i]] (op PB_CHAIN_AND) or A[a_idx[i]] minus
 * B[b_idx[i]] (op PB_CHAIN_SUB); b_idx[i] < 0 = empty B. */
#define PB_CHAIN_AND 0
#define PB_CHAIN_SUB 1
int pb_chain_union(const int64_t *bstart, const int64_t *bend, const int64_t *chain_off,
                   const int64_t *grp_off, const int64_t *members, int64_t n_grp,
                   int32_t *n_blk, const int64_t *out_off, int64_t *out_bstart, int64_t *out_bend, void *stream);
int pb_chain_binary(int op,
                    const int64_t *a_bstart, const int64_t *a_bend, const int64_t *a_off, const int64_t *a_idx,
                    const int64_t *b_bstart, const int64_t *b_bend, const int64_t *b_off, const int64_t *b_idx,
                    int64_t n_out, int32_t *n_blk, const int64_t *out_off, int64_t *out_bstart, int64_t *out_bend,
                    void *stream);

/* Roofline probe (the atomic peak a scatter-add design would be bound by; no
 * reference counterpart, not on the product path): n_updates `red.global.add.u32` into
 * bins[0..n_bins) — mode 0 uniformly random targets, mode 1 sorted targets with +-jitter.
 * bench.py --workload peaks times it with CUDA events. */
int pb_atomic_probe(uint32_t *bins, int64_t n_bins, int64_t n_updates, int mode, int jitter, void *stream);

/* Roofline probe for the region kernels (no reference counterpart, not on the product path): n_chunks warps each
 * read `chunk_bins` consecutive uint32 bins (a multiple of 4) from a pseudo-random 16-byte aligned place in
 * vec[0..n_bins) with 16-byte loads and write one word to out[n_chunks].  The bytes per second this launch achieves
 * are what HBM delivers for scattered kilobyte-sized segments — the access pattern of pb_region_sums /
 * pb_gather_windows over tens of thousands of exon blocks in a 12-25 GB plane. */
int pb_gather_probe(const uint32_t *vec, int64_t n_bins, int chunk_bins, int64_t n_chunks, uint32_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PLASTID_B200_H */
