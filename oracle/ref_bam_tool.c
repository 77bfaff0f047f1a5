/*
 * ref_bam_tool.c — SAM->BAM conversion and BAM dumping done by the REFERENCE's vendored htslib 1.3
 * (kent/src/htslib), compiled in place from /root/reference by oracle/Makefile into oracle/_ref/.
 * tests/golden/make_bam_golden.py uses it to produce the committed fixtures the BAM decoder
 * (plastid_b200/csrc/pb_bam.cpp) is checked against.  TEST INFRASTRUCTURE ONLY.
 *
 *   ref_bam_tool sam2bam in.sam out.bam
 *   ref_bam_tool dump in.bam          -> "tid pos flag n_cigar op:len,op:len,..." per record
 *   ref_bam_tool index in.bam         -> in.bam.bai written by htslib's own indexer (sam_index_build)
 *   ref_bam_tool fetch in.bam tid beg end [tid beg end ...]
 *                                     -> per region "# tid beg end", one "pos flag op:len,... endpos" line per record
 *                                        htslib's iterator returns (sam_itr_queryi: pysam's AlignmentFile.fetch), "= n"
 *   ref_bam_tool positions in.bam     -> "qname tid pos" for every reference position a read has an ALIGNED
 *                                        base at, as seen by htslib's own pileup engine (bam_plp_auto:
 *                                        the read is in the column and neither is_del nor is_refskip).
 *                                        Per read this is pysam's AlignedSegment.positions
 *                                        (= get_reference_positions()), the a1 row of SURVEY 8(a).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "htslib/sam.h"

typedef struct { samFile *in; bam_hdr_t *h; } plp_src;

static int plp_read(void *data, bam1_t *b)
{
    plp_src *s = (plp_src *)data;
    return sam_read1(s->in, s->h, b);
}

int main(int argc, char **argv)
{
    if (argc >= 4 && !strcmp(argv[1], "sam2bam")) {
        samFile *in = sam_open(argv[2], "r");
        if (!in) return 2;
        bam_hdr_t *h = sam_hdr_read(in);
        samFile *out = sam_open(argv[3], "wb");
        if (!out || sam_hdr_write(out, h) < 0) return 3;
        bam1_t *b = bam_init1();
        while (sam_read1(in, h, b) >= 0)
            if (sam_write1(out, h, b) < 0) return 4;
        sam_close(out);
        sam_close(in);
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "dump")) {
        samFile *in = sam_open(argv[2], "r");
        if (!in) return 2;
        bam_hdr_t *h = sam_hdr_read(in);
        for (int i = 0; i < h->n_targets; ++i) printf("@ %s %u\n", h->target_name[i], h->target_len[i]);
        bam1_t *b = bam_init1();
        while (sam_read1(in, h, b) >= 0) {
            printf("%d %d %d %d ", b->core.tid, b->core.pos, b->core.flag, b->core.n_cigar);
            const uint32_t *c = bam_get_cigar(b);
            for (int k = 0; k < b->core.n_cigar; ++k)
                printf("%s%d:%d", k ? "," : "", bam_cigar_op(c[k]), bam_cigar_oplen(c[k]));
            printf(" %d\n", bam_endpos(b));
        }
        sam_close(in);
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "positions")) {
        plp_src src;
        src.in = sam_open(argv[2], "r");
        if (!src.in) return 2;
        src.h = sam_hdr_read(src.in);
        bam_plp_t it = bam_plp_init(plp_read, &src);
        bam_plp_set_maxcnt(it, 1 << 30);
        int tid, pos, n;
        const bam_pileup1_t *col;
        while ((col = bam_plp_auto(it, &tid, &pos, &n)) != 0)
            for (int k = 0; k < n; ++k)
                if (!col[k].is_del && !col[k].is_refskip)
                    printf("%s %d %d\n", bam_get_qname(col[k].b), tid, pos);
        if (n < 0) return 5;
        bam_plp_destroy(it);
        sam_close(src.in);
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "index"))
        return sam_index_build(argv[2], 0) < 0 ? 6 : 0;
    if (argc >= 6 && !strcmp(argv[1], "fetch")) {
        samFile *in = sam_open(argv[2], "r");
        if (!in) return 2;
        bam_hdr_t *h = sam_hdr_read(in);
        hts_idx_t *idx = sam_index_load(in, argv[2]);
        if (!idx) return 7;
        bam1_t *b = bam_init1();
        for (int a = 3; a + 2 < argc; a += 3) {
            int tid = atoi(argv[a]), beg = atoi(argv[a + 1]), end = atoi(argv[a + 2]);
            hts_itr_t *it = sam_itr_queryi(idx, tid, beg, end);
            long n = 0;
            printf("# %d %d %d\n", tid, beg, end);
            while (it && sam_itr_next(in, it, b) >= 0) {
                printf("%d %d ", b->core.pos, b->core.flag);
                const uint32_t *c = bam_get_cigar(b);
                for (int k = 0; k < b->core.n_cigar; ++k)
                    printf("%s%d:%d", k ? "," : "", bam_cigar_op(c[k]), bam_cigar_oplen(c[k]));
                printf("%s %d\n", b->core.n_cigar ? "" : "*", bam_endpos(b));
                ++n;
            }
            printf("= %ld\n", n);
            if (it) hts_itr_destroy(it);
        }
        sam_close(in);
        return 0;
    }
    fprintf(stderr, "usage: ref_bam_tool sam2bam in.sam out.bam | dump in.bam | positions in.bam\n");
    return 1;
}
