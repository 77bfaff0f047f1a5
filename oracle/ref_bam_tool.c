/*
 * ref_bam_tool.c — SAM->BAM conversion and BAM dumping done by the REFERENCE's vendored htslib 1.3
 * (kent/src/htslib), compiled in place from /root/reference by oracle/Makefile into oracle/_ref/.
 * tests/golden/make_bam_golden.py uses it to produce the committed fixtures the BAM decoder
 * (plastid_b200/csrc/pb_bam.cpp) is checked against.  TEST INFRASTRUCTURE ONLY.
 *
 *   ref_bam_tool sam2bam in.sam out.bam
 *   ref_bam_tool dump in.bam          -> "tid pos flag n_cigar op:len,op:len,..." per record
 */
#include <stdio.h>
#include <string.h>
#include "htslib/sam.h"

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
    fprintf(stderr, "usage: ref_bam_tool sam2bam in.sam out.bam | dump in.bam\n");
    return 1;
}
