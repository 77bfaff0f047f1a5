/*
 * ref_cigar_table.c — prints the CIGAR consume table straight from the REFERENCE's
 * vendored htslib header (kent/src/htslib/htslib/sam.h:64-104), compiled in place from
 * /root/reference by oracle/Makefile into oracle/_ref/.  tests/test_oracle_kat.py
 * compares the committed copy of its output (tests/golden/cigar_consume_table.txt)
 * with the table oracle/pyoracle.py and the packer use.  TEST INFRASTRUCTURE ONLY.
 */
#include <stdio.h>
#include "htslib/sam.h"

int main(void)
{
    for (int op = 0; op < 10; ++op)
        printf("%c\t%d\t%d\t%d\n", BAM_CIGAR_STR[op], op, bam_cigar_type(op) & 1, (bam_cigar_type(op) >> 1) & 1);
    return 0;
}
