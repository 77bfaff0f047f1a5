"""Small shared helpers for the script entry points: alignment / annotation / table I/O and the
mapping-rule flags of ``plastid/util/scriptlib/argparsers.py:337-503`` (same names and defaults)."""
import numpy as np

from ..batch import AlignmentBatch
from ..genome_array import BAMGenomeArray
from ..map_factories import (CenterMapFactory, FivePrimeMapFactory, ThreePrimeMapFactory,
                             VariableFivePrimeMapFactory, SizeFilterFactory)
from ..roitools import GenomicSegment, SegmentChain, Transcript


def add_alignment_args(parser):
    g = parser.add_argument_group("alignment mapping options")
    g.add_argument("--count_files", nargs="+", required=True,
                   help="alignment batches (.npz written by save_batch) or sorted, indexed BAM files (needs pysam)")
    g.add_argument("--fiveprime", action="store_true")
    g.add_argument("--threeprime", action="store_true")
    g.add_argument("--center", action="store_true")
    g.add_argument("--fiveprime_variable", action="store_true")
    g.add_argument("--offset", default=0)
    g.add_argument("--nibble", type=int, default=0)
    g.add_argument("--min_length", type=int, default=25)
    g.add_argument("--max_length", type=int, default=100)
    g.add_argument("--sum", type=float, default=None)
    g.add_argument("--device", default="cuda")


def save_batch(path, batch):
    kw = dict(chroms=np.asarray(batch.chroms), chrom_len=batch.chrom_len, ref_start=batch.ref_start,
              meta=batch.meta, chrom_read_off=batch.chrom_read_off, max_span=batch.max_span, mapped=batch.mapped)
    if batch.blk is not None:
        kw.update(blk_off=batch.blk_off, blk=batch.blk)
    np.savez(path, **kw)


def load_batch(path):
    if str(path).endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        return AlignmentBatch([str(c) for c in z["chroms"]], z["chrom_len"], z["ref_start"], z["meta"],
                              z["chrom_read_off"], z["blk_off"] if "blk_off" in z else None,
                              z["blk"] if "blk" in z else None, int(z["max_span"]), int(z["mapped"]))
    from ..bam_io import batch_from_bam
    return batch_from_bam(path)


def genome_array_from_args(args):
    """argparsers.py:612-782: BAMGenomeArray + size filter + mapping factory (+ optional --sum)."""
    ga = BAMGenomeArray(*[load_batch(p) for p in args.count_files], device=args.device)
    ga.add_filter("size:%s-%s" % (args.min_length, args.max_length),
                  SizeFilterFactory(min=args.min_length, max=args.max_length))
    if args.fiveprime_variable:
        ga.set_mapping(VariableFivePrimeMapFactory.from_file(str(args.offset)))
    elif args.threeprime:
        ga.set_mapping(ThreePrimeMapFactory(offset=int(args.offset)))
    elif args.center:
        ga.set_mapping(CenterMapFactory(nibble=int(args.nibble)))
    else:
        ga.set_mapping(FivePrimeMapFactory(offset=int(args.offset)))
    if args.sum is not None:
        ga.set_sum(args.sum)
    return ga


def read_bed(path, as_transcripts=False):
    """BED3-BED12(+) -> list of SegmentChain (thin stand-in for plastid/readers/bed.py).  With
    ``as_transcripts`` every line becomes a :class:`Transcript` whose coding region is
    thickStart..thickEnd (none when they are equal, like ``Transcript.from_bed``); a 13th column, when
    present, is taken as ``gene_id``."""
    chains = []
    with open(path) as fh:
        for line in fh:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            f = line.rstrip("\n").split("\t")
            chrom, start, end = f[0], int(f[1]), int(f[2])
            name = f[3] if len(f) > 3 else "%s:%s-%s" % (chrom, start, end)
            strand = f[5] if len(f) > 5 and f[5] in ("+", "-", ".") else "."
            if len(f) >= 12:
                sizes = [int(x) for x in f[10].strip(",").split(",")]
                starts = [int(x) for x in f[11].strip(",").split(",")]
                segs = [GenomicSegment(chrom, start + a, start + a + n, strand) for a, n in zip(starts, sizes)]
            else:
                segs = [GenomicSegment(chrom, start, end, strand)]
            if as_transcripts:
                attr = dict(ID=name)
                if len(f) > 7 and int(f[6]) < int(f[7]):
                    attr.update(cds_genome_start=int(f[6]), cds_genome_end=int(f[7]))
                if len(f) > 12 and f[12]:
                    attr["gene_id"] = f[12]
                chains.append(Transcript(*segs, **attr))
            else:
                chains.append(SegmentChain(*segs, ID=name))
    return chains


def read_pl_table(path):
    """Tab-delimited table with '#' comments and a header row -> dict of column lists
    (plastid/util/io/filters.py read_pl_table stand-in, no pandas needed)."""
    cols, header = {}, None
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            f = line.rstrip("\n").split("\t")
            if header is None:
                header = f
                for h in header:
                    cols[h] = []
                continue
            for h, v in zip(header, f):
                cols[h].append(v)
    return cols
