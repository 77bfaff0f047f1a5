"""Shared plumbing of the command-line programs: the alignment / annotation / mask flags of the reference's parsers
(``plastid/util/scriptlib/argparsers.py:337-503, 845-975, 1190-1263`` — same names, defaults and exit messages, so that a
command line written for plastid runs unchanged), BED / table readers, and the multi-GPU start-up.

Multi-GPU: a program started under ``torchrun`` (``WORLD_SIZE`` > 1) calls :func:`init_distributed`; every rank
then builds a ``BAMGenomeArray`` that owns one position range of the genome (``plastid_b200.genome_array``), the
count tables are completed with all-reduces inside the library, and only rank 0 writes files (:func:`is_writer`).
"""
import os
import sys

import numpy as np

from ..batch import AlignmentBatch
from ..genome_array import BAMGenomeArray
from ..map_factories import (CenterMapFactory, FivePrimeMapFactory, ThreePrimeMapFactory,
                             VariableFivePrimeMapFactory, SizeFilterFactory)
from ..roitools import GenomicSegment, SegmentChain, Transcript


# ------------------------------------------------------------------------------------------------ multi-GPU
def init_distributed(device="cuda"):
    """Join the process group ``torchrun`` described in the environment (NCCL on GPUs, gloo otherwise) and return
    the device of this rank; a no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1") or 1)
    if world <= 1:
        return device
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # PB_DIST_BACKEND=gloo + PB_DIST_ONE_DEVICE=1: several ranks on ONE GPU, tables all-reduced through gloo — how the
    # single-GPU test box exercises the multi-rank code path (NCCL refuses two ranks on one device)
    backend = os.environ.get("PB_DIST_BACKEND")
    if str(device).startswith("cuda") and torch.cuda.is_available():
        if os.environ.get("PB_DIST_ONE_DEVICE"):
            local = 0
        torch.cuda.set_device(local)
        device = "cuda:%d" % local
        if not dist.is_initialized():
            if (backend or "nccl") == "nccl":
                dist.init_process_group("nccl", device_id=torch.device(device))
            else:
                dist.init_process_group(backend)
    elif not dist.is_initialized():
        dist.init_process_group(backend or "gloo")
    return device


def is_writer():
    """True on the one rank that writes output files."""
    from .. import dist as pdist
    return pdist.world()[0] == 0


def finish_distributed():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ parsers
def add_alignment_args(parser, disabled=()):
    """argparsers.py:337-503 (AlignmentParser): same flags; ``disabled`` like the reference's (e.g. "normalize")."""
    g = parser.add_argument_group("alignment mapping options")
    g.add_argument("--count_files", type=str, default=[], nargs="+",
                   help="One or more sorted BAM files (or .npz alignment batches written by save_batch) from a single "
                        "sample or set of samples to be pooled")
    g.add_argument("--countfile_format", choices=("BAM",), default="BAM",
                   help="Format of file containing alignments (the GPU path reads BAM; default: %(default)s)")
    if "normalize" not in disabled:
        g.add_argument("--normalize", action="store_true", default=False,
                       help="Whether counts should be normalized to counts per million (usually not. default: %(default)s)")
    if "sum" not in disabled:
        g.add_argument("--sum", type=float, default=None,
                       help="Sum used in normalization of counts and RPKM/RPNT calculations "
                            "(Default: total mapped reads/counts in dataset)")
    g.add_argument("--min_length", type=int, default=25, metavar="N",
                   help="Minimum read length required to be included (Default: %(default)s)")
    g.add_argument("--max_length", type=int, default=100, metavar="N",
                   help="Maximum read length permitted to be included (Default: %(default)s)")
    m = parser.add_argument_group("alignment mapping functions")
    # store_const into one destination, like the reference: the last flag given wins, none given is an error
    m.add_argument("--fiveprime_variable", action="store_const", const="fiveprime_variable", dest="mapping",
                   help="Map read alignment to a variable offset from 5' position of read, with offset determined by read "
                        "length. Requires `--offset` below")
    m.add_argument("--fiveprime", action="store_const", const="fiveprime", dest="mapping", help="Map read alignment to 5' position.")
    m.add_argument("--threeprime", action="store_const", const="threeprime", dest="mapping", help="Map read alignment to 3' position")
    m.add_argument("--center", action="store_const", const="center", dest="mapping",
                   help="Subtract N positions from each end of read, and add 1/(length-N), to each remaining position, "
                        "where N is specified by `--nibble`")
    m.add_argument("--offset", default=0, metavar="OFFSET",
                   help="For `--fiveprime` or `--threeprime`: integer offset into the read; for `--fiveprime_variable`: "
                        "two-column file of read length (or `default`) and offset (Default: %(default)s)")
    m.add_argument("--nibble", type=int, default=0, metavar="N",
                   help="For use with `--center` only. nt to remove from each end of read before mapping (Default: %(default)s)")
    g.add_argument("--device", default="cuda", help="CUDA device (under torchrun: one rank per GPU, chosen by LOCAL_RANK)")
    g.add_argument("--sharding", choices=("positions", "chromosomes"), default="positions",
                   help="under torchrun: cut the genome into per-GPU ranges anywhere (balanced by reads; default) or at "
                        "chromosome boundaries only")


def add_annotation_args(parser, prefix="", required=False):
    """argparsers.py:845-975 (AnnotationParser / MaskParser with ``prefix='mask_'``).  BED is the format on this path."""
    title = "mask file options (optional)" if prefix else "annotation file options (one or more annotation files required)"
    g = parser.add_argument_group(title)
    g.add_argument("--%sannotation_files" % prefix, type=str, nargs="+", default=[], metavar="infile.bed",
                   help="Zero or more annotation files (max 1 file if BigBed)")
    g.add_argument("--%sannotation_format" % prefix, choices=("BED", "GTF2", "GFF3", "BigBed"), default="GTF2",
                   help="Format of %sannotation_files (Default: %%(default)s; this path reads BED)" % prefix)
    g.add_argument("--%sadd_three" % prefix, default=False, action="store_true",
                   help="If supplied, coding regions will be extended by 3 nucleotides at their 3' ends")
    g.add_argument("--%stabix" % prefix, default=False, action="store_true", help="accepted for compatibility")
    g.add_argument("--%ssorted" % prefix, default=False, action="store_true", help="accepted for compatibility")
    g.add_argument("--%sbed_extra_columns" % prefix, default=0, nargs="+",
                   help="Number of extra columns in BED file, or list of names for those columns (Default: %(default)s)")


def add_mask_args(parser):
    add_annotation_args(parser, prefix="mask_")


def add_base_args(parser):
    """argparsers.py BaseParser: verbosity flags (accepted; warnings follow python's own filters here)."""
    g = parser.add_argument_group("warning/error options")
    g.add_argument("-q", "--quiet", action="count", default=0, help="Suppress all warning messages. Cannot use with '-v'.")
    g.add_argument("-v", "--verbose", action="count", default=0, help="Increase verbosity. With '-v', show every warning.")


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
                              z["blk"] if "blk" in z else None, int(z["max_span"]), int(z["mapped"])).pack()
    from ..bam_io import batch_from_bam
    return batch_from_bam(path)


def _message_and_exit(text):
    sys.stderr.write(text + "\n")
    sys.exit(1)


def mapping_from_args(args):
    """argparsers.py:656-700: the mapping factory the flags ask for, with the reference's exits."""
    rule = getattr(args, "mapping", None)
    if rule is None:
        _message_and_exit("Please specify a read mapping rule.")
    if rule == "fiveprime":
        return FivePrimeMapFactory(int(args.offset))
    if rule == "threeprime":
        return ThreePrimeMapFactory(int(args.offset))
    if rule == "center":
        return CenterMapFactory(args.nibble)
    if str(args.offset) == "0":
        _message_and_exit("Please specify a filename to use for fiveprime variable offsets in --offset.")
    return VariableFivePrimeMapFactory.from_file(str(args.offset))


def genome_array_from_args(args, disabled=()):
    """argparsers.py:612-782: BAMGenomeArray + size filter + mapping factory (+ ``--sum``, ``--normalize``)."""
    if len(args.count_files) == 0:
        _message_and_exit("Please include at least one input file.")
    mapping = mapping_from_args(args)
    device = init_distributed(getattr(args, "device", "cuda"))
    ga = BAMGenomeArray(*[load_batch(p) for p in args.count_files], device=device,
                        sharding=getattr(args, "sharding", "positions"))
    ga.add_filter("size:%s-%s" % (args.min_length, args.max_length),
                  SizeFilterFactory(min=args.min_length, max=args.max_length))
    ga.set_mapping(mapping)
    if "sum" not in disabled and getattr(args, "sum", None) is not None:
        ga.set_sum(args.sum)
    if "normalize" not in disabled and getattr(args, "normalize", False) is True:
        ga.set_normalize(True)
    return ga


def _extra_column_names(spec):
    """``--bed_extra_columns``: a number of unnamed columns or a list of names (readers/bed.py)."""
    if isinstance(spec, (list, tuple)):
        if len(spec) == 1 and str(spec[0]).isdigit():
            return ["custom%d" % i for i in range(int(spec[0]))]
        return [str(x) for x in spec]
    return ["custom%d" % i for i in range(int(spec))]


def read_bed(path, as_transcripts=False, extra_columns=None, add_three=False):
    """BED3-BED12(+) -> list of SegmentChain (stand-in for plastid/readers/bed.py).  With ``as_transcripts`` every
    line becomes a :class:`Transcript` whose coding region is thickStart..thickEnd (none when they are equal, like
    ``Transcript.from_bed``).  ``extra_columns``: names of the columns after the twelfth (``--bed_extra_columns``);
    default (None): a 13th column, when present, is taken as ``gene_id``.  ``add_three``: extend coding regions by
    three nucleotides at their 3' end (``--add_three``, readers/common.py add_three_for_stop_codon)."""
    names = None if extra_columns is None else _extra_column_names(extra_columns)
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
            attr = dict(ID=name)
            if names is None:
                if len(f) > 12 and f[12]:
                    attr["gene_id"] = f[12]
            else:
                for k, col in zip(names, f[12:]):
                    attr[k] = col
            if as_transcripts:
                if len(f) > 7 and int(f[6]) < int(f[7]):
                    attr.update(cds_genome_start=int(f[6]), cds_genome_end=int(f[7]))
                tx = Transcript(*segs, **attr)
                if add_three and tx.cds_genome_start is not None:
                    tx = _add_three(tx, segs, attr)
                chains.append(tx)
            else:
                chains.append(SegmentChain(*segs, **attr))
    return chains


def _add_three(tx, segs, attr):
    """Coding region + the three transcript positions after it (clamped to the transcript)."""
    new_end = min(tx.cds_end + 3, tx.length)
    if new_end == tx.cds_end:
        return tx
    a = dict(attr)
    last = tx.get_genomic_coordinate(new_end - 1)[1]
    if tx.strand == "-":
        a["cds_genome_start"] = last
    else:
        a["cds_genome_end"] = last + 1
    return Transcript(*segs, **a)


def chains_from_args(args, prefix="", as_transcripts=False):
    """argparsers.py:1014-1188: the regions named by ``--[mask_]annotation_files``."""
    files = getattr(args, prefix + "annotation_files")
    fmt = getattr(args, prefix + "annotation_format")
    if files and fmt != "BED" and not all(str(f).lower().endswith(".bed") for f in files):
        _message_and_exit("--%sannotation_format %s: this path reads BED (pass --%sannotation_format BED)" % (prefix, fmt, prefix))
    extra = getattr(args, prefix + "bed_extra_columns", 0)
    extra = None if extra in (0, "0", [], ["0"]) else extra
    out = []
    for fn in files:
        out.extend(read_bed(fn, as_transcripts=as_transcripts, extra_columns=extra if extra is not None else [],
                            add_three=bool(getattr(args, prefix + "add_three", False))))
    return out


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
