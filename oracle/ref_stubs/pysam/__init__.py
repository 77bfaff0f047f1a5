"""Stand-in for pysam (absent here): see ../README.md.  Test infrastructure only."""
import re

from pysam.libcalignmentfile import AlignedSegment

__version__ = "0.19.0"

_CIGAR_OPS = "MIDNSHP=X"
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")


def get_include():
    import os
    return [os.path.dirname(os.path.abspath(__file__))]


def get_defines():
    return []


def parse_cigar(text):
    return [(_CIGAR_OPS.index(op), int(n)) for n, op in _CIGAR_RE.findall(text)]


class AlignmentFile(object):
    """In-memory stand-in.  ``AlignmentFile(references, lengths, reads_by_ref)`` from objects, or
    ``AlignmentFile(path, "rb")`` from the plain-text alignment listing the golden-vector scripts write
    (``@SQ<TAB>name<TAB>length`` header lines, then ``name<TAB>0-based start<TAB>strand<TAB>CIGAR`` per read)."""

    def __init__(self, references, lengths="rb", reads_by_ref=None, mapped=None, filename="<memory>"):
        if isinstance(references, str):
            filename = references
            references, lengths, reads_by_ref = [], [], {}
            import gzip
            with (gzip.open(filename, "rt") if filename.endswith(".gz") else open(filename)) as fh:
                for k, line in enumerate(fh):
                    f = line.rstrip("\n").split("\t")
                    if f[0] == "@SQ":
                        references.append(f[1])
                        lengths.append(int(f[2]))
                    elif line.strip():
                        reads_by_ref.setdefault(f[0], []).append(
                            AlignedSegment(int(f[1]), parse_cigar(f[3]), f[2] == "-", "r%d" % k))
        self.references = tuple(references)
        self.lengths = tuple(int(x) for x in lengths)
        self.nreferences = len(self.references)
        self._reads = {r: sorted(reads_by_ref.get(r, ()), key=lambda x: x.reference_start) for r in self.references}
        n = sum(len(v) for v in self._reads.values())
        self.mapped = n if mapped is None else int(mapped)
        self.filename = filename

    def fetch(self, reference=None, start=None, end=None, until_eof=False, **kwargs):
        """Records whose reference span [reference_start, reference_end) overlaps [start, end)."""
        if reference is None:
            for r in self.references:
                for read in self._reads[r]:
                    yield read
            return
        if reference not in self._reads:
            raise ValueError("invalid reference `%s`" % reference)
        lo = 0 if start is None else start
        hi = float("inf") if end is None else end
        for read in self._reads[reference]:
            if read.reference_start >= hi:
                break
            if read.reference_end > lo:
                yield read

    def close(self):
        pass


Samfile = AlignmentFile
