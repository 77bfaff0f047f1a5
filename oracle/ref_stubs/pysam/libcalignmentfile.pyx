# cython: language_level=3
# Stand-in for pysam's AlignedSegment (see ../README.md): attributes the reference touches on the counting path.
cdef class AlignedSegment:
    def __init__(self, reference_start=0, cigartuples=(), is_reverse=False, query_name=None):
        self.reference_start = reference_start
        self.cigartuples = list(cigartuples)
        self.is_reverse = bool(is_reverse)
        self.query_name = query_name
        pos = []
        cdef long ref = reference_start
        for op, n in self.cigartuples:
            if op in (0, 7, 8):          # M = X: consume query and reference
                pos.extend(range(ref, ref + n))
                ref += n
            elif op in (2, 3):           # D N: reference only
                ref += n
        self._positions = pos
        self._reference_end = ref

    @property
    def positions(self):
        return list(self._positions)

    def get_reference_positions(self, full_length=False):
        return list(self._positions)

    @property
    def reference_end(self):
        return self._reference_end

    @property
    def pos(self):
        return self.reference_start

    @property
    def is_unmapped(self):
        return False

    def __repr__(self):
        return "<stub AlignedSegment %s start=%d %s>" % (self.query_name, self.reference_start, "-" if self.is_reverse else "+")
