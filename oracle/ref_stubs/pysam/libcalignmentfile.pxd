# Stand-in declaration for `from pysam.libcalignmentfile cimport AlignedSegment` (see ../README.md).
cdef class AlignedSegment:
    cdef public long reference_start
    cdef public bint is_reverse
    cdef public object cigartuples
    cdef public object query_name
    cdef public object _positions
    cdef public long _reference_end
