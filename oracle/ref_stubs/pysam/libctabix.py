"""Name-only stand-in for pysam.libctabix (tabix readers are not on the counting path)."""


def tabix_generic_iterator(*a, **k):
    raise IOError("pysam.libctabix stand-in: tabix is not available")


def tabix_file_iterator(*a, **k):
    raise IOError("pysam.libctabix stand-in: tabix is not available")


class asTuple(object):
    pass
