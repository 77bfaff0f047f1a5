"""Counting sub-programs of the plastid scripts, on the GPU path: ``counts_in_region``, ``cs count``,
``metagene count``, ``psite`` and ``phase_by_size`` (plastid/bin/*.py).  Annotation geometry
(``generate`` sub-programs), plotting and the BED/GTF/GFF parsers are out of scope."""
