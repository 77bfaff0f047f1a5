"""CPU oracle for the plastid read-to-coverage hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``plastid_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker.

Parity status
-------------
* ``<L>M`` alignments: pinned by the reference's own known-answer tests
  (``plastid/test/unit/genomics/test_map_factories.py:17-200``), transcribed in
  ``tests/test_oracle_kat.py``.
* CIGAR ops other than ``M``: **parity unpinned** against pysam itself (pysam
  0.19.0 is a third-party dependency that is absent from the reference tree and
  from this image).  The consume table is checked against the reference's
  vendored ``kent/src/htslib/htslib/sam.h:64-104`` by ``oracle/_ref`` (see
  ``oracle/Makefile``) and spliced goldens are regenerated analytically with the
  recipe of ``plastid/test/unit/genomics/test_genome_array.py:1776-1885``.
"""
