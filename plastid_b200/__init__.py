"""plastid_b200 — B200 (sm_100a) implementation of plastid's read-to-coverage hot path.

Alignments -> mapping rule -> per-position count vectors -> region / metagene reductions, behind
plastid's own mapping-rule operator API.  Host code is Python; the compute is hand-written CUDA in
``libplastid_b200.so`` (C-ABI in ``include/plastid_b200.h``), reached through ctypes with torch
tensors as device buffers.  There is no CPU fallback.
"""
from .map_factories import (CenterMapFactory, FivePrimeMapFactory, ThreePrimeMapFactory,
                            VariableFivePrimeMapFactory, StratifiedVariableFivePrimeMapFactory,
                            SizeFilterFactory, DataWarning, MalformedFileError)
from .roitools import GenomicSegment, SegmentChain, Transcript, positions_to_segments
from .batch import AlignmentBatch, GenomeLayout, pack_reads, batch_from_arrays
from .genome_array import BAMGenomeArray, GenomeArray, SparseGenomeArray
from .masks import GenomeHash

__version__ = "0.1.0"
