"""numpy front-end of the C oracle (``oracle/oracle.c``).  TEST INFRASTRUCTURE ONLY: imported by
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` — never by
``plastid_b200/``.  Parity status: see ``oracle/__init__.py``."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

STRAND = {"+": 1, "-": 2, ".": 3}
RULE = {"fiveprime": 0, "threeprime": 1, "variable": 2}


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _batch_args(hb):
    blk = None if hb.blk is None else np.ascontiguousarray(hb.blk, dtype=np.int32)
    return hb.ref_start, hb.meta, hb.blk_off, blk


def map_point(hb, i0, i1, rule, offset, luts, size_filter, strand, seg_start, seg_end, want_kept=False):
    """Operator on reads [i0,i1) -> (int64 counts, kept or None, dropped count, a dropped length)."""
    rs, meta, boff, blk = _batch_args(hb)
    n = max(seg_end - seg_start, 0)
    counts = np.zeros(n, dtype=np.int64)
    kept = np.zeros(max(i1 - i0, 1), dtype=np.uint8) if want_kept else None
    dropped = np.zeros(2, dtype=np.int64)
    fw, rc = luts if luts is not None else (None, None)
    smin, smax = size_filter if size_filter is not None else (0, -1)
    lib().or_map_point(_p(rs), _p(meta), _p(boff), _p(blk), C.c_int64(i0), C.c_int64(i1), RULE[rule], int(offset),
                       _p(fw), _p(rc), int(smin), int(smax), STRAND[strand], C.c_int64(seg_start),
                       C.c_int64(seg_end), _p(counts), _p(kept), _p(dropped))
    return counts, (None if kept is None else kept[:i1 - i0].astype(bool)), int(dropped[0]), int(dropped[1])


def map_center(hb, i0, i1, nibble, size_filter, strand, seg_start, seg_end, want_kept=False):
    rs, meta, boff, blk = _batch_args(hb)
    n = max(seg_end - seg_start, 0)
    counts = np.zeros(n, dtype=np.float64)
    kept = np.zeros(max(i1 - i0, 1), dtype=np.uint8) if want_kept else None
    dropped = np.zeros(2, dtype=np.int64)
    smin, smax = size_filter if size_filter is not None else (0, -1)
    lib().or_map_center(_p(rs), _p(meta), _p(boff), _p(blk), C.c_int64(i0), C.c_int64(i1), int(nibble),
                        int(smin), int(smax), STRAND[strand], C.c_int64(seg_start), C.c_int64(seg_end),
                        _p(counts), _p(kept), _p(dropped))
    return counts, (None if kept is None else kept[:i1 - i0].astype(bool)), int(dropped[0]), int(dropped[1])


def map_stratified(hb, i0, i1, luts, min_len, max_len, size_filter, strand, seg_start, seg_end, want_kept=False):
    rs, meta, boff, blk = _batch_args(hb)
    n = max(seg_end - seg_start, 0)
    counts = np.zeros((max_len - min_len + 1, n), dtype=np.int64)
    kept = np.zeros(max(i1 - i0, 1), dtype=np.uint8) if want_kept else None
    smin, smax = size_filter if size_filter is not None else (0, -1)
    lib().or_map_stratified(_p(rs), _p(meta), _p(boff), _p(blk), C.c_int64(i0), C.c_int64(i1), _p(luts[0]),
                            _p(luts[1]), int(min_len), int(max_len), int(smin), int(smax), STRAND[strand],
                            C.c_int64(seg_start), C.c_int64(seg_end), _p(counts), _p(kept))
    return counts, (None if kept is None else kept[:i1 - i0].astype(bool))


def genome_vector(hb, chrom_index, strand, rule=None, offset=0, luts=None, nibble=None, size_filter=None):
    """Whole-chromosome vector = the operator applied to ``GenomicSegment(chrom, 0, len, strand)``
    with every read of the chromosome (what ``fetch`` returns for that segment)."""
    i0, i1 = int(hb.chrom_read_off[chrom_index]), int(hb.chrom_read_off[chrom_index + 1])
    n = int(hb.chrom_len[chrom_index])
    if nibble is not None:
        return map_center(hb, i0, i1, nibble, size_filter, strand, 0, n)
    return map_point(hb, i0, i1, rule, offset, luts, size_filter, strand, 0, n)


def region_sums(vec, bstart, bend, chain_off, mask_bits=None, mask_off=None):
    bstart = np.ascontiguousarray(bstart, dtype=np.int64)
    bend = np.ascontiguousarray(bend, dtype=np.int64)
    chain_off = np.ascontiguousarray(chain_off, dtype=np.int64)
    n = len(chain_off) - 1
    sums = np.zeros(n, dtype=np.float64)
    live = np.zeros(n, dtype=np.int64)
    fns = {np.dtype(np.float64): "or_region_sums_f64", np.dtype(np.uint32): "or_region_sums_u32",
           np.dtype(np.int64): "or_region_sums_i64"}
    if vec.dtype not in fns:
        raise TypeError("vector must be uint32, int64 or float64")
    fn = getattr(lib(), fns[vec.dtype])
    fn(_p(vec), _p(bstart), _p(bend), _p(chain_off), C.c_int64(n), _p(mask_bits),
       _p(None if mask_off is None else np.ascontiguousarray(mask_off, dtype=np.int64)), _p(sums), _p(live))
    return sums, live
