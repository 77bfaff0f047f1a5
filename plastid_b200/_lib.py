"""ctypes binding of ``libplastid_b200.so`` (C-ABI declared in ``include/plastid_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C plastid_b200/csrc``.
There is no CPU fallback: if the library is missing, or no CUDA device is visible, every
compute call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplastid_b200.so")

PB_OK, PB_EINVAL, PB_ECUDA, PB_ENOSPACE = 0, -1, -2, -3
PB_LAYOUT_ALIGN = 16384
PB_LUT_SIZE = 10000
PB_PLANE_PLUS, PB_PLANE_MINUS, PB_PLANE_ANY = 1, 2, 4
PB_RULE_FIVEPRIME, PB_RULE_THREEPRIME, PB_RULE_VARIABLE, PB_RULE_CENTER, PB_RULE_STRATIFIED = range(5)
(PB_STAT_DROPPED_PLUS, PB_STAT_DROPPED_MINUS, PB_STAT_DROPPED_ANY, PB_STAT_DROPPED_LEN,
 PB_STAT_MAPPED_PLUS, PB_STAT_MAPPED_MINUS, PB_STAT_MAPPED_ANY) = range(7)
PB_NSTATS = 8
PB_WIN_HAS_REF, PB_WIN_INDEX_ERROR = 1, 2
PB_SPAN_NONE, PB_SPAN_WINDOW, PB_SPAN_REF_OUTSIDE = 0, 1, 2
PB_CHAIN_AND, PB_CHAIN_SUB = 0, 1

STRAND_PLANE = {"+": PB_PLANE_PLUS, "-": PB_PLANE_MINUS, ".": PB_PLANE_ANY}
PLANE_INDEX = {"+": 0, "-": 1, ".": 2}


class PbBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("ref_start", C.c_void_p), ("meta", C.c_void_p),
                ("blk_off", C.c_void_p), ("blk", C.c_void_p), ("chrom_read_off", C.c_void_p),
                ("n_chrom", C.c_int32), ("max_span", C.c_int32), ("n_blk", C.c_int64),
                ("max_block_len", C.c_int32), ("reserved", C.c_int32)]


class PbLayout(C.Structure):
    _fields_ = [("n_chrom", C.c_int32), ("reserved", C.c_int32), ("chrom_len", C.c_void_p),
                ("chrom_bin_off", C.c_void_p), ("total_bins", C.c_int64)]


class PbRule(C.Structure):
    _fields_ = [("kind", C.c_int32), ("param", C.c_int32), ("lut_fw", C.c_void_p), ("lut_rc", C.c_void_p),
                ("size_min", C.c_int32), ("size_max", C.c_int32), ("strat_min", C.c_int32),
                ("strat_max", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    "pb_version": (C.c_char_p, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_device_count": (C.c_int, []),
    "pb_bam_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "pb_bam_decode": (C.c_int, [_P, C.c_int]),
    "pb_bam_n_ref": (C.c_int, [_P]),
    "pb_bam_ref_name": (C.c_char_p, [_P, C.c_int]),
    "pb_bam_ref_len": (C.c_int64, [_P, C.c_int]),
    "pb_bam_n_reads": (C.c_int64, [_P]),
    "pb_bam_n_blk": (C.c_int64, [_P]),
    "pb_bam_n_mapped": (C.c_int64, [_P]),
    "pb_bam_n_skipped": (C.c_int64, [_P]),
    "pb_bam_max_span": (C.c_int32, [_P]),
    "pb_bam_copy": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "pb_bam_close": (None, [_P]),
    "pb_bai_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "pb_bai_close": (None, [_P]),
    "pb_bai_n_ref": (C.c_int, [_P]),
    "pb_bai_mapped": (C.c_int64, [_P, C.c_int]),
    "pb_bam_read_header": (C.c_int, [_P]),
    "pb_bam_fetch": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64]),
    "pb_bam_build_index": (C.c_int, [C.c_char_p, C.c_char_p]),
    "pb_meta_length_hist": (C.c_int, [_P, C.c_int64, C.c_int, _P]),
    "pb_inflate_raw": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t]),
    "pb_format_track_bound": (C.c_int64, [C.c_int, C.c_char_p, C.c_int64]),
    "pb_format_track": (C.c_int64, [C.c_int, C.c_char_p, _P, _P, _P, C.c_int, C.c_int64, _P, C.c_int64, C.c_int]),
    "pb_map_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "pb_unpack_wire16": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
    "pb_unpack_delta8": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
    "pb_pack_delta3": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pb_unpack_delta3": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
    "pb_unpack_blocks_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "pb_unpack_blocks": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "pb_unpack_blocks_range": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64,
                                         _P, _P, _P, C.c_size_t, _P]),
    "pb_map_point_range": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int,
                                     _P, _P, _P, _P, _P, C.c_size_t, C.c_int64, C.c_int64, C.c_int64, _P]),
    "pb_map_point": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int,
                               _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_map_center": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int,
                                _P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_map_center_fixed": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int,
                                      _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_map_center_range": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int,
                                      _P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_int64, _P]),
    "pb_map_center_fixed_range": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int,
                                            _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, C.c_int64, C.c_int64,
                                            C.c_int64, C.c_int64, _P]),
    "pb_enable_kernel_timing": (None, [C.c_int]),
    "pb_tiles_kernel_ms_total": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "pb_map_segment": (C.c_int, [C.POINTER(PbBatch), C.c_int64, C.c_int64, C.POINTER(PbRule), C.c_int, C.c_int,
                                 C.c_int64, C.c_int64, _P, _P, _P, _P]),
    "pb_length_hist": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbRule), C.c_int, _P, _P]),
    "pb_region_sums_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "pb_region_sums": (C.c_int, [C.POINTER(_P), C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, C.c_int64, C.c_int64,
                                 _P, _P, _P, C.c_size_t, _P]),
    "pb_gather_windows": (C.c_int, [C.POINTER(_P), C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32,
                                    _P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "pb_gather_chains": (C.c_int, [C.POINTER(_P), C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64,
                                   _P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "pb_chain_counts_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "pb_chain_counts": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), _P, _P, _P, _P, _P, _P, _P, C.c_int64,
                                  C.c_int64, _P, _P, C.c_int64, C.c_int64, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_stratified_windows_range": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int, C.c_int,
                                              _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int, C.c_int32, C.c_int32,
                                              _P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "pb_stratified_windows_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "pb_stratified_windows_ws": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int, C.c_int,
                                           _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int, C.c_int32, C.c_int32,
                                           _P, _P, C.c_int64, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "pb_phase_sums_range": (C.c_int, [C.POINTER(_P), C.c_int, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_int64, C.c_int64, _P, _P]),
    "pb_window_normalize": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                      _P, _P, _P, _P, _P]),
    "pb_column_profile_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "pb_stratified_windows": (C.c_int, [C.POINTER(PbBatch), C.POINTER(PbLayout), C.POINTER(PbRule), C.c_int, C.c_int,
                                        _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int, C.c_int32, C.c_int32,
                                        _P, _P, _P, _P, _P]),
    "pb_phase_sums": (C.c_int, [C.POINTER(_P), C.c_int, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "pb_mask_chains": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "pb_export_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "pb_export_runs": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_landmark_windows": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P]),
    "pb_spanning_windows": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32,
                                      _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pb_chain_union": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "pb_chain_binary": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "pb_atomic_probe": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int, C.c_int, _P]),
    "pb_gather_probe": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int64, _P, _P]),
    "pb_count_profiles_u32": (C.c_int, [_P, _P, C.c_int, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int,
                                        _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_column_profile_batched": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "pb_column_profile": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
}

_lib = None


class PlastidB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlastidB200Error(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C plastid_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc):
    if rc != PB_OK:
        raise PlastidB200Error("libplastid_b200 error %d: %s" % (rc, lib().pb_last_error().decode()))


def require_cuda():
    """The product path has no CPU fallback: fail loudly without a device."""
    import torch
    if not torch.cuda.is_available() or lib().pb_device_count() < 1:
        raise PlastidB200Error("plastid_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def device_key(device):
    """Canonical name of a torch device ("cuda" and "cuda:0" name the same GPU): cache key of per-device tables."""
    import torch
    dv = torch.device(device)
    if dv.type == "cuda" and dv.index is None and torch.cuda.is_available():
        dv = torch.device("cuda", torch.cuda.current_device())
    return str(dv)


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_threads(requested=0):
    """Worker threads for the host-side passes (BAM decode, transfer-format packing, track text): ``requested`` when
    positive, else the CPUs this process may run on (its affinity mask, not the machine's core count — a rank bound to
    its GPU's NUMA node must not start a thread per core of the whole box), shared out among the ranks of the node
    when torchrun started several and nobody narrowed the mask."""
    if requested and int(requested) > 0:
        return int(requested)
    import os
    try:
        allowed = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        allowed = os.cpu_count() or 1
    ranks = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    if ranks > 1 and allowed == (os.cpu_count() or allowed):
        allowed = max(1, allowed // ranks)
    return max(1, allowed)
