"""Drop-in mapping-rule factories backed by the sm_100a kernels.

Same names, constructor arguments, properties, warnings and call protocol as
``plastid/genomics/map_factories.pyx`` (operator spec ``:110-121``): every factory instance is a
callable ``fn(reads, seg) -> (reads_out, count_array)`` usable with
``BAMGenomeArray.set_mapping``; the count array's last axis has ``len(seg)`` entries in genomic
order.  The object call packs ``reads`` into an SoA batch and runs ``pb_map_segment`` on the GPU;
:class:`plastid_b200.genome_array.BAMGenomeArray` instead lowers the factory once into whole-genome
count planes (``pb_map_point`` / ``pb_map_center``).  No CPU path exists.
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib
from .batch import pack_reads

LUT_SIZE = _lib.PB_LUT_SIZE
_BAD_OFFSET = -1


class DataWarning(Warning):
    """Stand-in for ``plastid.util.services.exceptions.DataWarning``."""


class MalformedFileError(Exception):
    """Stand-in for ``plastid.util.services.exceptions.MalformedFileError``."""

    def __init__(self, filename, message, line_num=None):
        self.filename, self.msg, self.line_num = filename, message, line_num
        Exception.__init__(self, "Error opening file '%s': %s" % (filename, message))


_warned = set()


def warn_onceperfamily(message, category=DataWarning):
    """One warning per (category, message) — ``plastid/util/services/exceptions.py:146-232``."""
    key = (category, message)
    if key not in _warned:
        _warned.add(key)
        warnings.warn(message, category, stacklevel=3)


def _seg_fields(seg):
    strand = getattr(seg, "strand", ".")
    return int(seg.start), int(seg.end), strand


class _MapFactory(object):
    """Shared machinery: lowering to ``pb_rule`` and the per-segment operator call."""
    kind = None
    count_dtype = np.int64

    def _params(self):
        return 0

    def _luts(self):
        return None

    def _lut_tensors(self, device):
        luts = self._luts()
        if luts is None:
            return None, None
        import torch
        cache = self.__dict__.setdefault("_lut_dev", {})
        key = _lib.device_key(device)
        if key not in cache:
            cache[key] = (torch.from_numpy(luts[0]).to(device), torch.from_numpy(luts[1]).to(device))
        return cache[key]

    def pb_rule(self, device, size_filter=None):
        fw, rc = self._lut_tensors(device)
        smin, smax = (0, -1) if size_filter is None else (size_filter.min_, size_filter.max_)
        return _lib.PbRule(self.kind, int(self._params()),
                           None if fw is None else fw.data_ptr(), None if rc is None else rc.data_ptr(),
                           int(smin), int(smax), int(getattr(self, "min_length", 0)),
                           int(getattr(self, "max_length", 0)))

    def _leading_shape(self):
        return []

    def _warn_dropped(self, n_dropped, a_length):
        raise NotImplementedError

    def map_segment(self, dbatch, i0, i1, seg_start, seg_end, strand, size_filter=None, want_kept=True,
                    flags=3):
        """Run the operator on reads ``[i0,i1)`` of a device batch; returns (counts ndarray, kept
        ndarray of bool or None)."""
        import torch
        _lib.require_cuda()
        dev = dbatch.device
        n = max(int(seg_end) - int(seg_start), 0)
        shape = self._leading_shape() + [n]
        tdtype = torch.float64 if self.count_dtype == np.float64 else torch.int64
        counts = torch.zeros(shape, dtype=tdtype, device=dev)
        kept = torch.zeros(max(i1 - i0, 1), dtype=torch.uint8, device=dev) if want_kept else None
        stats = torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=dev)
        rule = self.pb_rule(dev, size_filter)
        b = dbatch.c_struct()
        _lib.check(_lib.lib().pb_map_segment(C.byref(b), i0, i1, C.byref(rule), _lib.STRAND_PLANE[strand],
                                             int(flags), int(seg_start), int(seg_end), _lib.ptr(counts), _lib.ptr(kept),
                                             _lib.ptr(stats), _lib.stream_ptr()))
        st = stats.cpu().numpy()
        dropped = int(st[{"+": 0, "-": 1, ".": 2}[strand]])
        if dropped:
            self._warn_dropped(dropped, int(st[_lib.PB_STAT_DROPPED_LEN]))
        out = counts.cpu().numpy()
        k = None if kept is None else kept[:i1 - i0].cpu().numpy().astype(bool)
        return out, k

    def __call__(self, reads, seg):
        if reads is None or seg is None:
            raise TypeError("reads and seg may not be None")
        _lib.require_cuda()
        reads = list(reads)
        start, end, strand = _seg_fields(seg)
        strand = strand if strand in ("+", "-") else "."
        hb = pack_reads({"_": reads}, {"_": max(end, 1)}, keep_objects=True)
        if len(hb) == 0:
            return [], np.zeros(self._leading_shape() + [max(end - start, 0)], dtype=self.count_dtype)
        counts, kept = self.map_segment(hb.to_device("cuda"), 0, len(hb), start, end, strand, flags=0)
        # pack_reads sorts by start; reads_out keeps the caller's order like the reference loop
        kept_ids = set(id(hb.objects[i]) for i in np.nonzero(kept)[0])
        reads_out = [r for r in reads if id(r) in kept_ids]
        return reads_out, counts.astype(self.count_dtype, copy=False)


class CenterMapFactory(_MapFactory):
    """``CenterMapFactory(nibble=0)`` — map_factories.pyx:167-275."""
    kind = _lib.PB_RULE_CENTER
    count_dtype = np.float64

    def __init__(self, nibble=0):
        if nibble < 0:
            raise ValueError("CenterMapFactory: `nibble` must be >= 0. Got %s." % nibble)
        self._nibble = int(nibble)

    @property
    def nibble(self):
        return self._nibble

    @nibble.setter
    def nibble(self, val):
        if val < 0:
            raise OverflowError("can't convert negative value to unsigned int")
        self._nibble = int(val)

    def _params(self):
        return self._nibble

    def _warn_dropped(self, n, length):
        warn_onceperfamily(
            "Data contains read alignments shorter than `2*nibble` value of '%s' nt. Ignoring these."
            % (2 * self._nibble), DataWarning)

    def slot_tables(self, length_hist):
        """From a histogram of aligned lengths build (slot_of_len int16[65536], inv_m float64[S]):
        one slot per distinct positive map length ``L - 2*nibble``, ascending."""
        lengths = np.nonzero(np.asarray(length_hist))[0]
        lengths = lengths[lengths - 2 * self._nibble > 0]
        slot_of_len = np.full(65536, -1, dtype=np.int16)
        slot_of_len[lengths] = np.arange(len(lengths), dtype=np.int16)
        inv_m = 1.0 / (lengths - 2 * self._nibble).astype(np.float64)
        return slot_of_len, inv_m


    FIXED_MAX_REL_ERR = 1e-9       # three orders below the north star's 1e-6

    def fixed_point_tables(self, length_hist):
        """Integer weights for ``pb_map_center_fixed``: ``(slot_of_len, w_fix int64[S], shift)`` or ``None``
        when one map length suffices (the exact kernel is used) or the weights cannot be made precise
        enough.  ``shift`` is the largest power of two for which the sum over ALL reads of
        ``round(2^shift / m)`` stays below 2^62 (no bin total can overflow); the relative error of a
        weight against ``1/m`` is at most ``m * 2^-(shift+1)``."""
        hist = np.asarray(length_hist, dtype=np.float64)
        lengths = np.nonzero(hist)[0]
        lengths = lengths[lengths - 2 * self._nibble > 0]
        if len(lengths) < 2:
            return None
        m = (lengths - 2 * self._nibble).astype(np.int64)
        bound = float((hist[lengths] / m).sum()) + 1.0          # sum over reads of 1/m
        shift = min(52, int(np.floor(61.0 - np.log2(bound))))
        if shift < 1 or float(m.max()) * 2.0 ** -(shift + 1) > self.FIXED_MAX_REL_ERR:
            return None
        slot_of_len = np.full(65536, -1, dtype=np.int16)
        slot_of_len[lengths] = np.arange(len(lengths), dtype=np.int16)
        w_fix = np.asarray([((1 << shift) + int(mm) // 2) // int(mm) for mm in m], dtype=np.int64)
        return slot_of_len, w_fix, shift


class FivePrimeMapFactory(_MapFactory):
    """``FivePrimeMapFactory(offset=0)`` — map_factories.pyx:278-374."""
    kind = _lib.PB_RULE_FIVEPRIME
    _name = "FivePrimeMapFactory"

    def __init__(self, offset=0):
        if offset < 0:
            raise ValueError("%s: `offset` must be <= 0. Got %s." % (self._name, offset))
        self._offset = int(offset)

    @property
    def offset(self):
        return self._offset

    @offset.setter
    def offset(self, val):
        self._offset = int(val)

    def _params(self):
        return self._offset

    def _warn_dropped(self, n, length):
        warn_onceperfamily("Data contains read alignments shorter than offset (%s nt). Ignoring."
                           % self._offset, DataWarning)


class ThreePrimeMapFactory(FivePrimeMapFactory):
    """``ThreePrimeMapFactory(offset=0)`` — map_factories.pyx:377-474."""
    kind = _lib.PB_RULE_THREEPRIME
    _name = "ThreePrimeMapFactory"


def _parse_variable_offset_file(fh):
    """Two tab-separated columns, read length (or ``default``) and 5' offset —
    ``plastid/util/scriptlib/argparsers.py:2505-2561``."""
    name = getattr(fh, "__name__", "Variable offset file")
    table = {}
    for line in fh:
        if line.startswith("#"):          # CommentReader, map_factories.pyx:580
            continue
        if line.startswith("length"):
            continue
        items = line.strip("\n").split("\t")
        if len(items) != 2:
            raise MalformedFileError(name, "More or fewer than two columns on line:\n\t%s" % line.strip("\n"))
        key = items[0]
        try:
            key = key if key == "default" else int(key)
        except ValueError:
            raise MalformedFileError(name, "Non integer value for key '%s' on line:\n\t%s" % (key, line.strip("\n")))
        if key in table:
            raise MalformedFileError(name, "multiple offsets defined for read length %s" % key)
        try:
            table[key] = int(items[1])
        except ValueError:
            raise MalformedFileError(name, "Non integer value for value '%s' on line:\n\t%s"
                                     % (items[1], line.strip("\n")))
    return table


def _build_luts(offset_dict):
    """Forward / reverse offset LUTs — map_factories.pyx:494-543."""
    fw = np.full(LUT_SIZE, _BAD_OFFSET, dtype=np.int32)
    rc = np.full(LUT_SIZE, _BAD_OFFSET, dtype=np.int32)
    if offset_dict is None:
        offset_dict = {"default": 0}
    default = None
    if "default" in offset_dict:
        default = int(offset_dict["default"])
        if default + 1 < LUT_SIZE:
            fw[default + 1:] = default
            rc[default + 1:] = np.arange(default + 1, LUT_SIZE, dtype=np.int32) - default - 1
    for read_length, offset in offset_dict.items():
        if read_length == "default":
            continue
        if offset >= read_length:
            if default is None:
                # the reference reads an unbound local here (map_factories.pyx:533)
                raise UnboundLocalError("local variable 'default' referenced before assignment")
            if read_length >= default:
                warn_onceperfamily(
                    "Given offset '%s' longer than read length '%s'. Falling back to default '%s'."
                    % (offset, read_length, default), DataWarning)
            else:
                warn_onceperfamily(
                    "Given offset '%s' and default '%s' are longer than read length '%s'. Ignoring %s-mers."
                    % (offset, default, read_length, read_length), DataWarning)
            continue
        if not 0 <= int(read_length) < LUT_SIZE:
            raise IndexError("read length %s outside the %d-entry offset table" % (read_length, LUT_SIZE))
        fw[read_length] = offset
        rc[read_length] = read_length - offset - 1
    return fw, rc


class VariableFivePrimeMapFactory(_MapFactory):
    """``VariableFivePrimeMapFactory(offset_dict)`` — map_factories.pyx:477-650."""
    kind = _lib.PB_RULE_VARIABLE

    def __init__(self, offset_dict, *args):
        if offset_dict is not None and not isinstance(offset_dict, dict):
            raise TypeError("offset_dict must be a dict")
        self.forward_offsets, self.reverse_offsets = _build_luts(offset_dict)

    @classmethod
    def from_file(cls, fn_or_fh):
        if isinstance(fn_or_fh, str):
            with open(fn_or_fh) as fh:
                return cls(_parse_variable_offset_file(fh))
        return cls(_parse_variable_offset_file(fn_or_fh))

    def _luts(self):
        return self.forward_offsets, self.reverse_offsets

    def _warn_dropped(self, n, length):
        warn_onceperfamily("No usable offset for reads of length %s nt in offset dict. Ignoring these."
                           % length, DataWarning)


class StratifiedVariableFivePrimeMapFactory(VariableFivePrimeMapFactory):
    """``StratifiedVariableFivePrimeMapFactory(offset_dict, min=25, max=35)`` —
    map_factories.pyx:653-791 (2-D output, one row per read length)."""
    kind = _lib.PB_RULE_STRATIFIED

    def __init__(self, offset_dict, min=25, max=35):
        VariableFivePrimeMapFactory.__init__(self, offset_dict)
        if max <= min:
            raise ValueError("Max length '%s' must be >= min length '%s'. " % (max, min))
        self.min_length, self.max_length = int(min), int(max)
        self._numlengths = self.max_length - self.min_length + 1

    @property
    def row_keys(self):
        return np.arange(self.min_length, self.max_length + 1)

    @property
    def shape(self):
        return [self._numlengths]

    def _leading_shape(self):
        return [self._numlengths]

    def _warn_dropped(self, n, length):
        pass


class SizeFilterFactory(object):
    """``SizeFilterFactory(min=1, max=-1)`` — map_factories.pyx:794-839.  Lowered into the kernels
    when added to a :class:`BAMGenomeArray`; also callable on a single read."""

    def __init__(self, min=1, max=-1):
        if max != -1 and max < min:
            raise ValueError("Alignment size filter: max read length must be >= min read length")
        if min < 1:
            raise ValueError("Alignment size filter: min read length must be >= 1. Got %s" % min)
        self.min_, self.max_ = int(min), int(max)

    def __call__(self, read):
        if read is None:
            raise TypeError("read may not be None")
        n = len(read.positions)
        return n >= self.min_ and (n <= self.max_ or self.max_ == -1)
