"""Host-side logic and the C-ABI surface (no GPU, no compute calls)."""
import ctypes as C
import os
import re

import warnings

import numpy as np
import pytest

import plastid_b200 as pb
from plastid_b200 import _lib, synth
from plastid_b200.batch import cigar_to_blocks, pack_reads, batch_from_arrays
from plastid_b200.genome_array import merge_batches
from plastid_b200.regions import ChainTable
from oracle import pyoracle as po
from helpers import random_cigar_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------- C-ABI surface
def declared_symbols():
    text = open(os.path.join(ROOT, "include", "plastid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 14
    handle = C.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), "libplastid_b200.so does not export %s" % name
    assert sorted(_lib.exported_symbols()) == names       # the ctypes table binds exactly the header


def test_abi_struct_layouts_match_header():
    assert C.sizeof(_lib.PbBatch) == 72 and C.sizeof(_lib.PbLayout) == 32 and C.sizeof(_lib.PbRule) == 40
    assert _lib.lib().pb_version().startswith(b"plastid_b200")


def test_argument_errors_are_reported_without_a_device():
    L = _lib.lib()
    assert L.pb_map_point(None, None, None, 3, None, None, None, None, None, 0, None) == _lib.PB_EINVAL
    assert b"null" in L.pb_last_error()
    assert L.pb_region_sums(None, 0, None, None, None, None, None, None, None, 0, 0, None, None, 0, 1, None, None, None, 0, None) == _lib.PB_EINVAL
    assert L.pb_map_workspace_bytes(16384 * 4, 0, 100) > 0
    # the annotation-side entry points validate before touching the device too
    assert L.pb_landmark_windows(None, None, None, None, None, None, 5, 50, 50, None, None, None) == _lib.PB_EINVAL
    assert L.pb_landmark_windows(None, None, None, None, None, None, 0, 50, 50, None, None, None) == _lib.PB_OK
    assert L.pb_spanning_windows(None, None, None, None, None, None, None, None, None, 3, 50, -1,
                                 None, None, None, None, None, None, None, None, None) == _lib.PB_EINVAL
    assert L.pb_chain_union(None, None, None, None, None, 2, None, None, None, None, None) == _lib.PB_EINVAL
    assert L.pb_chain_binary(7, None, None, None, None, None, None, None, None, 1, None, None, None, None, None) == _lib.PB_EINVAL
    assert b"pb_chain_binary" in L.pb_last_error()
    assert L.pb_map_workspace_bytes(16384 * 4, 1000, 100) >= L.pb_map_workspace_bytes(16384 * 4, 0, 100) + 16000


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(_lib.PlastidB200Error):
        _lib.require_cuda()
    reads = [po.Read(0, [(0, 30)], False)]
    with pytest.raises(_lib.PlastidB200Error):
        pb.FivePrimeMapFactory(0)(reads, pb.GenomicSegment("c", 0, 100, "+"))
    with pytest.raises(_lib.PlastidB200Error):
        pb.GenomeArray({"c": 100})


def test_product_never_imports_oracle():
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "plastid_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]|liboracle", text, flags=re.M), \
                    "%s reaches into oracle/" % fn


# ------------------------------------------------------------------------------- packing
def test_cigar_to_blocks():
    M, I, D, N, S, H, P, EQ, X = range(9)
    assert cigar_to_blocks([(M, 30)]) == ([(0, 30)], 30)
    assert cigar_to_blocks([(S, 3), (M, 10), (I, 2), (M, 10), (S, 4)]) == ([(0, 20)], 20)
    assert cigar_to_blocks([(M, 10), (D, 2), (M, 10)]) == ([(0, 10), (12, 10)], 22)
    assert cigar_to_blocks([(EQ, 5), (X, 1), (EQ, 4), (N, 100), (M, 7), (H, 9)]) == ([(0, 10), (110, 7)], 117)
    assert cigar_to_blocks([(M, 5), (P, 3), (M, 5)]) == ([(0, 10)], 10)
    assert cigar_to_blocks([(S, 30)]) == ([], 0)


def test_pack_reads_matches_oracle_positions():
    rng = np.random.default_rng(5)
    reads = {"a": random_cigar_reads(rng, 200, 5000, 3000), "b": random_cigar_reads(rng, 50, 2000, 1000)}
    hb = pack_reads(reads, {"a": 5000, "b": 2000})
    hb.check_sorted()
    assert len(hb) == 250 and list(hb.chrom_read_off) == [0, 200, 250]
    for i, r in enumerate(hb.objects):
        assert hb.positions_of(i) == r.positions
        assert hb.read_view(i) is r
        assert bool(hb.is_reverse[i]) == r.is_reverse and int(hb.aligned_len[i]) == len(r.positions)
    assert hb.max_span >= max(r.reference_end - r.positions[0] for r in hb.objects if r.positions)
    # a read whose CIGAR starts with a deletion: positions start after it
    odd = pack_reads({"a": [po.Read(100, [(2, 5), (0, 10)], False), po.Read(102, [(0, 4)], True)]}, {"a": 1000})
    assert list(odd.ref_start) == [102, 105] and odd.positions_of(1) == list(range(105, 115))
    with pytest.raises(ValueError):
        pack_reads({"a": [po.Read(0, [(0, 70000)], False)]}, {"a": 100000})
    with pytest.raises(ValueError):
        pack_reads({"a": [po.Read(0, [(0, 1), (3, 1)] * 300, False)]}, {"a": 100000})


def test_batch_from_arrays_and_merge():
    chroms, lens = ["x", "y"], [1000, 500]
    b1 = batch_from_arrays(chroms, lens, [1, 0, 0], [49, 999, 0], [1, 1, 30], [0, 1, 0])
    assert list(b1.ref_start) == [0, 999, 49] and list(b1.chrom_read_off) == [0, 2, 3]
    assert list(b1.aligned_len) == [30, 1, 1] and list(b1.is_reverse) == [False, True, False]
    nb = [1, 2]
    blk = [[0, 10], [0, 5], [50, 5]]
    b2 = batch_from_arrays(["y", "z"], [700, 50], [0, 0], [20, 5], [10, 10], [0, 1], blocks=(nb, blk))
    assert b2.blk is not None and b2.positions_of(0) == list(range(5, 10)) + list(range(55, 60))
    m = merge_batches([b1, b2])
    assert m.chroms == ["x", "y", "z"] and list(m.chrom_len) == [1000, 700, 50]
    assert list(m.chrom_read_off) == [0, 2, 5, 5] and m.mapped == 5
    assert list(m.ref_start) == [0, 999, 5, 20, 49]
    assert m.positions_of(2) == list(range(5, 10)) + list(range(55, 60)) and m.positions_of(3) == list(range(20, 30))
    d = m.with_drop_mask(np.array([0, 1, 0, 0, 1], dtype=bool))
    assert list((d.meta >> 17) & 1) == [0, 1, 0, 0, 1] and list(d.aligned_len) == list(m.aligned_len)
    empty = batch_from_arrays(chroms, lens, [], [], [], [])
    assert len(empty) == 0 and list(empty.chrom_read_off) == [0, 0, 0] and empty.max_span == 1


def test_genome_layout():
    lay = pb.GenomeLayout(["a", "b", "c"], [100, 16384, 16385])
    assert list(lay.chrom_bin_off) == [0, 16384, 32768, 65536] and lay.total_bins == 65536
    assert lay.bin_of("c", 7) == 32775
    with pytest.raises(ValueError):
        pb.GenomeLayout([], [])


# ------------------------------------------------------------------------------- factories (host side)
def test_factory_constructors_and_luts():
    with pytest.raises(ValueError):
        pb.FivePrimeMapFactory(-1)
    with pytest.raises(ValueError):
        pb.ThreePrimeMapFactory(-3)
    with pytest.raises(ValueError):
        pb.CenterMapFactory(-1)
    with pytest.raises(ValueError):
        pb.SizeFilterFactory(0, 10)
    with pytest.raises(ValueError):
        pb.SizeFilterFactory(30, 20)
    with pytest.raises(ValueError):
        pb.StratifiedVariableFivePrimeMapFactory({"default": 3}, 30, 30)
    f = pb.FivePrimeMapFactory(3)
    f.offset = 9
    assert f.offset == 9
    c = pb.CenterMapFactory(2)
    c.nibble = 5
    assert c.nibble == 5
    s = pb.StratifiedVariableFivePrimeMapFactory({"default": 3}, 26, 30)
    assert s.shape == [5] and list(s.row_keys) == [26, 27, 28, 29, 30]
    import warnings
    for d in ({"default": 0}, {25: 10, "default": 28}, {L: L // 2 for L in range(25, 40)}, dict(synth.RIBO_OFFSETS),
              {20: 25, 30: 40, "default": 27}, None):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fac = pb.VariableFivePrimeMapFactory(d)
            fw, rc = po.build_offset_luts(d)
        assert (fac.forward_offsets == fw).all() and (fac.reverse_offsets == rc).all()
    with pytest.raises(UnboundLocalError):
        pb.VariableFivePrimeMapFactory({20: 25})
    sf = pb.SizeFilterFactory(25, 30)
    assert [sf(po.Read(0, [(0, L)], False)) for L in (24, 25, 30, 31)] == [False, True, True, False]
    assert pb.SizeFilterFactory(25)(po.Read(0, [(0, 4000)], False))
    slot, inv = pb.CenterMapFactory(12).slot_tables(np.bincount([24, 25, 25, 30, 100], minlength=65536))
    assert list(inv) == [1.0, 1.0 / 6, 1.0 / 76] and slot[24] == -1 and list(slot[[25, 30, 100]]) == [0, 1, 2]


def test_offset_file_errors():
    import io
    good = "25\t12\n26\t13\ndefault\t14\n"
    fac = pb.VariableFivePrimeMapFactory.from_file(io.StringIO("# comment\n" + good))
    assert fac.forward_offsets[25] == 12 and fac.forward_offsets[40] == 14
    for bad in (good + good, "25\t12\t7\n", "25\n", "abc\t3\n", "25\t1.5\n"):
        with pytest.raises(pb.MalformedFileError):
            pb.VariableFivePrimeMapFactory.from_file(io.StringIO(bad))


# ------------------------------------------------------------------------------- SegmentChain
def test_segmentchain_layout_and_strings():
    ch = pb.SegmentChain(pb.GenomicSegment("chrA", 250, 300, "-"), pb.GenomicSegment("chrA", 100, 150, "-"),
                         pb.GenomicSegment("chrA", 150, 170, "-"), pb.GenomicSegment("chrA", 160, 180, "-"))
    assert str(ch) == "chrA:100-180^250-300(-)" and ch.length == 130 and len(ch) == 2
    assert str(pb.SegmentChain.from_str(str(ch))) == str(ch)
    assert str(pb.SegmentChain()) == "na" and len(pb.SegmentChain.from_str("na")) == 0
    o = po.Chain(po.Seg("chrA", 250, 300, "-"), po.Seg("chrA", 100, 150, "-"), po.Seg("chrA", 150, 170, "-"),
                 po.Seg("chrA", 160, 180, "-"))
    assert str(o) == str(ch) and o.position_list == ch.get_position_list()
    with pytest.raises(ValueError):
        pb.SegmentChain(pb.GenomicSegment("chrA", 0, 5, "+"), pb.GenomicSegment("chrB", 10, 15, "+"))
    with pytest.raises(ValueError):
        pb.SegmentChain(pb.GenomicSegment("chrA", 0, 5, "+"), pb.GenomicSegment("chrA", 10, 15, "-"))
    # coordinates (roitools.pyx:2957-3106)
    assert ch.get_genomic_coordinate(0) == ("chrA", 299, "-") and ch.get_genomic_coordinate(129) == ("chrA", 100, "-")
    assert ch.get_genomic_coordinate(50) == ("chrA", 179, "-")
    assert ch.get_genomic_coordinate(0, stranded=False) == ("chrA", 100, "-")
    assert ch.get_segmentchain_coordinate("chrA", 299, "-") == 0
    with pytest.raises(KeyError):
        ch.get_segmentchain_coordinate("chrA", 200, "-")
    with pytest.raises(IndexError):
        ch.get_genomic_coordinate(130)
    sub = ch.get_subchain(40, 60)
    assert str(sub) == "chrA:170-180^250-260(-)"
    assert str(pb.GenomicSegment.from_str("chrX:5-99(+)")) == "chrX:5-99(+)"


def test_add_masks_known_answers():
    """plastid/test/unit/genomics/test_roitools.py:976-1035 and :1382-1417, transcribed."""
    for strand, opp in (("+", "-"), ("-", "+")):
        for cls, seg in ((pb.SegmentChain, pb.GenomicSegment), (po.Chain, po.Seg)):
            ivc = cls(seg("chrA", 100, 150, strand), seg("chrA", 250, 300, strand))
            mask_a, mask_b, mask_c = seg("chrA", 125, 150, strand), seg("chrA", 275, 300, strand), seg("chrA", 275, 350, strand)

            def masks():
                if cls is pb.SegmentChain:
                    return [(m.start, m.end) for m in ivc.get_masks()]
                pm = ivc.position_mask or [0] * ivc.length
                cov = [p for p, m in zip(ivc.position_list, pm) if m]
                out = []
                for p in cov:
                    if out and out[-1][1] == p:
                        out[-1][1] = p + 1
                    else:
                        out.append([p, p + 1])
                return [tuple(x) for x in out]

            with pytest.raises(ValueError):
                ivc.add_masks(seg("chrA", 125, 150, opp))
            with pytest.raises(ValueError):
                ivc.add_masks(seg("chrB", 125, 150, strand), seg("chrB", 275, 300, strand))
            assert masks() == []
            ivc.add_masks(seg("chrA", 155, 245, strand))          # intron mask: nothing
            assert masks() == [] and ivc.masked_length == 100
            ivc.add_masks(mask_a)
            assert masks() == [(125, 150)]
            ivc.add_masks(mask_a)
            assert masks() == [(125, 150)]
            ivc.add_masks(mask_b)
            assert masks() == [(125, 150), (275, 300)] and ivc.masked_length == 50
            ivc = cls(seg("chrA", 100, 150, strand), seg("chrA", 250, 300, strand))
            ivc.add_masks(mask_a, mask_c)                          # trimmed to the chain
            assert masks() == [(125, 150), (275, 300)]
    ivc1 = pb.SegmentChain(pb.GenomicSegment("chrA", 100, 150, "+"), pb.GenomicSegment("chrA", 150, 200, "+"),
                           pb.GenomicSegment("chrA", 250, 350, "+"))
    assert (ivc1.length, ivc1.masked_length) == (200, 200)
    ivc1.add_masks(pb.GenomicSegment("chrA", 400, 500, "+"))
    assert ivc1.masked_length == 200
    ivc1.add_masks(pb.GenomicSegment("chrA", 50, 125, "+"))
    assert (ivc1.length, ivc1.masked_length) == (200, 175)
    assert ivc1.position_mask().sum() == 25 and ivc1.position_mask()[:25].all()


def test_chain_table_lowering():
    chroms, lens = ["a", "b"], [50000, 20000]
    lay = pb.GenomeLayout(chroms, lens)
    c1 = pb.SegmentChain(pb.GenomicSegment("a", 10, 20, "+"), pb.GenomicSegment("a", 30, 35, "+"))
    c2 = pb.SegmentChain(pb.GenomicSegment("b", 5, 9, "-"))
    c3 = pb.SegmentChain(pb.GenomicSegment("zz", 5, 9, "+"))
    c4 = pb.SegmentChain()
    c1.add_masks(pb.GenomicSegment("a", 18, 32, "+"))
    t = ChainTable.from_chains([c1, c2, c3, c4], lay)
    assert list(t.chain_off) == [0, 2, 3, 3, 3] and list(t.bstart) == [10, 30, 65536 + 5]
    assert list(t.chain_plane) == [0, 1, 0, 2] and list(t.chain_reverse) == [0, 1, 0, 0]
    assert list(t.chain_len) == [15, 4, 0, 0] and list(t.known) == [True, True, False, True]
    bits = np.unpackbits(t.mask_bits, bitorder="little")
    assert list(bits[:15]) == [0] * 8 + [1, 1] + [1, 1] + [0, 0, 0] and list(t.mask_off) == [0, 15, 19, 19]
    ann = synth.make_annotation(chroms, lens, 12, seed=3, exons=(1, 3), exon_len=(50, 200), intron_len=(20, 300))
    t1, t2 = synth.annotation_table(ann, lay), ChainTable.from_chains(ann.chains(), lay)
    for f in ("bstart", "bend", "chain_off", "chain_plane", "chain_reverse", "chain_len"):
        assert (getattr(t1, f) == getattr(t2, f)).all()


def test_synthetic_generators_are_seeded_and_sorted():
    chroms, lens = synth.yeast_like_genome(600_000, 4)
    ann = synth.make_annotation(chroms, lens, 60, seed=1, exons=(1, 2), exon_len=(200, 600), intron_len=(50, 400))
    a = synth.device_batch_to_host(synth.riboseq_reads(ann, 20000, seed=3, device="cpu"), chroms, lens)
    b = synth.device_batch_to_host(synth.riboseq_reads(ann, 20000, seed=3, device="cpu"), chroms, lens)
    a.check_sorted()
    assert (a.ref_start == b.ref_start).all() and (a.meta == b.meta).all()
    assert a.aligned_len.min() >= 25 and a.aligned_len.max() <= 35
    assert ((a.ref_start + a.aligned_len) <= np.repeat(lens, np.diff(a.chrom_read_off))).all() and a.ref_start.min() >= 0
    r = synth.device_batch_to_host(synth.rnaseq_reads(chroms, lens, 5000, seed=2, device="cpu"), chroms, lens)
    r.check_sorted()
    nb = (r.meta >> 24).astype(int)
    assert set(np.unique(nb)) == {1, 2, 3} and (np.diff(r.blk_off.astype(int)) == np.where(nb > 1, nb, 0)).all()
    assert all(len(r.positions_of(i)) == 100 for i in range(0, 5000, 97))


def test_wire16_host_format_round_trip():
    from plastid_b200.batch import Wire16Batch
    chroms, lens = ["a", "b", "c"], np.array([200_000, 65536, 70_000])
    ann = synth.make_annotation(chroms, lens, 30, seed=1, exons=(1, 2), exon_len=(200, 600), intron_len=(50, 400))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, 30000, seed=3, device="cpu"), chroms, lens)
    hb = hb.with_drop_mask(np.arange(len(hb)) % 17 == 0)
    w = Wire16Batch.from_batch(hb)
    assert w.nbytes < 0.55 * (hb.ref_start.nbytes + hb.meta.nbytes)
    assert len(w.seg_base) == 4 + 1 + 2 and list(w.seg_base) == [0, 65536, 131072, 196608, 0, 0, 65536]
    seg = np.searchsorted(w.seg_off, np.arange(len(w)), side="right") - 1
    assert (w.seg_base[seg].astype(np.int64) + w.start_lo == hb.ref_start).all()
    m = w.meta16.astype(np.uint32)
    meta = (m & 0x3FFF) | (((m >> 14) & 1) << 16) | (((m >> 15) & 1) << 17) | (1 << 24)
    assert (meta == hb.meta).all()
    spliced = synth.device_batch_to_host(synth.rnaseq_reads(chroms, lens, 2000, seed=2, device="cpu"), chroms, lens)
    with pytest.raises(ValueError):
        Wire16Batch.from_batch(spliced)


def _delta8_world(n_reads=30000, rare_meta=True):
    chroms, lens = ["a", "b", "c", "d"], np.array([200_000, 65536, 70_000, 3_000_000])
    ann = synth.make_annotation(chroms, lens, 40, seed=1, exons=(1, 2), exon_len=(200, 600), intron_len=(50, 400))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, n_reads, seed=3, device="cpu"), chroms, lens)
    hb = hb.with_drop_mask(np.arange(len(hb)) % 17 == 0)
    if rare_meta:      # > 255 distinct meta words: the rare ones must travel as exceptions
        meta = hb.meta.copy()
        odd = np.arange(len(hb)) % 29 == 0
        meta[odd] = (meta[odd] & ~np.uint32(0xFFFF)) | (300 + (np.arange(odd.sum()) % 700)).astype(np.uint32)
        from plastid_b200.batch import AlignmentBatch
        hb = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start, meta, hb.chrom_read_off, max_span=2000)
    return chroms, lens, hb


def test_delta8_host_format_round_trip():
    from plastid_b200.batch import Delta8Batch
    from helpers import delta8_decode
    chroms, lens, hb = _delta8_world()
    w = Delta8Batch.from_batch(hb)
    assert len(w.dstart) % 128 == 0 and len(w.dstart) == len(w.code) >= len(hb)
    assert len(np.unique(hb.meta)) > 255 and len(w.exc_start) == int(w.blk_exc_off[-1]) > 0
    start, meta = delta8_decode(w)
    assert (start == hb.ref_start).all() and (meta == hb.meta).all()
    # the common case is 2 bytes per read + block tables; exceptions stay a small share
    _c, _l, plain = _delta8_world(rare_meta=False)
    wp = Delta8Batch.from_batch(plain)
    assert wp.nbytes < 0.36 * (plain.ref_start.nbytes + plain.meta.nbytes)
    s2, m2 = delta8_decode(wp)
    assert (s2 == plain.ref_start).all() and (m2 == plain.meta).all()
    # degenerate batches: empty, one read, exactly one block
    for n in (0, 1, 128, 129):
        sub = pb.batch_from_arrays(["a", "b"], [1000, 1000], [0] * (n // 2) + [1] * (n - n // 2),
                                   sorted(range(n // 2)) + sorted(range(n - n // 2)), [30] * n, [i % 2 for i in range(n)])
        ws = Delta8Batch.from_batch(sub)
        s3, m3 = delta8_decode(ws)
        assert (s3 == sub.ref_start).all() and (m3 == sub.meta).all()
    spliced = synth.device_batch_to_host(synth.rnaseq_reads(chroms, lens, 2000, seed=2, device="cpu"), chroms, lens)
    with pytest.raises(ValueError):
        Delta8Batch.from_batch(spliced)


def test_delta3_host_format_round_trip():
    from plastid_b200.batch import Delta3Batch, Delta3Receiver
    from helpers import delta3_decode
    chroms, lens, hb = _delta8_world()                  # > 31 distinct meta words: rare ones become exceptions
    w = Delta3Batch.from_batch(hb)
    assert len(w.packed) % 128 == 0 and int(w.blk_exc_off[-1]) == len(w.exc_start) > 0 and int(w.blk_wide_off[-1]) == len(w.wide)
    start, meta = delta3_decode(w)
    assert (start == hb.ref_start).all() and (meta == hb.meta).all()
    _c, _l, plain = _delta8_world(rare_meta=False)
    wp = Delta3Batch.from_batch(plain)
    assert wp.nbytes < 0.25 * (plain.ref_start.nbytes + plain.meta.nbytes)       # < 2 bytes per read even on a sparse batch
    s2, m2 = delta3_decode(wp)
    assert (s2 == plain.ref_start).all() and (m2 == plain.meta).all()
    for n in (0, 1, 128, 129):
        sub = pb.batch_from_arrays(["a", "b"], [100000, 1000], [0] * (n // 2) + [1] * (n - n // 2),
                                   sorted(x * 97 % 5000 for x in range(n // 2)) + sorted(range(n - n // 2)), [30] * n,
                                   [i % 2 for i in range(n)])
        ws = Delta3Batch.from_batch(sub)
        s3, m3 = delta3_decode(ws)
        assert (s3 == sub.ref_start).all() and (m3 == sub.meta).all()
    lay = pb.GenomeLayout(chroms, lens)
    plan = Delta3Receiver.plan_chunks(w, lay, 8)
    assert plan[0][0] == 0 and plan[-1][1] == len(w) and plan[-1][3] == lay.total_bins
    # the library's multithreaded host encoder writes the same streams as the numpy statement of the format
    fields = ("packed", "wide", "blk_base", "blk_wide_off", "blk_exc_off", "exc_start", "exc_meta", "meta_dict",
              "blk_chrom", "blk_first_start")
    empty_chrom = pb.batch_from_arrays(["a", "e", "b"], [100000, 50, 1000], [0] * 300 + [2] * 200,
                                       sorted(x * 97 % 5000 for x in range(300)) + sorted(range(200)), [30] * 500, [0] * 500)
    for batch in (hb, plain, empty_chrom, pb.batch_from_arrays(["a"], [10], [], [], [], [])):
        for threads in (1, 3):
            nat, ref = Delta3Batch.from_batch(batch, native=True, threads=threads), Delta3Batch.from_batch(batch, native=False)
            for f in fields:
                assert np.array_equal(getattr(nat, f), getattr(ref, f)), f
            assert nat.nbytes == ref.nbytes


def test_delta8_chunk_plan_covers_reads_and_bins():
    from plastid_b200.batch import Delta8Batch, Delta8Receiver
    chroms, lens = synth.human_like_genome(0.004)
    ann = synth.make_annotation(chroms, lens, 300, seed=1, exons=(1, 2), exon_len=(200, 600), intron_len=(50, 400))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, 50000, seed=3, device="cpu"), chroms, lens)
    wire, lay = Delta8Batch.from_batch(hb), pb.GenomeLayout(chroms, lens)
    c_of = np.searchsorted(hb.chrom_read_off, np.arange(len(hb)), side="right") - 1
    g = lay.chrom_bin_off[c_of] + hb.ref_start
    for k in (1, 2, 8, 64):
        plan = Delta8Receiver.plan_chunks(wire, lay, k)
        assert 1 <= len(plan) <= k and plan[0][0] == 0 and plan[0][2] == 0
        assert plan[-1][1] == len(wire) and plan[-1][3] == lay.total_bins
        for (a, b, x, y), (a2, b2, x2, y2) in zip(plan[:-1], plan[1:]):
            assert b == a2 and y == x2 and x % _lib.PB_LAYOUT_ALIGN == 0 and a % 128 == 0 and a < b and x < y
        for a, b, x, y in plan:
            # bins below y are final once reads [0, b) have landed: nothing later starts below y
            assert (g[b:] >= y).all()


def test_wire16_chunk_plan_covers_reads_and_bins():
    from plastid_b200.batch import Wire16Batch, Wire16Receiver
    chroms, lens = synth.human_like_genome(0.004)
    ann = synth.make_annotation(chroms, lens, 300, seed=1, exons=(1, 2), exon_len=(200, 600), intron_len=(50, 400))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, 50000, seed=3, device="cpu"), chroms, lens)
    wire, lay = Wire16Batch.from_batch(hb), pb.GenomeLayout(chroms, lens)
    for k in (1, 2, 8, 64):
        plan = Wire16Receiver.plan_chunks(wire, lay, k)
        assert 1 <= len(plan) <= k and plan[0][0] == 0 and plan[0][2] == 0
        assert plan[-1][1] == len(wire) and plan[-1][3] == lay.total_bins
        for (a, b, x, y), (a2, b2, x2, y2) in zip(plan[:-1], plan[1:]):
            assert b == a2 and y == x2 and x % _lib.PB_LAYOUT_ALIGN == 0 and a <= b and x < y
        for a, b, x, y in plan:
            c = np.searchsorted(hb.chrom_read_off, np.arange(a, b), side="right") - 1
            g = lay.chrom_bin_off[c] + hb.ref_start[a:b]
            assert (g >= x).all() and (g < y).all()


def test_mask_index_merges_per_strand_in_global_bins():
    from plastid_b200.masks import MaskIndex
    lay = pb.GenomeLayout(["a", "b"], [50_000, 20_000])
    seg = pb.GenomicSegment
    feats = [pb.SegmentChain(seg("b", 10, 20, "+"), seg("b", 100, 120, "+")), pb.SegmentChain(seg("a", 5, 15, "-")),
             pb.SegmentChain(seg("b", 15, 30, "+")), pb.SegmentChain(seg("a", 15, 18, "-")),      # touching: merged
             pb.SegmentChain(seg("zz", 1, 2, "+")), pb.SegmentChain(seg("a", 40, 50, "+"))]
    mi = MaskIndex(feats, lay)
    base_b = int(lay.chrom_bin_off[1])
    assert list(mi.class_off) == [0, 3, 4, 4]
    assert list(zip(mi.mask_start, mi.mask_end)) == [(40, 50), (base_b + 10, base_b + 30), (base_b + 100, base_b + 120), (5, 18)]
    with pytest.raises(KeyError):
        MaskIndex([pb.SegmentChain(seg("a", 1, 2, "."))], lay)
    assert len(MaskIndex([], lay)) == 0


def test_oracle_genome_hash_overlap_query():
    """genome_hash.py:259-436 restated: bins by chromosome and strand, true position overlap, all
    segments of a hit feature are returned (and masked only where they meet the region)."""
    from oracle import pyoracle as po
    S = po.Seg
    roi = po.Chain(S("c", 100, 200, "+"), S("c", 300, 400, "+"))
    feats = [po.Chain(S("c", 150, 160, "+")),                       # inside exon 1
             po.Chain(S("c", 200, 300, "+")),                       # intron only: no shared position
             po.Chain(S("c", 390, 30000, "+"), S("c", 50, 60, "+")),  # second exon + a far segment
             po.Chain(S("c", 150, 160, "-")),                       # other strand
             po.Chain(S("d", 150, 160, "+")),                       # other chromosome
             po.Chain(S("c", 45000, 45010, "+"))]                   # same hash neighbourhood rules, no overlap
    gh = po.GenomeHash(feats)
    hits = gh.get_overlapping_features(roi)
    assert hits == [feats[0], feats[2]]
    roi.add_masks(*[s for f in hits for s in f.segments])
    assert roi.masked_length == 200 - 10 - 10 and sum(roi.position_mask[50:60]) == 10 and sum(roi.position_mask[190:]) == 10
    assert gh.get_overlapping_features(po.Chain(S("c", 100, 200, "."))) == []
    with pytest.raises(KeyError):
        po.GenomeHash([po.Chain(S("c", 1, 5, "."))])


# ---------------------------------------------------------------------------------------------
# metagene generate: host objects (Transcript, window functions, GenomeHash, lowering)
# ---------------------------------------------------------------------------------------------
def _golden_transcripts():
    from helpers import metagene_generate_golden, gff3_transcript_records
    gold = metagene_generate_golden()
    txs = {}
    for name, r in gff3_transcript_records(gold["transcripts_gff"]).items():
        segs = [pb.GenomicSegment(r["chrom"], s, e, r["strand"]) for s, e in r["segments"]]
        txs[name] = pb.Transcript(*segs, ID=name, gene_id=r["gene_id"], cds_genome_start=r["cds_genome_start"],
                                  cds_genome_end=r["cds_genome_end"])
    return gold, txs


@pytest.mark.parametrize("which", ["cds_start", "cds_stop", "cds_stop_with_delta"])
def test_window_functions_match_reference_tables(which):
    """plastid/test/unit/bin/test_metagene.py:118-176 (data in tests/golden/metagene_generate.json)."""
    from plastid_b200.bin import metagene as mg
    gold, txs = _golden_transcripts()
    func = mg.window_cds_start if which == "cds_start" else mg.window_cds_stop
    for up, down in gold["flanks"]:
        for txid in gold["cds_start_queries" if which == "cds_start" else "cds_stop_queries"]:
            known = gold[which + "_results"]["%s_%s_%s" % (txid, up, down)]
            roi, offset, ref = func(txs[txid], up, down, ref_delta=3 if which.endswith("delta") else 0)
            assert str(roi) == str(pb.SegmentChain.from_str(known[0]))
            if known[1] is None or known[2] is None:
                assert np.isnan(offset) and np.isnan(ref)
            else:
                assert offset == known[1] and tuple(ref) == tuple(known[2])


def test_transcript_cds_coordinates_and_subregions():
    """Transcript._update_cds (roitools.pyx:3883-3913), get_cds / get_utr5 / get_utr3 (:4005-4130)."""
    for strand in "+-":
        tx = pb.Transcript(pb.GenomicSegment("chrA", 100, 200, strand), pb.GenomicSegment("chrA", 300, 400, strand),
                           ID="tx", cds_genome_start=150, cds_genome_end=350)
        assert (tx.cds_start, tx.cds_end) == (50, 150)
        assert tx.get_cds().length == 100 and tx.get_utr5().length == 50 and tx.get_utr3().length == 50
        assert str(tx.get_cds()) == "chrA:150-200^300-350(%s)" % strand
        assert str(tx.get_utr5()) == ("chrA:100-150(+)" if strand == "+" else "chrA:350-400(-)")
        assert tx.get_gene() == "gene_tx" and tx.get_cds().get_name() == "tx_CDS"
    # half-open CDS end on an exon end (the KeyError branch of _update_cds)
    tx = pb.Transcript(pb.GenomicSegment("chrA", 100, 200, "+"), pb.GenomicSegment("chrA", 300, 400, "+"),
                       cds_genome_start=120, cds_genome_end=200)
    assert (tx.cds_start, tx.cds_end) == (20, 100)
    nc = pb.Transcript(pb.GenomicSegment("chrA", 100, 200, "+"), ID="nc")
    assert nc.cds_start is None and len(nc.get_cds()) == 0 and len(nc.get_utr3()) == 0
    assert pb.Transcript(pb.GenomicSegment("c", 1, 5, "+"), Parent=["g2", "g1"]).get_gene() == "g1,g2"
    # get_subchain slices like the reference's position hash: out-of-range bounds clamp
    ch = pb.SegmentChain(pb.GenomicSegment("chrA", 100, 150, "+"), ID="x")
    assert str(ch.get_subchain(40, 80)) == "chrA:140-150(+)" and len(ch.get_subchain(60, 80)) == 0
    assert ch.get_subchain(0, 10).get_name() == "x_subchain"
    assert [str(s) for s in pb.positions_to_segments("c", "+", {5, 3, 4, 9, 10, 20})] == ["c:3-6(+)", "c:9-11(+)", "c:20-21(+)"]
    bed = pb.SegmentChain(pb.GenomicSegment("c", 10, 20, "-"), pb.GenomicSegment("c", 30, 45, "-"), ID="w", thickstart=12,
                          thickend=13).as_bed()
    assert bed == "c\t10\t45\tw\t0\t-\t12\t13\t0,0,0\t2\t10,15,\t0,20,\n"


def test_genome_hash_and_transcript_table_lowering():
    from plastid_b200.windows import TranscriptTable, layout_for_features
    from plastid_b200.masks import mask_intervals_of_chains
    from plastid_b200.regions import ChainTable
    import torch
    gold, txs = _golden_transcripts()
    masks = [pb.SegmentChain.from_str(m) for m in gold["masks"]]
    gh = pb.GenomeHash(masks)
    roi = pb.SegmentChain.from_str("2L:7985674-7985768^7985833-7985839(+)")
    assert [str(m) for m in gh[roi]] == ["2L:7985694-7985744(+)"]
    assert gh.get_overlapping_features(pb.SegmentChain.from_str("2L:7985674-7985768(-)")) == []
    with pytest.raises(KeyError):
        pb.GenomeHash([pb.SegmentChain.from_str("2L:5-10(.)")])
    names = ["FBtr0079531", "FBtr0081950", "FBtr0081950_no_cds"]
    layout = layout_for_features([txs[n] for n in names], masks)
    assert layout.chroms == ["2L", "3R", "4"]
    table = TranscriptTable.from_transcripts([txs[n] for n in names], layout,
                                             [txs[n].cds_start for n in names])
    assert table.n_tx == 3 and list(table.reverse) == [0, 1, 1] and list(table.landmark[:2]) == [txs[names[0]].cds_start, txs[names[1]].cds_start]
    assert table.landmark[2] == -1
    for t, n in enumerate(names):                                  # bcum restarts per transcript
        k0, k1 = int(table.tx_off[t]), int(table.tx_off[t + 1])
        lens = table.bend[k0:k1] - table.bstart[k0:k1]
        assert list(table.bcum[k0:k1]) == list(np.cumsum(lens) - lens) and lens.sum() == txs[n].length
    # an unstranded transcript gets orientation code 2: coordinates like '+', window columns like '-' (metagene.py:443-455)
    assert list(TranscriptTable.from_transcripts([pb.SegmentChain(pb.GenomicSegment("2L", 5, 9, "."))], layout, [None]).reverse) == [2]
    # mask bits -> intervals, in the layout pb_mask_chains writes (bit mask_off[c] + j, genomic order)
    chains = [roi, pb.SegmentChain.from_str("3R:4519776-4519894(-)")]
    ctable = ChainTable.from_chains(chains, layout)
    flat = np.zeros(((roi.length + chains[1].length + 31) // 32) * 32 + 32, dtype=np.uint8)
    flat[20:70] = 1                                                # 2L:7985694-7985744
    flat[roi.length + 103:roi.length + 115] = 1                    # 3R:4519879-4519891
    bits = torch.from_numpy(np.packbits(flat, bitorder="little"))
    assert mask_intervals_of_chains(ctable, bits) == [[(7985694, 7985744)], [(4519879, 4519891)]]


def test_block_words_encode_the_block_table():
    """Delta3SplicedBatch: 4-byte block words (12-bit length, 20-bit gap, exception rows) decode back to
    the SoA block table (host statement of what pb_unpack_blocks does on the device)."""
    from plastid_b200.batch import Delta3SplicedBatch
    rng = np.random.default_rng(12)
    reads = random_cigar_reads(rng, 2000, 400_000, 300_000)
    reads.append(po.Read(1000, [(0, 30), (3, 2_000_000), (0, 20)], False))          # intron beyond 20 bits
    reads.append(po.Read(2000, [(0, 5000), (3, 70), (0, 4095), (2, 1), (0, 4096)], True))   # lengths around 12 bits
    hb = pack_reads({"chrA": reads}, {"chrA": 5_000_000}, keep_objects=False)
    wire = Delta3SplicedBatch.from_batch(hb)
    assert len(wire.bwords) == len(hb.blk) and 2 <= len(wire.bexc_row) <= 4
    assert wire.nbytes == wire.base.nbytes + 4 * len(hb.blk) + 12 * len(wire.bexc_row)
    exc = dict(zip(wire.bexc_row.tolist(), wire.bexc.tolist()))
    k, out = 0, []
    for i in range(len(hb)):
        nb = int(hb.meta[i] >> 24)
        if nb <= 1:
            continue
        pos = 0
        for _ in range(nb):
            word = int(wire.bwords[k])
            gap, ln = exc[k] if word == 0xFFFFFFFF else (word >> 12, word & 0xFFF)
            out.append((pos + gap, ln))
            pos += gap + ln
            k += 1
    assert out == [tuple(x) for x in hb.blk.tolist()]
    # a batch without multi-block reads ships no block words
    plain = pack_reads({"chrA": [po.Read(5, [(0, 30)], False), po.Read(9, [(0, 28)], True)]}, {"chrA": 1000}, keep_objects=False)
    assert len(Delta3SplicedBatch.from_batch(plain).bwords) == 0


def test_phase_table_reproduces_the_rows_printed_in_the_reference_docs():
    """docs/source/examples/phasing.rst:194-205 prints a `phase_by_size` table (SRR609197): with the
    per-length phase counts that table implies, `phase_table` reproduces every printed
    ``fraction_reads_counted`` and ``phase0..2`` value in the reference's ``%.6f`` format
    (phase_by_size.py:220-232, 243-251)."""
    from plastid_b200.bin.phase_by_size import phase_table
    doc = """25 6511 0.009640 0.326832 0.327599 0.345569
             26 9952 0.014735 0.385953 0.295217 0.318830
             27 17636 0.026111 0.320934 0.282717 0.396348
             28 42976 0.063629 0.251792 0.381794 0.366414
             29 93754 0.138809 0.309309 0.370971 0.319720
             30 148400 0.219716 0.318733 0.367635 0.313632
             31 155684 0.230501 0.336624 0.421713 0.241663
             32 118565 0.175543 0.445578 0.374141 0.180281
             33 58761 0.087000 0.511121 0.299076 0.189803
             34 18818 0.027861 0.508237 0.276597 0.215166
             35 4360 0.006455 0.514450 0.236468 0.249083"""
    rows = [line.split() for line in doc.splitlines()]
    sums = {}
    for length, reads, _frac, p0, p1, p2 in rows:
        counts = np.array([round(float(p) * int(reads)) for p in (p0, p1, p2)], dtype=np.float64)
        assert counts.sum() == int(reads)                      # the printed fractions imply integer counts
        sums[int(length)] = counts
    lengths, counted, frac, phases = phase_table(sums)
    for i, (length, reads, f, p0, p1, p2) in enumerate(rows):
        assert lengths[i] == int(length) and counted[i] == int(reads)
        assert "%.6f" % frac[i] == f
        assert ["%.6f" % x for x in phases[i]] == [p0, p1, p2]


def test_merge_and_positions_to_segments_reference_known_answers():
    """test_roitools.py:212-268 and :356-398 for the host objects and the oracle's restatements."""
    from helpers import merge_segments_known_answers, positions_to_segments_known_answers
    from oracle import generate as og
    for strand in "+-.":
        for positions, expected in positions_to_segments_known_answers():
            got = pb.positions_to_segments("chrA", strand, positions)
            assert [(s.chrom, s.start, s.end, s.strand) for s in got] == [("chrA", a, b, strand) for a, b in expected]
            got = og.positions_to_segments("chrA", strand, positions)
            assert [(s.chrom, s.start, s.end, s.strand) for s in got] == [("chrA", a, b, strand) for a, b in expected]
    for segs, expected in merge_segments_known_answers():
        chain = pb.SegmentChain(*[pb.GenomicSegment("chrA", a, b, "+") for a, b in segs])
        assert [(s.start, s.end) for s in chain] == expected
        ochain = po.Chain(*[po.Seg("chrA", a, b, "+") for a, b in segs])
        assert [(s.start, s.end) for s in ochain.segments] == expected


def test_segmentchain_coordinate_known_answers_from_the_reference_tests():
    """test_roitools.py:1120-1273 (get_segmentchain_coordinate, get_genomic_coordinate, get_subchain —
    incl. the sub-range 2..27 of a 10-nt chain, which clamps like a python slice), transcribed; run
    against the host SegmentChain and the oracle's restatement."""
    from oracle import generate as og
    for strand in "+-":
        ivc = pb.SegmentChain(*[pb.GenomicSegment("chrA", a, b, strand) for a, b in ((20, 40), (60, 70), (80, 90))])
        och = po.Chain(*[po.Seg("chrA", a, b, strand) for a, b in ((20, 40), (60, 70), (80, 90))])
        c = 0
        for iv in ivc:
            for x in range(iv.start, iv.end):
                expected = c if strand == "+" else ivc.length - 1 - c
                assert ivc.get_segmentchain_coordinate("chrA", x, strand, stranded=True) == expected
                assert ivc.get_segmentchain_coordinate("chrA", x, strand, stranded=False) == c
                assert og.get_segmentchain_coordinate(och, x, stranded=True) == expected
                assert og.get_segmentchain_coordinate(och, x, stranded=False) == c
                c += 1
    blocks = ((2, 3), (15, 19), (20, 24), (29, 30))
    ivca = pb.SegmentChain(*[pb.GenomicSegment("chrA", a, b, "+") for a, b in blocks])
    nvca = pb.SegmentChain(*[pb.GenomicSegment("chrA", a, b, "-") for a, b in blocks])
    oca = po.Chain(*[po.Seg("chrA", a, b, "+") for a, b in blocks])
    ocn = po.Chain(*[po.Seg("chrA", a, b, "-") for a, b in blocks])
    assert ivca.get_genomic_coordinate(0)[1] == 2 and ivca.get_genomic_coordinate(ivca.length - 1)[1] == 29
    assert ivca.get_genomic_coordinate(0, stranded=False)[1] == 2
    assert nvca.get_genomic_coordinate(0)[1] == 29 and nvca.get_genomic_coordinate(nvca.length - 1)[1] == 2
    assert nvca.get_genomic_coordinate(0, stranded=False)[1] == 2 and nvca.get_genomic_coordinate(9, stranded=False)[1] == 29
    for ivc, och in ((ivca, oca), (nvca, ocn)):
        positions = set(ivc.get_genomic_coordinate(i)[1] for i in range(ivc.length))
        assert [(s.start, s.end) for s in pb.positions_to_segments(ivc.chrom, ivc.strand, positions)] == list(blocks)
        for i in range(ivc.length):
            for stranded in (True, False):
                x = ivc.get_genomic_coordinate(i, stranded=stranded)[1]
                assert ivc.get_segmentchain_coordinate(ivc.chrom, x, ivc.strand, stranded=stranded) == i
                assert og.get_genomic_coordinate(och, i, stranded=stranded)[1] == x
    assert str(ivca.get_subchain(0, ivca.length)) == str(ivca) and str(nvca.get_subchain(0, nvca.length)) == str(nvca)
    assert ivca.get_subchain(2, 27).get_position_set() == {16, 17, 18, 20, 21, 22, 23, 29}
    assert nvca.get_subchain(2, 27).get_position_set() == {2, 15, 16, 17, 18, 20, 21, 22}
    assert set(og.get_subchain(oca, 2, 27).position_list) == {16, 17, 18, 20, 21, 22, 23, 29}
    assert set(og.get_subchain(ocn, 2, 27).position_list) == {2, 15, 16, 17, 18, 20, 21, 22}


def test_spliced_chunk_plan_and_read_windows():
    """Host side of map_center_streamed: the chunk plan of a spliced batch (bins below bin_b are final once
    reads [0, read_b) have landed) and the read window of every chunk (no read before it can reach the
    chunk's first bin), checked against the reads themselves."""
    from plastid_b200.batch import Delta3SplicedBatch, Delta3SplicedReceiver
    chroms, lens = synth.human_like_genome(0.005)
    hb = synth.device_batch_to_host(synth.rnaseq_reads(chroms, lens, 300_000, seed=100, device="cpu"), chroms, lens)
    wire = Delta3SplicedBatch.from_batch(hb)
    lay = pb.GenomeLayout(chroms, lens)
    assert wire.length_hist.sum() == len(hb) and (wire.row_of_read == hb.blk_off).all()
    c_of = np.searchsorted(hb.chrom_read_off, np.arange(len(hb)), side="right") - 1
    g = lay.chrom_bin_off[c_of] + hb.ref_start                      # global bin of every read's start

    class Stub(object):                                             # first_read_reaching needs only .wire
        pass
    rx = Stub()
    rx.wire = wire
    for n in (1, 5, 12):
        chunks = Delta3SplicedReceiver.plan_chunks(wire, lay, n)
        assert chunks[0][0] == 0 and chunks[0][2] == 0 and chunks[-1][1] == len(hb) and chunks[-1][3] == lay.total_bins
        for (a, b, bin_a, bin_b), nxt in zip(chunks, chunks[1:] + [None]):
            assert a % 128 == 0 and bin_a % 16384 == 0 and bin_a < bin_b and a < b
            if nxt is not None:
                assert nxt[0] == b and nxt[2] == bin_b and g[b:].min() >= bin_b      # later reads start beyond bin_b
            i0 = Delta3SplicedReceiver.first_read_reaching(rx, bin_a, lay)
            reaching = np.flatnonzero(g + hb.max_span > bin_a)
            assert i0 % 128 == 0 and (len(reaching) == 0 or i0 <= reaching[0])
            assert i0 <= a


def test_mask_accessors_known_answers_from_the_reference_tests():
    """test_roitools.py:1037-1095 (get_masks, get_masks_as_segmentchain, reset_masks), transcribed."""
    for strand in ("+", "-"):
        def chain():
            return pb.SegmentChain(pb.GenomicSegment("chrA", 100, 150, strand), pb.GenomicSegment("chrA", 250, 300, strand))
        mask_a, mask_b = pb.GenomicSegment("chrA", 125, 150, strand), pb.GenomicSegment("chrA", 275, 300, strand)
        ch = chain()
        assert ch.get_masks() == []
        assert len(ch.get_masks_as_segmentchain()) == 0 and isinstance(ch.get_masks_as_segmentchain(), pb.SegmentChain)
        ch.add_masks(mask_a)
        assert ch.get_masks() == [mask_a]
        assert str(ch.get_masks_as_segmentchain()) == str(pb.SegmentChain(mask_a))
        ch.reset_masks()
        assert ch.get_masks() == []
        ch.add_masks(mask_a, mask_b)
        assert ch.get_masks() == [mask_a, mask_b]
        assert str(ch.get_masks_as_segmentchain()) == str(pb.SegmentChain(mask_a, mask_b))
        pre_length = ch.length
        assert ch.masked_length == pre_length - len(mask_a) - len(mask_b)
        ch.reset_masks()
        assert ch.get_masks() == [] and ch.masked_length == pre_length


class _HostGenomeArray(object):
    """The smallest valid `ga` of the reference's contract: anything with ``get(seg, roi_order)``
    (`roitools.pyx:3259-3268` calls exactly ``ga.get(seg, roi_order=False)``).  Optional leading axis."""

    def __init__(self, length, rows=None):
        shape = (length,) if rows is None else (rows, length)
        self.data = {s: np.zeros(shape) for s in "+-"}

    def set(self, seg, val):
        self.data[seg.strand][..., seg.start:seg.end] = val

    def get(self, seg, roi_order=True):
        out = self.data[seg.strand][..., seg.start:seg.end].copy()
        return out[..., ::-1] if (roi_order and seg.strand == "-") else out


@pytest.mark.parametrize("strand", ["+", "-"])
def test_get_counts_and_masked_counts_over_any_ga_object(strand):
    """test_roitools.py:1274-1337 (get_masked_counts plus / minus) and :1340-1417 (counts 200 -> 175 under a mask)
    through the generic per-exon path of SegmentChain.get_counts — a host object stands in for the genome array."""
    S = pb.GenomicSegment
    ga = _HostGenomeArray(2000)
    ga.set(S("chrA", 100, 200, strand), 1)
    ga.set(S("chrA", 250, 350, strand), 5)
    chain = pb.SegmentChain(S("chrA", 100, 150, strand), S("chrA", 150, 200, strand), S("chrA", 250, 350, strand))
    unmasked = np.zeros(chain.length)
    if strand == "+":
        unmasked[:100], unmasked[100:200] = 1, 5
    else:
        unmasked[-100:], unmasked[-200:-100] = 1, 5
    assert (chain.get_counts(ga) == unmasked).all() and chain.get_counts(ga).dtype == np.float64
    assert (chain.get_counts(ga, stranded=False) == (unmasked if strand == "+" else unmasked[::-1])).all()
    assert (chain.get_masked_counts(ga) == unmasked).all()
    chain.add_masks(S("chrA", 400, 500, strand))
    assert (chain.get_masked_counts(ga) == unmasked).all() and chain.masked_length == 200
    chain.add_masks(S("chrA", 50, 125, strand))
    mask = np.tile(False, chain.length)
    if strand == "+":
        mask[:25] = True
    else:
        mask[-25:] = True
    found = chain.get_masked_counts(ga)
    assert (found.mask == mask).all() and (found.data == unmasked).all()
    assert found.sum() == 575 and chain.get_counts(ga).sum() == 600 and chain.masked_length == 175
    # `stranded` is ignored by get_masked_counts, like the reference (roitools.pyx:3301)
    assert (chain.get_masked_counts(ga, stranded=False).data == unmasked).all()
    # leading (stratification) axes are kept, the mask is broadcast over them (roitools.pyx:3259-3268, 3308-3313)
    ga2 = _HostGenomeArray(2000, rows=3)
    for r in range(3):
        ga2.data[strand][r, 100:200] = r + 1
    found2 = chain.get_masked_counts(ga2)
    assert found2.shape == (3, 200) and (found2.mask == mask[None, :]).all()
    assert list(found2.sum(axis=1)) == [75.0, 150.0, 225.0]


def test_empty_chain_counts_warn_and_return_an_empty_vector():
    chain = pb.SegmentChain()
    with pytest.warns(pb.DataWarning, match="zero-length"):
        out = chain.get_counts(_HostGenomeArray(10))
    assert out.shape == (0,) and out.dtype == np.float64


def test_genomic_segment_ordering_and_hashing():
    S = pb.GenomicSegment
    a, b, c = S("chrA", 10, 20, "+"), S("chrA", 10, 20, "+"), S("chrA", 10, 25, "-")
    assert a == b and not (a != b) and hash(a) == hash(b) and a != c and len({a, b, c}) == 2
    segs = [S("chrB", 0, 5, "+"), S("chrA", 50, 60, "-"), S("chrA", 50, 60, "+"), S("chrA", 5, 100, "+")]
    assert [str(s) for s in sorted(segs)] == ["chrA:5-100(+)", "chrA:50-60(+)", "chrA:50-60(-)", "chrB:0-5(+)"]
    assert repr(a).endswith("chrA:10-20(+)>") or "chrA:10-20(+)" in repr(a)
    assert S.from_str("chrA:10-20(+)") == a


def test_center_fixed_point_weights_bound_their_error():
    """CenterMapFactory.fixed_point_tables: integer weights round(2^shift / m) whose relative error against 1/m
    is at most m * 2^-(shift+1), with a shift that keeps the sum over all reads below 2^62."""
    fac = pb.CenterMapFactory(nibble=3)
    hist = np.zeros(65536, dtype=np.int64)
    assert fac.fixed_point_tables(hist) is None
    hist[30] = 10**8
    assert fac.fixed_point_tables(hist) is None              # one map length: the exact kernel
    hist[25:36] = 10**7
    hist[6] = 5                                                # L - 2*nibble == 0: no slot
    hist[4] = 7                                                # negative map length: no slot
    slot_of_len, w_fix, shift = fac.fixed_point_tables(hist)
    lengths = np.arange(25, 36)
    assert (slot_of_len[lengths] == np.arange(11)).all() and slot_of_len[6] == -1 and slot_of_len[4] == -1
    assert (slot_of_len >= 0).sum() == 11 and 1 <= shift <= 52
    m = lengths - 6
    rel = np.abs(w_fix.astype(np.float64) / 2.0 ** shift - 1.0 / m) * m
    assert (rel <= m * 2.0 ** -(shift + 1) + 1e-18).all() and rel.max() < fac.FIXED_MAX_REL_ERR
    total = sum(int(hist[L]) * int(w) for L, w in zip(lengths, w_fix))
    assert total < 2 ** 62
    s2, inv_m = fac.slot_tables(hist)
    assert (s2 == slot_of_len).all() and np.array_equal(inv_m, 1.0 / m)
    # so many reads that no shift keeps 1e-9: refused, the caller falls back to exact passes
    hist[:] = 0
    hist[2000:2003] = 2 ** 40
    assert pb.CenterMapFactory(0).fixed_point_tables(hist) is None


def test_script_table_io_round_trips(tmp_path):
    """bin/_cli.py: alignment batches as .npz, BED3-12(+gene_id) as chains / transcripts, plastid's tab tables."""
    from plastid_b200.bin import _cli
    rng = np.random.default_rng(5)
    lens = {"chrA": 30_000, "chrB": 4000}
    for reads in ({"chrA": random_cigar_reads(rng, 300, 30_000, 25_000), "chrB": random_cigar_reads(rng, 40, 4000, 2000)},
                  {"chrA": [po.Read(int(p), [(0, 30)], bool(p % 2)) for p in range(0, 900, 7)]}):
        hb = pack_reads(reads, lens)
        path = str(tmp_path / "b.npz")
        _cli.save_batch(path, hb)
        back = _cli.load_batch(path)
        assert back.chroms == hb.chroms and back.mapped == hb.mapped and back.max_span == hb.max_span
        for f in ("ref_start", "meta", "chrom_read_off", "chrom_len"):
            assert (getattr(back, f) == getattr(hb, f)).all(), f
        assert (back.blk is None) == (hb.blk is None) and (hb.blk is None or ((back.blk == hb.blk).all() and (back.blk_off == hb.blk_off).all()))
    bed = tmp_path / "a.bed"
    bed.write_text("track name=x\n# comment\n"
                   "chrA\t100\t400\ttxA\t0\t+\t150\t350\t0\t2\t100,100,\t0,200,\tgeneA\n"
                   "chrA\t1000\t1100\ttxB\t0\t-\t1000\t1000\t0\t1\t100,\t0,\n"
                   "chrB\t5\t50\n")
    chains = _cli.read_bed(str(bed))
    assert [str(c) for c in chains] == ["chrA:100-200^300-400(+)", "chrA:1000-1100(-)", "chrB:5-50(.)"]
    assert [c.get_name() for c in chains] == ["txA", "txB", "chrB:5-50"]
    txs = _cli.read_bed(str(bed), as_transcripts=True)
    assert txs[0].cds_genome_start == 150 and txs[0].cds_genome_end == 350 and txs[0].get_gene() == "geneA"
    assert str(txs[0].get_cds()) == "chrA:150-200^300-350(+)" and txs[1].cds_genome_start is None
    assert txs[0].as_bed().rstrip("\n").split("\t") == ["chrA", "100", "400", "txA", "0", "+", "150", "350", "0,0,0", "2", "100,100,", "0,200,"]
    assert txs[1].as_bed().split("\t")[6:8] == ["1000", "1000"]           # no coding region: thick columns = chain start
    again = tmp_path / "again.bed"
    again.write_text("".join(t.as_bed() for t in txs))
    assert [(str(t), t.cds_genome_start, t.cds_genome_end) for t in _cli.read_bed(str(again), as_transcripts=True)] == \
        [(str(t), t.cds_genome_start, t.cds_genome_end) for t in txs]
    tab = tmp_path / "t.txt"
    tab.write_text("## note\nregion_name\tregion\tcounts\n#x\ntxA\tchrA:100-200^300-400(+)\t5\ntxB\tchrA:1000-1100(-)\t7\n")
    cols = _cli.read_pl_table(str(tab))
    assert cols == {"region_name": ["txA", "txB"], "region": ["chrA:100-200^300-400(+)", "chrA:1000-1100(-)"], "counts": ["5", "7"]}
    assert [str(pb.SegmentChain.from_str(r)) for r in cols["region"]] == cols["region"]


class _HostCountingArray(object):
    """Stand-in for the device genome array in the script loops: ``sum()``, ``get(seg, roi_order)`` answered by the
    oracle's BAMGenomeArray, ``count_chains`` = masked sum + unmasked length per chain through the PRODUCT's chain
    objects (add_masks / get_masked_counts / masked_length) — so the host side of count_regions can be held against
    the oracle's restatement of the script without a GPU."""

    def __init__(self, oga):
        self.oga = oga

    def sum(self):
        return self.oga.sum()

    def get(self, seg, roi_order=True):
        return self.oga.get(po.Seg(seg.chrom, seg.start, seg.end, seg.strand), roi_order=roi_order)

    def count_chains(self, chains, use_masks=True):
        if not use_masks:
            return (np.asarray([float(ch.get_counts(self).sum()) if ch.length else 0.0 for ch in chains]),
                    np.asarray([ch.length for ch in chains]))
        sums = [float(np.nansum(ch.get_masked_counts(self).filled(0.0))) if ch.length else 0.0 for ch in chains]
        return np.asarray(sums), np.asarray([ch.masked_length for ch in chains])


def _host_world(mapping):
    chroms, lens = synth.yeast_like_genome(total=120_000, n_chrom=3)
    ann = synth.make_annotation(chroms, lens, 30, seed=8, exons=(1, 3), exon_len=(150, 400), intron_len=(40, 300))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, 6000, seed=4, device="cpu", lengths=range(24, 36)), chroms, lens)
    reads = {c: [] for c in chroms}
    for i in range(len(hb)):
        c = int(np.searchsorted(hb.chrom_read_off, i, side="right")) - 1
        reads[chroms[c]].append(po.Read(int(hb.ref_start[i]), [(0, int(hb.meta[i] & 0xFFFF))], bool((hb.meta[i] >> 16) & 1)))
    oga = po.OracleBAMGenomeArray(po.ReadStore(dict(zip(chroms, [int(x) for x in lens])), reads), mapping=mapping)
    oga.add_filter("size", po.SizeFilter(25, 100))
    return ann, oga


def test_counts_in_region_host_side_against_the_oracle_script(tmp_path):
    from oracle import scripts as osc
    from plastid_b200.bin import counts_in_region
    ann, oga = _host_world(po.FivePrimeMap(14))
    ga = _HostCountingArray(oga)

    def both_kinds():
        chains = ann.chains()
        chains.append(pb.SegmentChain(pb.GenomicSegment("chrUnknown", 10, 500, "+"), ID="nowhere"))
        ochains = []
        for ch in chains:
            oc = po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in ch])
            oc.name = ch.get_name()
            ochains.append(oc)
        return chains, ochains

    # masks handed over per region (counts_in_region.py:115-116), one region fully masked
    chains, ochains = both_kinds()
    masks = synth.make_masks(ann, frac=0.25, seed=3) + [[]]
    span = chains[5].spanning_segment
    masks[5] = [pb.GenomicSegment(span.chrom, span.start - 10, span.end + 10, span.strand)]
    omasks = [[po.Seg(m.chrom, m.start, m.end, m.strand) for m in ms] for ms in masks]
    exp = osc.counts_in_region_rows(oga, ochains, omasks)
    ga_sum, got = counts_in_region.count_regions(ga, chains, masks)
    assert ga_sum == oga.sum() and got == exp
    assert got[5][2] == "nan" and got[5][3] == "nan" and got[5][5] == "0" and got[-1][2] == "0.00000000e+00"
    assert sum(float(r[2]) for r in got if r[2] != "nan") > 0
    out = tmp_path / "counts.txt"
    with open(out, "w") as fh:
        counts_in_region.write_table(fh, ga_sum, got)
    text = out.read_text().split("\n")
    assert text[0] == "## total_dataset_counts: %s" % oga.sum()
    assert text[1].split("\t") == ["region_name", "region", "counts", "counts_per_nucleotide", "rpkm", "length"]
    assert [line.split("\t") for line in text[2:-1]] == exp

    # masks found by the overlap query of a mask annotation (counts_in_region.py:114-115 through GenomeHash)
    chains, ochains = both_kinds()
    rng = np.random.default_rng(11)
    feats = []
    for k in range(60):
        ch = chains[int(rng.integers(len(chains) - 1))]
        sp = ch.spanning_segment
        a = int(rng.integers(max(sp.start - 300, 0), sp.end + 100))
        segs = [pb.GenomicSegment(ch.chrom, a, a + int(rng.integers(1, 400)), ch.strand if k % 5 else "+-"[k % 2])]
        if k % 3 == 0:
            b = segs[0].end + int(rng.integers(1, 2000))
            segs.append(pb.GenomicSegment(ch.chrom, b, b + int(rng.integers(1, 300)), segs[0].strand))
        feats.append(pb.SegmentChain(*segs, ID="mask%d" % k))
    feats += feats[:5]
    ofeats = [po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in f]) for f in feats]
    exp = osc.counts_in_region_rows(oga, ochains, crossmap=po.GenomeHash(ofeats))
    _sum, got = counts_in_region.count_regions(ga, chains, masks=counts_in_region.overlapping_masks(chains, feats))
    assert got == exp and sum(int(r[5]) for r in got) < sum(ch.length for ch in chains)
    with pytest.raises(KeyError):
        counts_in_region.overlapping_masks(chains[:2], [pb.SegmentChain(pb.GenomicSegment(chains[0].chrom, 1, 9, "."))])


def test_psite_offset_choice_and_offset_file_round_trip(tmp_path):
    """psite.py:462-521 by hand: the offset of a read length is the distance from the profile's highest allowed
    column to the landmark; all-NaN or all-zero profiles take the default; --constrain / --require_upstream limit
    the columns.  The table psite writes is what VariableFivePrimeMapFactory.from_file reads (p_site.rst:128-151)."""
    from plastid_b200.bin import psite
    from oracle import scripts as osc
    x = np.arange(-50, 100)
    profiles = {}
    for k, peak in ((28, -12), (29, -13), (30, -13), (31, 5), (32, None), (33, "zero")):
        y = np.full(len(x), np.nan) if peak is None else np.zeros(len(x))
        if isinstance(peak, int):
            y[:] = 0.1
            y[x == peak] = 3.0
            y[x == -30] = 2.0                                   # a lower, farther peak
        profiles[k] = y
    free = psite.pick_offsets(x, profiles, default=13)
    assert free == {28: 12, 29: 13, 30: 13, 31: -5, 32: 13, 33: 13}
    assert psite.pick_offsets(x, profiles, default=13, require_upstream=True)[31] == 30        # columns x < 0 only
    tight = psite.pick_offsets(x, profiles, default=14, constrain=(25, 35))                    # 25..35 nt upstream
    assert tight == {28: 30, 29: 30, 30: 30, 31: 30, 32: 14, 33: 14}
    for kw in (dict(), dict(require_upstream=True), dict(constrain=(35, 25)), dict(constrain=(0, 20))):
        assert psite.pick_offsets(x, profiles, 13, **kw) == osc.psite_pick_offsets(x, profiles, 13, **kw)
    path = tmp_path / "p_offsets.txt"
    with open(path, "w") as fh:
        psite.write_offsets(fh, {k: v for k, v in free.items() if v >= 0}, 13)
    assert path.read_text().split("\n")[0] == "length\tp_offset" and path.read_text().endswith("default\t13")
    fac = pb.VariableFivePrimeMapFactory.from_file(str(path))
    fw, rc = po.build_offset_luts({28: 12, 29: 13, 30: 13, 32: 13, 33: 13, "default": 13})
    assert (fac.forward_offsets == fw).all() and (fac.reverse_offsets == rc).all()
    assert fac.forward_offsets[28] == 12 and fac.forward_offsets[31] == 13 and fac.forward_offsets[13] == -1


def test_cs_count_host_side_against_the_oracle_script(tmp_path):
    """cs.py:682-714 (`cs count`): per gene and per class reads / length / RPKM, nan where a class is empty,
    formatted with %.8f — the product's do_count + write_table over a host stand-in vs the oracle's loop."""
    import warnings
    from oracle import scripts as osc
    from plastid_b200.bin import cs
    ann, oga = _host_world(po.ThreePrimeMap(0))
    ga = _HostCountingArray(oga)
    pos = {"region": [], "exon": [], "utr5": [], "cds": [], "utr3": []}
    for i, ch in enumerate(ann.chains()):
        n = ch.length
        a, b = n // 5, n - n // 4
        pos["region"].append(ch.get_name())
        pos["exon"].append(str(ch))
        pos["utr5"].append(str(ch.get_subchain(0, a)) if i % 7 else "na")          # an empty class now and then
        pos["cds"].append(str(ch.get_subchain(a, b)))
        pos["utr3"].append(str(ch.get_subchain(b, n)))
    exp = osc.cs_count(oga, pos)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        order, cols = cs.do_count(ga, pos)
    assert order[0] == "region" and len(order) == 13 and cols["region"] == exp["region"]
    for key in order[1:]:
        got, want = np.asarray(cols[key], dtype=float), np.asarray(exp[key], dtype=float)
        assert np.array_equal(got, want, equal_nan=True), key
    assert np.isnan(cols["utr5_rpkm"][0]) and cols["utr5_length"][0] == 0 and cols["utr5_reads"][0] == 0
    assert sum(cols["exon_reads"]) > 0
    out = tmp_path / "cs.txt"
    with open(out, "w") as fh:
        cs.write_table(fh, order, cols)
    lines = out.read_text().rstrip("\n").split("\n")
    assert lines[0].split("\t") == order and len(lines) == 1 + len(pos["region"])
    first = dict(zip(order, lines[1].split("\t")))
    # the reference's own output (tests/golden/ref_scripts/out/cs_count_*.txt): the integer zero of an empty chain sits
    # in a float column of the pandas table and is written as 0.00000000
    assert first["utr5_rpkm"] == "nan" and first["utr5_reads"] == "0.00000000" and first["exon_rpkm"] == "%.8f" % exp["exon_rpkm"][0]


def test_mapping_rule_plugin_descriptors(tmp_path):
    """The `plastid.mapping_rules` entry-point contract (argparsers.py:505-534, docs/source/devinfo/entrypoints.rst):
    a dict with name / bamfunc / help whose bamfunc is called as functools.partial(bamfunc, args=args)()
    (argparsers.py:695-696) and has to hand back a mapping rule."""
    import argparse
    import functools
    from plastid_b200 import plugins
    path = tmp_path / "p_offsets.txt"
    path.write_text("length\tp_offset\n28\t12\n29\t13\ndefault\t13")
    args = argparse.Namespace(offset=14, nibble=12)
    made = {}
    for desc in (plugins.fiveprime, plugins.threeprime, plugins.center, plugins.fiveprime_variable):
        assert set(desc) >= {"name", "bamfunc", "help"} and desc["name"].startswith("b200_") and desc["help"]
        if desc is plugins.fiveprime_variable:
            args = argparse.Namespace(offset=str(path), nibble=0)
        made[desc["name"]] = functools.partial(desc["bamfunc"], args=args)()
    assert isinstance(made["b200_fiveprime"], pb.FivePrimeMapFactory) and made["b200_fiveprime"].offset == 14
    assert type(made["b200_threeprime"]) is pb.ThreePrimeMapFactory and made["b200_threeprime"].offset == 14
    assert isinstance(made["b200_center"], pb.CenterMapFactory) and made["b200_center"].nibble == 12
    var = made["b200_fiveprime_variable"]
    assert isinstance(var, pb.VariableFivePrimeMapFactory) and var.forward_offsets[28] == 12 and var.forward_offsets[40] == 13
    assert plugins.device_option["name"] == "device" and plugins.device_option["default"] == "cuda"
    for rule in made.values():
        assert callable(rule)                                   # fn(reads, seg) -> (reads_out, counts)


def test_batch_rows_viewed_as_read_objects():
    """`reads_out` of a batch decoded from a file holds BatchRead views (no pysam objects exist there): the
    attributes a mapping rule or a user filter reads (map_factories.pyx:243,345; genome_array.py:811-818)."""
    rng = np.random.default_rng(3)
    lens = {"chrA": 20_000}
    reads = {"chrA": random_cigar_reads(rng, 200, 20_000, 15_000)}
    packed = pack_reads(reads, lens, keep_objects=False)
    assert packed.objects is None
    ordered = sorted(reads["chrA"], key=lambda r: r.positions[0])
    for i in (0, 1, 57, 199):
        view = packed.read_view(i)
        want = po.positions_from_cigar(ordered[i].reference_start, ordered[i].cigartuples)
        assert view.positions == want == view.get_reference_positions()
        assert view.reference_start == want[0] and view.is_reverse == ordered[i].is_reverse
        assert view == packed.read_view(i) and hash(view) == hash(packed.read_view(i)) and view != packed.read_view(i + 1 if i < 199 else 0)
        assert ("#%d" % i) in repr(view)
    assert len({packed.read_view(i) for i in range(200)} | {packed.read_view(0)}) == 200
    kept = pack_reads(reads, lens)                              # the packer's own objects when they are kept
    assert kept.read_view(0) is kept.objects[0] and kept.read_view(0).positions == packed.read_view(0).positions


def _python_track_lines(kind, chrom, st, en, vals):
    """The reference's own per-line writes (genome_array.py:1030-1037, 1096-1111) on numpy scalars."""
    if kind == 0:
        return "".join("%s\t%s\n" % (p + 1, v) for p, v in zip(st.tolist(), vals))
    return "".join("%s\t%s\t%s\t%s\n" % (chrom, a, b, v) for a, b, v in zip(st.tolist(), en.tolist(), vals))


def test_track_text_is_formatted_like_python_does():
    """pb_format_track (what to_variable_step / to_bedgraph write with) against Python's `%s` of numpy scalars:
    integer counts, normalised counts, Center-rule fractions, and the corners of float repr (exponent thresholds at
    1e-4 and 1e16, subnormals, 17-digit values, integers-valued floats, nan / inf / signed zero)."""
    import io
    rng = np.random.default_rng(13)
    n = 50_000
    st = np.sort(rng.integers(0, 2_000_000_000, n)).astype(np.int64)
    en = st + rng.integers(1, 5000, n)
    corner = np.array([1e-4, 9.999e-5, 1e-5, 1.5e-7, 1e16, 9999999999999998.0, 1.2345678901234567e16, 123456789.125, 0.1 + 0.2, 1 / 3.0,
                       5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1e22, 1e23, 100000.0, 5.0, 0.5, 1e15, 123456.0,
                       0.0001234, 12.0 / 76, 1e6 / 3e8, 0.0, -0.0, -2.5, float("nan"), float("inf"), float("-inf"), 1e100, 4.35, 0.3])
    families = [
        rng.integers(1, 100_000, n).astype(np.int64),                                     # raw counts
        np.array([0, 1, 9, 10, 99, 2**31, 2**40, 2**62, -7], dtype=np.int64),             # integer corners
        rng.integers(1, 5000, n) / 500477.0 * 1e6,                                         # normalised counts
        rng.integers(1, 4000, n) / 76.0,                                                   # Center fractions, one map length
        (rng.integers(1, 300, n) / rng.integers(1, 40, n)) + (rng.integers(0, 50, n) / 13.0),
        rng.random(n) * 10.0 ** rng.integers(-12, 20, n),                                 # every magnitude
        corner,
        np.ldexp(rng.random(n), rng.integers(-1070, 1020, n)),                            # down to subnormals
    ]
    for vals in families:
        m = len(vals)
        for kind, chrom in ((0, "chrI"), (1, "chrI"), (1, "a_rather_long_contig_name|with.odd-chars")):
            fh = io.StringIO()
            pb.BAMGenomeArray._write_records(fh, kind, chrom, st[:m], en[:m] if kind else None, vals, chunk=7001)
            assert fh.getvalue() == _python_track_lines(kind, chrom, st[:m], en[:m], vals), (kind, vals.dtype)
    fh = io.StringIO()
    pb.BAMGenomeArray._write_records(fh, 1, "c", st[:0], en[:0], families[0][:0])
    assert fh.getvalue() == ""
    big = np.arange(300_000, dtype=np.int64)                                               # several threads, one chunk
    fh = io.StringIO()
    pb.BAMGenomeArray._write_records(fh, 0, "c", big, None, big % 977 / 7.0)
    assert fh.getvalue() == _python_track_lines(0, "c", big, None, big % 977 / 7.0)
    L = _lib.lib()
    assert L.pb_format_track(2, b"c", None, None, None, 0, 0, None, 0, 1) == -1
    one = np.zeros(1, dtype=np.int64)
    p = one.ctypes.data_as(C.c_void_p)
    assert L.pb_format_track(0, None, p, None, p, 0, 1, p, 8, 1) == -1 and b"bound" in L.pb_last_error()


def test_track_writers_around_the_formatter(monkeypatch):
    """to_variable_step / to_bedgraph with the device compaction replaced by given records: header line, track
    keywords in sorted order, chromosomes in sorted order, `variableStep` lines, one formatted line per record
    (genome_array.py:990-1111)."""
    import io
    records = {"chrB": (np.array([4, 9, 10], dtype=np.int64), np.array([6, 10, 40], dtype=np.int64), np.array([2, 1, 7], dtype=np.int64)),
               "chrA": (np.array([0, 99], dtype=np.int64), np.array([3, 100], dtype=np.int64), np.array([0.5, 1e-05])),
               "chrC": (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))}
    ga = object.__new__(pb.BAMGenomeArray)
    monkeypatch.setattr(pb.BAMGenomeArray, "strands", lambda self: ("+", "-", "."))
    monkeypatch.setattr(pb.BAMGenomeArray, "chroms", lambda self: ["chrB", "chrC", "chrA"])
    monkeypatch.setattr(pb.BAMGenomeArray, "_is_lowerable", lambda self: True)
    monkeypatch.setattr(pb.BAMGenomeArray, "_export_records",
                        lambda self, chrom, strand, mode, window: (records[chrom][0], records[chrom][1] if mode == 1 else None, records[chrom][2]))
    fh = io.StringIO()
    ga.to_variable_step(fh, "tr", "+", color="0,0,255", autoScale="on")
    assert fh.getvalue() == ("track type=wiggle_0 name=tr autoScale=on color=0,0,255\n"
                             "variableStep chrom=chrA span=1\n1\t0.5\n100\t1e-05\n"
                             "variableStep chrom=chrB span=1\n5\t2\n10\t1\n11\t7\n"
                             "variableStep chrom=chrC span=1\n")
    fh = io.StringIO()
    ga.to_bedgraph(fh, "tr", "-")
    assert fh.getvalue() == ("track type=bedGraph name=tr\n"
                             "chrA\t0\t3\t0.5\nchrA\t99\t100\t1e-05\n"
                             "chrB\t4\t6\t2\nchrB\t9\t10\t1\nchrB\t10\t40\t7\n")


def _spliced_arrays(rng, n, n_chrom, with_blocks=True):
    cid = rng.integers(0, n_chrom, n)
    start = rng.integers(0, 3000, n)                        # many ties: the merge has to be stable
    nb = rng.choice([1, 1, 1, 2, 3, 5], n) if with_blocks else np.ones(n, dtype=np.int64)
    rows, L = [], np.zeros(n, dtype=np.int64)
    for i in range(n):
        at = 0
        for _k in range(int(nb[i])):
            ln = int(rng.integers(1, 40))
            rows.append((at, ln))
            L[i] += ln
            at += ln + int(rng.integers(1, 500))
    return cid, start, L, rng.integers(0, 2, n), nb, np.asarray(rows, dtype=np.int32)


def test_batches_from_arrays_and_merges_keep_every_read_and_block():
    """batch_from_arrays / merge_batches on random spliced reads (vectorised block gathers) against a plain
    per-read statement: stable order by (chromosome, start), every read's blocks behind it, file order for ties
    (the reference chains `fetch` over the files in the order given, genome_array.py:800-809)."""
    rng = np.random.default_rng(17)
    chroms, lens = ["a", "b", "c"], [10**6] * 3
    parts, flat = [], []
    for f, (n, with_blocks) in enumerate(((700, True), (300, False), (500, True), (0, True))):
        cid, start, L, rev, nb, rows = _spliced_arrays(rng, n, 3, with_blocks)
        b = batch_from_arrays(chroms, lens, cid, start, L, rev, blocks=(nb, rows) if with_blocks else None)
        b.check_sorted()
        at = np.concatenate([[0], np.cumsum(nb)])
        reads = [(int(cid[i]), int(start[i]), int(L[i]), bool(rev[i]), [tuple(r) for r in rows[at[i]:at[i + 1]].tolist()]) for i in range(n)]
        want = sorted(reads, key=lambda r: (r[0], r[1]))                      # sorted() is stable
        assert len(b) == n
        for i, (c, s, ln, rv, blks) in enumerate(want):
            assert int(b.ref_start[i]) == s and int(b.aligned_len[i]) == ln and bool(b.is_reverse[i]) == rv
            assert b.positions_of(i) == [s + a + k for a, m in blks for k in range(m)]
        assert list(np.diff(b.chrom_read_off)) == [sum(1 for r in reads if r[0] == c) for c in range(3)]
        parts.append(b)
        flat.extend((r, f) for r in want)
    m = merge_batches(parts)
    m.check_sorted()
    want = sorted(flat, key=lambda x: (x[0][0], x[0][1]))                      # ties: first file first
    assert len(m) == len(want) == 1500 and m.mapped == 1500
    for i, ((c, s, ln, rv, blks), _f) in enumerate(want):
        assert int(m.ref_start[i]) == s and int(m.aligned_len[i]) == ln and bool(m.is_reverse[i]) == rv
        assert m.positions_of(i) == [s + a + k for a, mm in blks for k in range(mm)]
    assert (m.meta >> 24).tolist() == [len(r[0][4]) for r in want]
    assert merge_batches(parts[1:2]).blk is None and merge_batches([parts[1], parts[1]]).blk is None


def test_host_thread_count_follows_affinity_and_ranks(monkeypatch):
    import os
    assert _lib.host_threads(3) == 3
    allowed = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    assert _lib.host_threads() == _lib.host_threads(0) == allowed
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    if allowed == os.cpu_count():
        assert _lib.host_threads() == max(1, allowed // 8)
    if hasattr(os, "sched_setaffinity") and allowed > 1:
        before = os.sched_getaffinity(0)
        try:
            os.sched_setaffinity(0, set(sorted(before)[:1]))
            assert _lib.host_threads() == 1                    # a narrowed mask is taken as it is, ranks or not
        finally:
            os.sched_setaffinity(0, before)


@pytest.mark.parametrize("prog", ["counts_in_region", "cs", "metagene", "psite", "phase_by_size", "make_wiggle"])
def test_script_command_lines_parse_without_a_device(prog, capsys):
    """Every sub-program builds its parser and prints its usage without touching CUDA; the reference's mapping flags
    (argparsers.py:337-503: --fiveprime --threeprime --center --fiveprime_variable --offset --nibble --min_length
    --max_length --sum) are there wherever alignments are read."""
    import importlib
    mod = importlib.import_module("plastid_b200.bin.%s" % prog)
    with pytest.raises(SystemExit) as e:
        mod.main(["--help"])
    assert e.value.code == 0
    usage = capsys.readouterr().out
    sub = {"cs": "count", "metagene": "count"}.get(prog)
    if sub is not None:
        assert "generate" in usage and "count" in usage
        with pytest.raises(SystemExit) as e:
            mod.main([sub, "--help"])
        assert e.value.code == 0
        usage = capsys.readouterr().out
    for flag in ("--count_files", "--fiveprime", "--threeprime", "--center", "--fiveprime_variable", "--offset", "--nibble",
                 "--min_length", "--max_length", "--sum"):
        assert flag in usage, (prog, flag)
    with pytest.raises(SystemExit) as e:                      # required arguments missing: argparse's exit status 2
        mod.main([sub] if sub else [])
    assert e.value.code == 2


# ---------------------------------------------------------------------------------------------
# round 2: region lowering, chromosome cuts, command-line plumbing, --keep matrices
# ---------------------------------------------------------------------------------------------
def test_regions_are_clipped_to_their_chromosome_when_lowered():
    """blocks past a chromosome's end must not alias the next chromosome's bins; they become virtual blocks
    (coordinates beyond every layout) that keep length, mask bits and window columns."""
    from plastid_b200.regions import ChainTable, lower_segment, VIRTUAL_BIN as V
    assert lower_segment(10, 20, 1000, 100) == [(1010, 1020)]
    assert lower_segment(90, 120, 1000, 100) == [(1090, 1100), (V + 100, V + 120)]
    assert lower_segment(150, 170, 1000, 100) == [(V + 150, V + 170)]
    assert lower_segment(-5, 120, 1000, 100) == [(V, V + 5), (1000, 1100), (V + 100, V + 120)]
    layout = pb.GenomeLayout(["a", "b"], [20000, 500])
    ch = pb.SegmentChain(pb.GenomicSegment("a", 19990, 20040, "+"), pb.GenomicSegment("a", 20100, 20110, "+"))
    ch.add_masks(pb.GenomicSegment("a", 19995, 20005, "+"))
    t = ChainTable.from_chains([ch], layout)
    assert t.chain_len[0] == ch.length == 60 and list(t.bend - t.bstart) == [10, 40, 10]
    assert t.bstart[0] == 19990 and (t.bstart[1:] >= V).all() and (t.bend[:1] <= layout.chrom_bin_off[1]).all()
    assert np.unpackbits(t.mask_bits, bitorder="little")[:60].sum() == 10        # mask bits keep their chain positions


def test_chromosome_cuts_fall_on_chromosome_boundaries():
    from plastid_b200 import dist as pd
    rng = np.random.default_rng(3)
    lens = [50000, 20000, 90000, 16384, 70000, 30000]
    chroms = ["c%d" % i for i in range(len(lens))]
    n = 40000
    cid = np.sort(rng.integers(0, len(lens), n))
    start = np.array([rng.integers(0, lens[c] - 40) for c in cid])
    hb = pb.batch_from_arrays(chroms, lens, cid, start, np.full(n, 30), np.zeros(n, dtype=bool))
    layout = pb.GenomeLayout(chroms, lens)
    for world in (2, 3, 4, 8):
        cuts = pd.position_cuts(hb, layout, world, snap="chromosomes")
        assert cuts[0] == 0 and cuts[-1] == layout.total_bins and (np.diff(cuts) >= 0).all() and len(cuts) == world + 1
        assert set(int(c) for c in cuts) <= set(int(x) for x in layout.chrom_bin_off)
        held = 0
        for r in range(world):
            sub, lo, hi = pd.shard_positions(hb, layout, r, world, snap="chromosomes")
            held += len(sub)
        assert held == len(hb)          # whole chromosomes: no halo read is needed twice


def test_chain_table_live_lengths_and_owned_rows():
    """``ChainTable.live_lengths`` = ``chain.masked_length`` (geometry, host side) and ``owned_rows`` = the blocks that
    overlap a rank's bins with chain offsets that address them — what a position-sharded rank hands to pb_region_sums."""
    from plastid_b200.regions import ChainTable
    layout = pb.GenomeLayout(["c1", "c2"], [50000, 30000])
    rng = np.random.default_rng(9)
    chains = []
    for i in range(60):
        c = "c1" if i % 3 else "c2"
        a = int(rng.integers(0, 20000))
        segs = [pb.GenomicSegment(c, a, a + 300, "+-"[i % 2]), pb.GenomicSegment(c, a + 1000, a + 1400, "+-"[i % 2])]
        if i % 4 == 0:
            segs.append(pb.GenomicSegment(c, a + 5000, a + 5100, "+-"[i % 2]))
        ch = pb.SegmentChain(*segs)
        if i % 5 == 0:
            ch.add_masks(pb.GenomicSegment(c, a + 250, a + 1100, "+-"[i % 2]))
        chains.append(ch)
    chains.append(pb.SegmentChain())
    t = ChainTable.from_chains(chains, layout)
    assert list(t.live_lengths()) == [ch.masked_length for ch in chains]
    total = 0
    for lo, hi in ((0, 16384), (16384, 49152), (49152, layout.total_bins)):
        idx, sub_off = t.owned_rows(lo, hi)
        total += len(idx)
        assert ((t.bend[idx] > lo) & (t.bstart[idx] < hi)).all() and len(sub_off) == t.n_chains + 1 and sub_off[-1] == len(idx)
        chain_of = np.repeat(np.arange(t.n_chains), np.diff(t.chain_off))
        for c in range(t.n_chains):
            assert (chain_of[idx[sub_off[c]:sub_off[c + 1]]] == c).all()
    assert total >= len(t.bstart)                      # a block across a cut belongs to both sides


def test_position_cuts_balance_reads_plus_plane_bins():
    """Cuts give every rank the same cost ``reads + bins`` (bytes streamed + plane bytes written, dist.balanced_cuts):
    a skewed batch (most reads on one chromosome) no longer leaves one rank with most of the genome to write;
    ``weights=(1, 0)`` is the equal-read-count rule."""
    from plastid_b200 import dist as pd, _lib
    rng = np.random.default_rng(5)
    lens = [400000, 900000, 250000, 16384, 700000]
    chroms = ["c%d" % i for i in range(len(lens))]
    n = 300000
    cid = np.sort(rng.choice(len(lens), n, p=[0.7, 0.05, 0.1, 0.05, 0.1]))
    start = np.array([rng.integers(0, lens[c] - 40) for c in cid])
    hb = pb.batch_from_arrays(chroms, lens, cid, start, np.full(n, 30), np.zeros(n, dtype=bool))
    layout = pb.GenomeLayout(chroms, lens)
    g = layout.chrom_bin_off[cid] + hb.ref_start                                       # global bin of every read's start
    for world in (2, 3, 8):
        for w in ((1.0, 1.0), (1.0, 0.0), (1.0, 4.0)):
            cuts = pd.position_cuts(hb, layout, world, weights=w)
            assert cuts[0] == 0 and cuts[-1] == layout.total_bins and (np.diff(cuts) >= 0).all() and len(cuts) == world + 1
            assert (cuts % _lib.PB_LAYOUT_ALIGN == 0).all()
            reads = np.array([np.count_nonzero((g >= cuts[r]) & (g < cuts[r + 1])) for r in range(world)])
            cost = w[0] * reads + w[1] * np.diff(cuts)
            assert reads.sum() == n
            # one grid step of slack: PB_LAYOUT_ALIGN bins and the reads that start in them
            slack = w[1] * _lib.PB_LAYOUT_ALIGN + w[0] * np.bincount(g // _lib.PB_LAYOUT_ALIGN).max()
            assert cost.max() - cost.min() <= 2 * slack, (world, w, cost)
        eq = pd.position_cuts(hb, layout, world, weights=(1.0, 0.0))
        reads = np.array([np.count_nonzero((g >= eq[r]) & (g < eq[r + 1])) for r in range(world)])
        bal = pd.position_cuts(hb, layout, world)
        assert np.diff(bal).max() < np.diff(eq).max()                   # the sparse stretch is shared out


def test_mapping_flags_behave_like_the_reference_parser(capsys):
    """argparsers.py:439-470, 656-700: store_const into one destination (last flag wins), no flag -> message + exit 1,
    --fiveprime_variable without an offset file -> message + exit 1; --normalize / --sum handled as in :775-780."""
    import argparse
    from plastid_b200.bin import _cli
    def parse(argv, disabled=()):
        p = argparse.ArgumentParser()
        _cli.add_alignment_args(p, disabled=disabled)
        return p.parse_args(argv)
    a = parse(["--count_files", "x.bam", "--fiveprime", "--threeprime", "--offset", "3"])
    assert a.mapping == "threeprime" and isinstance(_cli.mapping_from_args(a), pb.ThreePrimeMapFactory)
    assert _cli.mapping_from_args(parse(["--center", "--nibble", "7"])).nibble == 7
    with pytest.raises(SystemExit) as e:
        _cli.mapping_from_args(parse(["--count_files", "x.bam"]))
    assert e.value.code == 1 and "Please specify a read mapping rule." in capsys.readouterr().err
    with pytest.raises(SystemExit):
        _cli.mapping_from_args(parse(["--fiveprime_variable"]))
    assert "Please specify a filename to use for fiveprime variable offsets in --offset." in capsys.readouterr().err
    assert parse(["--normalize", "--sum", "5"]).normalize is True
    with pytest.raises(SystemExit):
        parse(["--normalize"], disabled=("normalize",))        # counts_in_region.py:60 disables it


def test_keep_matrices_follow_numpy_ma_division():
    """metagene.py:918-932: numpy.savetxt writes the DATA of the masked arrays, and numpy.ma's division leaves the
    numerator wherever it masks — the --keep files of the programs must carry exactly that."""
    from plastid_b200.bin.metagene import keep_matrices
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 5, (12, 30)).astype(float)
    raw[:, :4] = np.nan
    mask = rng.random((12, 30)) < 0.2
    mask[:, :4] = True
    mask[3, 8:20] = True                                  # a row whose normalisation window is fully masked
    raw[5, 8:20] = 0                                      # a row with denominator 0
    counts = np.ma.MaskedArray(raw.copy(), mask=mask.copy())
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        denominator = np.nansum(counts[:, 8:20], axis=1)
        norm_counts = (counts.T.astype(float) / denominator).T
        norm_counts = np.ma.MaskedArray(norm_counts, mask=counts.mask)
        norm_counts.mask[np.isnan(norm_counts)] = True
        norm_counts.mask[np.isinf(norm_counts)] = True
    den = np.ma.filled(denominator.astype(float), np.nan)
    got_raw, got_norm, got_mask = keep_matrices(dict(counts=counts, denominator=den))
    assert np.array_equal(got_raw, np.ma.getdata(counts), equal_nan=True)
    assert np.array_equal(got_norm, np.ma.getdata(norm_counts), equal_nan=True)
    assert np.array_equal(got_mask, np.ma.getmaskarray(norm_counts))


def test_entry_points_declared_in_pyproject_resolve():
    """the `plastid.mapping_rules` / `plastid.mapping_options` entry points and the console scripts are
    declared in pyproject.toml; every target imports and honours the contract of docs/source/devinfo/entrypoints.rst
    (a dict with name / bamfunc / help whose bamfunc(args=...) returns the mapping callable)."""
    import argparse
    import importlib
    import tomllib
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyproject.toml"), "rb") as fh:
        proj = tomllib.load(fh)["project"]

    def resolve(target):
        mod, attr = target.split(":")
        return getattr(importlib.import_module(mod), attr)
    rules = proj["entry-points"]["plastid.mapping_rules"]
    assert set(rules) == {"b200_fiveprime", "b200_threeprime", "b200_center", "b200_fiveprime_variable"}
    for name, target in rules.items():
        d = resolve(target)
        assert d["name"] == name and callable(d["bamfunc"]) and isinstance(d["help"], str)
    ns = argparse.Namespace(offset=7, nibble=3)
    assert resolve(rules["b200_fiveprime"])["bamfunc"](args=ns).offset == 7
    assert resolve(rules["b200_center"])["bamfunc"](args=ns).nibble == 3
    assert resolve(proj["entry-points"]["plastid.mapping_options"]["device"])["name"] == "device"
    for name, target in proj["scripts"].items():
        assert callable(resolve(target)), name


def test_indexed_container_bookkeeping_without_a_device():
    """``BAMGenomeArray(path, indexed=True)``: chromosomes, lengths and ``sum()`` come from the header and the ``.bai``
    statistics (``bamfile.mapped``, genome_array.py:690) without decoding a record; the files are decoded when
    something asks for every read, and a ``set_sum`` made before that survives it."""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden", "htslib_allops.bam")
    from plastid_b200.bam_io import batch_from_bam
    whole = batch_from_bam(gold)
    ga = pb.BAMGenomeArray(gold, indexed=True, device="cpu")
    assert ga.is_lazy and ga.sum() == whole.mapped
    assert ga.chroms() == sorted(whole.chroms) and ga.lengths() == dict(zip(whole.chroms, (int(x) for x in whole.chrom_len)))
    ga.set_mapping(pb.FivePrimeMapFactory(3))
    ga.add_filter("size", pb.SizeFilterFactory(10, 200))
    assert ga.is_lazy and ga.sum() == whole.mapped                       # nothing above needs the reads
    ga.set_sum(1234)
    assert len(ga.batch) == len(whole) and not ga.is_lazy and ga.sum() == 1234
    assert (ga.batch.ref_start == whole.ref_start).all() and (ga.batch.meta == whole.meta).all()
    with pytest.raises(TypeError):
        pb.BAMGenomeArray(whole, indexed=True, device="cpu")
    with pytest.raises(IOError):
        pb.BAMGenomeArray(os.path.join(os.path.dirname(__file__), "golden", "missing.bam"), indexed=True, device="cpu")
