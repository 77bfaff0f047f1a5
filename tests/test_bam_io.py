"""Host BAM decoder (pb_bam_*, no GPU): against the reference's vendored htslib through committed
fixtures (tests/golden/htslib_allops.*, made by tests/golden/make_bam_golden.py with
oracle/_ref/ref_bam_tool), against the packer on random CIGARs, and on malformed input."""
import ctypes as C
import os

import numpy as np
import pytest

from plastid_b200 import _lib, bam_io
from plastid_b200.batch import cigar_to_blocks, pack_reads
from oracle import pyoracle as po
from helpers import random_cigar_reads

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def parse_dump(path):
    refs, recs = [], []
    for line in open(path):
        f = line.split()
        if f[0] == "@":
            refs.append((f[1], int(f[2])))
            continue
        tid, pos, flag, n_cigar = (int(x) for x in f[:4])
        cigar = [tuple(int(v) for v in c.split(":")) for c in f[4].split(",")] if n_cigar else []
        recs.append((tid, pos, flag, cigar, int(f[-1])))
    return refs, recs


def test_decoder_matches_reference_htslib_dump():
    refs, recs = parse_dump(os.path.join(GOLD, "htslib_allops.dump.txt"))
    hb = bam_io.batch_from_bam(os.path.join(GOLD, "htslib_allops.bam"), threads=3)
    assert hb.chroms == [r[0] for r in refs] and list(hb.chrom_len) == [r[1] for r in refs]
    mapped = [r for r in recs if r[0] >= 0 and not (r[2] & 4) and r[3]]
    assert len(hb) == len(mapped) == hb.mapped and len(recs) - len(mapped) == 5 + sum(1 for r in recs if r[0] >= 0 and r[2] & 4)
    assert list(np.diff(hb.chrom_read_off)) == [sum(1 for r in mapped if r[0] == t) for t in range(len(refs))]
    hb.check_sorted()
    # rows are in file order within a chromosome (stable); compare field by field with htslib's view
    i = 0
    for tid, pos, flag, cigar, endpos in mapped:
        blocks, span = cigar_to_blocks(cigar)
        shift = blocks[0][0] if blocks else 0
        assert int(hb.ref_start[i]) == pos + shift
        m = int(hb.meta[i])
        assert (m & 0xFFFF) == sum(n for _a, n in blocks) and ((m >> 16) & 1) == ((flag >> 4) & 1) and (m >> 24) == len(blocks)
        assert hb.positions_of(i) == po.positions_from_cigar(pos, cigar)
        assert pos + span == endpos                    # htslib's bam_endpos == our reference span
        i += 1
    assert hb.max_span == max(cigar_to_blocks(c)[1] - (cigar_to_blocks(c)[0][0][0]) for _t, _p, _f, c, _e in mapped)


@pytest.mark.parametrize("window,aligned", [(None, False), ("70000", False), (None, True), ("70000", True)])
def test_decoder_matches_packer_on_random_cigars(tmp_path, monkeypatch, window, aligned):
    """Windows of 70 kB exercise the carry-over of partial members / records; PB_BAM_WALK_MIN=0 switches the
    speculative parallel record walk on for small inputs: with record-aligned members (htslib's layout) every
    guess is a record boundary, with members cut anywhere every guess misses and is walked again."""
    if window:
        monkeypatch.setenv("PB_BAM_WINDOW", window)
    monkeypatch.setenv("PB_BAM_WALK_MIN", "0")
    rng = np.random.default_rng(9)
    lens = {"chrA": 300_000, "chrB": 40_000, "chrNone": 1000}
    reads = {"chrA": random_cigar_reads(rng, 30_000, 300_000, 290_000), "chrB": random_cigar_reads(rng, 4000, 40_000, 30_000)}
    recs = []
    for ci, c in enumerate(lens):
        for k, r in enumerate(sorted(reads.get(c, []), key=lambda x: x.reference_start)):
            recs.append((ci, r.reference_start, 16 if r.is_reverse else 0, r.cigartuples))
            if k % 500 == 0:
                recs.append((ci, r.reference_start, 4, []))
    recs.append((-1, -1, 4, []))
    path = str(tmp_path / "t.bam")
    bam_io.write_bam(path, lens, recs, block_bytes=20_000 if window else 60_000, record_aligned=aligned)
    assert os.path.getsize(path) > (3 * 70_000 if window else 0)
    hb = bam_io.batch_from_bam(path, threads=4)
    ref = pack_reads(reads, lens)
    assert hb.chroms == ref.chroms and hb.mapped == len(ref) == len(hb)
    for f in ("ref_start", "meta", "chrom_read_off", "blk_off", "blk"):
        assert (getattr(hb, f) == getattr(ref, f)).all(), f
    assert hb.max_span == ref.max_span and hb.max_block_len == ref.max_block_len
    single = bam_io.batch_from_bam(path, threads=1)
    assert (single.ref_start == hb.ref_start).all() and (single.blk == hb.blk).all()


def test_parallel_record_walk_with_records_spanning_members(tmp_path, monkeypatch):
    """Record-aligned members of 2 kB with a 3000-nt read now and then: the long records span several members, so
    some of the walk's guesses are record boundaries and some fall inside a record; the result has to be the
    serial walk's either way."""
    monkeypatch.setenv("PB_BAM_WALK_MIN", "0")
    rng = np.random.default_rng(4)
    pos = np.sort(rng.integers(0, 900_000, 6000))
    recs = [(0, int(p), int(rng.integers(0, 2)) * 16,
             [(0, 3000)] if i % 97 == 5 else [(0, 12), (3, int(rng.integers(50, 900))), (0, 18)] if i % 5 == 0 else [(0, int(rng.integers(25, 36)))])
            for i, p in enumerate(pos)]
    path = str(tmp_path / "long.bam")
    bam_io.write_bam(path, {"c": 1_000_000}, recs, block_bytes=2000, record_aligned=True)
    one = bam_io.batch_from_bam(path, threads=1)
    assert len(one) == len(recs) and int(one.aligned_len.max()) == 3000
    assert [int(x) for x in one.ref_start[:50]] == [r[1] for r in recs[:50]]
    for threads in (2, 5, 8):
        many = bam_io.batch_from_bam(path, threads=threads)
        for f in ("ref_start", "meta", "chrom_read_off", "blk_off", "blk"):
            assert (getattr(many, f) == getattr(one, f)).all(), (threads, f)


def _pb_inflate(comp, n):
    """pb_inflate_raw into a buffer with 64 guard bytes behind it, which have to survive."""
    out = (C.c_uint8 * (n + 64))()
    C.memset(C.byref(out, n), 0xA5, 64)
    rc = _lib.lib().pb_inflate_raw(comp, len(comp), out, n)
    raw = bytes(out)
    assert raw[n:] == b"\xa5" * 64, "pb_inflate_raw wrote past the end of its output"
    return rc, raw[:n]


def _deflate_payload(rng, kind, n):
    if kind == 0:                                                # incompressible: stored blocks
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 1:                                                # four symbols: short codes, literal pairs
        return rng.integers(0, 4, n, dtype=np.uint8).tobytes()
    if kind == 2:                                                # one byte: distance-1 runs of 258
        return bytes(n)
    if kind == 3:                                                # period 50: far matches, chained
        return (rng.integers(0, 256, 50, dtype=np.uint8).tobytes() * (n // 50 + 1))[:n]
    if kind == 4:                                                # short periods: distances 1..7
        period = int(rng.integers(1, 8))
        return (rng.integers(0, 256, period, dtype=np.uint8).tobytes() * (n // period + 1))[:n]
    if kind == 5:                                                # halving frequencies: code lengths up to 15 (sub-tables)
        return np.minimum(rng.geometric(0.5, n) + rng.integers(0, 2, n) * 40, 255).astype(np.uint8).tobytes()
    if kind == 6:                                                # skewed qualities + random bases, like a BAM record
        q = 40 - np.minimum(rng.geometric(0.35, n), 38)
        noisy = rng.random(n) < 0.3
        q[noisy] = rng.integers(0, 256, int(noisy.sum()))
        return q.astype(np.uint8).tobytes()
    words = [rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8).tobytes() for _ in range(2000)]
    return b" ".join(words[int(i)] for i in rng.integers(0, 2000, n // 4 + 4))[:n]      # text: all of it mixed


def test_inflater_matches_zlib_on_every_block_shape():
    """pb_inflate_raw against zlib: stored / fixed / dynamic blocks, all levels and strategies, several blocks per
    stream (sync and full flushes), sizes around the fast loop's 320-byte margin and at a BGZF member's maximum;
    wrong sizes, truncated and corrupted input are refused without a write outside the output."""
    import zlib
    rng = np.random.default_rng(11)
    sizes = [0, 1, 2, 7, 100, 319, 320, 321, 1000, 5000, 65280]
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    for it in range(600):
        n = int(sizes[it % len(sizes)]) if it % 3 else int(rng.integers(0, 65281))
        data = _deflate_payload(rng, it % 8, n)
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, -15, int(rng.integers(1, 10)), int(rng.choice(strategies)))
        if n > 10 and rng.random() < 0.3:
            k = int(rng.integers(1, n))
            comp = co.compress(data[:k]) + co.flush(zlib.Z_SYNC_FLUSH if rng.random() < 0.5 else zlib.Z_FULL_FLUSH) \
                + co.compress(data[k:]) + co.flush()
        else:
            comp = co.compress(data) + co.flush()
        rc, out = _pb_inflate(comp, n)
        assert rc == 0 and out == data, (it, n)
        if n:
            assert _pb_inflate(comp, n - 1)[0] == -1 and _pb_inflate(comp, n + 1)[0] == -1
            assert _pb_inflate(comp[:int(rng.integers(0, len(comp) - 1))], n)[0] == -1
        if len(comp) > 4:
            bad = bytearray(comp)
            for _ in range(3):
                bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
            rc, out = _pb_inflate(bytes(bad), n)
            assert rc == -1 or zlib.decompress(bytes(bad), -15) == out       # accepted only if zlib reads the same


def test_inflater_refuses_malformed_headers():
    # block type 3; stored block with a bad complement; dynamic header with HLIT = 31 (288 codes)
    assert _pb_inflate(bytes([0b111]), 0)[0] == -1
    assert _pb_inflate(bytes([0b001, 5, 0, 0, 0]) + b"hello", 5)[0] == -1
    assert _pb_inflate(bytes([0b001, 5, 0, 0xFA, 0xFF]) + b"hello", 5) == (0, b"hello")
    assert _pb_inflate(bytes([0b11111101, 0xFF, 0xFF, 0xFF]), 10)[0] == -1
    assert _pb_inflate(b"", 0)[0] == -1 and _pb_inflate(bytes([0b011, 0]), 0) == (0, b"")      # empty fixed block


def test_inflater_fuzz_with_poisoned_tables(tmp_path):
    """The C++ fuzz driver (profiles/scripts/inflate_fuzz.cpp) built with PB_INFLATE_POISON: decode tables are filled
    with stale-looking entries before every build, so an entry a code does not define itself cannot go unnoticed
    (this is what caught literal pairing reading a previous block's entries)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "inflate_fuzz")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DPB_INFLATE_POISON", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "profiles", "scripts", "inflate_fuzz.cpp"),
                           os.path.join(root, "plastid_b200", "csrc", "pb_inflate.cpp"), "-o", exe, "-lz"])
    out = subprocess.run([exe, "1200"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok 1200"), out.stdout + out.stderr
    assert "disagreements with zlib 0" in out.stdout          # damaged streams: accepted exactly when zlib accepts them


def test_decoder_same_arrays_with_zlib(tmp_path, monkeypatch):
    rng = np.random.default_rng(2)
    pos = np.sort(rng.integers(0, 90_000, 20_000))
    recs = [(0, int(p), int(rng.integers(0, 2)) * 16, [(0, int(rng.integers(20, 40)))]) for p in pos]
    path = str(tmp_path / "p.bam")
    bam_io.write_bam(path, {"c": 100_000}, recs, record_aligned=True, payload_rng=rng)
    own = bam_io.batch_from_bam(path, threads=3)
    monkeypatch.setenv("PB_BAM_ZLIB", "1")
    ref = bam_io.batch_from_bam(path, threads=3)
    assert len(own) == len(recs) and (own.ref_start == ref.ref_start).all() and (own.meta == ref.meta).all()
    # and through read() into a window buffer instead of the mapped file (what a pipe gets), in small windows
    monkeypatch.delenv("PB_BAM_ZLIB")
    monkeypatch.setenv("PB_BAM_NOMMAP", "1")
    monkeypatch.setenv("PB_BAM_WINDOW", "70000")
    buffered = bam_io.batch_from_bam(path, threads=3)
    assert (buffered.ref_start == own.ref_start).all() and (buffered.meta == own.meta).all() and buffered.mapped == own.mapped


class _FakePysamRead(object):
    """The attributes of pysam.AlignedSegment the adaptor reads (genome_array.py:800-815, map_factories.pyx:243)."""

    def __init__(self, tid, read, unmapped=False):
        self.reference_id, self.reference_start = tid, read.reference_start
        self.cigartuples, self.is_reverse, self.is_unmapped = (None if unmapped else read.cigartuples), read.is_reverse, unmapped


class _FakePysamFile(object):
    def __init__(self, lens, reads_by_chrom):
        self.references, self.lengths = tuple(lens), tuple(lens.values())
        self._reads = []
        for tid, c in enumerate(lens):
            for k, r in enumerate(reads_by_chrom.get(c, [])):          # file order: unsorted on purpose
                self._reads.append(_FakePysamRead(tid, r))
                if k % 100 == 0:
                    self._reads.append(_FakePysamRead(tid, r, unmapped=True))
        self._reads.append(_FakePysamRead(-1, reads_by_chrom["chrA"][0]))
        self.mapped = sum(len(v) for v in reads_by_chrom.values())

    def fetch(self, until_eof=False):
        assert until_eof
        return iter(self._reads)


def test_open_pysam_handles_are_accepted_like_paths():
    """`BAMGenomeArray(pysam.AlignmentFile(...))` users: anything that is not a path goes through the pysam adaptor
    (pysam itself is not installed here, so a stand-in with the same attributes): same batch as the packer's."""
    rng = np.random.default_rng(21)
    lens = {"chrA": 50_000, "chrB": 9000, "chrEmpty": 100}
    reads = {"chrA": random_cigar_reads(rng, 3000, 50_000, 45_000), "chrB": random_cigar_reads(rng, 500, 9000, 6000)}
    hb = bam_io.batch_from_bam(_FakePysamFile(lens, reads))
    ref = pack_reads(reads, lens)
    assert hb.chroms == ref.chroms and len(hb) == len(ref) and hb.mapped == ref.mapped
    for f in ("ref_start", "meta", "chrom_read_off", "blk_off", "blk"):
        assert (getattr(hb, f) == getattr(ref, f)).all(), f
    assert hb.max_span == ref.max_span
    hb.check_sorted()


def test_unspliced_file_has_no_block_table(tmp_path):
    lens = {"c": 5000}
    recs = [(0, p, (p % 2) * 16, [(4, 2), (0, 20 + p % 7), (1, 3), (0, 5)]) for p in range(0, 4000, 3)]
    path = str(tmp_path / "u.bam")
    bam_io.write_bam(path, lens, recs)
    hb = bam_io.batch_from_bam(path)
    assert hb.blk is None and hb.blk_off is None and len(hb) == len(recs)
    assert list(hb.aligned_len[:3]) == [25 + 0, 25 + 3, 25 + 6] and hb.mapped == len(recs)


def test_malformed_input_is_reported(tmp_path):
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"this is not a bam file at all" * 10)
    with pytest.raises(_lib.PlastidB200Error, match="BGZF"):
        bam_io.batch_from_bam(str(bad))
    with pytest.raises(_lib.PlastidB200Error, match="cannot open"):
        bam_io.batch_from_bam(str(tmp_path / "missing.bam"))
    good = tmp_path / "g.bam"
    bam_io.write_bam(str(good), {"c": 1000}, [(0, 5, 0, [(0, 30)])] * 50)
    data = good.read_bytes()
    trunc = tmp_path / "trunc.bam"
    trunc.write_bytes(data[:len(data) // 2])
    with pytest.raises(_lib.PlastidB200Error):
        bam_io.batch_from_bam(str(trunc))
    # one flipped bit in the first member's CRC32 field (the member ends 8 bytes before the next one's magic)
    first_len = int.from_bytes(data[16:18], "little") + 1
    flipped = bytearray(data)
    flipped[first_len - 8] ^= 0x10
    crc = tmp_path / "crc.bam"
    crc.write_bytes(bytes(flipped))
    with pytest.raises(_lib.PlastidB200Error, match="CRC32"):
        bam_io.batch_from_bam(str(crc))
    # and one in its payload: either the inflater or the CRC refuses it
    flipped = bytearray(data)
    flipped[first_len - 12] ^= 0x04
    crc.write_bytes(bytes(flipped))
    with pytest.raises(_lib.PlastidB200Error, match="corrupt BGZF"):
        bam_io.batch_from_bam(str(crc))
    unsorted = tmp_path / "unsorted.bam"
    bam_io.write_bam(str(unsorted), {"a": 1000, "b": 1000}, [(1, 5, 0, [(0, 30)]), (0, 5, 0, [(0, 30)])])
    with pytest.raises(_lib.PlastidB200Error, match="sorted"):
        bam_io.batch_from_bam(str(unsorted))
    empty = tmp_path / "empty.bam"
    bam_io.write_bam(str(empty), {"a": 1000}, [])
    hb = bam_io.batch_from_bam(str(empty))
    assert len(hb) == 0 and hb.chroms == ["a"] and hb.mapped == 0


def test_corrupt_member_headers_are_errors_not_crashes(tmp_path):
    """BSIZE was never checked against XLEN — an 18-byte header with BSIZE 5 put the CRC / ISIZE reads before
    the mapping and crashed the process.  Every damaged header field must come back as an error."""
    good = tmp_path / "g.bam"
    bam_io.write_bam(str(good), {"c": 100000}, [(0, 5 + i, 0, [(0, 30)]) for i in range(3000)], block_bytes=8000)
    data = good.read_bytes()
    f = tmp_path / "h.bam"
    # the reported reproducer: a lone 18-byte header whose BSIZE field says 5
    hdr = bytearray(data[:18])
    hdr[16:18] = (5).to_bytes(2, "little")
    f.write_bytes(bytes(hdr))
    with pytest.raises(_lib.PlastidB200Error, match="corrupt BGZF member header"):
        bam_io.batch_from_bam(str(f))
    bsize = int.from_bytes(data[16:18], "little") + 1
    for name, off, value in (("BSIZE too small", 16, 10), ("BSIZE = header only", 16, 17), ("XLEN larger than BSIZE", 10, 60000),
                             ("SLEN past XLEN", 14, 40)):
        bad = bytearray(data)
        bad[off:off + 2] = value.to_bytes(2, "little")
        f.write_bytes(bytes(bad))
        with pytest.raises(_lib.PlastidB200Error):
            bam_io.batch_from_bam(str(f))
    bad = bytearray(data)                       # ISIZE beyond the 64 KiB the format allows: no multi-GB allocation
    bad[bsize - 4:bsize] = (0x7fff0000).to_bytes(4, "little")
    f.write_bytes(bytes(bad))
    with pytest.raises(_lib.PlastidB200Error, match="ISIZE"):
        bam_io.batch_from_bam(str(f))
    # fuzz: random bytes in the headers of the first members — an error or (when the mutation is harmless) the same
    # batch, never a crash
    ref = bam_io.batch_from_bam(str(good), pack=False)
    rng = np.random.default_rng(7)
    offs, o = [], 0
    while o + 18 <= len(data) and len(offs) < 6:
        offs.append(o)
        o += int.from_bytes(data[o + 16:o + 18], "little") + 1
    for _ in range(300):
        bad = bytearray(data)
        base = offs[int(rng.integers(len(offs)))]
        for _k in range(int(rng.integers(1, 4))):
            bad[base + int(rng.integers(0, 18))] = int(rng.integers(0, 256))
        f.write_bytes(bytes(bad))
        try:
            hb = bam_io.batch_from_bam(str(f), pack=False)
        except _lib.PlastidB200Error:
            continue
        assert len(hb) == len(ref) and (hb.ref_start == ref.ref_start).all()


def test_records_the_batch_cannot_hold_are_refused_or_skipped(tmp_path):
    """a negative position, a reference span beyond int32, or a CIGAR without aligned bases must not turn into
    garbage rows.  `mapped` is the index statistic (placed records without the unmapped flag), whatever the CIGAR."""
    f = str(tmp_path / "r.bam")
    bam_io.write_bam(f, {"c": 1000}, [(0, 5, 0, [(0, 30)]), (0, 7, 0, [(4, 20), (1, 3)]), (0, 9, 16, [(0, 25)]), (0, 11, 0, [])])
    hb = bam_io.batch_from_bam(f)
    assert len(hb) == 2 and list(hb.ref_start) == [5, 9] and hb.mapped == 4        # S/I-only and CIGAR-less records: no positions
    bam_io.write_bam(f, {"c": 1000}, [(0, -3, 0, [(0, 30)])])
    with pytest.raises(_lib.PlastidB200Error, match="negative position"):
        bam_io.batch_from_bam(f)
    bam_io.write_bam(f, {"c": 1000}, [(0, 2_000_000_000, 0, [(0, 30), (3, 200_000_000), (0, 10)])])
    with pytest.raises(_lib.PlastidB200Error, match="2\\^31"):
        bam_io.batch_from_bam(f)


def test_positions_match_reference_htslib_pileup():
    """SURVEY 8(a) row a1 for EVERY CIGAR op: the reference's vendored htslib pileup engine
    (kent/src/htslib/sam.c bam_plp_auto via oracle/_ref/ref_bam_tool positions; committed output
    tests/golden/htslib_allops.positions.txt) reports, per read, the reference positions that carry an
    aligned base.  The oracle's CIGAR walk, the host packer and the BAM decoder must give the same
    positions for all 2300 reads (M, I, D, N, S, H, P, =, X all occur)."""
    refs, recs = parse_dump(os.path.join(GOLD, "htslib_allops.dump.txt"))
    mapped = [r for r in recs if r[0] >= 0 and not (r[2] & 4) and r[3]]
    gold = {}
    for line in open(os.path.join(GOLD, "htslib_allops.positions.txt")):
        name, tid, runs = line.split()
        pos = []
        for run in runs.split(","):
            a, b = run.split("-")
            pos.extend(range(int(a), int(b)))
        gold[int(name[1:])] = (int(tid), pos)
    assert sorted(gold) == list(range(len(mapped)))             # reads are named r<n> in file order
    hb = bam_io.batch_from_bam(os.path.join(GOLD, "htslib_allops.bam"), threads=2)
    ops_seen = set()
    for i, (tid, pos, flag, cigar, endpos) in enumerate(mapped):
        ops_seen.update(op for op, _n in cigar)
        assert gold[i][0] == tid
        assert po.positions_from_cigar(pos, cigar) == gold[i][1]          # oracle == reference htslib
        blocks, _span = cigar_to_blocks(cigar)
        assert [pos + a + k for a, n in blocks for k in range(n)] == gold[i][1]   # host packer
        assert hb.positions_of(i) == gold[i][1]                                   # BAM decoder
    assert ops_seen == set(range(9))


def test_leading_deletions_reorder_rows_and_their_blocks(tmp_path, monkeypatch):
    """A leading D / N moves a read's first aligned position past the file's POS, possibly past its successors:
    rows (and the block rows of spliced reads among them) are put back in start order, per chromosome, stably."""
    monkeypatch.setenv("PB_BAM_WALK_MIN", "0")
    rng = np.random.default_rng(6)
    lens = {"chrA": 100_000, "chrB": 50_000}
    reads, recs = {"chrA": [], "chrB": []}, []
    for ci, c in enumerate(lens):
        for pos in np.sort(rng.integers(0, 40_000, 4000)).tolist():
            kind = int(rng.integers(0, 6))
            if kind == 0:
                cigar = [(2, int(rng.integers(1, 300))), (0, 25)]                             # leading deletion
            elif kind == 1:
                cigar = [(4, 3), (3, int(rng.integers(1, 200))), (0, 10), (3, 700), (0, 12)]    # clip, leading skip, spliced
            elif kind == 2:
                cigar = [(0, 14), (3, int(rng.integers(50, 900))), (0, 16), (2, 4), (0, 5)]
            else:
                cigar = [(0, int(rng.integers(20, 40)))]
            rev = bool(rng.integers(0, 2))
            reads[c].append(po.Read(pos, cigar, rev))
            recs.append((ci, pos, 16 if rev else 0, cigar))
    path = str(tmp_path / "lead.bam")
    bam_io.write_bam(path, lens, recs, block_bytes=30_000, record_aligned=True)
    ref = pack_reads(reads, lens)
    for threads in (1, 4):
        hb = bam_io.batch_from_bam(path, threads=threads)
        hb.check_sorted()
        assert len(hb) == len(ref) == 8000
        for f in ("ref_start", "meta", "chrom_read_off", "blk_off", "blk"):
            assert (getattr(hb, f) == getattr(ref, f)).all(), (threads, f)
    assert (np.diff(hb.ref_start[:4000]) >= 0).all() and int(hb.ref_start[0]) >= 0
    starts_in_file = np.array([r[1] for r in recs[:4000]])
    assert not np.array_equal(np.sort(hb.ref_start[:4000]), starts_in_file)                    # something did move


def test_speculative_walk_holds_on_a_file_written_by_the_reference_htslib(tmp_path):
    """The record walk guesses that BGZF members start at record boundaries.  That is how htslib writes (bgzf_flush_try
    before a record that does not fit): a 60 k-record BAM written by the reference's vendored htslib 1.3
    (oracle/_ref/ref_bam_tool sam2bam; only where /root/reference was there to build it) is walked without a single
    stretch done twice, and decodes to what htslib's own dump of it says."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_bam_tool")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/ref_bam_tool not built (needs the reference tree)")
    rng = np.random.default_rng(77)
    lens = {"chrA": 900_000, "chrB": 300_000}
    lines = ["@HD\tVN:1.4\tSO:coordinate"] + ["@SQ\tSN:%s\tLN:%d" % kv for kv in lens.items()]
    n = 0
    for chrom, count in (("chrA", 45_000), ("chrB", 15_000)):
        for pos in np.sort(rng.integers(0, lens[chrom] - 2000, count)).tolist():
            L = int(rng.integers(25, 76))
            cigar = "%dM" % L if n % 4 else "%dM%dN%dM" % (L // 2, int(rng.integers(50, 900)), L - L // 2)
            seq = "".join("ACGT"[k] for k in rng.integers(0, 4, L))
            qual = "".join(chr(33 + int(q)) for q in 40 - np.minimum(rng.geometric(0.35, L), 38))
            lines.append("read%07d\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t%s\tNM:i:%d" % (n, 16 * (n % 2), chrom, pos + 1, cigar, seq, qual, n % 3))
            n += 1
    sam, bam = str(tmp_path / "h.sam"), str(tmp_path / "h.bam")
    with open(sam, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    subprocess.check_call([tool, "sam2bam", sam, bam])
    assert os.path.getsize(bam) > 20 * 65536 // 4                                   # dozens of members
    code = ("import sys; sys.path.insert(0, %r); from plastid_b200.bam_io import batch_from_bam; "
            "b = batch_from_bam(%r, threads=6); print(len(b), b.mapped)" % (os.path.dirname(os.path.dirname(tool)), bam))
    env = dict(os.environ, PB_BAM_DEBUG="1", PB_BAM_WALK_MIN="0")
    run = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.split() == ["60000", "60000"], run.stderr
    walk = run.stderr.split("record walk:")[1].split()
    assert int(walk[0]) >= 4 and int(walk[2]) == 0, run.stderr                      # several stretches, none walked again
    hb = bam_io.batch_from_bam(bam, threads=6)
    dump = subprocess.check_output([tool, "dump", bam]).decode().splitlines()
    recs = [d.split() for d in dump if not d.startswith("@")]
    assert len(recs) == len(hb) == 60_000
    for i in (0, 1, 2, 29_999, 45_000, 59_999):
        tid, pos, flag = int(recs[i][0]), int(recs[i][1]), int(recs[i][2])
        cig = [tuple(int(v) for v in c.split(":")) for c in recs[i][4].split(",")]
        assert int(hb.ref_start[i]) == pos and bool((int(hb.meta[i]) >> 16) & 1) == bool(flag & 16)
        assert hb.positions_of(i) == po.positions_from_cigar(pos, cig)
    starts = np.array([int(r[1]) for r in recs])
    assert (hb.ref_start == starts).all()


# ------------------------------------------------------------------------------------------------------------------
# indexed access: pb_bai_open / pb_bam_fetch / pb_bam_build_index against the reference's vendored htslib
# (tests/golden/htslib_allops.bam.bai written by its indexer, htslib_allops.fetch.txt.gz by its iterator)
# ------------------------------------------------------------------------------------------------------------------
def parse_fetches(path):
    import gzip
    regions = []
    with gzip.open(path, "rt") as fh:
        for line in fh:
            f = line.split()
            if f[0] == "#":
                regions.append([tuple(int(x) for x in f[1:4]), []])
            elif f[0] == "=":
                assert int(f[1]) == len(regions[-1][1])
            else:
                cigar = [] if f[2] == "*" else [tuple(int(v) for v in c.split(":")) for c in f[2].split(",")]
                regions[-1][1].append((int(f[0]), int(f[1]), cigar, int(f[3])))
    return regions


def expected_rows(records):
    """What a batch holds of the records htslib's iterator returned: placed, not flagged unmapped, at least one
    aligned base; rows by first aligned position (stable)."""
    rows = []
    for pos, flag, cigar, _endpos in records:
        blocks, _span = cigar_to_blocks(cigar) if cigar else ([], 0)
        if flag & 4 or not blocks:
            continue
        rows.append((pos + blocks[0][0], sum(n for _a, n in blocks), (flag >> 4) & 1, len(blocks)))
    return sorted(rows, key=lambda r: r[0])


def batch_rows(hb):
    return [(int(hb.ref_start[i]), int(hb.meta[i]) & 0xFFFF, (int(hb.meta[i]) >> 16) & 1, int(hb.meta[i]) >> 24) for i in range(len(hb))]


def test_index_statistics_and_header_without_decoding():
    refs, recs = parse_dump(os.path.join(GOLD, "htslib_allops.dump.txt"))
    f = bam_io.IndexedBam(os.path.join(GOLD, "htslib_allops.bam"))
    assert list(f.references) == [r[0] for r in refs] and list(f.lengths) == [r[1] for r in refs]
    # `bamfile.mapped`: placed records without the unmapped flag, from the index's metadata pseudo-bins
    assert f.mapped == sum(1 for r in recs if r[0] >= 0 and not (r[2] & 4))
    L = _lib.lib()
    for t in range(len(refs)):
        assert L.pb_bai_mapped(f._idx, t) == sum(1 for r in recs if r[0] == t and not (r[2] & 4))
    f.close()
    with pytest.raises(ValueError):
        f.fetch(refs[0][0], 0, 10)


def test_fetch_returns_what_the_reference_htslib_iterator_returns():
    regions = parse_fetches(os.path.join(GOLD, "htslib_allops.fetch.txt.gz"))
    assert len(regions) >= 300 and sum(len(r) for _k, r in regions) > 4000
    f = bam_io.IndexedBam(os.path.join(GOLD, "htslib_allops.bam"))
    whole = bam_io.batch_from_bam(os.path.join(GOLD, "htslib_allops.bam"))
    for (tid, beg, end), records in regions:
        hb = f.fetch(f.references[tid], beg, end)
        assert batch_rows(hb) == expected_rows(records), (tid, beg, end)
        assert list(np.diff(hb.chrom_read_off)) == [len(hb) if t == tid else 0 for t in range(len(f.references))]
        for i in range(len(hb)):                      # block rows travel with their reads
            r0, r1 = int(whole.chrom_read_off[tid]), int(whole.chrom_read_off[tid + 1])
            cands = [j for j in range(r0, r1) if whole.ref_start[j] == hb.ref_start[i] and whole.meta[j] == hb.meta[i]]
            assert any(whole.positions_of(j) == hb.positions_of(i) for j in cands)
    # regions outside the file's references or past a chromosome's end are empty, as pysam's fetch is
    assert len(f.fetch("nope", 0, 100)) == 0
    assert len(f.fetch(f.references[2], f.lengths[2], f.lengths[2] + 50)) == 0
    assert len(f.fetch(f.references[0], 50, 50)) == 0


def _parse_bai(path):
    import struct
    raw = open(path, "rb").read()
    assert raw[:4] == b"BAI\1"
    n_ref, = struct.unpack_from("<i", raw, 4)
    p, refs = 8, []
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", raw, p); p += 4
        bins = {}
        for _b in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", raw, p); p += 8
            bins[b] = [struct.unpack_from("<QQ", raw, p + 16 * k) for k in range(n_chunk)]
            p += 16 * n_chunk
        n_intv, = struct.unpack_from("<i", raw, p); p += 4
        ioff = list(struct.unpack_from("<%dQ" % n_intv, raw, p)); p += 8 * n_intv
        refs.append((bins, ioff))
    tail = struct.unpack_from("<Q", raw, p)[0] if p + 8 <= len(raw) else None
    return refs, tail


def test_own_index_serves_the_same_fetches_as_the_reference_htslib_index(tmp_path):
    bam = os.path.join(GOLD, "htslib_allops.bam")
    mine = bam_io.build_index(bam, str(tmp_path / "mine.bai"))
    got, got_tail = _parse_bai(mine)
    want, want_tail = _parse_bai(bam + ".bai")
    assert len(got) == len(want) and got_tail == want_tail               # records without coordinates
    for (gb, gi), (wb, wi) in zip(got, want):
        assert gb.get(37450) == wb.get(37450)                            # file range + mapped / unmapped counts
        assert gi == wi                                                   # linear index, window by window
        # htslib moves the chunks of small bins into their parents; every record offset it lists must still be
        # covered by a chunk of ours in the same bin or a descendant of it — checked through the fetches below
    a, b = bam_io.IndexedBam(bam, index=mine), bam_io.IndexedBam(bam)
    assert a.mapped == b.mapped
    for (tid, beg, end), records in parse_fetches(os.path.join(GOLD, "htslib_allops.fetch.txt.gz")):
        assert batch_rows(a.fetch(a.references[tid], beg, end)) == expected_rows(records)


def test_build_index_refuses_unsorted_files_and_fetch_reports_bad_indexes(tmp_path):
    path = str(tmp_path / "u.bam")
    bam_io.write_bam(path, {"c": 1000}, [(0, 500, 0, [(0, 20)]), (0, 100, 0, [(0, 20)])])
    with pytest.raises(_lib.PlastidB200Error, match="not coordinate-sorted"):
        bam_io.build_index(path)
    open(str(tmp_path / "bad.bai"), "wb").write(b"BAI\1\x01\x00\x00\x00\x05")
    with pytest.raises(_lib.PlastidB200Error, match="truncated index"):
        bam_io.IndexedBam(os.path.join(GOLD, "htslib_allops.bam"), index=str(tmp_path / "bad.bai"))
    with pytest.raises(IOError):
        bam_io.IndexedBam(path)                                           # no index beside the file


@pytest.mark.parametrize("aligned", [False, True])
def test_fetch_equals_filtering_the_whole_file(tmp_path, aligned):
    """Many BGZF members, records cut across members (``aligned=False``), spliced reads longer than a 16 kb index
    window: every fetch equals the rows of the decoded file whose reference span overlaps the region."""
    rng = np.random.default_rng(21)
    lens = {"c1": 400000, "c2": 90000}
    recs, spans = [], []
    for ci, c in enumerate(lens):
        reads = sorted(random_cigar_reads(rng, 6000 if ci == 0 else 1500, lens[c], lens[c] - 30000), key=lambda r: r.reference_start)
        for k, r in enumerate(reads):
            cig = list(r.cigartuples)
            if k % 400 == 0:
                cig = cig + [(3, 20000), (0, 10)]                        # an intron across index windows
            recs.append((ci, r.reference_start, 16 if r.is_reverse else 0, cig))
    path = str(tmp_path / "big.bam")
    bam_io.write_bam(path, lens, recs, block_bytes=3000, record_aligned=aligned)
    bam_io.build_index(path)
    f = bam_io.IndexedBam(path)
    whole = bam_io.batch_from_bam(path)
    assert f.mapped == whole.mapped == len(recs)
    per = [[], []]
    for tid, pos, flag, cig in recs:
        blocks, span = cigar_to_blocks(cig)
        per[tid].append((pos, pos + max(span, 1), pos + blocks[0][0], sum(n for _a, n in blocks), (flag >> 4) & 1, len(blocks)))
    for _ in range(120):
        tid = int(rng.integers(0, 2))
        n = lens[f.references[tid]]
        beg = int(rng.integers(0, n))
        end = min(n, beg + int(rng.choice([1, 50, 3000, 40000])))
        want = sorted((r[2:] for r in per[tid] if r[0] < end and r[1] > beg), key=lambda r: r[0])
        assert batch_rows(f.fetch(f.references[tid], beg, end)) == want, (tid, beg, end)


def test_index_without_statistics_falls_back_to_decoding(tmp_path):
    """An index written without the metadata pseudo-bin (old indexers) has no ``mapped`` statistic: ``IndexedBam.mapped``
    is None, fetches still work, and ``BAMGenomeArray(indexed=True)`` decodes the file at once to know its sum."""
    import shutil
    import struct
    import plastid_b200 as pb
    bam = str(tmp_path / "x.bam")
    shutil.copy(os.path.join(GOLD, "htslib_allops.bam"), bam)
    raw = open(os.path.join(GOLD, "htslib_allops.bam.bai"), "rb").read()
    n_ref, = struct.unpack_from("<i", raw, 4)
    out, p = [raw[:8]], 8
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", raw, p); p += 4
        kept = []
        for _b in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", raw, p)
            size = 8 + 16 * n_chunk
            if b != 37450:
                kept.append(raw[p:p + size])
            p += size
        n_intv, = struct.unpack_from("<i", raw, p)
        out.append(struct.pack("<i", len(kept)) + b"".join(kept) + raw[p:p + 4 + 8 * n_intv])
        p += 4 + 8 * n_intv
    open(bam + ".bai", "wb").write(b"".join(out) + raw[p:])
    f = bam_io.IndexedBam(bam)
    assert f.mapped is None
    regions = parse_fetches(os.path.join(GOLD, "htslib_allops.fetch.txt.gz"))
    for (tid, beg, end), records in regions[:40]:
        assert batch_rows(f.fetch(f.references[tid], beg, end)) == expected_rows(records)
    ga = pb.BAMGenomeArray(bam, indexed=True, device="cpu")
    assert not ga.is_lazy and ga.sum() == bam_io.batch_from_bam(bam).mapped


def test_damaged_indexes_are_errors_or_answers_never_crashes(tmp_path):
    """Random byte damage in the .bai (bin numbers, chunk counts, virtual offsets, linear index): opening and fetching
    either raises PlastidB200Error or returns a batch; a damaged offset may lose reads but must not crash or hang."""
    rng = np.random.default_rng(99)
    bam = os.path.join(GOLD, "htslib_allops.bam")
    raw = bytearray(open(bam + ".bai", "rb").read())
    f_ok = bam_io.IndexedBam(bam)
    names, lens = f_ok.references, f_ok.lengths
    outcomes = {"error": 0, "answer": 0}
    for trial in range(150):
        bad = bytearray(raw)
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(8, len(bad)))
            bad[at] = int(rng.integers(0, 256))
        if trial % 10 == 0:
            bad = bad[:int(rng.integers(8, len(bad)))]                   # truncated file
        path = str(tmp_path / ("bad%d.bai" % trial))
        open(path, "wb").write(bytes(bad))
        try:
            f = bam_io.IndexedBam(bam, index=path)
            for _ in range(4):
                tid = int(rng.integers(0, len(names)))
                beg = int(rng.integers(0, max(lens[tid], 1)))
                hb = f.fetch(names[tid], beg, beg + int(rng.integers(1, 5000)))
                assert len(hb) >= 0
            f.close()
            outcomes["answer"] += 1
        except _lib.PlastidB200Error:
            outcomes["error"] += 1
    assert outcomes["answer"] + outcomes["error"] == 150 and outcomes["error"] > 0
    for n_ref in (2 ** 31 - 1, 10 ** 6, -5):                             # a reference count the file cannot hold
        bad = bytearray(raw)
        bad[4:8] = int(n_ref).to_bytes(4, "little", signed=True)
        path = str(tmp_path / "nref.bai")
        open(path, "wb").write(bytes(bad))
        with pytest.raises(_lib.PlastidB200Error):
            bam_io.IndexedBam(bam, index=path)
