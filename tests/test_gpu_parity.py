"""GPU parity: the CUDA path (through the C-ABI) against the oracles, bit-exact for integer rules,
rel <= 1e-6 (north_star; observed ~1e-15) for CenterMapFactory's fractional weights."""
import io
import warnings

import numpy as np
import pytest

import plastid_b200 as pb
from plastid_b200 import synth, _lib
from plastid_b200.genome_array import map_batch, region_sums, gather_windows, window_normalize, column_profile
from plastid_b200.regions import ChainTable
from oracle import pyoracle as po
from oracle import coracle
from helpers import kat_reads, kat_expected, random_cigar_reads

pytestmark = pytest.mark.gpu

CENTER_RTOL = 1e-6      # north_star tolerance for fractional weights


def plane_chrom(planes, layout, strand, c):
    base = int(layout.chrom_bin_off[c])
    t = planes.planes[strand][base:base + int(layout.chrom_len[c])].cpu().numpy()
    return t.view(np.uint32).astype(np.int64) if planes.dtype == "u32" else t


def padding_is_zero(planes, layout, strand):
    t = planes.planes[strand].cpu().numpy()
    t = t.view(np.uint32) if planes.dtype == "u32" else t
    for c in range(len(layout.chroms)):
        a = int(layout.chrom_bin_off[c]) + int(layout.chrom_len[c])
        if np.any(t[a:int(layout.chrom_bin_off[c + 1])] != 0):
            return False
    return True


# ---------------------------------------------------------------------------------------------
# the reference's own known answers, through the drop-in factories
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mapping", ["fiveprime", "threeprime", "center"])
@pytest.mark.parametrize("param", [0, 10])
@pytest.mark.parametrize("strand", ["+", "-"])
def test_kat_factories(cuda_device, mapping, param, strand):
    fac = {"fiveprime": pb.FivePrimeMapFactory, "threeprime": pb.ThreePrimeMapFactory,
           "center": pb.CenterMapFactory}[mapping](param)
    reads = kat_reads()[strand]
    reads_out, counts = fac(reads, pb.GenomicSegment("mock", 0, 2000, strand))
    assert reads_out == reads
    exp = kat_expected()[(mapping, param, strand)]
    if mapping == "center":
        np.testing.assert_allclose(counts, exp, rtol=CENTER_RTOL, atol=0)
        assert counts.dtype == np.float64
    else:
        assert (counts == exp).all() and counts.dtype == np.int64


@pytest.mark.parametrize("strand", ["+", "-"])
def test_kat_variable_and_from_file(cuda_device, strand):
    reads = kat_reads()[strand]
    seg = pb.GenomicSegment("mock", 0, 2000, strand)
    fancy = {L: L // 2 for L in range(25, 40)}
    expected = np.zeros(2000)
    for r in reads:
        idx = fancy[len(r.positions)]
        expected[r.positions[idx] if strand == "+" else r.positions[-idx - 1]] += 1
    for dict_, exp in (({"default": 0}, kat_expected()[("fiveprime", 0, strand)]), (fancy, expected)):
        _, counts = pb.VariableFivePrimeMapFactory(dict_)(reads, seg)
        assert (counts == exp).all()
        fh = io.StringIO("\n".join("%s\t%s" % kv for kv in dict_.items()))
        _, counts = pb.VariableFivePrimeMapFactory.from_file(fh)(reads, seg)
        assert (counts == exp).all()


@pytest.mark.parametrize("strand", ["+", "-"])
def test_kat_unmappable(cuda_device, strand):
    reads = kat_reads()[strand]
    seg = pb.GenomicSegment("mock", 0, 2000, strand)
    lens = [len(r.positions) for r in reads]
    cases = {"fiveprime": (pb.FivePrimeMapFactory(30), sum(L > 30 for L in lens)),
             "threeprime": (pb.ThreePrimeMapFactory(30), sum(L > 30 for L in lens)),
             "center": (pb.CenterMapFactory(15), sum(L > 30 for L in lens)),
             "variable": (pb.VariableFivePrimeMapFactory({25: 10, "default": 28}),
                          sum(L > 28 or L == 25 for L in lens))}
    for name, (fn, n_exp) in cases.items():
        pb.map_factories._warned.clear()
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            reads_out, counts = fn(reads, seg)
        assert len(reads_out) == n_exp, name
        assert abs(counts.sum() - n_exp) < 1e-9, name
        assert any(issubclass(x.category, pb.DataWarning) for x in w), name


# ---------------------------------------------------------------------------------------------
# operator on random CIGARs vs the Python oracle (every op, spliced, stratified quirk)
# ---------------------------------------------------------------------------------------------
def test_segment_operator_random_cigars(cuda_device):
    rng = np.random.default_rng(11)
    reads = random_cigar_reads(rng, 500, 8000, max_start=4000)
    offs = {L: L // 3 for L in range(10, 60)}
    offs["default"] = 2
    for strand in ("+", "-", "."):
        seg_o = po.Seg("c", 700, 4200, strand)
        seg = pb.GenomicSegment("c", 700, 4200, strand)
        pairs = [(po.FivePrimeMap(11), pb.FivePrimeMapFactory(11)), (po.ThreePrimeMap(4), pb.ThreePrimeMapFactory(4)),
                 (po.VariableFivePrimeMap(offs), pb.VariableFivePrimeMapFactory(offs)),
                 (po.StratifiedVariableFivePrimeMap({20: 3, 21: 20, "default": 30}, 18, 40),
                  pb.StratifiedVariableFivePrimeMapFactory({20: 3, 21: 20, "default": 30}, 18, 40)),
                 (po.CenterMap(0), pb.CenterMapFactory(0)), (po.CenterMap(7), pb.CenterMapFactory(7))]
        for ofn, gfn in pairs:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kept_o, exp = ofn(reads, seg_o)
                kept_g, got = gfn(reads, seg)
            assert kept_g == kept_o, (type(gfn).__name__, strand)
            assert got.shape == exp.shape
            if isinstance(gfn, pb.CenterMapFactory):
                np.testing.assert_allclose(got, exp, rtol=CENTER_RTOL, atol=1e-300)
                assert ((got == 0) == (exp == 0)).all()
            else:
                assert (got == exp).all(), (type(gfn).__name__, strand)


# ---------------------------------------------------------------------------------------------
# whole-genome planes vs the C oracle
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def small_world(cuda_device):
    chroms, lens = synth.yeast_like_genome(total=1_200_000, n_chrom=5)
    lens[2] = 16384 * 3                  # a chromosome ending exactly on the layout alignment
    lens[3] = 100_001
    ann = synth.make_annotation(chroms, lens, 200, seed=5, exons=(1, 3), exon_len=(200, 600), intron_len=(50, 400))
    dbatch = synth.riboseq_reads(ann, 400_000, seed=9, device=cuda_device, lengths=range(18, 41))
    hb = synth.device_batch_to_host(dbatch, chroms, lens)
    hb.check_sorted()
    layout = pb.GenomeLayout(chroms, lens)
    return dict(chroms=chroms, lens=lens, ann=ann, dbatch=dbatch, hb=hb, layout=layout)


POINT_CASES = [("fiveprime", pb.FivePrimeMapFactory(14), dict(rule="fiveprime", offset=14)),
               ("fiveprime0", pb.FivePrimeMapFactory(0), dict(rule="fiveprime", offset=0)),
               ("fiveprime30", pb.FivePrimeMapFactory(30), dict(rule="fiveprime", offset=30)),
               ("threeprime", pb.ThreePrimeMapFactory(15), dict(rule="threeprime", offset=15)),
               ("threeprime0", pb.ThreePrimeMapFactory(0), dict(rule="threeprime", offset=0))]


@pytest.mark.parametrize("name,factory,okw", POINT_CASES, ids=[c[0] for c in POINT_CASES])
@pytest.mark.parametrize("size_filter", [None, (25, 35)])
def test_point_planes_bit_exact(small_world, name, factory, okw, size_filter):
    w = small_world
    sf = None if size_filter is None else pb.SizeFilterFactory(*size_filter)
    planes = map_batch(w["dbatch"], w["layout"], factory, sf, strands=("+", "-", "."))
    total = {"+": 0, "-": 0, ".": 0}
    dropped = {}
    for strand in ("+", "-", "."):
        dropped[strand] = 0
        for c in range(len(w["chroms"])):
            exp, _, d, _ = coracle.genome_vector(w["hb"], c, strand, size_filter=size_filter, **okw)
            got = plane_chrom(planes, w["layout"], strand, c)
            assert (got == exp).all(), (name, strand, c)
            total[strand] += int(exp.sum())
            dropped[strand] += d
        assert padding_is_zero(planes, w["layout"], strand)
    st = planes.stats
    assert [int(st[_lib.PB_STAT_MAPPED_PLUS]), int(st[_lib.PB_STAT_MAPPED_MINUS]), int(st[_lib.PB_STAT_MAPPED_ANY])] \
        == [total["+"], total["-"], total["."]]
    assert [int(st[0]), int(st[1]), int(st[2])] == [dropped["+"], dropped["-"], dropped["."]]


@pytest.mark.parametrize("with_default", [True, False])
def test_variable_planes_bit_exact(small_world, with_default):
    w = small_world
    offs = dict(synth.RIBO_OFFSETS)
    if with_default:
        offs[20] = 25        # offset >= length with default: entry skipped, falls back to the default fill
    else:
        del offs["default"]  # lengths outside 25..35 have no offset: dropped + DataWarning
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fac = pb.VariableFivePrimeMapFactory(offs)
        luts = po.build_offset_luts(offs)
    assert (fac.forward_offsets == luts[0]).all() and (fac.reverse_offsets == luts[1]).all()
    planes = map_batch(w["dbatch"], w["layout"], fac, None, strands=("+", "-", "."))
    n_dropped = {"+": 0, "-": 0, ".": 0}
    for strand in ("+", "-", "."):
        for c in range(len(w["chroms"])):
            exp, _, d, _ = coracle.genome_vector(w["hb"], c, strand, rule="variable", luts=luts)
            assert (plane_chrom(planes, w["layout"], strand, c) == exp).all(), (strand, c)
            n_dropped[strand] += d
    assert (n_dropped["."] > 0) == (not with_default)
    assert [int(x) for x in planes.stats[:3]] == [n_dropped["+"], n_dropped["-"], n_dropped["."]]


@pytest.mark.parametrize("nibble", [0, 12, 20])
def test_center_planes(small_world, nibble):
    w = small_world
    planes = map_batch(w["dbatch"], w["layout"], pb.CenterMapFactory(nibble), None, strands=("+", "-", "."))
    for strand in ("+", "-", "."):
        for c in range(len(w["chroms"])):
            exp, _, _, _ = coracle.genome_vector(w["hb"], c, strand, nibble=nibble)
            got = plane_chrom(planes, w["layout"], strand, c)
            np.testing.assert_allclose(got, exp, rtol=CENTER_RTOL, atol=0)
            assert ((got == 0) == (exp == 0)).all()
        assert padding_is_zero(planes, w["layout"], strand)
    # each mapped read contributes exactly 1.0 in total
    tot = sum(float(planes.planes["."].sum().item()) for _ in [0])
    assert abs(tot - int(planes.stats[_lib.PB_STAT_MAPPED_ANY])) < 1e-6 * max(tot, 1)


def test_spliced_reads_planes(cuda_device):
    chroms, lens = ["a", "b", "c"], np.array([400_000, 16384, 250_000])
    dbatch = synth.rnaseq_reads(chroms, lens, 150_000, seed=4, device=cuda_device, intron=(100, 30000))
    hb = synth.device_batch_to_host(dbatch, chroms, lens)
    hb.check_sorted()
    assert hb.blk is not None and hb.max_span > 10_000
    layout = pb.GenomeLayout(chroms, lens)
    for fac, okw in ((pb.FivePrimeMapFactory(70), dict(rule="fiveprime", offset=70)),
                     (pb.ThreePrimeMapFactory(3), dict(rule="threeprime", offset=3))):
        planes = map_batch(dbatch, layout, fac, None, strands=("+", "-", "."))
        for strand in ("+", "-", "."):
            for c in range(3):
                exp, _, _, _ = coracle.genome_vector(hb, c, strand, **okw)
                assert (plane_chrom(planes, layout, strand, c) == exp).all(), (okw, strand, c)
    planes = map_batch(dbatch, layout, pb.CenterMapFactory(12), None, strands=("+", "-"))
    for strand in ("+", "-"):
        for c in range(3):
            exp, _, _, _ = coracle.genome_vector(hb, c, strand, nibble=12)
            got = plane_chrom(planes, layout, strand, c)
            np.testing.assert_allclose(got, exp, rtol=CENTER_RTOL, atol=0)
            assert ((got == 0) == (exp == 0)).all()


def test_empty_and_tiny_batches(cuda_device):
    import torch
    chroms, lens = ["x", "y"], np.array([1000, 50])
    layout = pb.GenomeLayout(chroms, lens)
    empty = pb.batch_from_arrays(chroms, lens, [], [], [], [])
    for fac in (pb.FivePrimeMapFactory(0), pb.CenterMapFactory(0)):
        planes = map_batch(empty.to_device(cuda_device), layout, fac, None, strands=("+", "-", "."))
        for s in ("+", "-", "."):
            assert float(planes.planes[s].double().abs().sum().item()) == 0.0
    one = pb.batch_from_arrays(chroms, lens, [1, 0, 0], [49, 999, 0], [1, 1, 30], [0, 1, 0])
    planes = map_batch(one.to_device(cuda_device), layout, pb.FivePrimeMapFactory(0), None, strands=("+", "-", "."))
    got = plane_chrom(planes, layout, ".", 0)
    assert got[0] == 1 and got[999] == 1 and got.sum() == 2
    assert plane_chrom(planes, layout, "-", 0)[999] == 1 and plane_chrom(planes, layout, "+", 1)[49] == 1


# ---------------------------------------------------------------------------------------------
# containers and reductions
# ---------------------------------------------------------------------------------------------
def test_bam_genome_array_matches_reference_semantics(cuda_device):
    rng = np.random.default_rng(21)
    lens = {"chrA": 5000, "chrB": 3000}
    reads = {"chrA": random_cigar_reads(rng, 700, 5000, max_start=4000),
             "chrB": random_cigar_reads(rng, 300, 3000, max_start=2000)}
    hb = pb.pack_reads(reads, lens)
    store = po.ReadStore(lens, reads)
    for ofn, gfn in ((po.FivePrimeMap(5), pb.FivePrimeMapFactory(5)), (po.CenterMap(3), pb.CenterMapFactory(3)),
                     (po.ThreePrimeMap(0), pb.ThreePrimeMapFactory(0))):
        oga = po.OracleBAMGenomeArray(store, mapping=ofn)
        ga = pb.BAMGenomeArray(hb, mapping=gfn, device=cuda_device)
        oga.add_filter("size", po.SizeFilter(15, 80))
        ga.add_filter("size", pb.SizeFilterFactory(15, 80))
        assert ga.sum() == oga.sum() and ga.chroms() == oga.chroms()
        for strand in ("+", "-", "."):
            for chrom, a, b in (("chrA", 0, 5000), ("chrA", 1234, 2345), ("chrB", 2900, 3000), ("chrZ", 5, 10)):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    exp = oga[po.Seg(chrom, a, b, strand)]
                    got = ga[pb.GenomicSegment(chrom, a, b, strand)]
                assert got.shape == exp.shape
                np.testing.assert_allclose(got, exp, rtol=CENTER_RTOL, atol=0)
                if not isinstance(gfn, pb.CenterMapFactory):
                    assert (got == exp).all()
        # get_reads_and_counts returns the reads the reference returns
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r_exp, c_exp = oga.get_reads_and_counts(po.Seg("chrA", 1000, 1500, "+"))
            r_got, c_got = ga.get_reads_and_counts(pb.GenomicSegment("chrA", 1000, 1500, "+"))
        assert r_got == r_exp
        np.testing.assert_allclose(c_got, c_exp, rtol=CENTER_RTOL, atol=0)
        # normalisation (genome_array.py:826-827)
        ga.set_normalize(True)
        oga.set_normalize(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np.testing.assert_allclose(ga[pb.GenomicSegment("chrA", 0, 5000, "-")],
                                       oga[po.Seg("chrA", 0, 5000, "-")], rtol=1e-12)
        # generic python filter is honoured (host keep-mask)
        ga.set_normalize(False)
        oga.set_normalize(False)
        ga.add_filter("odd", lambda r: r.reference_start % 2 == 1)
        oga.add_filter("odd", lambda r: r.reference_start % 2 == 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np.testing.assert_allclose(ga[pb.GenomicSegment("chrA", 0, 5000, ".")],
                                       oga[po.Seg("chrA", 0, 5000, ".")], rtol=CENTER_RTOL)


def test_segmentchain_counts_and_masks(small_world, cuda_device):
    w = small_world
    ga = pb.BAMGenomeArray(w["hb"], mapping=pb.FivePrimeMapFactory(14), device=cuda_device)
    ga.add_filter("size", pb.SizeFilterFactory(25, 100))
    chains = w["ann"].chains()
    masks = synth.make_masks(w["ann"], frac=0.2, seed=2)
    for ch, m in zip(chains, masks):
        ch.add_masks(*m)
    sums, live = ga.count_chains(chains)
    # oracle: per chromosome/strand vectors + python chain walk
    vecs = {}
    for c, chrom in enumerate(w["chroms"]):
        for strand in ("+", "-"):
            vecs[(chrom, strand)] = coracle.genome_vector(w["hb"], c, strand, rule="fiveprime", offset=14,
                                                          size_filter=(25, 100))[0]
    for i in list(range(0, len(chains), 7)) + [len(chains) - 1]:
        ch = chains[i]
        pos = np.asarray(ch.get_position_list())
        keep = ch.position_mask() == 0
        exp = vecs[(ch.chrom, ch.strand)][pos][keep].sum()
        assert sums[i] == exp and live[i] == keep.sum() == ch.masked_length, i
        # the per-chain object path gives the same masked vector, 5'->3'
        mc = ch.get_masked_counts(ga)
        ref = vecs[(ch.chrom, ch.strand)][pos].astype(float)
        refm = ~keep
        if ch.strand == "-":
            ref, refm = ref[::-1], refm[::-1]
        assert (mc.data == ref).all() and (mc.mask == refm).all()
    # C oracle for the whole table, unmasked
    table = ChainTable.from_chains(chains, ga.layout, use_masks=False)
    s2, l2 = ga.count_chains(table)
    for strand, pidx in (("+", 0), ("-", 1)):
        sel = np.nonzero(table.chain_plane == pidx)[0]
        big = np.zeros(ga.layout.total_bins, dtype=np.uint32)
        for c, chrom in enumerate(w["chroms"]):
            base = int(ga.layout.chrom_bin_off[c])
            big[base:base + int(w["lens"][c])] = vecs[(chrom, strand)]
        for i in sel[::11]:
            a, b = table.chain_off[i], table.chain_off[i + 1]
            es, el = coracle.region_sums(big, table.bstart[a:b], table.bend[a:b], [0, b - a])
            assert s2[i] == es[0] and l2[i] == el[0]


def test_window_matrix_and_profiles(small_world, cuda_device):
    import torch
    w = small_world
    ga = pb.BAMGenomeArray(w["hb"], mapping=pb.FivePrimeMapFactory(12), device=cuda_device)
    chains = w["ann"].chains()
    width, flank = 120, 30
    wins, cols = [], []
    rng = np.random.default_rng(0)
    for ch in chains:
        n = min(ch.length, int(rng.integers(40, width - 10)))
        win = ch.get_subchain(0, n)
        if rng.random() < 0.3:
            a = win.get_genomic_coordinate(int(rng.integers(0, n)))[1]
            win.add_masks(pb.GenomicSegment(win.chrom, a, a + 25, win.strand))
        wins.append(win)
        cols.append(int(rng.integers(0, width - n + 1)))
    planes = ga.count_planes(("+", "-"))
    table = ChainTable.from_chains(wins, ga.layout)
    mat, mmask = gather_windows(planes, table, cols, width)
    # oracle: metagene.py:895-914
    exp = np.ma.MaskedArray(np.tile(np.nan, (len(wins), width)), mask=np.tile(True, (len(wins), width)))
    for i, (win, col) in enumerate(zip(wins, cols)):
        mvec = win.get_masked_counts(ga)
        exp.data[i, col:col + win.length] = mvec.data
        exp.mask[i, col:col + win.length] = mvec.mask
    got = mat.cpu().numpy()
    gm = mmask.cpu().numpy().astype(bool)
    assert (gm == exp.mask).all()
    assert (got[~gm] == exp.data[~gm]).all() and np.isnan(got[np.isnan(exp.data)]).all()
    # metagene.py:918-953
    ns, ne, min_counts = flank, flank + 40, 8
    denom = np.nansum(exp[:, ns:ne], axis=1)
    row_select = denom >= min_counts
    norm = (exp.T.astype(float) / denom).T
    norm = np.ma.MaskedArray(norm, mask=exp.mask)
    with np.errstate(all="ignore"):
        norm.mask[np.isnan(norm)] = True
        norm.mask[np.isinf(norm)] = True
    d, sel, nmat, nmask = window_normalize(mat, mmask, ns, ne, min_counts)
    sel_h = sel.cpu().numpy().astype(bool)
    rs = np.ma.filled(row_select, False).astype(bool)
    assert (sel_h == rs).all()
    assert (nmask.cpu().numpy().astype(bool)[rs] == np.ma.getmaskarray(norm)[rs]).all()
    for mode, fn in (("median", np.ma.median), ("mean", np.ma.mean)):
        prof, nreg, csum = column_profile(nmat, nmask, sel, mode)
        e = fn(norm[rs], axis=0)
        np.testing.assert_allclose(prof.cpu().numpy(), np.ma.filled(e, np.nan), rtol=1e-12, equal_nan=True)
        assert (nreg.cpu().numpy() == (~np.ma.getmaskarray(norm))[rs].sum(0)).all()
    prof, nreg, csum = column_profile(mat, mmask, torch.ones_like(sel), "sum")
    np.testing.assert_allclose(prof.cpu().numpy(), np.ma.filled(exp.sum(0), 0.0), rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# GenomeArray / SparseGenomeArray: the reference's own known answers (test_roitools.py:1274-1379)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls", [pb.GenomeArray, pb.SparseGenomeArray])
def test_genome_array_masked_counts_known_answers(cuda_device, cls):
    for strand in ("+", "-"):
        ga = cls({"chrA": 2000}, device=cuda_device)
        ga[pb.GenomicSegment("chrA", 100, 200, strand)] = 1
        ga[pb.GenomicSegment("chrA", 250, 350, strand)] = 5
        chain = pb.SegmentChain(pb.GenomicSegment("chrA", 100, 150, strand), pb.GenomicSegment("chrA", 150, 200, strand),
                                pb.GenomicSegment("chrA", 250, 350, strand))
        unmasked = np.zeros(chain.length)
        if strand == "+":
            unmasked[:100], unmasked[100:200] = 1, 5
        else:
            unmasked[-100:], unmasked[-200:-100] = 1, 5
        assert (chain.get_masked_counts(ga) == unmasked).all()
        chain.add_masks(pb.GenomicSegment("chrA", 400, 500, strand))
        assert (chain.get_masked_counts(ga) == unmasked).all()
        chain.add_masks(pb.GenomicSegment("chrA", 50, 125, strand))
        mask = np.tile(False, chain.length)
        if strand == "+":
            mask[:25] = True
        else:
            mask[-25:] = True
        found = chain.get_masked_counts(ga)
        assert (found.mask == mask).all() and (found.data == unmasked).all()
        assert ga.sum() == 600.0
        sums, live = ga.count_chains([chain])
        assert sums[0] == 575.0 and live[0] == 175
    ga = cls({"chrA": 2000}, device=cuda_device)
    ga[pb.GenomicSegment("chrA", 100, 200, "+")] = 1
    ga[pb.GenomicSegment("chrA", 250, 350, "+")] = 1
    ivc = pb.SegmentChain(pb.GenomicSegment("chrA", 100, 150, "+"), pb.GenomicSegment("chrA", 150, 200, "+"),
                          pb.GenomicSegment("chrA", 250, 350, "+"))
    assert sum(ivc.get_counts(ga)) == 200 and ivc.get_masked_counts(ga).sum() == 200
    ivc.add_masks(pb.GenomicSegment("chrA", 50, 125, "+"))
    assert ivc.get_masked_counts(ga).sum() == 175 and sum(ivc.get_counts(ga)) == 200
    # vectors are set 5'->3' and read back the same way; chains set exon by exon (genome_array.py:1561-1575)
    vec = np.arange(30, dtype=float)
    rc = pb.SegmentChain(pb.GenomicSegment("chrA", 500, 510, "-"), pb.GenomicSegment("chrA", 600, 620, "-"))
    ga2 = cls({"chrA": 2000}, strands=("+", "-"), device=cuda_device)
    ga2[rc] = vec
    assert (ga2[rc] == vec).all() and (ga2[pb.GenomicSegment("chrA", 600, 620, "-")] == vec[:20]).all()
    assert (ga2.get(pb.GenomicSegment("chrA", 500, 510, "-"), roi_order=False) == vec[20:][::-1]).all()
    assert (ga2[pb.GenomicSegment("chrB", 0, 7, "+")] == 0).all()


def test_to_genome_array_drops_last_base_like_the_reference(cuda_device):
    hb = pb.batch_from_arrays(["c"], [100], [0, 0, 0], [0, 50, 99], [1, 1, 1], [0, 0, 1])
    ga = pb.BAMGenomeArray(hb, mapping=pb.FivePrimeMapFactory(0), device=cuda_device)
    dense = ga.to_genome_array()
    assert dense.sum() == 4.0            # '+': 0,50   '-': (99 dropped)   '.': 0,50 (99 dropped) -- genome_array.py:985
    assert dense[pb.GenomicSegment("c", 0, 100, "+")].sum() == 2 and dense[pb.GenomicSegment("c", 0, 100, "-")].sum() == 0
    assert ga[pb.GenomicSegment("c", 0, 100, "-")].sum() == 1


def test_wire16_unpack_on_device(small_world, cuda_device):
    from plastid_b200.batch import Wire16Batch, Wire16Receiver
    hb = small_world["hb"]
    wire = Wire16Batch.from_batch(hb)
    rx = Wire16Receiver(wire, cuda_device)
    db = rx.receive(wire.pinned())
    assert (db.ref_start.cpu().numpy() == hb.ref_start).all()
    assert (db.meta.cpu().numpy().view(np.uint32) == hb.meta).all()
    assert (db.chrom_read_off.cpu().numpy() == hb.chrom_read_off).all()
    planes = map_batch(db, small_world["layout"], pb.FivePrimeMapFactory(14), None, strands=("+",))
    exp = coracle.genome_vector(hb, 0, "+", rule="fiveprime", offset=14)[0]
    assert (plane_chrom(planes, small_world["layout"], "+", 0) == exp).all()


@pytest.mark.parametrize("n_chunks", [1, 3, 8])
def test_streamed_upload_matches_whole_batch(small_world, cuda_device, n_chunks):
    """wire16 chunks uploaded on a copy stream + pb_map_point_range per chunk == one whole launch."""
    import torch
    from plastid_b200.batch import Wire16Batch, Wire16Receiver
    from plastid_b200.genome_array import map_wire16_streamed
    w = small_world
    wire = Wire16Batch.from_batch(w["hb"])
    rx = Wire16Receiver(wire, cuda_device)
    rx.batch.ref_start.fill_(-7)                  # poison: nothing may be read before it has landed
    rx.batch.meta.fill_(0)
    chunks = Wire16Receiver.plan_chunks(wire, w["layout"], n_chunks)
    assert chunks[0][0] == 0 and chunks[-1][1] == len(wire) and chunks[-1][3] == w["layout"].total_bins
    fac = pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS)
    planes = map_wire16_streamed(rx, wire.pinned(), chunks, w["layout"], fac, pb.SizeFilterFactory(20, 38),
                                 ("+", "-", "."))
    torch.cuda.synchronize()
    ref = map_batch(w["dbatch"], w["layout"], fac, pb.SizeFilterFactory(20, 38), strands=("+", "-", "."))
    for s in ("+", "-", "."):
        assert torch.equal(planes.planes[s], ref.planes[s])
    assert (planes.stats_dev.cpu().numpy() == ref.stats).all()
    exp = coracle.genome_vector(w["hb"], 1, "-", rule="variable", luts=(fac.forward_offsets, fac.reverse_offsets),
                                size_filter=(20, 38))[0]
    assert (plane_chrom(planes, w["layout"], "-", 1) == exp).all()


def test_delta8_unpack_on_device(cuda_device):
    """2-byte transfer format: device expansion == the batch it was packed from, incl. exceptions
    (large gaps, chromosome changes inside a block, > 255 distinct meta words) and a ragged tail."""
    from plastid_b200.batch import Delta8Batch, Delta8Receiver
    from test_host_logic import _delta8_world
    for n_reads, rare in ((30000, True), (30001, False), (127, False), (3, False)):
        chroms, lens, hb = _delta8_world(n_reads, rare)
        wire = Delta8Batch.from_batch(hb)
        rx = Delta8Receiver(wire, cuda_device)
        rx.batch.ref_start.fill_(-7)
        db = rx.receive(wire.pinned())
        assert (db.ref_start.cpu().numpy() == hb.ref_start).all()
        assert (db.meta.cpu().numpy().view(np.uint32) == hb.meta).all()
        assert (db.chrom_read_off.cpu().numpy() == hb.chrom_read_off).all()


def test_delta3_unpack_on_device(cuda_device):
    """1-byte transfer format: device expansion == the batch it was packed from (wide deltas,
    exceptions, chromosome changes inside a block, > 31 distinct meta words, ragged tail)."""
    from plastid_b200.batch import Delta3Batch, Delta3Receiver
    from test_host_logic import _delta8_world
    for n_reads, rare in ((30000, True), (30001, False), (127, False), (3, False)):
        chroms, lens, hb = _delta8_world(n_reads, rare)
        wire = Delta3Batch.from_batch(hb)
        rx = Delta3Receiver(wire, cuda_device)
        rx.batch.ref_start.fill_(-7)
        db = rx.receive(wire.pinned())
        assert (db.ref_start.cpu().numpy() == hb.ref_start).all()
        assert (db.meta.cpu().numpy().view(np.uint32) == hb.meta).all()
        assert (db.chrom_read_off.cpu().numpy() == hb.chrom_read_off).all()


@pytest.mark.parametrize("fmt", ["delta8", "delta3"])
@pytest.mark.parametrize("n_chunks", [1, 3, 8])
def test_delta8_streamed_upload_matches_whole_batch(small_world, cuda_device, n_chunks, fmt):
    import torch
    from plastid_b200 import batch as pbatch
    from plastid_b200.genome_array import map_wire16_streamed
    Delta8Batch, Delta8Receiver = ((pbatch.Delta8Batch, pbatch.Delta8Receiver) if fmt == "delta8"
                                   else (pbatch.Delta3Batch, pbatch.Delta3Receiver))
    w = small_world
    wire = Delta8Batch.from_batch(w["hb"])
    rx = Delta8Receiver(wire, cuda_device)
    rx.batch.ref_start.fill_(-7)                  # poison: nothing may be read before it has landed
    rx.batch.meta.fill_(0)
    chunks = Delta8Receiver.plan_chunks(wire, w["layout"], n_chunks)
    fac = pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS)
    planes = map_wire16_streamed(rx, wire.pinned(), chunks, w["layout"], fac, pb.SizeFilterFactory(20, 38),
                                 ("+", "-", "."))
    torch.cuda.synchronize()
    ref = map_batch(w["dbatch"], w["layout"], fac, pb.SizeFilterFactory(20, 38), strands=("+", "-", "."))
    for s in ("+", "-", "."):
        assert torch.equal(planes.planes[s], ref.planes[s])
    assert (planes.stats_dev.cpu().numpy() == ref.stats).all()


@pytest.mark.parametrize("world_size", [2, 3, 7])
@pytest.mark.parametrize("spliced", [False, True])
def test_position_range_sharding_matches_whole_batch(small_world, cuda_device, world_size, spliced):
    """every rank maps only its bin range from the reads starting in it + a halo, into
    range-only planes; clipped region tables of all ranks add up to the whole table (one all-reduce)."""
    import torch
    from plastid_b200 import dist as pd
    from plastid_b200.batch import DeviceBatch
    from plastid_b200.genome_array import region_sums
    from plastid_b200.regions import ChainTable
    w = small_world
    if spliced:
        hb = synth.device_batch_to_host(synth.rnaseq_reads(w["chroms"], w["lens"], 60_000, seed=4, device="cpu",
                                                           intron=(50, 3000)), w["chroms"], w["lens"])
        fac, sf = pb.ThreePrimeMapFactory(3), None
    else:
        hb, fac, sf = w["hb"], pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS), pb.SizeFilterFactory(20, 38)
    lay = w["layout"]
    chains = w["ann"].chains()
    masks = synth.make_masks(w["ann"], frac=0.3, seed=2)
    for ch, m in zip(chains, masks):
        if m:
            ch.add_masks(*m)
    table = ChainTable.from_chains(chains, lay)
    whole = map_batch(DeviceBatch.from_host(hb, cuda_device), lay, fac, sf, strands=("+", "-"))
    ref_sums, ref_live = region_sums(whole, table)
    cuts = pd.position_cuts(hb, lay, world_size)
    assert cuts[0] == 0 and cuts[-1] == lay.total_bins and (np.diff(cuts) >= 0).all() and (cuts % 16384 == 0).all()
    sums = torch.zeros_like(ref_sums)
    live = torch.zeros_like(ref_live)
    mapped = np.zeros(_lib.PB_NSTATS, dtype=np.int64)
    for rank in range(world_size):
        sub, lo, hi = pd.shard_positions(hb, lay, rank, world_size, cuts)
        if hi == lo:
            continue
        planes = map_batch(DeviceBatch.from_host(sub, cuda_device), lay, fac, sf, strands=("+", "-"), bin_range=(lo, hi))
        for s in ("+", "-"):
            assert planes.planes[s].numel() == hi - lo
            assert torch.equal(planes.planes[s], whole.planes[s][lo:hi])
        s_r, l_r = region_sums(planes, pd.clip_table(table, lo, hi))
        sums += s_r
        live += l_r
        mapped += planes.stats
    assert torch.equal(sums, ref_sums) and torch.equal(live, ref_live)
    if not spliced:      # the device-side sharder picks the same cuts and the same reads
        dfull = DeviceBatch.from_host(hb, cuda_device)
        for rank in range(world_size):
            sub, lo, hi = pd.shard_positions(hb, lay, rank, world_size, cuts)
            dsub, dlo, dhi, dcuts = pd.shard_positions_device(dfull, lay, rank, world_size)
            assert (dlo, dhi) == (lo, hi) and (dcuts == cuts).all() and dsub.n_reads == len(sub)
            assert (dsub.ref_start.cpu().numpy() == sub.ref_start).all() and (dsub.chrom_read_off.cpu().numpy() == sub.chrom_read_off).all()
    for k in (_lib.PB_STAT_MAPPED_PLUS, _lib.PB_STAT_MAPPED_MINUS, _lib.PB_STAT_DROPPED_PLUS, _lib.PB_STAT_DROPPED_MINUS):
        assert mapped[k] == whole.stats[k]        # halo reads are not counted twice


def test_cuda_graph_replay_matches_eager_pass(small_world, cuda_device):
    import torch
    from plastid_b200.genome_array import GraphedCount
    w = small_world
    fac, sf = pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS), pb.SizeFilterFactory(20, 38)
    table = ChainTable.from_chains(w["ann"].chains(), w["layout"])
    eager = map_batch(w["dbatch"], w["layout"], fac, sf, strands=("+", "-"))
    e_sums, e_live = region_sums(eager, table)
    g = GraphedCount(w["dbatch"], w["layout"], fac, sf, table)
    for _ in range(3):
        g.planes.planes["+"].fill_(-1)            # every bin is rewritten by each replay
        sums, live = g.replay()
        torch.cuda.synchronize()
        assert torch.equal(sums, e_sums) and torch.equal(live, e_live)
        assert torch.equal(g.planes.planes["+"], eager.planes["+"]) and torch.equal(g.planes.planes["-"], eager.planes["-"])
    with pytest.raises(TypeError):
        GraphedCount(w["dbatch"], w["layout"], pb.CenterMapFactory(12), None, table)


def test_pileup_tile_is_split_into_overflow_jobs(cuda_device):
    """> 32768 candidate reads in one 4096-bin tile: the tile job keeps the first slice, the rest are
    added by pb_point_overflow_kernel with TMA bulk reductions — still bit-exact, stats included."""
    rng = np.random.default_rng(3)
    chroms, lens = ["a", "b"], np.array([30_000, 9_000])
    n_hot, n_bg = 150_000, 20_000
    start = np.concatenate([rng.integers(5000, 5300, n_hot), rng.integers(0, 29_900, n_bg), rng.integers(0, 8_900, 5000)])
    cid = np.concatenate([np.zeros(n_hot + n_bg, dtype=int), np.ones(5000, dtype=int)])
    L = rng.integers(20, 41, len(start))
    rev = rng.integers(0, 2, len(start))
    hb = pb.batch_from_arrays(chroms, lens, cid, start, L, rev)
    layout = pb.GenomeLayout(chroms, lens)
    offs = {k: k // 2 for k in range(24, 41)}            # lengths 20..23 have no offset: dropped
    fac = pb.VariableFivePrimeMapFactory(offs)
    planes = map_batch(hb.to_device(cuda_device), layout, fac, None, strands=("+", "-", "."))
    luts = (fac.forward_offsets, fac.reverse_offsets)
    dropped = {}
    for strand in ("+", "-", "."):
        dropped[strand] = 0
        for c in range(2):
            exp, _, d, _ = coracle.genome_vector(hb, c, strand, rule="variable", luts=luts)
            assert (plane_chrom(planes, layout, strand, c) == exp).all(), (strand, c)
            dropped[strand] += d
    assert [int(x) for x in planes.stats[:3]] == [dropped["+"], dropped["-"], dropped["."]] and dropped["."] > 1000
    assert int(planes.stats[_lib.PB_STAT_MAPPED_ANY]) == len(hb) - dropped["."]


@pytest.mark.parametrize("exact", [False, True])
def test_center_many_lengths_and_pileup_tiles(cuda_device, monkeypatch, exact):
    """Center rule on reads of 21 different lengths: the one-pass 64-bit fixed-point kernel (sparse tiles:
    plain shared atomics; pile-up tiles: run-aggregated adds) and, forced by PB_CENTER_EXACT, the exact
    multi-pass kernel — both within the north star's 1e-6 of the oracle with the same zero pattern,
    statistics included; repeated launches are bit-identical (integer accumulation)."""
    import torch
    if exact:
        monkeypatch.setenv("PB_CENTER_EXACT", "1")
    rng = np.random.default_rng(5)
    chroms, lens = ["a", "b"], np.array([60_000, 9_000])
    n_hot, n_bg = 60_000, 20_000
    start = np.concatenate([rng.integers(5000, 5200, n_hot), rng.integers(0, 59_900, n_bg), rng.integers(0, 8_900, 5000)])
    cid = np.concatenate([np.zeros(n_hot + n_bg, dtype=int), np.ones(5000, dtype=int)])
    L = rng.integers(20, 41, len(start))
    rev = rng.integers(0, 2, len(start))
    hb = pb.batch_from_arrays(chroms, lens, cid, start, L, rev)
    layout = pb.GenomeLayout(chroms, lens)
    fac = pb.CenterMapFactory(11)                      # L = 20, 21 are shorter than 2*nibble: dropped; L = 22: m = 0
    db = hb.to_device(cuda_device)
    planes = map_batch(db, layout, fac, None, strands=("+", "-", "."))
    first = {s: planes.planes[s].clone() for s in ("+", "-", ".")}
    dropped = {}
    for strand in ("+", "-", "."):
        dropped[strand] = 0
        for c in range(2):
            exp, _, d, _ = coracle.genome_vector(hb, c, strand, nibble=11)
            got = plane_chrom(planes, layout, strand, c)
            assert ((got == 0) == (exp == 0)).all(), (strand, c)
            np.testing.assert_allclose(got, exp, rtol=CENTER_RTOL, atol=0)
            if not exact:
                np.testing.assert_allclose(got, exp, rtol=1e-9, atol=0)     # the bound map_batch demands of the weights
            dropped[strand] += d
    assert [int(x) for x in planes.stats[:3]] == [dropped["+"], dropped["-"], dropped["."]] and dropped["."] > 1000
    again = map_batch(db, layout, fac, None, strands=("+", "-", "."))
    for s in ("+", "-", "."):
        assert torch.equal(again.planes[s], first[s])


def _pileup_batch(rng, lengths, n_hot=70_000, n_bg=15_000):
    """Reads piled up on two places of chromosome `a` (one of them across a tile boundary), a third of them spliced
    with their second block landing in ANOTHER pile (binned records), plus background."""
    chroms, lens = ["a", "b"], np.array([80_000, 9_000])
    start = np.concatenate([rng.integers(5000, 5150, n_hot), rng.integers(2040, 2060, n_hot // 2),
                            rng.integers(0, 69_000, n_bg), rng.integers(0, 8_800, 3000)])
    cid = np.concatenate([np.zeros(n_hot + n_hot // 2 + n_bg, dtype=int), np.ones(3000, dtype=int)])
    n = len(start)
    L = rng.choice(np.asarray(lengths), n)
    rev = rng.integers(0, 2, n)
    spliced = rng.random(n) < 0.33
    cut = np.minimum(rng.integers(5, 30, n), L - 3)
    gap = np.where(cid == 0, rng.integers(9000, 9040, n), rng.integers(30, 60, n))     # pile -> pile at +9000
    nb = np.where(spliced, 2, 1)
    rows = []
    for i in range(n):
        if spliced[i]:
            rows.append((0, cut[i])); rows.append((cut[i] + gap[i], L[i] - cut[i]))
        else:
            rows.append((0, L[i]))
    hb = pb.batch_from_arrays(chroms, lens, cid, start, L, rev, blocks=(nb, np.asarray(rows, dtype=np.int32)))
    return chroms, lens, hb


@pytest.mark.parametrize("mode", ["one_length", "fixed_point", "exact_multi"])
def test_center_pileup_tiles_are_split_bit_identically(cuda_device, monkeypatch, mode):
    """Pile-up tiles of the Center rule: candidate reads AND binned records of a tile beyond the split
    become overflow jobs whose integer partial difference arrays are reduced into a scratch tile.  Planes and
    statistics are bit-identical to the unsplit run for the exact kernel (one map length / several passes) and the
    64-bit fixed-point kernel, with small and default splits, whole genome and as bin ranges; and right against the oracle."""
    import torch
    rng = np.random.default_rng(17)
    lengths = [40] if mode == "one_length" else list(range(24, 45))
    if mode == "exact_multi":
        monkeypatch.setenv("PB_CENTER_EXACT", "1")
    chroms, lens, hb = _pileup_batch(rng, lengths)
    assert hb.blk is not None
    layout = pb.GenomeLayout(chroms, lens)
    fac = pb.CenterMapFactory(12)
    db = hb.to_device(cuda_device)
    monkeypatch.setenv("PB_CENTER_SPLIT", "0")                     # never split: one CTA walks every pile
    ref = map_batch(db, layout, fac, None, strands=("+", "-", "."))
    ref2 = map_batch(db, layout, fac, None, strands=("+", "-"))   # (several passes group the map lengths by plane count)
    ref_r = map_batch(db, layout, fac, None, strands=("+", "."))
    exp = coracle.genome_vector(hb, 0, "+", nibble=12)[0]
    got = plane_chrom(ref, layout, "+", 0)
    assert ((got == 0) == (exp == 0)).all()
    np.testing.assert_allclose(got, exp, rtol=CENTER_RTOL, atol=0)
    for split in ("256", "1000", None):
        if split is None:
            monkeypatch.delenv("PB_CENTER_SPLIT")
        else:
            monkeypatch.setenv("PB_CENTER_SPLIT", split)
        out = map_batch(db, layout, fac, None, strands=("+", "-", "."))
        for s in "+-.":
            assert torch.equal(out.planes[s], ref.planes[s]), (mode, split, s)
        assert (np.asarray(out.stats) == np.asarray(ref.stats)).all(), (mode, split)
        two = map_batch(db, layout, fac, None, strands=("+", "-"))          # the specialised two-plane kernels
        for s in "+-":
            assert torch.equal(two.planes[s], ref2.planes[s]), (mode, split, s)
        total, cut = int(layout.total_bins), 16384                           # bin ranges (position sharding)
        parts = [map_batch(db, layout, fac, None, strands=("+", "."), bin_range=r) for r in ((0, cut), (cut, total))]
        for s in "+.":
            assert torch.equal(torch.cat([q.planes[s] for q in parts]), ref_r.planes[s]), (mode, split, s, "ranges")


def test_gpu_reproduces_the_reference_golden_vector_recipe(cuda_device):
    """The CUDA path against count vectors built by the reference's own recipe
    (test_genome_array.py:1832-1866; tests/helpers.py:genome_array_recipe): 5'/3' at offsets 0 and 15 bit
    for bit, center 0/12 within the reference's 1e-8, spliced reads included — whole-genome planes
    and the per-segment operator."""
    from helpers import genome_array_recipe
    reads, vectors = genome_array_recipe(seed=11)
    n = len(next(iter(vectors.values())))
    hb = pb.pack_reads({"chrA": reads}, {"chrA": n})
    assert hb.blk is not None
    layout = pb.GenomeLayout(["chrA"], [n])
    db = hb.to_device(cuda_device)
    facs = {"fiveprime": pb.FivePrimeMapFactory, "threeprime": pb.ThreePrimeMapFactory, "center": pb.CenterMapFactory}
    for (rule, par, strand), exp in vectors.items():
        planes = map_batch(db, layout, facs[rule](par), None, strands=(strand,))
        got = plane_chrom(planes, layout, strand, 0)
        ga = pb.BAMGenomeArray(hb, mapping=facs[rule](par), device=cuda_device)
        seg = ga.get(pb.GenomicSegment("chrA", 100, n - 100, strand), roi_order=False)
        if rule == "center":
            np.testing.assert_allclose(got, exp, rtol=0, atol=1e-8)
            np.testing.assert_allclose(seg, exp[100:n - 100], rtol=0, atol=1e-8)
        else:
            assert (got == exp).all() and (seg == exp[100:n - 100]).all(), (rule, par, strand)


def test_bam_genome_array_from_bam_file(tmp_path, cuda_device):
    """BAMGenomeArray("x.bam"): decoded by the library's own BGZF/BAM reader, no pysam."""
    from plastid_b200 import bam_io
    rng = np.random.default_rng(8)
    lens = {"chrA": 9000, "chrB": 4000}
    reads = {"chrA": random_cigar_reads(rng, 1500, 9000, 7000), "chrB": random_cigar_reads(rng, 500, 4000, 3000)}
    recs = []
    for ci, c in enumerate(lens):
        for r in sorted(reads[c], key=lambda x: x.reference_start):
            recs.append((ci, r.reference_start, 16 if r.is_reverse else 0, r.cigartuples))
    path = str(tmp_path / "x.bam")
    bam_io.write_bam(path, lens, recs)
    ga = pb.BAMGenomeArray(path, mapping=pb.ThreePrimeMapFactory(2), device=cuda_device)
    oga = po.OracleBAMGenomeArray(po.ReadStore(lens, reads), mapping=po.ThreePrimeMap(2))
    assert ga.sum() == oga.sum() == 2000 and ga.chroms() == ["chrA", "chrB"] and ga.lengths() == lens
    for strand in ("+", "-", "."):
        for chrom, a, b in (("chrA", 0, 9000), ("chrB", 100, 3900)):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                exp = oga[po.Seg(chrom, a, b, strand)]
                got = ga[pb.GenomicSegment(chrom, a, b, strand)]
            assert (got == exp).all(), (strand, chrom)


def test_indexed_bam_genome_array_seeks_and_equals_the_decoded_one(tmp_path, cuda_device):
    """``BAMGenomeArray(path, indexed=True)``: region queries go through the .bai (the reference's access pattern,
    genome_array.py:800-809) and give what the whole-file container gives — counts, reads returned, sum from the
    index statistic, normalisation, generic filters — for point, Center and stratified rules over two files; a
    whole-genome consumer (``count_chains``) decodes the files on first use."""
    from plastid_b200 import bam_io
    rng = np.random.default_rng(18)
    lens = {"chrA": 60000, "chrB": 20000}
    paths = []
    for k in range(2):
        recs = []
        for ci, c in enumerate(lens):
            reads = sorted(random_cigar_reads(rng, 3000 if ci == 0 else 800, lens[c], lens[c] - 2000), key=lambda x: x.reference_start)
            recs += [(ci, r.reference_start, 16 if r.is_reverse else 0, r.cigartuples) for r in reads]
        paths.append(str(tmp_path / ("f%d.bam" % k)))
        bam_io.write_bam(paths[-1], lens, recs, block_bytes=4000)
        bam_io.build_index(paths[-1])
    segs = []
    for _ in range(25):
        c = "chrA" if rng.random() < 0.7 else "chrB"
        a = int(rng.integers(0, lens[c] - 10))
        segs.append(pb.GenomicSegment(c, a, min(lens[c], a + int(rng.choice([1, 30, 700, 9000]))), "+-."[int(rng.integers(0, 3))]))
    segs.append(pb.GenomicSegment("chrB", 19990, 20040, "+"))               # past the chromosome's end
    segs.append(pb.GenomicSegment("nope", 0, 10, "+"))
    for mapping in (pb.FivePrimeMapFactory(3), pb.CenterMapFactory(4), pb.VariableFivePrimeMapFactory({20: 2, 30: 5, "default": 1}),
                    pb.StratifiedVariableFivePrimeMapFactory({20: 2, 30: 5, "default": 1}, 15, 40)):
        lazy = pb.BAMGenomeArray(*paths, mapping=mapping, device=cuda_device, indexed=True)
        eager = pb.BAMGenomeArray(*paths, mapping=mapping, device=cuda_device)
        assert lazy.is_lazy and lazy.sum() == eager.sum() == 7600
        assert lazy.chroms() == eager.chroms() and lazy.lengths() == eager.lengths()
        lazy.add_filter("size", pb.SizeFilterFactory(12, 200)); eager.add_filter("size", pb.SizeFilterFactory(12, 200))
        lazy.add_filter("fw5", lambda r: r.reference_start % 5 != 0); eager.add_filter("fw5", lambda r: r.reference_start % 5 != 0)
        lazy.set_normalize(True); eager.set_normalize(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for seg in segs:
                ra, ca = lazy.get_reads_and_counts(seg)
                rb, cb = eager.get_reads_and_counts(seg)
                # integer rules: bit-exact; Center: float64 atomics in whatever order the reads arrive
                same = np.allclose(ca, cb, rtol=1e-8, atol=0) if isinstance(mapping, pb.CenterMapFactory) else (ca == cb).all()
                assert ca.shape == cb.shape and same, (type(mapping).__name__, str(seg))
                key = lambda r: (r.reference_start, r.is_reverse, tuple(r.positions))
                assert sorted(map(key, ra)) == sorted(map(key, rb))
                ga, gb = lazy[seg], eager[seg]
                assert ga.shape == gb.shape and np.allclose(ga, gb, rtol=1e-8, atol=0)      # eager Center planes: fixed-point weights
        assert lazy.is_lazy                                                  # nothing decoded so far
    lazy = pb.BAMGenomeArray(*paths, mapping=pb.FivePrimeMapFactory(3), device=cuda_device, indexed=True)
    eager = pb.BAMGenomeArray(*paths, mapping=pb.FivePrimeMapFactory(3), device=cuda_device)
    lazy.set_sum(123)
    chains = [pb.SegmentChain(pb.GenomicSegment("chrA", 100, 900, "+"), pb.GenomicSegment("chrA", 2000, 2500, "+")),
              pb.SegmentChain(pb.GenomicSegment("chrB", 50, 15000, "-"))]
    sa, la = lazy.count_chains(chains)
    sb, lb = eager.count_chains(chains)
    assert not lazy.is_lazy and (sa == sb).all() and (la == lb).all() and sa.sum() > 0 and lazy.sum() == 123


def test_staged_upload_of_large_pageable_arrays(cuda_device, monkeypatch):
    """Plain SoA batches above 64 MB per array are uploaded through a ring of pinned buffers filled by host threads
    (batch._staged_upload); with the thresholds turned down, a spliced batch arrives bit for bit — 1-D and 2-D arrays,
    sizes that are not a multiple of the chunk."""
    import torch
    from plastid_b200 import batch as pbatch
    chroms, lens = synth.yeast_like_genome(total=800_000, n_chrom=3)
    hb = synth.device_batch_to_host(synth.rnaseq_reads(chroms, lens, 70_001, seed=5, device="cpu", intron=(50, 900)), chroms, lens)
    monkeypatch.setattr(pbatch, "_STAGED_MIN_BYTES", 1 << 10)
    monkeypatch.setattr(pbatch, "_STAGED_CHUNK", (1 << 16) + 24)
    pbatch._staging.clear()
    try:
        db = pbatch.DeviceBatch.from_host(hb, cuda_device)
        torch.cuda.synchronize()
        assert (db.ref_start.cpu().numpy() == hb.ref_start).all()
        assert (db.meta.cpu().numpy().view(np.uint32) == hb.meta).all()
        assert (db.blk_off.cpu().numpy().view(np.uint32) == hb.blk_off).all() and db.blk.shape == hb.blk.shape
        assert (db.blk.cpu().numpy() == hb.blk).all() and (db.chrom_read_off.cpu().numpy() == hb.chrom_read_off).all()
    finally:
        pbatch._staging.clear()


def test_golden_bam_count_vectors_from_reference_htslib_positions(cuda_device):
    """End of the chain for non-M CIGARs without the oracle in between: the committed golden BAM (every
    CIGAR op; written and piled up by the reference's vendored htslib, tests/golden/htslib_allops.*) is
    decoded and mapped on the device; the expected vectors are built directly from the pileup-derived
    ``positions`` of each read with the rules' definitions (map_factories.pyx:345-353, 444-452, 246-254)."""
    import os
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    lens, flags = [], []
    for line in open(os.path.join(gold_dir, "htslib_allops.dump.txt")):
        f = line.split()
        if f[0] == "@":
            lens.append((f[1], int(f[2])))
        elif int(f[0]) >= 0 and not (int(f[2]) & 4) and int(f[3]):
            flags.append(int(f[2]))
    reads = []
    for line in open(os.path.join(gold_dir, "htslib_allops.positions.txt")):
        name, tid, runs = line.split()
        pos = [p for run in runs.split(",") for p in range(int(run.split("-")[0]), int(run.split("-")[1]))]
        reads.append((int(tid), pos, bool(flags[int(name[1:])] & 16)))
    assert len(reads) == len(flags) == 2300

    def expected(rule, param, strand):
        out = {c: np.zeros(n, dtype=np.float64) for c, n in lens}
        for tid, pos, is_rev in reads:
            if strand != "." and is_rev != (strand == "-"):
                continue
            vec, L = out[lens[tid][0]], len(pos)
            if rule == "center":
                m = L - 2 * param
                if m > 0:
                    for p in pos[param:L - param]:
                        vec[p] += 1.0 / m
            elif param < L:
                from_left = (rule == "fiveprime") == (strand != "-")
                vec[pos[param] if from_left else pos[L - 1 - param]] += 1
        return out

    path = os.path.join(gold_dir, "htslib_allops.bam")
    for rule, param, factory in (("fiveprime", 0, pb.FivePrimeMapFactory(0)), ("fiveprime", 7, pb.FivePrimeMapFactory(7)),
                                 ("threeprime", 3, pb.ThreePrimeMapFactory(3)), ("center", 5, pb.CenterMapFactory(5))):
        ga = pb.BAMGenomeArray(path, mapping=factory, device=cuda_device)
        for strand in ("+", "-", "."):
            exp = expected(rule, param, strand)
            for chrom, n in lens:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    got = ga.get(pb.GenomicSegment(chrom, 0, n, strand), roi_order=False)
                if rule == "center":
                    assert ((got == 0) == (exp[chrom] == 0)).all()
                    np.testing.assert_allclose(got, exp[chrom], rtol=1e-6, atol=0)      # north-star tolerance
                else:
                    assert (got == exp[chrom]).all(), (rule, param, strand, chrom)
                assert exp[chrom].sum() > 0 or chrom == "chrEmpty"


def test_spliced_transfer_format_round_trip(cuda_device):
    """delta3 + block words (pb_unpack_delta3, pb_unpack_blocks): the device batch rebuilt from the
    4-byte-per-block transfer format equals the SoA batch, incl. exception rows, and maps identically."""
    import torch
    from plastid_b200.batch import Delta3SplicedBatch, Delta3SplicedReceiver, DeviceBatch
    rng = np.random.default_rng(13)
    lens = {"chrA": 3_000_000, "chrB": 50_000}
    reads = {"chrA": random_cigar_reads(rng, 20_000, 3_000_000, 900_000), "chrB": random_cigar_reads(rng, 3000, 50_000, 40_000)}
    reads["chrA"].append(po.Read(1000, [(0, 30), (3, 1_500_000), (0, 20)], False))
    reads["chrA"].append(po.Read(2000, [(0, 5000), (3, 70), (0, 4095), (2, 1), (0, 4096)], True))
    hb = pb.pack_reads(reads, lens, keep_objects=False)
    wire = Delta3SplicedBatch.from_batch(hb)
    assert len(wire.bexc_row) >= 2
    rx = Delta3SplicedReceiver(wire, cuda_device)
    pinned = wire.pinned()
    for _ in range(2):                                              # buffers are reusable
        db = rx.receive(pinned)
        torch.cuda.synchronize()
        assert (db.ref_start.cpu().numpy() == hb.ref_start).all()
        assert (db.meta.cpu().numpy().view(np.uint32) == hb.meta).all()
        assert (db.chrom_read_off.cpu().numpy() == hb.chrom_read_off).all()
        assert (db.blk_off.cpu().numpy().view(np.uint32) == hb.blk_off).all()
        assert (db.blk.cpu().numpy() == hb.blk).all()
        assert db.n_blk == len(hb.blk) and db.max_block_len == hb.max_block_len
    layout = pb.GenomeLayout(list(lens), list(lens.values()))
    ref = DeviceBatch.from_host(hb, cuda_device)
    for fac in (pb.FivePrimeMapFactory(3), pb.CenterMapFactory(2)):
        a = map_batch(db, layout, fac, None, strands=("+", "-"))
        b = map_batch(ref, layout, fac, None, strands=("+", "-"))
        for strand in ("+", "-"):
            assert torch.equal(a.planes[strand], b.planes[strand])


def test_genome_array_setitem_known_answers_from_the_reference_tests(cuda_device):
    """plastid/test/unit/genomics/test_genome_array.py:811-898 (scalar / vector ``__setitem__`` over
    segments and chains, both strands, incl. a chain running past the declared chromosome end) and the
    auto-grow / ``+=`` part of ``test_setters_and_getters`` (:1146-1160), for GenomeArray and
    SparseGenomeArray."""
    rng = np.random.default_rng(17)
    for cls in (pb.GenomeArray, pb.SparseGenomeArray):
        ga = cls({"chrA": 2000}, device=cuda_device)
        segplus, segminus = pb.GenomicSegment("chrA", 50, 100, "+"), pb.GenomicSegment("chrA", 50, 100, "-")
        ga[segplus] = 52
        ga[segminus] = 342
        assert (ga.get(segplus, roi_order=False) == 52).all() and (ga.get(segminus, roi_order=False) == 342).all()
        assert ga.sum() == 52 * len(segplus) + 342 * len(segminus)

        ga = cls({"chrA": 2000}, device=cuda_device)
        r1, r2 = rng.integers(0, 242, 50), rng.integers(0, 242, 50)
        ga[segplus] = r1
        ga[segminus] = r2
        assert (ga.get(segplus, roi_order=False) == r1).all()
        assert (ga.get(segminus, roi_order=False) == r2[::-1]).all() and (ga[segminus] == r2).all()
        assert ga.sum() == r1.sum() + r2.sum()

        blocks = ((50, 100), (150, 732), (1800, 2500))
        pluschain = pb.SegmentChain(*[pb.GenomicSegment("chrA", a, b, "+") for a, b in blocks])
        minuschain = pb.SegmentChain(*[pb.GenomicSegment("chrA", a, b, "-") for a, b in blocks])
        ga = cls({"chrA": 2000}, device=cuda_device)
        ga[pluschain] = 31
        ga[minuschain] = 424
        for seg in pluschain:
            assert (ga.get(seg, roi_order=False) == 31).all()
        for seg in minuschain:
            assert (ga.get(seg, roi_order=False) == 424).all()
        assert ga.sum() == 31 * pluschain.length + 424 * minuschain.length

        ga = cls({"chrA": 2000}, device=cuda_device)
        plusvec, minusvec = rng.integers(0, 250, pluschain.length), rng.integers(0, 250, minuschain.length)
        ga[pluschain] = plusvec
        ga[minuschain] = minusvec
        x = 0
        for seg in pluschain:
            sub = ga.get(seg, roi_order=False)
            assert (sub == plusvec[x:x + len(sub)]).all()
            x += len(sub)
        x = 0
        for seg in minuschain:
            sub = ga.get(seg, roi_order=False)[::-1]
            assert (sub == minusvec[len(minusvec) - x - len(sub):len(minusvec) - x]).all()
            x += len(sub)
        assert ga.sum() == plusvec.sum() + minusvec.sum()
        assert (pluschain.get_counts(ga) == plusvec).all() and (minuschain.get_counts(ga) == minusvec).all()

        gnd = cls({"chrA": 1000, "chrB": 10000}, device=cuda_device)
        iv1, iv2 = pb.GenomicSegment("chrA", 10000, 11000, "+"), pb.GenomicSegment("chrA", 10500, 11000, "+")
        iv3 = pb.GenomicSegment("chrA", 500000 + 10500, 500000 + 11000, "+")
        iv4 = pb.GenomicSegment("chrB", 500000 + 10500, 500000 + 11000, "+")
        gnd[iv1] = 1
        assert sum(gnd[iv1]) == 1000
        gnd[iv2] += 1
        assert sum(gnd[iv1]) == 1500 and sum(gnd[iv2]) == 1000 and gnd.lengths()["chrA"] > 1000
        gnd[iv3] += 1
        assert sum(gnd[iv3]) == 500 and sum(gnd[iv4]) == 0


@pytest.mark.parametrize("world_size", [2, 5])
@pytest.mark.parametrize("kind", ["spliced_one_length", "ribo_many_lengths", "ribo_exact"])
def test_position_range_sharding_center_rule(small_world, cuda_device, world_size, kind, monkeypatch):
    """SURVEY 8e for the Center rule (pb_map_center_range / pb_map_center_fixed_range): range-only fp64
    planes of every rank equal the whole-genome planes bit for bit — intervals that start before a range
    and reach into it come through the bin kernel's look-back records and the halo reads — and the
    statistics add up.  Ranks derive their slot / fixed-point tables from the histogram of the whole batch."""
    import torch
    from plastid_b200 import dist as pd
    from plastid_b200.batch import DeviceBatch
    from plastid_b200.genome_array import length_histogram
    w = small_world
    lay = w["layout"]
    if kind == "spliced_one_length":
        hb = synth.device_batch_to_host(synth.rnaseq_reads(w["chroms"], w["lens"], 60_000, seed=4, device="cpu",
                                                           intron=(50, 3000)), w["chroms"], w["lens"])
        fac = pb.CenterMapFactory(12)
    else:
        hb, fac = w["hb"], pb.CenterMapFactory(3 if kind == "ribo_exact" else 0)
        if kind == "ribo_exact":
            monkeypatch.setenv("PB_CENTER_EXACT", "1")
    dfull = DeviceBatch.from_host(hb, cuda_device)
    hist = length_histogram(dfull, fac)
    whole = map_batch(dfull, lay, fac, None, strands=("+", "-"), length_hist=hist)
    assert float(whole.planes["+"].sum().item()) > 0
    cuts = pd.position_cuts(hb, lay, world_size)
    mapped = np.zeros(_lib.PB_NSTATS, dtype=np.int64)
    for rank in range(world_size):
        sub, lo, hi = pd.shard_positions(hb, lay, rank, world_size, cuts)
        if hi == lo:
            continue
        planes = map_batch(DeviceBatch.from_host(sub, cuda_device), lay, fac, None, strands=("+", "-"), bin_range=(lo, hi),
                           length_hist=hist)
        for s in ("+", "-"):
            assert planes.planes[s].numel() == hi - lo
            assert torch.equal(planes.planes[s], whole.planes[s][lo:hi]), (kind, rank, s)
        mapped += planes.stats
    for k in (_lib.PB_STAT_MAPPED_PLUS, _lib.PB_STAT_MAPPED_MINUS, _lib.PB_STAT_DROPPED_PLUS, _lib.PB_STAT_DROPPED_MINUS):
        assert mapped[k] == whole.stats[k]


def test_center_tables_from_batch_metadata_equal_device_histogram(small_world, cuda_device):
    """A batch that carries its length histogram (decoder / receiver metadata) skips the device
    histogram + host round trip of the Center rule; the planes are identical, with and without a size
    filter and with host-evaluated drop bits."""
    import torch
    from plastid_b200.batch import DeviceBatch
    w = small_world
    # the histogram is metadata of the PACKED batch (AlignmentBatch.pack: what the decoder emits); a plain SoA batch
    # carries none and the Center rule measures it on the device
    hb = pb.AlignmentBatch(w["hb"].chroms, w["hb"].chrom_len, w["hb"].ref_start, w["hb"].meta, w["hb"].chrom_read_off,
                           max_span=w["hb"].max_span).pack()
    assert DeviceBatch.from_host(w["hb"], cuda_device).length_hist is None or w["hb"].transfer is not None
    d_meta = DeviceBatch.from_host(hb, cuda_device)
    assert d_meta.length_hist is not None and d_meta.length_hist.sum() == len(hb)
    for sf in (None, pb.SizeFilterFactory(24, 33), pb.SizeFilterFactory(30, -1)):
        for fac in (pb.CenterMapFactory(0), pb.CenterMapFactory(11)):
            with_meta = map_batch(d_meta, w["layout"], fac, sf, strands=("+", "-"))
            d_plain = DeviceBatch.from_host(hb, cuda_device)
            d_plain.length_hist = None
            measured = map_batch(d_plain, w["layout"], fac, sf, strands=("+", "-"))
            for s in ("+", "-"):
                assert torch.equal(with_meta.planes[s], measured.planes[s])
            assert (with_meta.stats == measured.stats).all()


def test_center_streamed_upload_many_chunks_repeated(cuda_device):
    """The same at a size where chunks of both compute lanes are in flight together: 2 M spliced reads, 12
    and 24 chunks, repeated — every repeat must reproduce the whole-batch planes bit for bit."""
    import torch
    from plastid_b200.batch import Delta3SplicedBatch, Delta3SplicedReceiver, DeviceBatch
    from plastid_b200.genome_array import map_center_streamed
    chroms, lens = synth.human_like_genome(0.02)
    lay = pb.GenomeLayout(chroms, lens)
    dbatch = synth.rnaseq_reads(chroms, lens, 2_000_000, seed=21, device=cuda_device)
    hb = synth.device_batch_to_host(dbatch, chroms, lens)
    fac = pb.CenterMapFactory(12)
    whole = map_batch(dbatch, lay, fac, None, strands=("+", "-"))
    wire = Delta3SplicedBatch.from_batch(hb)
    rx = Delta3SplicedReceiver(wire, cuda_device)
    pinned = wire.pinned()
    planes = None
    for n_chunks in (12, 24):
        chunks = Delta3SplicedReceiver.plan_chunks(wire, lay, n_chunks)
        assert len(chunks) == n_chunks
        for rep in range(4):
            planes = map_center_streamed(rx, pinned, chunks, lay, fac, None, ("+", "-"), planes)
            torch.cuda.synchronize()
            for s in ("+", "-"):
                assert torch.equal(planes.planes[s], whole.planes[s]), (n_chunks, rep, s)
            assert (planes.stats_dev.cpu().numpy()[:7] == whole.stats[:7]).all()


@pytest.mark.parametrize("n_chunks", [1, 3, 7])
def test_center_streamed_upload_matches_whole_batch(small_world, cuda_device, n_chunks):
    """map_center_streamed: a spliced batch uploaded chunk by chunk (delta3 + block words), every chunk's
    final bin range produced with pb_map_center_range from a read window — planes, statistics and the
    rebuilt block table equal the whole-batch path bit for bit."""
    import torch
    from plastid_b200.batch import Delta3SplicedBatch, Delta3SplicedReceiver, DeviceBatch
    from plastid_b200.genome_array import map_center_streamed
    w = small_world
    lay = w["layout"]
    hb = synth.device_batch_to_host(synth.rnaseq_reads(w["chroms"], w["lens"], 80_000, seed=6, device="cpu",
                                                       intron=(50, 4000)), w["chroms"], w["lens"])
    fac = pb.CenterMapFactory(12)
    whole = map_batch(DeviceBatch.from_host(hb, cuda_device), lay, fac, None, strands=("+", "-"))
    wire = Delta3SplicedBatch.from_batch(hb)
    rx = Delta3SplicedReceiver(wire, cuda_device)
    pinned = wire.pinned()
    chunks = Delta3SplicedReceiver.plan_chunks(wire, lay, n_chunks)
    assert len(chunks) >= min(n_chunks, 2) - 1 and chunks[0][0] == 0 and chunks[-1][1] == len(hb) and chunks[-1][3] == lay.total_bins
    for _ in range(2):                                              # buffers and planes are reusable
        planes = map_center_streamed(rx, pinned, chunks, lay, fac, None, ("+", "-"))
        torch.cuda.synchronize()
        for s in ("+", "-"):
            assert torch.equal(planes.planes[s], whole.planes[s]), (n_chunks, s)
        assert (planes.stats_dev.cpu().numpy()[:7] == whole.stats[:7]).all()
        assert (rx.batch.blk_off.cpu().numpy().view(np.uint32) == hb.blk_off).all()
        assert (rx.batch.blk.cpu().numpy() == hb.blk).all()
