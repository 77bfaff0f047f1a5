"""GPU parity of the script-level counting loops (counts_in_region, cs count, metagene count, psite,
phase_by_size) against the oracle's statement-by-statement restatements of the reference scripts."""
import warnings

import numpy as np
import pytest

import plastid_b200 as pb
from plastid_b200 import synth
from plastid_b200.bin import counts_in_region, cs, metagene, psite, phase_by_size
from oracle import pyoracle as po
from oracle import scripts as osc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(cuda_device):
    chroms, lens = synth.yeast_like_genome(total=700_000, n_chrom=4)
    ann = synth.make_annotation(chroms, lens, 120, seed=8, exons=(1, 3), exon_len=(150, 500), intron_len=(40, 300))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, 60_000, seed=4, device="cpu", lengths=range(22, 38)),
                                    chroms, lens)
    reads = {c: [] for c in chroms}
    for i in range(len(hb)):
        c = int(np.searchsorted(hb.chrom_read_off, i, side="right")) - 1
        reads[chroms[c]].append(po.Read(int(hb.ref_start[i]), [(0, int(hb.meta[i] & 0xFFFF))], bool((hb.meta[i] >> 16) & 1)))
    store = po.ReadStore(dict(zip(chroms, [int(x) for x in lens])), reads)
    hb2 = pb.pack_reads(reads, dict(zip(chroms, [int(x) for x in lens])))
    return dict(chroms=chroms, lens=lens, ann=ann, store=store, hb=hb2, dev=cuda_device)


def make_gas(w, ofn, gfn, size=(25, 100)):
    oga = po.OracleBAMGenomeArray(w["store"], mapping=ofn)
    ga = pb.BAMGenomeArray(w["hb"], mapping=gfn, device=w["dev"])
    if size is not None:
        oga.add_filter("size", po.SizeFilter(*size))
        ga.add_filter("size", pb.SizeFilterFactory(*size))
    return oga, ga


def test_counts_in_region(world):
    w = world
    oga, ga = make_gas(w, po.FivePrimeMap(14), pb.FivePrimeMapFactory(14))
    chains = w["ann"].chains()
    masks = synth.make_masks(w["ann"], frac=0.25, seed=3)
    masks[5] = [pb.GenomicSegment(chains[5].chrom, chains[5].spanning_segment.start - 10,
                                  chains[5].spanning_segment.end + 10, chains[5].strand)]      # fully masked
    chains.append(pb.SegmentChain(pb.GenomicSegment("chrUnknown", 10, 500, "+"), ID="nowhere"))
    masks.append([])
    ochains = []
    for ch in chains:
        oc = po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in ch])
        oc.name = ch.get_name()
        ochains.append(oc)
    omasks = [[po.Seg(m.chrom, m.start, m.end, m.strand) for m in ms] for ms in masks]
    exp = osc.counts_in_region_rows(oga, ochains, omasks)
    ga_sum, got = counts_in_region.count_regions(ga, chains, masks)
    assert ga_sum == oga.sum()
    assert got == exp
    assert got[5][2] == "nan" and got[5][5] == "0"


def _mask_features(w, rng, n=150):
    """Mask annotation as a crossmap would give it: one- and two-segment features on both strands,
    some inside transcripts, some spanning exon boundaries, some far from any region, a few duplicated."""
    chains = w["ann"].chains()
    feats = []
    for k in range(n):
        ch = chains[int(rng.integers(len(chains)))]
        span = ch.spanning_segment
        a = int(rng.integers(max(span.start - 300, 0), span.end + 100))
        segs = [pb.GenomicSegment(ch.chrom, a, a + int(rng.integers(1, 400)), ch.strand if k % 5 else "+-"[k % 2])]
        if k % 3 == 0:
            b = segs[0].end + int(rng.integers(1, 2000))
            segs.append(pb.GenomicSegment(ch.chrom, b, b + int(rng.integers(1, 300)), segs[0].strand))
        feats.append(pb.SegmentChain(*segs, ID="mask%d" % k))
    feats.append(pb.SegmentChain(pb.GenomicSegment("chrUnknown", 0, 1000, "+"), ID="elsewhere"))
    return feats + feats[:7]


def test_counts_in_region_with_device_mask_pipeline(world):
    """Mask annotation -> pb_mask_chains == per-region GenomeHash.get_overlapping_features + add_masks."""
    w = world
    oga, ga = make_gas(w, po.FivePrimeMap(14), pb.FivePrimeMapFactory(14))
    chains = w["ann"].chains()
    chains.append(pb.SegmentChain(pb.GenomicSegment("chrUnknown", 10, 500, "+"), ID="nowhere"))
    chains.append(pb.SegmentChain(pb.GenomicSegment(chains[0].chrom, 10, 500, "."), ID="unstranded"))
    feats = _mask_features(w, np.random.default_rng(11))
    # one region also carries a mask of its own from add_masks: both must apply
    own = pb.GenomicSegment(chains[3].chrom, chains[3].spanning_segment.start, chains[3].spanning_segment.start + 40, chains[3].strand)
    chains[3].add_masks(own)
    ochains = []
    for ch in chains:
        oc = po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in ch])
        oc.name = ch.get_name()
        ochains.append(oc)
    ochains[3].add_masks(po.Seg(own.chrom, own.start, own.end, own.strand))
    ofeats = [po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in f]) for f in feats]
    exp = osc.counts_in_region_rows(oga, ochains, crossmap=po.GenomeHash(ofeats))
    ga_sum, got = counts_in_region.count_regions(ga, chains, mask_features=feats)
    assert got == exp
    assert sum(int(r[5]) for r in got) < sum(ch.length for ch in chains)       # something was masked
    # the host statement of the query (per-chain add_masks) gives the same table
    chains2 = w["ann"].chains() + chains[-2:]
    chains2[3].add_masks(own)
    _s, got2 = counts_in_region.count_regions(ga, chains2, masks=counts_in_region.overlapping_masks(chains2, feats))
    assert got2 == exp
    with pytest.raises(KeyError):
        counts_in_region.count_regions(ga, chains[:2], mask_features=[pb.SegmentChain(pb.GenomicSegment(chains[0].chrom, 1, 9, "."))])


def test_cs_count(world):
    w = world
    oga, ga = make_gas(w, po.ThreePrimeMap(0), pb.ThreePrimeMapFactory(0))
    chains = w["ann"].chains()
    pos = {"region": [], "exon": [], "utr5": [], "cds": [], "utr3": []}
    for ch in chains[:60]:
        n = ch.length
        a, b = n // 5, n - n // 4
        pos["region"].append(ch.get_name())
        pos["exon"].append(str(ch))
        pos["utr5"].append(str(ch.get_subchain(0, a)))
        pos["cds"].append(str(ch.get_subchain(a, b)))
        pos["utr3"].append(str(ch.get_subchain(b, n)) if len(pos["region"]) % 7 else "na")
    exp = osc.cs_count(oga, pos)
    order, got = cs.do_count(ga, pos)
    assert order[0] == "region" and len(order) == 13
    for col in order:
        for e, g in zip(exp[col], got[col]):
            if isinstance(e, float) and np.isnan(e):
                assert np.isnan(g), col
            elif col.endswith("_rpkm"):
                assert g == pytest.approx(e, rel=1e-14), col
            else:
                assert g == e, col


def roi_rows(w, flank=50, down=100, mask_every=4):
    rows = {"region": [], "masked": [], "alignment_offset": [], "window_size": [], "zero_point": []}
    for i, ch in enumerate(w["ann"].chains()):
        start = min(ch.length // 3, 200)
        lo, hi = max(0, start - flank), min(ch.length, start + down)
        win = ch.get_subchain(lo, hi)
        rows["region"].append(str(win))
        if i % mask_every == 0:
            g = win.get_genomic_coordinate(min(60, win.length - 1))[1]
            rows["masked"].append("%s:%d-%d(%s)" % (win.chrom, g, g + 12, win.strand))
        else:
            rows["masked"].append("na")
        rows["alignment_offset"].append(flank - (start - lo))
        rows["window_size"].append(flank + down)
        rows["zero_point"].append(flank)
    return rows


def as_oracle_rows(rows):
    return [dict(region=r, masked=m, alignment_offset=o)
            for r, m, o in zip(rows["region"], rows["masked"], rows["alignment_offset"])]


@pytest.mark.parametrize("use_mean", [False, True])
def test_metagene_count(world, use_mean):
    w = world
    oga, ga = make_gas(w, po.FivePrimeMap(12), pb.FivePrimeMapFactory(12))
    rows = roi_rows(w)
    counts, norm, profile, num_genes, row_select = osc.metagene_count(oga, as_oracle_rows(rows), 150, 70, 100, 3, use_mean)
    out = metagene.do_count(ga, rows, 70, 100, 3, use_mean, keep=True)
    assert (np.ma.getmaskarray(out["counts"]) == np.ma.getmaskarray(counts)).all()
    assert (out["counts"].compressed() == counts.compressed()).all()
    rs = np.ma.filled(row_select, False).astype(bool)
    assert (out["row_select"] == rs).all() and rs.sum() > 10
    assert (np.ma.getmaskarray(out["norm_counts"])[rs] == np.ma.getmaskarray(norm)[rs]).all()
    np.testing.assert_allclose(out["norm_counts"][rs].compressed(), norm[rs].compressed(), rtol=1e-15)
    np.testing.assert_allclose(out["metagene_average"], np.ma.filled(profile, np.nan), rtol=1e-12, equal_nan=True)
    assert (out["regions_counted"] == num_genes).all()
    assert (out["x"] == np.arange(-50, 100)).all()


@pytest.mark.parametrize("aggregate", [False, True])
def test_psite(world, aggregate):
    w = world
    oga, ga = make_gas(w, po.FivePrimeMap(0), pb.FivePrimeMapFactory(0), size=None)     # psite.py:357-383
    rows = roi_rows(w)
    raw, profiles, regions = osc.psite_count(oga, as_oracle_rows(rows), 150, 70, 100, 2, 26, 33, aggregate)
    out = psite.do_count(ga, rows, 70, 100, 2, 26, 33, aggregate, keep=True)
    for k in range(26, 34):
        assert (np.ma.getmaskarray(out["raw"][k]) == np.ma.getmaskarray(raw[k])).all(), k
        assert (out["raw"][k].compressed() == raw[k].compressed()).all(), k
        np.testing.assert_allclose(out["profiles"][k], np.ma.filled(profiles[k], np.nan), rtol=1e-12, equal_nan=True)
        assert (out["regions_counted"][k] == regions[k]).all(), k
    if not aggregate:
        # the default call (no raw matrices kept) runs the fused normalise+keys kernel: same numbers
        fused = psite.do_count(ga, rows, 70, 100, 2, 26, 33, aggregate, keep=False)
        for k in range(26, 34):
            np.testing.assert_allclose(fused["profiles"][k], np.ma.filled(profiles[k], np.nan), rtol=1e-12, equal_nan=True)
            assert (fused["profiles"][k] == out["profiles"][k]).all() or np.isnan(out["profiles"][k]).any()
            assert (fused["regions_counted"][k] == regions[k]).all(), k
        # a threshold nothing reaches: no row is selected, numpy.ma.median of the empty selection is all masked -> nan
        none = psite.do_count(ga, rows, 70, 100, 10**9, 26, 33, aggregate, keep=False)
        ref0 = psite.do_count(ga, rows, 70, 100, 10**9, 26, 33, aggregate, keep=True)
        for k in range(26, 34):
            assert np.array_equal(none["profiles"][k], ref0["profiles"][k], equal_nan=True) and (none["regions_counted"][k] == ref0["regions_counted"][k]).all()
    x = np.arange(-50, 100)
    for kw in (dict(), dict(require_upstream=True), dict(constrain=(5, 25))):
        assert psite.pick_offsets(x, out["profiles"], 13, **kw) == osc.psite_pick_offsets(x, profiles, 13, **kw)


@pytest.mark.parametrize("back", [-1, -5, 0])
def test_phase_by_size(world, back):
    w = world
    offs = dict(synth.RIBO_OFFSETS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oga, ga = make_gas(w, po.VariableFivePrimeMap(offs), pb.VariableFivePrimeMapFactory(offs))
    chains = w["ann"].chains()
    cds = [ch.get_subchain(ch.length // 6, ch.length - ch.length // 8) for ch in chains[:80]]
    ocds = [po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in ch]) for ch in cds]
    lengths = list(range(24, 36))
    exp = osc.phase_by_size(oga, ocds, lengths, 5, back)
    got = phase_by_size.do_phase(ga, cds, lengths, 5, back)
    for k in lengths:
        assert (got[k] == exp[k]).all(), k
    if back != 0:
        assert sum(v.sum() for v in got.values()) > 1000
        assert got[24].sum() == 0          # below the 25-100 size filter


def test_phase_sums_on_planes_agree_with_single_launch_path(world):
    """pb_phase_sums over per-length count planes (one map pass per length) == pb_stratified_windows."""
    from plastid_b200.genome_array import map_batch, phase_sums
    from plastid_b200.regions import ChainTable
    w = world
    oga, ga = make_gas(w, po.FivePrimeMap(3), pb.FivePrimeMapFactory(3))
    cds = [ch.get_subchain(ch.length // 6, ch.length - ch.length // 8) for ch in w["ann"].chains()[:80]]
    got = phase_by_size.do_phase(ga, cds, [27, 28, 31], 5, -1)
    table = ChainTable.from_chains(cds, ga.layout, use_masks=False)
    for k in (27, 28, 31):
        planes = map_batch(ga._device_batch(), ga.layout, ga.map_fn, pb.SizeFilterFactory(k, k), strands=("+", "-"))
        exp = phase_sums(planes, table, 5, -1).sum(dim=0).cpu().numpy()
        assert (got[k] == exp).all() and exp.sum() > 0


@pytest.mark.parametrize("window", [100000, 1000, 7])
def test_track_export_matches_reference_loops(world, window):
    """to_variable_step / to_bedgraph (genome_array.py:990-1111): identical files, incl. runs cut at
    window boundaries, both strands and the merged '.' strand, raw and normalised."""
    import io
    w = world
    oga, ga = make_gas(w, po.FivePrimeMap(14), pb.FivePrimeMapFactory(14))
    for strand in ("+", "-", "."):
        for norm in (False, True):
            oga.set_normalize(norm)
            ga.set_normalize(norm)
            a, b = io.StringIO(), io.StringIO()
            oga.to_bedgraph(a, "trk", strand, window_size=window, color="0,0,255", alwaysZero="on")
            ga.to_bedgraph(b, "trk", strand, window_size=window, color="0,0,255", alwaysZero="on")
            assert a.getvalue() == b.getvalue() and a.getvalue().count("\n") > 100
            a, b = io.StringIO(), io.StringIO()
            oga.to_variable_step(a, "trk", strand, window_size=window)
            ga.to_variable_step(b, "trk", strand, window_size=window)
            assert a.getvalue() == b.getvalue()
    oga.set_normalize(False)
    ga.set_normalize(False)


def test_track_export_center_rule_within_tolerance(world):
    import io
    w = world
    oga, ga = make_gas(w, po.CenterMap(12), pb.CenterMapFactory(12))
    a, b = io.StringIO(), io.StringIO()
    oga.to_variable_step(a, "trk", "+", window_size=5000)
    ga.to_variable_step(b, "trk", "+", window_size=5000)
    la, lb = a.getvalue().splitlines(), b.getvalue().splitlines()
    assert len(la) == len(lb) > 100
    for x, y in zip(la, lb):
        if "\t" not in x:
            assert x == y
            continue
        (px, vx), (py, vy) = x.split("\t"), y.split("\t")
        assert px == py and abs(float(vx) - float(vy)) <= 1e-6 * abs(float(vx))      # north-star tolerance
    # bedGraph of fractional coverage: same run boundaries wherever the values are exactly representable
    a, b = io.StringIO(), io.StringIO()
    oga.to_bedgraph(a, "trk", "-", window_size=100000)
    ga.to_bedgraph(b, "trk", "-", window_size=100000)
    ra = [l.split("\t") for l in a.getvalue().splitlines()[1:]]
    rb = [l.split("\t") for l in b.getvalue().splitlines()[1:]]
    cov = lambda rows: sum((int(r[2]) - int(r[1])) * float(r[3]) for r in rows)
    assert abs(cov(ra) - cov(rb)) <= 1e-6 * cov(ra) and rb[0][:2] == ra[0][:2]


def test_make_wiggle_and_counts_in_region_command_lines(world, tmp_path):
    """The script entry points end to end: batch file -> tracks / region table on disk, against the
    oracle's restatements of the reference programs (make_wiggle.py:172-209, counts_in_region.py:107-125)."""
    import io
    from plastid_b200.bin import _cli, make_wiggle
    w = world
    batch = str(tmp_path / "reads.npz")
    _cli.save_batch(batch, w["hb"])
    out = str(tmp_path / "trk")
    make_wiggle.main(["--count_files", batch, "--fiveprime", "--offset", "14", "--device", w["dev"], "-o", out,
                      "--color", "#0000FF", "--window_size", "5000"])
    oga, _ga = make_gas(w, po.FivePrimeMap(14), pb.FivePrimeMapFactory(14))
    for suffix, strand in (("fw", "+"), ("rc", "-")):
        exp = io.StringIO()
        oga.to_bedgraph(exp, "%s_%s" % (out, suffix), strand, window_size=5000, color="0,0,255")
        assert open("%s_%s.wig" % (out, suffix)).read() == exp.getvalue()
    make_wiggle.main(["--count_files", batch, "--threeprime", "--device", w["dev"], "-o", out, "-t", "mytrack",
                      "--output_format", "variable_step"])
    oga3, _ = make_gas(w, po.ThreePrimeMap(0), pb.ThreePrimeMapFactory(0))
    exp = io.StringIO()
    oga3.to_variable_step(exp, "mytrack_rc", "-", color="0,0,0")
    assert open(out + "_rc.wig").read() == exp.getvalue()
    # counts_in_region from BED files, with a mask annotation
    chains = w["ann"].chains()[:40]
    bed, mbed, table = str(tmp_path / "a.bed"), str(tmp_path / "m.bed"), str(tmp_path / "out.txt")
    with open(bed, "w") as fh:
        for ch in chains:
            sp = ch.spanning_segment
            sizes = ",".join(str(len(s)) for s in ch)
            starts = ",".join(str(s.start - sp.start) for s in ch)
            fh.write("\t".join([ch.chrom, str(sp.start), str(sp.end), ch.get_name(), "0", ch.strand, str(sp.start), str(sp.end),
                                 "0,0,0", str(len(ch)), sizes, starts]) + "\n")
    feats = _mask_features(w, np.random.default_rng(4), n=60)[:-8]
    with open(mbed, "w") as fh:
        for f in feats:
            for s in f:
                fh.write("\t".join([s.chrom, str(s.start), str(s.end), f.get_name(), "0", s.strand]) + "\n")
    counts_in_region.main(["--annotation_files", bed, "--mask_annotation_files", mbed, "--count_files", batch,
                           "--fiveprime", "--offset", "14", "--device", w["dev"], table])
    ochains = []
    for ch in chains:
        oc = po.Chain(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in ch])
        oc.name = ch.get_name()
        ochains.append(oc)
    ofeats = [po.Chain(po.Seg(s.chrom, s.start, s.end, s.strand)) for f in feats for s in f]
    exp_rows = osc.counts_in_region_rows(oga, ochains, crossmap=po.GenomeHash(ofeats))
    lines = open(table).read().splitlines()
    assert lines[0] == "## total_dataset_counts: %s" % oga.sum()
    assert [l.split("\t") for l in lines[2:]] == exp_rows


# ---------------------------------------------------------------------------------------------
# metagene generate: pb_landmark_windows + pb_spanning_windows + pb_mask_chains
# ---------------------------------------------------------------------------------------------
def _transcripts(records, product=True):
    from oracle import generate as og
    out = []
    for name, r in records.items():
        kw = dict(ID=name, gene_id=r["gene_id"], cds_genome_start=r["cds_genome_start"], cds_genome_end=r["cds_genome_end"])
        if product:
            out.append(pb.Transcript(*[pb.GenomicSegment(r["chrom"], s, e, r["strand"]) for s, e in r["segments"]], **kw))
        else:
            out.append(og.Tx(*[po.Seg(r["chrom"], s, e, r["strand"]) for s, e in r["segments"]], **kw))
    return out


def _check_reference_rows(rows, result_groups, up):
    result_groups = sorted(result_groups, key=lambda x: x[0])
    rows = sorted(rows, key=lambda r: r["region"])
    c = 0
    for n, group in enumerate(result_groups):
        if group[1] is None or group[2] is None:
            c += 1
            continue
        row = rows[n - c]
        assert str(pb.SegmentChain.from_str(group[0])) == row["region"]
        assert group[1] == row["alignment_offset"] and group[2] == row["zero_point"] == up
        if len(group) == 4:
            assert group[3] == row["masked"]
    assert len(result_groups) - c == len(rows)


@pytest.mark.parametrize("masked", [False, True])
def test_metagene_generate_reproduces_the_reference_tables(cuda_device, masked):
    """plastid/test/unit/bin/test_metagene.py:292-326 through the device path, incl. a custom window
    function (evaluated per region on the host, solved on the device)."""
    from helpers import metagene_generate_golden, gff3_transcript_records
    gold = metagene_generate_golden()
    txs = {t.get_name(): t for t in _transcripts(gff3_transcript_records(gold["transcripts_gff"]))}
    mask_hash = pb.GenomeHash([pb.SegmentChain.from_str(m) for m in gold["masks"]] if masked else [])
    results = gold["do_generate_max_window_results_masked" if masked else "do_generate_max_window_results"]
    custom = lambda region, up, down: metagene.window_cds_start(region, up, down)       # noqa: E731
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for up, down in gold["flanks"]:
            for name, group in gold["do_generate_max_window"].items():
                for func in (metagene.window_cds_start, custom):
                    df = metagene.group_regions_make_windows([txs[t] for t in group], mask_hash, up, down, func,
                                                             device=cuda_device)
                    _check_reference_rows(df.to_dict("records"), [results["%s_%s_%s" % (name, up, down)]], up)
        for name, group in gold["do_generate_multi_gene"].items():
            df = metagene.group_regions_make_windows([txs[t] for t in group], mask_hash, 50, 100,
                                                     metagene.window_cds_start, device=cuda_device)
            _check_reference_rows(df.to_dict("records"), gold["do_generate_multi_gene_results"]["%s_50_100" % name], 50)
        # single-gene entry point
        roi, offset = metagene.maximal_spanning_window([txs[t] for t in gold["do_generate_max_window"]["3_same_start_plus"]],
                                                       mask_hash, 50, 100, device=cuda_device)
        assert str(roi) == "2L:7985664-7985768^7985833-7985839(+)" and offset == 40
        assert roi.attr["thickstart"] == 7985674 and roi.masked_length == roi.length - (50 if masked else 0)
        roi, offset = metagene.maximal_spanning_window([txs[t] for t in gold["do_generate_max_window"]["3_diff_start_plus"]],
                                                       mask_hash, 50, 100, device=cuda_device)
        assert len(roi) == 0 and np.isnan(offset)


@pytest.mark.parametrize("landmark", ["cds_start", "cds_stop"])
def test_metagene_generate_random_gene_models_match_oracle(cuda_device, landmark):
    from helpers import random_gene_models
    from oracle import generate as og
    rng = np.random.default_rng(11)
    recs = random_gene_models(rng, 400)
    txs, otxs = _transcripts(recs), _transcripts(recs, product=False)
    masks = []
    for r in list(recs.values())[::3]:                               # masks placed on the genes themselves
        if r["cds_genome_start"] is not None:
            edge = r["cds_genome_start"] if (landmark == "cds_start") == (r["strand"] == "+") else r["cds_genome_end"]
            s = max(edge + int(rng.integers(-60, 60)), 0)
            masks.append((r["chrom"], s, s + int(rng.integers(1, 40)), r["strand"] if rng.random() < 0.8 else "+"))
    mh = pb.GenomeHash([pb.SegmentChain(pb.GenomicSegment(*m)) for m in masks])
    omh = po.GenomeHash([po.Chain(po.Seg(*m)) for m in masks])
    ofunc = og.window_cds_start if landmark == "cds_start" else og.window_cds_stop
    n_masked = n_spliced = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for up, down in ((50, 100), (0, 30), (100, 0), (20, 400), (3, 3), (33, 31)):
            df = metagene.do_generate(txs, mh, landmark, up, down, device=cuda_device)
            exp = og.group_regions_make_windows(otxs, omh, up, down, ofunc)
            got = df.to_dict("records")
            assert len(got) == len(exp) and len(exp) > 200
            for a, b in zip(got, exp):
                for k in ("region_id", "region", "masked", "alignment_offset", "zero_point", "region_length",
                          "threeprime_offset", "window_size"):
                    assert a[k] == b[k], (k, a, b)
                n_masked += a["masked"] != "na"
                n_spliced += "^" in a["region"]
    assert n_masked > 50 and n_spliced > 200


def test_metagene_generate_then_count(world, tmp_path):
    """generate -> ROI file -> count, as the two sub-programs are chained (metagene.py:1269-1330)."""
    w = world
    ann = w["ann"]
    txs = []
    for t, ch in enumerate(ann.chains()):
        a = ch.get_genomic_coordinate(ch.length // 5)[1]
        b = ch.get_genomic_coordinate(ch.length - ch.length // 5)[1]
        lo, hi = (a, b + 1) if ch.strand == "+" else (b, a + 1)
        txs.append(pb.Transcript(*ch.segments, ID=ch.get_name(), gene_id="gene%d" % t, cds_genome_start=lo, cds_genome_end=hi))
    bed = tmp_path / "tx.bed"
    with open(bed, "w") as fh:
        for tx in txs:
            line = tx.as_bed(thickstart=tx.cds_genome_start, thickend=tx.cds_genome_end).rstrip("\n")
            fh.write(line + "\t" + tx.attr["gene_id"] + "\n")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        metagene.main(["generate", "--annotation_files", str(bed), "--upstream", "30", "--downstream", "120",
                       str(tmp_path / "mg")])
    from plastid_b200.bin import _cli
    roi = _cli.read_pl_table(str(tmp_path / "mg_rois.txt"))
    assert len(roi["region"]) == len(txs) and set(roi["zero_point"]) == {"30"}
    from oracle import generate as og
    otxs = [og.Tx(*[po.Seg(s.chrom, s.start, s.end, s.strand) for s in tx], ID=tx.get_name(), gene_id=tx.attr["gene_id"],
                  cds_genome_start=tx.cds_genome_start, cds_genome_end=tx.cds_genome_end) for tx in txs]
    exp = og.group_regions_make_windows(otxs, po.GenomeHash([]), 30, 120, og.window_cds_start)
    assert [r["region"] for r in exp] == roi["region"]
    assert [str(r["alignment_offset"]) for r in exp] == roi["alignment_offset"]
    oga, ga = make_gas(w, po.FivePrimeMap(12), pb.FivePrimeMapFactory(12))
    out = metagene.do_count(ga, roi, 40, 90, min_counts=5)
    counts, norm, profile, num_genes, row_select = osc.metagene_count(oga, exp, 150, 40, 90, 5)
    assert np.ma.filled(row_select, False).sum() > 10
    np.testing.assert_allclose(out["metagene_average"], np.ma.filled(profile, np.nan), rtol=1e-12, equal_nan=True)
    assert (out["regions_counted"] == num_genes).all()


# ---------------------------------------------------------------------------------------------
# cs generate: pb_chain_union / pb_chain_binary
# ---------------------------------------------------------------------------------------------
def _random_chains(rng, n, span=4000):
    chains = []
    for _ in range(n):
        k = int(rng.integers(0, 9))
        cuts = np.sort(rng.choice(span, size=2 * k, replace=False)) if k else np.zeros(0, dtype=np.int64)
        blocks = [(int(a), int(b)) for a, b in zip(cuts[0::2], cuts[1::2]) if b > a]
        merged = []
        for a, b in blocks:                                          # normalise: non-touching
            if merged and a <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(b, merged[-1][1]))
            else:
                merged.append((a, b))
        chains.append(merged)
    return chains


def _to_set(device, chains):
    from plastid_b200.chains import ChainSet
    off = np.zeros(len(chains) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in chains], out=off[1:])
    flat = [b for c in chains for b in c]
    return ChainSet.from_numpy([a for a, _ in flat], [b for _, b in flat], off, device)


def _positions(chain):
    return set(p for a, b in chain for p in range(a, b))


def _runs(positions):
    out = []
    for p in sorted(positions):
        if out and out[-1][1] == p:
            out[-1][1] = p + 1
        else:
            out.append([p, p + 1])
    return [tuple(x) for x in out]


def test_chain_set_algebra_matches_python_sets(cuda_device):
    from plastid_b200.chains import chain_union, chain_binary
    rng = np.random.default_rng(31)
    a_chains, b_chains = _random_chains(rng, 300), _random_chains(rng, 200)
    A, B = _to_set(cuda_device, a_chains), _to_set(cuda_device, b_chains)
    a_idx = rng.integers(0, 300, 500)
    b_idx = rng.integers(-1, 200, 500)
    for op in ("and", "sub"):
        bs, be, off = chain_binary(op, A, a_idx, B, b_idx).numpy()
        for i in range(500):
            pa = _positions(a_chains[a_idx[i]])
            pb_ = _positions(b_chains[b_idx[i]]) if b_idx[i] >= 0 else set()
            exp = _runs(pa & pb_ if op == "and" else pa - pb_)
            assert list(zip(bs[off[i]:off[i + 1]], be[off[i]:off[i + 1]])) == exp, (op, i)
    sizes = rng.integers(0, 70, 150)                                  # groups of 0..69 members (more than a warp)
    grp_off = np.zeros(151, dtype=np.int64)
    np.cumsum(sizes, out=grp_off[1:])
    members = rng.integers(0, 300, int(grp_off[-1]))
    bs, be, off = chain_union(A, grp_off, members).numpy()
    for g in range(150):
        exp = _runs(set().union(*[_positions(a_chains[m]) for m in members[grp_off[g]:grp_off[g + 1]]]))
        assert list(zip(bs[off[g]:off[g + 1]], be[off[g]:off[g + 1]])) == exp, g
    # the reference's own merge_segments known answers (test_roitools.py:356-398): union of one-block chains
    from helpers import merge_segments_known_answers
    kat = merge_segments_known_answers()
    singles = [[seg] for segs, _ in kat for seg in segs]
    S = _to_set(cuda_device, singles)
    k_off = np.zeros(len(kat) + 1, dtype=np.int64)
    np.cumsum([len(segs) for segs, _ in kat], out=k_off[1:])
    bs, be, off = chain_union(S, k_off, np.arange(len(singles), dtype=np.int64)).numpy()
    for g, (_segs, expected) in enumerate(kat):
        assert list(zip(bs[off[g]:off[g + 1]], be[off[g]:off[g + 1]])) == expected, g
    # touching blocks of different members merge; empty inputs give empty chains
    T = _to_set(cuda_device, [[(0, 10)], [(10, 20)], [], [(25, 30)]])
    bs, be, off = chain_union(T, [0, 4, 4, 5], [0, 1, 2, 3, 2]).numpy()
    assert list(zip(bs, be)) == [(0, 20), (25, 30)] and list(off) == [0, 2, 2, 2]


@pytest.mark.parametrize("seed,spacing", [(3, 1500), (4, 700), (5, 400)])
def test_cs_generate_matches_oracle(cuda_device, seed, spacing):
    from helpers import random_gene_models, add_shared_exon_genes
    from oracle import generate as og
    rng = np.random.default_rng(seed)
    recs = add_shared_exon_genes(random_gene_models(rng, 150, spacing=spacing), rng)
    txs = {t.get_name(): t for t in _transcripts(recs)}
    otxs = {t.get_name(): t for t in _transcripts(recs, product=False)}
    masks = []
    for i in range(80):
        s = int(rng.integers(0, 75 * spacing))
        masks.append(("chrA" if i % 2 else "chrB", s, s + int(rng.integers(5, 300)), "+-"[i % 3 == 0]))
    mh = pb.GenomeHash([pb.SegmentChain(pb.GenomicSegment(*m)) for m in masks])
    omh = po.GenomeHash([po.Chain(po.Seg(*m)) for m in masks])
    gene_df, tx_df, merged = cs.process_partial_group(txs, mh, device=cuda_device)
    exp_genes, exp_txs, exp_merged = og.cs_process_partial_group(otxs, omh)
    assert merged == exp_merged and len(set(merged.values())) < len(merged)
    got_genes, got_txs = gene_df.to_dict("records"), tx_df.to_dict("records")
    assert len(got_genes) == len(exp_genes) and len(got_txs) == len(exp_txs)
    for got, exp in ((got_genes, exp_genes), (got_txs, exp_txs)):
        for a, b in zip(got, exp):
            for k in b:
                assert a[k] == b[k], (k, a, b)
    assert sum(r["masked"] != "na" for r in got_genes) > 20
    twins = {r["region"]: r for r in got_genes if r["region"].startswith("twin")}
    assert twins["twinA"]["masked"] == twins["twinB"]["masked"] == "chrA:10000015-10000020(+)"   # by twinC only


def test_cs_generate_hand_case_and_command_line(world, tmp_path):
    from helpers import cs_generate_hand_case
    records, masks, genes, transcripts = cs_generate_hand_case()
    txs = _transcripts(records)
    gene_df, tx_df, merged = cs.do_generate(txs, pb.GenomeHash([pb.SegmentChain(pb.GenomicSegment(*m)) for m in masks]),
                                            device=world["dev"])
    assert {r["region"]: {k: r[k] for k in genes[r["region"]]} for r in gene_df.to_dict("records")} == genes
    assert {r["region"]: {k: r[k] for k in transcripts[r["region"]]} for r in tx_df.to_dict("records")} == transcripts
    assert gene_df["exon_bed"].iloc[0] == "c\t100\t380\tA\t0\t+\t100\t100\t0,0,0\t3\t20,70,80,\t0,30,200,\n"
    # generate -> positions file -> count on the synthetic world's annotation
    w = world
    bed = tmp_path / "tx.bed"
    with open(bed, "w") as fh:
        for t, ch in enumerate(w["ann"].chains()):
            a = ch.get_genomic_coordinate(ch.length // 5)[1]
            b = ch.get_genomic_coordinate(ch.length - ch.length // 5)[1]
            lo, hi = (a, b + 1) if ch.strand == "+" else (b, a + 1)
            fh.write(ch.as_bed(thickstart=lo, thickend=hi).rstrip("\n") + "\tgene%d\n" % t)
    cs.main(["generate", str(tmp_path / "cs"), "--annotation_files", str(bed)])
    from plastid_b200.bin import _cli
    pos = _cli.read_pl_table(str(tmp_path / "cs_gene.positions"))
    assert len(pos["region"]) == len(w["ann"].chains())
    assert len(open(str(tmp_path / "cs_merged.txt")).read().splitlines()) == len(pos["region"])
    oga, ga = make_gas(w, po.ThreePrimeMap(0), pb.ThreePrimeMapFactory(0))
    order, cols = cs.do_count(ga, pos)
    ref = osc.cs_count(oga, pos)
    for k in ("exon_reads", "cds_reads", "utr5_length", "utr3_rpkm"):
        np.testing.assert_allclose(np.asarray(cols[k], dtype=float), np.asarray(ref[k], dtype=float), rtol=1e-12, equal_nan=True)
    assert sum(cols["cds_reads"]) > 0
