"""Pins the oracles against the reference's own known-answer tests (no GPU).

Transcribed from plastid/test/unit/genomics/test_map_factories.py:17-200 and
plastid/test/unit/util/scriptlib/test_argparsers.py:69-130; the C oracle is cross-checked against
the Python restatement on random CIGARs, and the CIGAR consume table against the table printed from
the reference's vendored htslib header (tests/golden/cigar_consume_table.txt, made by
`make -C oracle ref`)."""
import io
import os
import warnings

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import coracle
from plastid_b200.batch import pack_reads, cigar_to_blocks
from helpers import kat_reads, kat_expected, random_cigar_reads

SEG = {s: po.Seg("mock", 0, 2000, s) for s in ("+", "-")}
FACTORY = {"fiveprime": po.FivePrimeMap, "threeprime": po.ThreePrimeMap, "center": po.CenterMap}


@pytest.mark.parametrize("mapping", ["fiveprime", "threeprime", "center"])
@pytest.mark.parametrize("param", [0, 10])
@pytest.mark.parametrize("strand", ["+", "-"])
def test_kat_fiveprime_threeprime_center(mapping, param, strand):
    reads = kat_reads()[strand]
    kept, counts = FACTORY[mapping](param)(reads, SEG[strand])
    assert kept == reads
    assert (counts == kat_expected()[(mapping, param, strand)]).all()      # exact, as the reference test


@pytest.mark.parametrize("strand", ["+", "-"])
def test_kat_variable(strand):
    reads = kat_reads()[strand]
    fancy = {L: L // 2 for L in range(25, 40)}
    expected = np.zeros(2000)
    for r in reads:
        idx = fancy[len(r.positions)]
        expected[r.positions[idx] if strand == "+" else r.positions[-idx - 1]] += 1
    for dict_, exp in (({"default": 0}, kat_expected()[("fiveprime", 0, strand)]), (fancy, expected)):
        _, counts = po.VariableFivePrimeMap(dict_)(reads, SEG[strand])
        assert (counts == exp).all()
        text = "\n".join("%s\t%s" % kv for kv in dict_.items())
        _, counts = po.VariableFivePrimeMap(po.parse_offset_file(io.StringIO(text)))(reads, SEG[strand])
        assert (counts == exp).all()


@pytest.mark.parametrize("strand", ["+", "-"])
def test_kat_unmappable(strand):
    reads = kat_reads()[strand]
    lens = [len(r.positions) for r in reads]
    cases = {
        "fiveprime": (po.FivePrimeMap(30), sum(L > 30 for L in lens)),
        "threeprime": (po.ThreePrimeMap(30), sum(L > 30 for L in lens)),
        "center": (po.CenterMap(15), sum(L > 30 for L in lens)),
        "variable": (po.VariableFivePrimeMap({25: 10, "default": 28}), sum(L > 28 or L == 25 for L in lens)),
    }
    for name, (fn, n_exp) in cases.items():
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            kept, counts = fn(reads, SEG[strand])
        assert len(kept) == n_exp, name
        assert abs(counts.sum() - n_exp) < 1e-9, name
        assert any(issubclass(x.category, po.DataWarning) for x in w), name


def test_offset_file_parser():
    lines = ["26\t12", "27\t12", "28\t13", "29\t13", "30\t14", "31\t13", "default\t13"]
    exp = {26: 12, 27: 12, 28: 13, 29: 13, 30: 14, 31: 13, "default": 13}
    assert po.parse_offset_file(io.StringIO("\n".join(lines))) == exp
    assert po.parse_offset_file(io.StringIO("length\tp_offset\n" + "\n".join(lines))) == exp
    assert po.parse_offset_file(io.StringIO(lines[-1])) == {"default": 13}
    with pytest.raises(po.MalformedFileError):
        po.parse_offset_file(io.StringIO("\n".join(lines + lines)))
    with pytest.raises(po.MalformedFileError):
        po.parse_offset_file(io.StringIO("25\t12\n28\t15\t52\n"))
    with pytest.raises(po.MalformedFileError):
        po.parse_offset_file(io.StringIO("25\t12\n27\n"))
    with pytest.raises(po.MalformedFileError):
        po.parse_offset_file(io.StringIO("25\t12\nabc\t13\n"))
    with pytest.raises(po.MalformedFileError):
        po.parse_offset_file(io.StringIO("25\t12\n26\t1.5\n"))


def test_cigar_table_matches_reference_header():
    path = os.path.join(os.path.dirname(__file__), "golden", "cigar_consume_table.txt")
    table = {}
    for line in open(path):
        ch, op, q, r = line.split("\t")
        table[int(op)] = (int(q), int(r))
    for op in range(10):
        pos = po.positions_from_cigar(100, [(op, 7)])
        end = po.Read(100, [(op, 7)], False).reference_end
        q, r = table[op]
        assert (len(pos) == 7) == (q == 1 and r == 1)          # only ops consuming both emit positions
        assert (end == 107) == (r == 1)
        blocks, span = cigar_to_blocks([(op, 7)])
        assert (span == 7) == (r == 1) and (len(blocks) == 1) == (q == 1 and r == 1)


def test_c_oracle_matches_python_oracle_on_random_cigars():
    rng = np.random.default_rng(7)
    reads = random_cigar_reads(rng, 400, 6000, max_start=3000)
    hb = pack_reads({"c": reads}, {"c": 6000})
    luts = po.build_offset_luts({L: L // 3 for L in range(10, 60)} | {"default": 2})
    for strand in ("+", "-", "."):
        seg = po.Seg("c", 500, 3500, strand)
        sel = [r for r in hb.objects if strand == "." or r.is_reverse == (strand == "-")]
        for name, pyfn, kw in (
                ("fiveprime", po.FivePrimeMap(11), dict(rule="fiveprime", offset=11)),
                ("threeprime", po.ThreePrimeMap(4), dict(rule="threeprime", offset=4)),
                ("variable", po.VariableFivePrimeMap({L: L // 3 for L in range(10, 60)} | {"default": 2}),
                 dict(rule="variable", luts=luts))):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kept_py, exp = pyfn(sel, seg)
            got, kept, _d, _l = coracle.map_point(hb, 0, len(hb), kw["rule"], kw.get("offset", 0), kw.get("luts"),
                                                  None, strand, seg.start, seg.end, want_kept=True)
            assert (got == exp).all(), (name, strand)
            assert [hb.objects[i] for i in np.nonzero(kept)[0]] == kept_py, (name, strand)
        for nib in (0, 5, 12):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kept_py, exp = po.CenterMap(nib)(sel, seg)
            got, kept, _d, _l = coracle.map_center(hb, 0, len(hb), nib, None, strand, seg.start, seg.end, True)
            assert (got == exp).all(), (nib, strand)            # same order of fp64 adds => bit equal
            assert [hb.objects[i] for i in np.nonzero(kept)[0]] == kept_py
        strat = po.StratifiedVariableFivePrimeMap({20: 3, 21: 20, "default": 30}, 18, 40)
        kept_py, exp = strat(sel, seg)
        got, kept = coracle.map_stratified(hb, 0, len(hb), (strat.fw, strat.rc), 18, 40, None, strand,
                                           seg.start, seg.end, True)
        assert (got == exp).all(), strand
        assert [hb.objects[i] for i in np.nonzero(kept)[0]] == kept_py


def test_size_filter_and_strand_filter_in_c_oracle():
    rng = np.random.default_rng(3)
    reads = random_cigar_reads(rng, 300, 5000, max_start=2500)
    hb = pack_reads({"c": reads}, {"c": 5000})
    store = po.ReadStore({"c": 5000}, {"c": hb.objects})
    ga = po.OracleBAMGenomeArray(store, mapping=po.FivePrimeMap(3))
    ga.add_filter("size", po.SizeFilter(20, 45))
    for strand in ("+", "-", "."):
        seg = po.Seg("c", 100, 2900, strand)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exp = ga.get(seg, roi_order=False)
        got, _, _, _ = coracle.map_point(hb, 0, len(hb), "fiveprime", 3, None, (20, 45), strand, 100, 2900)
        assert (got == exp).all()


DOC_ROWS = """ORFL1W_(RL1)    merlin:1316-2398(+)     1.14000000e+02  1.05360444e-01          2.10520051e+02  1082
ORFL2C          merlin:2401-2772(-)     1.00000000e+01  2.69541779e-02          5.38569762e+01  371
ORFL3C          merlin:2834-3064(-)     1.50000000e+01  6.52173913e-02          1.30310466e+02  230
ORFL4C          merlin:2929-3201(-)     1.40000000e+01  5.14705882e-02          1.02843064e+02  272
ORFL5C          merlin:4074-4307(-)     2.30000000e+01  9.87124464e-02          1.97236729e+02  233
ORFL6C          merlin:4078-4488(-)     6.10000000e+01  1.48780488e-01          2.97277373e+02  410
ORFL7C          merlin:4335-4739(-)     6.20000000e+01  1.53465347e-01          3.06638160e+02  404"""


def test_counts_in_region_rows_printed_in_the_reference_docs():
    """docs/source/examples/gene_expression.rst:75-83 (total_dataset_counts: 500477): the oracle's
    and the product's row arithmetic + formatting reproduce every printed value."""
    from oracle import scripts as osc
    from plastid_b200.bin.counts_in_region import format_row
    for line in DOC_ROWS.split("\n"):
        name, region, counts, rpnt, rpkm, length = line.split()
        exp = [name, region, counts, rpnt, rpkm, length]
        assert osc.format_counts_row(name, region, float(counts), int(length), 500477) == exp
        assert format_row(name, region, float(counts), int(length), 1000.0 * 1e6 / 500477) == exp
        chain = po.Chain.from_str(region)
        assert chain.length == int(length) and str(chain) == region


def test_offsets_table_from_reference_docs():
    """docs/source/examples/p_site.rst:143-151 style table round-trips through both parsers."""
    import plastid_b200 as pb
    text = "length\tp_offset\n29\t12\n30\t12\n31\t13\n32\t14\n33\t14\n34\t14\n35\t14\ndefault\t14"
    exp = {29: 12, 30: 12, 31: 13, 32: 14, 33: 14, 34: 14, 35: 14, "default": 14}
    assert po.parse_offset_file(io.StringIO(text)) == exp
    fac = pb.VariableFivePrimeMapFactory.from_file(io.StringIO(text))
    fw, rc = po.build_offset_luts(exp)
    assert (fac.forward_offsets == fw).all() and (fac.reverse_offsets == rc).all()
    assert fw[28] == 14 and fw[14] == -1 and rc[30] == 17


def test_oracle_reproduces_the_reference_golden_vector_recipe():
    """count vectors built by the recipe of test_genome_array.py:1832-1866 — written in
    transcript coordinates, no CIGAR involved — against the oracle's map factories on the same reads,
    including reads spliced over one or two exon junctions."""
    from helpers import genome_array_recipe
    reads, vectors = genome_array_recipe(seed=7)
    assert sum(1 for r in reads if len(r.cigartuples) > 1) > 100          # spliced reads are in play
    n = len(next(iter(vectors.values())))
    rules = {"fiveprime": po.FivePrimeMap, "threeprime": po.ThreePrimeMap, "center": po.CenterMap}
    for (rule, par, strand), exp in vectors.items():
        mine = [r for r in reads if r.is_reverse is (strand == "-")]      # genome_array.py:811-815
        _kept, got = rules[rule](par)(mine, po.Seg("chrA", 0, n, strand))
        if rule == "center":
            np.testing.assert_allclose(got, exp, rtol=0, atol=1e-8)        # the reference's own tolerance (test_genome_array.py:254)
        else:
            assert (got == exp).all(), (rule, par, strand)


# ---------------------------------------------------------------------------------------------
# metagene generate geometry: oracle/generate.py against the reference's own tables
# (plastid/test/unit/bin/test_metagene.py:118-326, data transcribed to tests/golden/metagene_generate.json)
# ---------------------------------------------------------------------------------------------
from oracle import generate as og                              # noqa: E402
from helpers import metagene_generate_golden, gff3_transcript_records   # noqa: E402


def _oracle_transcripts(gold):
    out = {}
    for name, rec in gff3_transcript_records(gold["transcripts_gff"]).items():
        segs = [po.Seg(rec["chrom"], s, e, rec["strand"]) for s, e in rec["segments"]]
        out[name] = og.Tx(*segs, ID=name, gene_id=rec["gene_id"], cds_genome_start=rec["cds_genome_start"],
                          cds_genome_end=rec["cds_genome_end"])
    return out


def _check_window(result, known):
    roi, offset, ref_point = result
    known_roi, known_offset, known_ref = known
    assert str(roi) == str(po.Chain.from_str(known_roi))
    if known_offset is None or known_ref is None:                # nan in the reference's tables
        assert np.isnan(offset) and ref_point is np.nan
    else:
        assert offset == known_offset and tuple(ref_point) == tuple(known_ref)


@pytest.mark.parametrize("which", ["cds_start", "cds_stop", "cds_stop_with_delta"])
def test_oracle_window_functions_match_reference_tables(which):
    gold = metagene_generate_golden()
    txs = _oracle_transcripts(gold)
    func = og.window_cds_start if which == "cds_start" else og.window_cds_stop
    queries = gold["cds_start_queries" if which == "cds_start" else "cds_stop_queries"]
    n = 0
    for up, down in gold["flanks"]:
        for txid in queries:
            known = gold[which + "_results"]["%s_%s_%s" % (txid, up, down)]
            _check_window(func(txs[txid], up, down, ref_delta=3 if which.endswith("delta") else 0), known)
            n += 1
    assert n == 40


def test_oracle_window_landmark_properties():
    """test_metagene.py:179-216 (check_window_landmark) on the same two spliced chains."""
    for strand in "+-":
        chain = po.Chain(po.Seg("chrA", 50, 350, strand), po.Seg("chrA", 500, 900, strand))
        for landmark in range(0, 700, 50):
            roi, offset, ref = og.window_landmark(chain, 50, 100, landmark=landmark)
            assert ref == og.get_genomic_coordinate(chain, landmark)
            assert ref[1] in roi.position_list
            if landmark + 100 <= roi.length:
                assert offset + roi.length == 150
            else:
                assert offset + roi.length <= 150
            assert og.get_segmentchain_coordinate(roi, ref[1]) + offset == 50


def _check_maximal_windows(rows, result_groups, up):
    result_groups = sorted(result_groups, key=lambda x: x[0])
    rows = sorted(rows, key=lambda r: r["region"])
    c = 0
    for n, group in enumerate(result_groups):
        if group[1] is None or group[2] is None:
            c += 1
            continue
        row = rows[n - c]
        assert str(po.Chain.from_str(group[0])) == row["region"]
        assert group[1] == row["alignment_offset"] and group[2] == row["zero_point"] == up
        if len(group) == 4:
            assert group[3] == row["masked"]
    assert len(result_groups) - c == len(rows)


@pytest.mark.parametrize("masked", [False, True])
def test_oracle_maximal_spanning_windows_match_reference_tables(masked):
    gold = metagene_generate_golden()
    txs = _oracle_transcripts(gold)
    mask_hash = po.GenomeHash([po.Chain.from_str(m) for m in gold["masks"]] if masked else [])
    results = gold["do_generate_max_window_results_masked" if masked else "do_generate_max_window_results"]
    for up, down in gold["flanks"]:
        for name, group in gold["do_generate_max_window"].items():
            rows = og.group_regions_make_windows([txs[t] for t in group], mask_hash, up, down, og.window_cds_start)
            _check_maximal_windows(rows, [results["%s_%s_%s" % (name, up, down)]], up)
    if not masked:
        for name, group in gold["do_generate_multi_gene"].items():
            rows = og.group_regions_make_windows([txs[t] for t in group], mask_hash, 50, 100, og.window_cds_start)
            _check_maximal_windows(rows, gold["do_generate_multi_gene_results"]["%s_50_100" % name], 50)


def test_oracle_cs_generate_hand_derived_case():
    """``cs generate`` has no in-tree known answers in the reference (parity unpinned): the oracle is
    held to a hand-derived case instead (tests/helpers.py:cs_generate_hand_case)."""
    from helpers import cs_generate_hand_case
    records, masks, genes, transcripts = cs_generate_hand_case()
    txs = {n: og.Tx(*[po.Seg(r["chrom"], s, e, r["strand"]) for s, e in r["segments"]], ID=n, gene_id=r["gene_id"],
                    cds_genome_start=r["cds_genome_start"], cds_genome_end=r["cds_genome_end"]) for n, r in records.items()}
    g_rows, t_rows, merged = og.cs_process_partial_group(txs, po.GenomeHash([po.Chain(po.Seg(*m)) for m in masks]))
    assert merged == {"A": "A", "B": "B"}
    assert {r["region"]: {k: r[k] for k in genes[r["region"]]} for r in g_rows} == genes
    assert {r["region"]: {k: r[k] for k in transcripts[r["region"]]} for r in t_rows} == transcripts


def test_merge_genes_host_matches_oracle():
    """plastid_b200.bin.cs.merge_genes (union-find) against the restated merge_sets loop (cs.py:190-239)."""
    from helpers import random_gene_models, add_shared_exon_genes
    import plastid_b200 as pb
    from plastid_b200.bin import cs
    rng = np.random.default_rng(21)
    recs = add_shared_exon_genes(random_gene_models(rng, 120, spacing=900), rng)
    otx = {n: og.Tx(*[po.Seg(r["chrom"], s, e, r["strand"]) for s, e in r["segments"]], ID=n, gene_id=r["gene_id"])
           for n, r in recs.items()}
    ptx = {n: pb.Transcript(*[pb.GenomicSegment(r["chrom"], s, e, r["strand"]) for s, e in r["segments"]], ID=n,
                            gene_id=r["gene_id"]) for n, r in recs.items()}
    exp = og.merge_genes(otx)
    assert cs.merge_genes(ptx) == exp
    assert len(set(exp.values())) < len(exp) and any(v.count(",") == 2 for v in exp.values())   # pairs and a chain of three


def test_merge_sets_known_answers_from_the_reference_tests():
    """plastid/test/unit/util/services/test_sets.py:14-83, transcribed: the grouping step of `cs generate`
    (merge_genes: genes sharing an exon are merged transitively).  Checked for the oracle's merge_sets and,
    with every set turned into an exon shared by its member genes, for the product's union-find merge_genes."""
    import plastid_b200 as pb
    from plastid_b200.bin import cs
    a = ["h", "abm", "c", "c", "bj", "i", "ko", "n", "go", "ik", "ei", "a", "dh", "l", "gjk", "f", "b", "gn", "dmp", "in"]
    b = ["gm", "o", "ag", "f", "h", "h", "o", "bl", "e", "p", "j", "p", "k", "cf", "bc", "b", "ik", "g", "jm", "bn", "i",
         "do", "bl", "a", "ap"]
    exp_a = ["c", "f", "l", "abdeghijkmnop"]
    exp_b = ["h", "e", "ik", "do", "bcfln", "agjmp"]
    tform = lambda groups: frozenset(frozenset(g) for g in groups)          # noqa: E731
    for sets, expected in ((a, exp_a), (b, exp_b)):
        assert tform(og.merge_sets([set(s) for s in sets])) == tform(expected)
        txs = {}
        for k, members in enumerate(sets):
            for gene in members:                                    # exon k belongs to a transcript of every member gene
                name = "%s.%d" % (gene, k)
                txs[name] = pb.Transcript(pb.GenomicSegment("chrA", 1000 * k, 1000 * k + 100, "+"), ID=name, gene_id=gene)
        merged = cs.merge_genes(txs)
        assert tform(v.split(",") for v in merged.values()) == tform(expected)
        assert all(merged[g] == ",".join(sorted(merged[g].split(","))) for g in merged)
