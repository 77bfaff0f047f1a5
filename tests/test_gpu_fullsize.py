"""Size-independent properties at BASELINE.json's full C2 size (200 M reads, hg38-scale genome): the
oracle cannot run there, so parity rests on conservation, linearity, idempotence and agreement with
an oracle-checked sample."""
import numpy as np
import pytest
import torch

import plastid_b200 as pb
from plastid_b200 import synth, _lib, dist as pdist
from plastid_b200.batch import DeviceBatch
from plastid_b200.genome_array import map_batch, region_sums, CountPlanes
from oracle import coracle

pytestmark = pytest.mark.gpu

N_READS = 200_000_000


@pytest.fixture(scope="module")
def c2(cuda_device):
    chroms, lens = synth.human_like_genome(1.0)
    ann = synth.make_annotation(chroms, lens, 60_000, seed=0, exons=(1, 3), exon_len=(150, 600), intron_len=(100, 3000))
    layout = pb.GenomeLayout(chroms, lens)
    dbatch = synth.riboseq_reads(ann, N_READS, seed=100, device=cuda_device, frac_in=0.85, lengths=range(22, 39))
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    offs = dict(synth.RIBO_OFFSETS)
    del offs["default"]                      # lengths outside 25..35 are dropped (and counted)
    fac = pb.VariableFivePrimeMapFactory(offs)
    return dict(chroms=chroms, lens=lens, ann=ann, layout=layout, dbatch=dbatch, fac=fac,
                table=synth.annotation_table(ann, layout))


def slice_batch(db, a, b, off):
    return DeviceBatch(b - a, db.n_chrom, db.max_span, db.ref_start[a:b], db.meta[a:b], off, None, None, db.max_block_len)


def test_conservation_idempotence_linearity_at_full_size(c2):
    db, layout, fac = c2["dbatch"], c2["layout"], c2["fac"]
    planes = map_batch(db, layout, fac, None, strands=("+", "-", "."))
    st = planes.stats
    L = (db.meta & 0xFFFF)
    n_drop = int(((L < 25) | (L > 35)).sum().item())
    n_rev = int((((db.meta >> 16) & 1) == 1).sum().item())
    n_drop_rev = int(((((db.meta >> 16) & 1) == 1) & ((L < 25) | (L > 35))).sum().item())
    # every read is either mapped exactly once per plane it belongs to, or dropped
    assert int(st[_lib.PB_STAT_DROPPED_ANY]) == n_drop > 0
    assert int(st[_lib.PB_STAT_DROPPED_MINUS]) == n_drop_rev and int(st[_lib.PB_STAT_DROPPED_PLUS]) == n_drop - n_drop_rev
    assert int(st[_lib.PB_STAT_MAPPED_ANY]) == N_READS - n_drop
    assert int(st[_lib.PB_STAT_MAPPED_MINUS]) == n_rev - n_drop_rev
    assert int(st[_lib.PB_STAT_MAPPED_PLUS]) == (N_READS - n_rev) - (n_drop - n_drop_rev)
    sums = {s: int((planes.planes[s].to(torch.int64) & 0xFFFFFFFF).sum().item()) for s in "+-."}
    assert sums["."] == N_READS - n_drop and sums["+"] + sums["-"] == N_READS - n_drop
    # forward reads map to the same site under '+' and '.' queries
    fwd_only = map_batch(db, layout, fac, None, strands=("+",))
    assert torch.equal(fwd_only.planes["+"], planes.planes["+"])
    # idempotence: a second run gives bit-identical planes
    again = map_batch(db, layout, fac, None, strands=("+", "-", "."))
    for s in "+-.":
        assert torch.equal(again.planes[s], planes.planes[s])
    # linearity: planes of two read-range shards add up to the planes of the whole batch
    a, b = pdist.read_range(N_READS, 0, 2)[1], N_READS
    off = db.chrom_read_off
    acc = None
    for lo, hi in ((0, a), (a, b)):
        shard = slice_batch(db, lo, hi, torch.clamp(off, lo, hi) - lo)
        part = map_batch(shard, layout, fac, None, strands=("-",))
        acc = part.planes["-"].clone() if acc is None else acc + part.planes["-"]
    assert torch.equal(acc, planes.planes["-"])
    # region table == sums of plane slices, and is linear too
    table = c2["table"]
    rs, live = region_sums(planes, table)
    rs = rs.cpu().numpy()
    assert (live.cpu().numpy() == table.chain_len).all()
    for i in range(0, table.n_chains, 997):
        pl = planes.planes["+-"[table.chain_plane[i]]]
        tot = 0
        for k in range(table.chain_off[i], table.chain_off[i + 1]):
            tot += int((pl[table.bstart[k]:table.bend[k]].to(torch.int64) & 0xFFFFFFFF).sum().item())
        assert rs[i] == tot


def test_oracle_checked_sample_at_full_size(c2):
    """chr21 of the full-size run (the smallest chromosome) against the C oracle, bit for bit."""
    db, layout, fac = c2["dbatch"], c2["layout"], c2["fac"]
    planes = map_batch(db, layout, fac, pb.SizeFilterFactory(25, 100), strands=("+", "-"))
    c = c2["chroms"].index("chr21")
    off = db.chrom_read_off.cpu().numpy()
    a, b = int(off[c]), int(off[c + 1])
    new_off = np.zeros(len(c2["chroms"]) + 1, dtype=np.int64)
    new_off[c + 1:] = b - a
    hb = pb.AlignmentBatch(c2["chroms"], c2["lens"], db.ref_start[a:b].cpu().numpy(),
                           db.meta[a:b].cpu().numpy().view(np.uint32), new_off, max_span=db.max_span)
    base = int(layout.chrom_bin_off[c])
    for strand in "+-":
        exp = coracle.genome_vector(hb, c, strand, rule="variable", luts=(fac.forward_offsets, fac.reverse_offsets),
                                    size_filter=(25, 100))[0]
        got = planes.planes[strand][base:base + int(c2["lens"][c])].cpu().numpy().view(np.uint32)
        assert (got == exp).all()
        assert not planes.planes[strand][base + int(c2["lens"][c]):int(layout.chrom_bin_off[c + 1])].any()


def test_center_rule_at_full_size(cuda_device):
    """BASELINE config 3 at full size (100 M spliced 100-nt reads, CenterMapFactory(12), fp64 planes):
    conservation (every read that lies inside its chromosome adds exactly 1 in total), bit-identical
    repeats, padding stays zero, and chr21 against the C oracle within the north star's tolerance."""
    chroms, lens = synth.human_like_genome(1.0)
    layout = pb.GenomeLayout(chroms, lens)
    db = synth.rnaseq_reads(chroms, lens, 100_000_000, seed=100, device=cuda_device)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    fac = pb.CenterMapFactory(12)
    planes = map_batch(db, layout, fac, None, strands=("+", "-"))
    st = planes.stats
    n_rev = int((((db.meta >> 16) & 1) == 1).sum().item())
    assert int(st[_lib.PB_STAT_DROPPED_ANY]) == 0
    assert int(st[_lib.PB_STAT_MAPPED_MINUS]) == n_rev and int(st[_lib.PB_STAT_MAPPED_PLUS]) == db.n_reads - n_rev
    tot = {s: float(planes.planes[s].sum().item()) for s in "+-"}
    assert abs(tot["+"] - (db.n_reads - n_rev)) <= 1e-6 * db.n_reads and abs(tot["-"] - n_rev) <= 1e-6 * db.n_reads
    first = planes.planes["-"].clone()
    again = map_batch(db, layout, fac, None, strands=("+", "-"), planes=planes)
    assert torch.equal(again.planes["-"], first)
    del first
    hb = synth.device_batch_to_host(db, chroms, lens)
    c = chroms.index("chr21")
    sub = pdist.shard_chromosomes(hb, [c])
    base = int(layout.chrom_bin_off[c])
    for strand in "+-":
        exp = coracle.genome_vector(sub, 0, strand, nibble=12)[0]
        got = planes.planes[strand][base:base + int(lens[c])].cpu().numpy()
        assert ((got == 0) == (exp == 0)).all()
        np.testing.assert_allclose(got, exp, rtol=1e-6, atol=0)
        assert not planes.planes[strand][base + int(lens[c]):int(layout.chrom_bin_off[c + 1])].any()


def test_delta8_transfer_format_at_scale(c2, cuda_device):
    """50 M reads of the C2 batch through the 2-byte transfer format: the device expansion returns the
    batch bit for bit, at under 2.3 bytes per read."""
    from plastid_b200.batch import AlignmentBatch, Delta8Batch, Delta8Receiver
    db = c2["dbatch"]
    n = 50_000_000
    off = torch.clamp(db.chrom_read_off, 0, n).cpu().numpy()
    hb = AlignmentBatch(c2["chroms"], c2["lens"], db.ref_start[:n].cpu().numpy(), db.meta[:n].cpu().numpy().view(np.uint32),
                        off, max_span=db.max_span)
    wire = Delta8Batch.from_batch(hb)
    assert wire.nbytes < 2.3 * n
    rx = Delta8Receiver(wire, cuda_device)
    out = rx.receive(wire.pinned())
    assert torch.equal(out.ref_start, db.ref_start[:n]) and torch.equal(out.meta, db.meta[:n])


def test_plane_free_counts_and_api_path_at_full_size(c2, cuda_device):
    """VERDICT r1 items 1 and 4 at BASELINE's size: the region table counted straight from the 200 M sorted reads
    (pb_chain_counts, no planes) equals the sums over the dense planes bit for bit, and so does the table the drop-in
    container produces from the transfer format the decoder emits (first 60 M reads: the host copy is the slow part)."""
    from plastid_b200.genome_array import chain_counts
    db, layout, fac, table = c2["dbatch"], c2["layout"], c2["fac"], c2["table"]
    sf = pb.SizeFilterFactory(25, 100)
    planes = map_batch(db, layout, fac, sf, strands=("+", "-"))
    want_s, want_l = region_sums(planes, table)
    got_s, got_l = chain_counts(db, layout, fac, sf, table)
    assert torch.equal(got_s, want_s) and torch.equal(got_l, want_l) and float(want_s.sum()) > 1e8
    # position ranges add up (8 "ranks")
    cuts = np.linspace(0, layout.total_bins // 16384, 9).astype(np.int64) * 16384
    cuts[-1] = layout.total_bins
    acc = torch.zeros_like(want_s)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        acc += chain_counts(db, layout, fac, sf, table, (int(lo), int(hi)))[0]
    assert torch.equal(acc, want_s)
    del planes, acc
    torch.cuda.empty_cache()
    n = 60_000_000
    off = torch.clamp(db.chrom_read_off, 0, n).cpu().numpy()
    hb = pb.AlignmentBatch(c2["chroms"], c2["lens"], db.ref_start[:n].cpu().numpy(), db.meta[:n].cpu().numpy().view(np.uint32),
                           off, max_span=db.max_span).pack()
    assert hb.transfer.nbytes < 1.6 * n
    ga = pb.BAMGenomeArray(hb, mapping=fac, device=cuda_device)
    ga.add_filter("size", sf)
    api_s, api_l = ga.count_chains(table, planes=True)
    sub = slice_batch(db, 0, n, torch.from_numpy(off).to(cuda_device))
    ref = map_batch(sub, layout, fac, sf, strands=("+", "-"))
    ref_s, ref_l = region_sums(ref, table)
    assert (api_s == ref_s.cpu().numpy()).all() and (api_l == ref_l.cpu().numpy()).all()
    for s in "+-":
        assert torch.equal(ga._planes.planes[s], ref.planes[s])
    direct_s, _ = pb.BAMGenomeArray(hb, mapping=fac, device=cuda_device).count_chains(table, planes=False)
    nofilter = region_sums(map_batch(sub, layout, fac, None, strands=("+", "-")), table)[0]
    assert (direct_s == nofilter.cpu().numpy()).all()


def test_c1_at_its_real_size_against_the_oracle(cuda_device):
    """BASELINE config 1 as stated: 12 Mb yeast-scale genome, 2 M reads, 6 k transcripts, 10 % of positions masked,
    FivePrimeMapFactory(14) + size filter 25-100 — every region's masked count and length against the C oracle."""
    chroms, lens = synth.yeast_like_genome()
    ann = synth.make_annotation(chroms, lens, 6000, seed=0, exons=(1, 2), exon_len=(300, 700), intron_len=(80, 200))
    db = synth.riboseq_reads(ann, 2_000_000, seed=100, device=cuda_device, frac_in=0.9)
    hb = synth.device_batch_to_host(db, chroms, lens).pack()
    chains = ann.chains()
    for ch, m in zip(chains, synth.make_masks(ann, frac=0.10, block=200, seed=1)):
        if m:
            ch.add_masks(*m)
    ga = pb.BAMGenomeArray(hb, mapping=pb.FivePrimeMapFactory(14), device=cuda_device)
    ga.add_filter("size", pb.SizeFilterFactory(25, 100))
    table = ga.chain_table(chains)
    direct = ga.count_chains(table, planes=False)
    dense = ga.count_chains(table, planes=True)
    assert (direct[0] == dense[0]).all() and (direct[1] == dense[1]).all()
    layout = ga.layout
    vec = {}
    for strand in "+-":
        big = np.zeros(layout.total_bins, dtype=np.int64)
        for c in range(len(chroms)):
            base = int(layout.chrom_bin_off[c])
            big[base:base + int(lens[c])] = coracle.genome_vector(hb, c, strand, rule="fiveprime", offset=14, size_filter=(25, 100))[0]
        vec[strand] = big
    for pidx, strand in enumerate("+-"):
        sel = np.nonzero(table.chain_plane == pidx)[0]
        offs = np.concatenate([[0], np.cumsum(table.chain_off[sel + 1] - table.chain_off[sel])])
        idx = np.concatenate([np.arange(table.chain_off[i], table.chain_off[i + 1]) for i in sel])
        mask_off = table.mask_off[sel] if table.mask_bits is not None else None
        es, el = coracle.region_sums(vec[strand], table.bstart[idx], table.bend[idx], offs, table.mask_bits, mask_off)
        assert (dense[0][sel] == es).all() and (dense[1][sel] == el).all()
    assert dense[0].sum() > 1e6 and (dense[1] < table.chain_len).any()


def test_c5_threeprime_sample_and_c4_windows_at_full_size(c2, cuda_device):
    """BASELINE config 5's rule (ThreePrimeMapFactory 0 and 15) on the full-size batch against the C oracle on chr21, and
    config 4's 60 k x 350 window matrix: row sums equal the sums of the same chains, all 21 M cells are accounted for."""
    from plastid_b200.genome_array import gather_windows
    db, layout = c2["dbatch"], c2["layout"]
    c = c2["chroms"].index("chr21")
    off = db.chrom_read_off.cpu().numpy()
    a, b = int(off[c]), int(off[c + 1])
    new_off = np.zeros(len(c2["chroms"]) + 1, dtype=np.int64)
    new_off[c + 1:] = b - a
    hb = pb.AlignmentBatch(c2["chroms"], c2["lens"], db.ref_start[a:b].cpu().numpy(), db.meta[a:b].cpu().numpy().view(np.uint32),
                           new_off, max_span=db.max_span)
    base = int(layout.chrom_bin_off[c])
    for offset in (0, 15):
        planes = map_batch(db, layout, pb.ThreePrimeMapFactory(offset), pb.SizeFilterFactory(25, 100), strands=("+", "-"))
        for strand in "+-":
            exp = coracle.genome_vector(hb, c, strand, rule="threeprime", offset=offset, size_filter=(25, 100))[0]
            got = planes.planes[strand][base:base + int(c2["lens"][c])].cpu().numpy().view(np.uint32)
            assert (got == exp).all()
    wtable, cols = synth.window_table(c2["ann"], layout, width=350)
    mat, mmask = gather_windows(planes, wtable, cols, 350)
    row_sums = torch.nan_to_num(mat, nan=0.0).sum(dim=1)
    import dataclasses  # noqa: F401  (keeps the import block of this module flat)
    unmasked = type(wtable)(layout, wtable.bstart, wtable.bend, wtable.chain_off, wtable.chain_plane, wtable.chain_reverse,
                            wtable.chain_len)
    sums, live = region_sums(planes, unmasked)
    assert torch.equal(row_sums, sums)
    assert int((~torch.isnan(mat)).sum().item()) == int(wtable.chain_len.sum()) > 15_000_000
    assert int((mmask == 0).sum().item()) < int(wtable.chain_len.sum())        # some positions are masked
