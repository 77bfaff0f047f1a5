"""GPU parity of the region kernels of round 2 (``pb_regions.cu``) and of the drop-in container on the fast path:
flattened 16-byte region sums, column-major window rows, the plane-free ``pb_chain_counts``, position-range
ownership inside ``BAMGenomeArray`` and the transfer-format upload it uses."""
import warnings

import numpy as np
import pytest
import torch

import plastid_b200 as pb
from plastid_b200 import synth, _lib
from plastid_b200 import dist as pdist
from plastid_b200.batch import DeviceBatch
from plastid_b200.genome_array import (CountPlanes, map_batch, region_sums, gather_windows, gather_chains, chain_counts,
                                       map_wire16_streamed)
from plastid_b200.regions import ChainTable, VIRTUAL_BIN
from oracle import coracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(cuda_device):
    chroms, lens = synth.yeast_like_genome(total=2_500_000, n_chrom=5)
    lens[1] = 16384 * 8 + 5
    lens[3] = 199_999
    ann = synth.make_annotation(chroms, lens, 250, seed=11, exons=(1, 4), exon_len=(90, 500), intron_len=(40, 300))
    dbatch = synth.riboseq_reads(ann, 500_000, seed=3, device=cuda_device, lengths=range(20, 40))
    hb = synth.device_batch_to_host(dbatch, chroms, lens)
    layout = pb.GenomeLayout(chroms, lens)
    return dict(chroms=chroms, lens=lens, ann=ann, dbatch=dbatch, hb=hb, layout=layout)


def host_planes(planes, dtype):
    out = {}
    for s, t in planes.planes.items():
        a = t.cpu().numpy()
        out[s] = a.view(np.uint32) if dtype == "u32" else a
    return out


def oracle_table(vecs, table):
    """C oracle over whole-layout host vectors: (sums, live) for every chain."""
    sums = np.zeros(table.n_chains)
    live = np.zeros(table.n_chains, dtype=np.int64)
    for pidx, strand in enumerate("+-."):
        sel = np.nonzero(table.chain_plane == pidx)[0]
        if not len(sel) or strand not in vecs:
            continue
        for i in sel:
            a, b = table.chain_off[i], table.chain_off[i + 1]
            bs, be = np.clip(table.bstart[a:b], 0, len(vecs[strand])), np.clip(table.bend[a:b], 0, len(vecs[strand]))
            virt = table.bstart[a:b] >= VIRTUAL_BIN
            if virt.any():          # positions beyond the chromosome count zero but keep their place: sum the real blocks
                keep_bits = None
                if table.mask_bits is not None:
                    bits = np.unpackbits(table.mask_bits, bitorder="little")[int(table.mask_off[i]):int(table.mask_off[i]) + int(table.chain_len[i])]
                    keep_bits = bits == 0
                pos, tot, lv = 0, 0.0, 0
                for k in range(a, b):
                    n = int(table.bend[k] - table.bstart[k])
                    kb = np.ones(n, dtype=bool) if keep_bits is None else keep_bits[pos:pos + n]
                    if table.bstart[k] < VIRTUAL_BIN:
                        tot += float(vecs[strand][int(table.bstart[k]):int(table.bend[k])][kb].sum())
                    lv += int(kb.sum())
                    pos += n
                sums[i], live[i] = tot, lv
                continue
            mb = mo = None
            if table.mask_bits is not None:
                mb, mo = table.mask_bits, table.mask_off[i:i + 1]
            s, l = coracle.region_sums(vecs[strand], bs, be, [0, b - a], mb, mo)
            sums[i], live[i] = s[0], l[0]
    return sums, live


def masked_chains(w, frac=0.25, seed=2):
    chains = w["ann"].chains()
    for ch, m in zip(chains, synth.make_masks(w["ann"], frac=frac, seed=seed)):
        if m:
            ch.add_masks(*m)
    return chains


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("kind", ["u32", "f64"])
def test_region_sums_flattened_groups_match_oracle(world, cuda_device, masked, kind):
    """Blocks start and end at every alignment (mod 4 / mod 2), chains have 1..4 blocks; sums are exact for the
    integer planes and within 1e-12 for fp64 (fixed-shape tree, not the reference's sequential order)."""
    w = world
    fac = pb.FivePrimeMapFactory(13) if kind == "u32" else pb.CenterMapFactory(6)
    planes = map_batch(w["dbatch"], w["layout"], fac, None, strands=("+", "-"))
    chains = masked_chains(w) if masked else w["ann"].chains()
    table = ChainTable.from_chains(chains, w["layout"])
    sums, live = region_sums(planes, table)
    exp_s, exp_l = oracle_table(host_planes(planes, kind), table)
    assert (live.cpu().numpy() == exp_l).all()
    if kind == "u32":
        assert (sums.cpu().numpy() == exp_s).all()
    else:
        np.testing.assert_allclose(sums.cpu().numpy(), exp_s, rtol=1e-12, atol=0)
    assert {int(b) % 4 for b in table.bstart} == {0, 1, 2, 3}


def test_region_sums_chain_with_many_blocks_and_tiny_blocks(world, cuda_device):
    """More than 32 blocks in one chain (several bound batches per warp), 1-nt blocks, a chain of one position."""
    w = world
    planes = map_batch(w["dbatch"], w["layout"], pb.FivePrimeMapFactory(0), None, strands=("+", "-", "."))
    chrom = w["chroms"][0]
    segs = [pb.GenomicSegment(chrom, 1000 + 37 * k, 1000 + 37 * k + 1 + (k % 9), "+") for k in range(75)]
    big = pb.SegmentChain(*segs)
    big.add_masks(pb.GenomicSegment(chrom, 1100, 1400, "+"), pb.GenomicSegment(chrom, 3000, 3001, "+"))
    one = pb.SegmentChain(pb.GenomicSegment(chrom, 5003, 5004, "-"))
    unstranded = pb.SegmentChain(pb.GenomicSegment(chrom, 2001, 2777, "."), pb.GenomicSegment(chrom, 9000, 9003, "."))
    table = ChainTable.from_chains([big, one, unstranded, pb.SegmentChain()], w["layout"])
    sums, live = region_sums(planes, table)
    exp_s, exp_l = oracle_table(host_planes(planes, "u32"), table)
    assert (sums.cpu().numpy() == exp_s).all() and (live.cpu().numpy() == exp_l).all()
    assert live.cpu().numpy()[0] == big.masked_length and exp_s.sum() > 0


def test_region_sums_staged_by_the_copy_engine_give_the_same_table(world, cuda_device, monkeypatch):
    """``PB_REGION_TMA=1``: the blocks are staged in shared memory by ``cp.async.bulk`` + mbarrier instead of being read
    with 16-byte loads (the measured-slower variant kept for A/B, profiles/NOTES_r02.md 7.14): same sums and lengths on
    uint32 and float64 planes, with masks, blocks longer than a staging slot, tiny blocks and an empty chain."""
    w = world
    chrom = w["chroms"][0]
    chains = [pb.SegmentChain(*[pb.GenomicSegment(chrom, 1000 + 37 * k, 1000 + 37 * k + 1 + (k % 9), "+") for k in range(75)]),
              pb.SegmentChain(pb.GenomicSegment(chrom, 20_001, 23_777, "-"), pb.GenomicSegment(chrom, 30_003, 30_500, "-")),   # > one slot
              pb.SegmentChain(pb.GenomicSegment(chrom, 5003, 5004, "-")),
              pb.SegmentChain(pb.GenomicSegment(chrom, 2001, 2777, "."), pb.GenomicSegment(chrom, 9000, 9003, ".")),
              pb.SegmentChain()]
    chains[0].add_masks(pb.GenomicSegment(chrom, 1100, 1400, "+"), pb.GenomicSegment(chrom, 3000, 3001, "+"))
    chains[1].add_masks(pb.GenomicSegment(chrom, 21_000, 22_501, "-"))
    chains += w["ann"].chains()[:200]
    table = ChainTable.from_chains(chains, w["layout"])
    for fac in (pb.FivePrimeMapFactory(0), pb.CenterMapFactory(3)):
        planes = map_batch(w["dbatch"], w["layout"], fac, None, strands=("+", "-", "."))
        monkeypatch.delenv("PB_REGION_TMA", raising=False)
        s0, l0 = region_sums(planes, table)
        s0, l0 = s0.clone(), l0.clone()
        monkeypatch.setenv("PB_REGION_TMA", "1")
        s1, l1 = region_sums(planes, table)
        assert torch.equal(l0, l1) and float(s0.sum()) > 0
        if planes.dtype == "u32":
            assert torch.equal(s0, s1)
        else:
            assert torch.allclose(s0, s1, rtol=1e-12, atol=0)          # the lanes meet the groups in another order
    monkeypatch.delenv("PB_REGION_TMA", raising=False)


def test_regions_beyond_the_chromosome_count_zero_there(world, cuda_device):
    """a region reaching past its chromosome's end (annotation / BAM length mismatch) must not read the next
    chromosome's bins.  The reference returns zeros for those positions (fetch yields no reads) and keeps the length."""
    w = world
    lay = w["layout"]
    ga = pb.BAMGenomeArray(w["hb"], mapping=pb.FivePrimeMapFactory(0), device=cuda_device)
    c0, n0 = w["chroms"][0], int(w["lens"][0])
    last, nl = w["chroms"][-1], int(w["lens"][-1])
    straddle = pb.SegmentChain(pb.GenomicSegment(c0, n0 - 300, n0 + 40_000, "+"))
    beyond = pb.SegmentChain(pb.GenomicSegment(last, nl + 10, nl + 500, "-"))
    tail = pb.SegmentChain(pb.GenomicSegment(last, nl - 200, nl + 20_000, "+"))
    inside = pb.SegmentChain(pb.GenomicSegment(c0, n0 - 300, n0, "+"))
    for planes in (True, False):
        sums, live = ga.count_chains([straddle, beyond, tail, inside], planes=planes)
        assert sums[0] == sums[3] and live[0] == 40_300 and sums[1] == 0 and live[1] == 490 and live[2] == 20_200
    vec = ga[pb.GenomicSegment(c0, n0 - 300, n0 + 500, "+")]
    assert len(vec) == 800 and (vec[300:] == 0).all() and vec[:300].sum() == sums[3]
    assert (ga[pb.GenomicSegment(last, nl + 5, nl + 50, "-")] == 0).all()
    mc = straddle.get_masked_counts(ga)
    assert len(mc) == 40_300 and mc.sum() == sums[0]
    table = ChainTable.from_chains([straddle], lay)
    assert table.bstart[1] >= VIRTUAL_BIN and table.chain_len[0] == 40_300


@pytest.mark.parametrize("kind", ["u32", "f64"])
def test_window_rows_column_major_match_python(world, cuda_device, kind):
    """Rows laid 5'->3' at a column offset, reversed for '-' chains, masked cells flagged, NaN where no position."""
    w = world
    fac = pb.FivePrimeMapFactory(13) if kind == "u32" else pb.CenterMapFactory(6)
    planes = map_batch(w["dbatch"], w["layout"], fac, None, strands=("+", "-"))
    vecs = host_planes(planes, kind)
    table, cols = synth.window_table(w["ann"], w["layout"], width=210, mask_frac=0.1)
    mat, mmask = gather_windows(planes, table, cols, 210)
    mat, mmask = mat.cpu().numpy(), mmask.cpu().numpy()
    bits = np.unpackbits(table.mask_bits, bitorder="little")
    for i in list(range(0, table.n_chains, 5)) + [table.n_chains - 1]:
        a, b = table.chain_off[i], table.chain_off[i + 1]
        pos = np.concatenate([np.arange(s, e) for s, e in zip(table.bstart[a:b], table.bend[a:b])])
        vals = vecs["+-"[table.chain_plane[i]]][pos].astype(float)
        mk = bits[int(table.mask_off[i]):int(table.mask_off[i]) + len(pos)].astype(bool)
        if table.chain_reverse[i]:
            vals, mk = vals[::-1], mk[::-1]
        row = np.full(210, np.nan)
        mrow = np.ones(210, dtype=bool)
        row[cols[i]:cols[i] + len(pos)] = vals
        mrow[cols[i]:cols[i] + len(pos)] = mk
        assert np.array_equal(mat[i], row, equal_nan=True) and (mmask[i].astype(bool) == mrow).all(), i


def test_ragged_count_vectors_match_object_path(world, cuda_device):
    w = world
    ga = pb.BAMGenomeArray(w["hb"], mapping=pb.FivePrimeMapFactory(13), device=cuda_device)
    chains = masked_chains(w)[:60]
    table = ChainTable.from_chains(chains, w["layout"])
    values, masked, row_off = gather_chains(ga.count_planes(("+", "-")), table)
    values, masked = values.cpu().numpy(), masked.cpu().numpy().astype(bool)
    for i in range(0, 60, 7):
        mc = chains[i].get_masked_counts(ga)
        assert (values[row_off[i]:row_off[i + 1]] == np.ma.getdata(mc)).all()
        assert (masked[row_off[i]:row_off[i + 1]] == np.ma.getmaskarray(mc)).all()


# ---------------------------------------------------------------------------------------------
# plane-free region counts
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rule", ["fiveprime", "threeprime", "variable"])
@pytest.mark.parametrize("spliced", [False, True])
def test_plane_free_counts_equal_sums_over_planes(world, cuda_device, rule, spliced):
    """pb_chain_counts == pb_region_sums(pb_map_point(...)) for every chain: both strands and '.', masks, size filter,
    reads the rule cannot place, spliced reads whose site lies in a later block."""
    w = world
    if spliced:
        dbatch = DeviceBatch.from_host(synth.device_batch_to_host(
            synth.rnaseq_reads(w["chroms"], w["lens"], 80_000, seed=4, device="cpu", intron=(50, 3000)), w["chroms"], w["lens"]),
            cuda_device)
        sf = None
    else:
        dbatch, sf = w["dbatch"], pb.SizeFilterFactory(22, 36)
    fac = {"fiveprime": pb.FivePrimeMapFactory(24), "threeprime": pb.ThreePrimeMapFactory(3),
           "variable": pb.VariableFivePrimeMapFactory({25: 12, 26: 12, 27: 13, 28: 13, 29: 14, 30: 14, 31: 14, 100: 47})}[rule]
    chains = masked_chains(w)
    chrom = w["chroms"][2]
    chains.append(pb.SegmentChain(pb.GenomicSegment(chrom, 100, 30_000, "."), pb.GenomicSegment(chrom, 31_000, 45_000, ".")))
    chains[-1].add_masks(pb.GenomicSegment(chrom, 5000, 9000, "."))
    table = ChainTable.from_chains(chains, w["layout"])
    planes = map_batch(dbatch, w["layout"], fac, sf, strands=("+", "-", "."))
    want_s, want_l = region_sums(planes, table)
    import torch
    stats = torch.zeros(_lib.PB_NSTATS, dtype=torch.int64, device=cuda_device)
    got_s, got_l = chain_counts(dbatch, w["layout"], fac, sf, table, stats=stats)
    assert torch.equal(got_s, want_s) and torch.equal(got_l, want_l)
    assert float(want_s.sum()) > 0
    if rule != "threeprime" and not spliced:          # 20-24 nt reads cannot be placed by these rules: the warning paths
        assert stats.cpu().numpy()[:3].sum() > 0
    # position ranges: sites are counted by the rank owning them, partial tables add up
    cuts = [0, 16384 * 9, 16384 * 40, int(w["layout"].total_bins)]
    acc = torch.zeros_like(want_s)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        s, l = chain_counts(dbatch, w["layout"], fac, sf, table, (lo, hi))
        acc += s
        assert torch.equal(l, want_l)               # lengths are geometry: whole on every rank
    assert torch.equal(acc, want_s)


def test_range_kernels_add_up_over_ranks(world, cuda_device):
    """Region sums, window matrices and phase sums over range-only planes: every position is owned by one rank."""
    import torch
    from plastid_b200.genome_array import phase_sums, stratified_windows
    w = world
    lay, fac = w["layout"], pb.FivePrimeMapFactory(13)
    whole = map_batch(w["dbatch"], lay, fac, None, strands=("+", "-"))
    table = ChainTable.from_chains(masked_chains(w), lay)
    wtable, cols = synth.window_table(w["ann"], lay, width=180, mask_frac=0.1)
    ref_s, ref_l = region_sums(whole, table)
    ref_m, ref_mm = gather_windows(whole, wtable, cols, 180)
    ref_p = phase_sums(whole, table, 2, -2)
    ref_strat, ref_smask = stratified_windows(w["dbatch"], lay, fac, None, wtable, cols, 180, 25, 34)
    hb = w["hb"]
    world_size = 3
    cuts = pdist.position_cuts(hb, lay, world_size)
    acc_s, acc_m, acc_p = torch.zeros_like(ref_s), torch.zeros_like(ref_m), torch.zeros_like(ref_p)
    acc_strat = torch.zeros_like(ref_strat)
    for rank in range(world_size):
        sub, lo, hi = pdist.shard_positions(hb, lay, rank, world_size, cuts)
        dsub = DeviceBatch.from_host(sub, cuda_device)
        planes = map_batch(dsub, lay, fac, None, strands=("+", "-"), bin_range=(lo, hi))
        s, l = region_sums(planes, table)
        m, mm = gather_windows(planes, wtable, cols, 180)
        acc_s += s
        acc_m += m
        acc_p += phase_sums(planes, table, 2, -2)
        st, sm = stratified_windows(dsub, lay, fac, None, wtable, cols, 180, 25, 34, bin_range=(lo, hi))
        acc_strat += st
        assert torch.equal(l, ref_l) and torch.equal(mm, ref_mm) and torch.equal(sm, ref_smask)
    assert torch.equal(acc_s, ref_s) and torch.equal(acc_p, ref_p) and torch.equal(acc_strat, ref_strat)
    assert torch.equal(torch.nan_to_num(acc_m, nan=-1.0), torch.nan_to_num(ref_m, nan=-1.0))
    # chromosome cuts (BASELINE config 5): every cut is a chromosome boundary
    ccuts = pdist.position_cuts(hb, lay, 3, snap="chromosomes")
    assert set(int(c) for c in ccuts) <= set(int(x) for x in lay.chrom_bin_off)


# ---------------------------------------------------------------------------------------------
# the container on the fast path
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rule", ["variable", "center_spliced", "threeprime_spliced", "center_unspliced"])
def test_api_ships_the_transfer_format_and_equals_the_soa_path(world, cuda_device, rule):
    """``BAMGenomeArray(batch).count_chains(table)`` on a batch that carries its transfer format
    (what the decoder emits) goes through the chunked delta3 upload + range mapping; tables and planes are identical
    to the plain SoA upload and to the engine called directly."""
    import torch
    w = world
    if "spliced" in rule and rule != "center_unspliced":
        hb = synth.device_batch_to_host(synth.rnaseq_reads(w["chroms"], w["lens"], 120_000, seed=8, device="cpu", intron=(50, 3000)),
                                        w["chroms"], w["lens"])
        sf = None
    else:
        hb, sf = w["hb"], pb.SizeFilterFactory(22, 38)
    fac = {"variable": pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS), "center_spliced": pb.CenterMapFactory(12),
           "threeprime_spliced": pb.ThreePrimeMapFactory(2), "center_unspliced": pb.CenterMapFactory(3)}[rule]
    from plastid_b200.batch import AlignmentBatch
    packed = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start, hb.meta, hb.chrom_read_off, hb.blk_off, hb.blk,
                            max_span=hb.max_span, mapped=hb.mapped).pack()
    plain = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start, hb.meta, hb.chrom_read_off, hb.blk_off, hb.blk,
                           max_span=hb.max_span, mapped=hb.mapped)
    assert packed.transfer is not None and plain.transfer is None
    table = ChainTable.from_chains(masked_chains(w), w["layout"])
    out = {}
    for name, batch in (("packed", packed), ("plain", plain)):
        ga = pb.BAMGenomeArray(batch, mapping=fac, device=cuda_device)
        if sf is not None:
            ga.add_filter("size", sf)
        planes = ga.count_planes(("+", "-"))
        sums, live = ga.count_chains(table, planes=True)
        out[name] = (planes, sums, live, ga)
    assert out["packed"][3]._receiver is not None and out["plain"][3]._receiver is None
    for s in ("+", "-"):
        assert torch.equal(out["packed"][0].planes[s], out["plain"][0].planes[s])
    assert np.array_equal(out["packed"][1], out["plain"][1]) and np.array_equal(out["packed"][2], out["plain"][2])
    assert (out["packed"][0].stats == out["plain"][0].stats).all()
    # the engine called directly on the SoA
    ref = map_batch(DeviceBatch.from_host(hb, cuda_device), w["layout"], fac, sf, strands=("+", "-"))
    for s in ("+", "-"):
        assert torch.equal(ref.planes[s], out["packed"][0].planes[s])
    if not isinstance(fac, pb.CenterMapFactory):        # plane-free path from the uploaded transfer format
        ga = pb.BAMGenomeArray(packed, mapping=fac, device=cuda_device)
        if sf is not None:
            ga.add_filter("size", sf)
        s2, l2 = ga.count_chains(table)                 # no planes yet: pb_chain_counts
        assert ga._planes is None and np.array_equal(s2, out["packed"][1]) and np.array_equal(l2, out["packed"][2])


def test_decoder_emits_the_transfer_format(tmp_path, cuda_device):
    """``batch_from_bam`` packs while decoding; ``BAMGenomeArray(path)`` therefore uploads delta3, not the SoA."""
    from plastid_b200.bam_io import write_bam, batch_from_bam
    rng = np.random.default_rng(5)
    lens = {"chrA": 70_000, "chrB": 40_000}
    recs = []
    for tid, n in enumerate(lens.values()):
        starts = np.sort(rng.integers(0, n - 200, 3000))
        for s in starts:
            L = int(rng.integers(24, 36))
            cigar = [(0, L)] if rng.random() < 0.8 else [(0, 10), (3, int(rng.integers(20, 90))), (0, L - 10)]
            recs.append((tid, int(s), 16 if rng.random() < 0.5 else 0, cigar))
    path = str(tmp_path / "x.bam")
    write_bam(path, lens, recs)
    hb = batch_from_bam(path)
    assert hb.transfer is not None and hb.transfer.nbytes < 8 * len(hb)
    ga = pb.BAMGenomeArray(path, mapping=pb.FivePrimeMapFactory(5), device=cuda_device)
    plain = batch_from_bam(path, pack=False)
    gb = pb.BAMGenomeArray(plain, mapping=pb.FivePrimeMapFactory(5), device=cuda_device)
    seg = pb.GenomicSegment("chrA", 0, 70_000, "+")
    assert (ga[seg] == gb[seg]).all() and ga[seg].sum() > 0 and ga._receiver is not None and gb._receiver is None


@pytest.mark.parametrize("world_size", [2, 5])
@pytest.mark.parametrize("sharding", ["positions", "chromosomes"])
def test_sharded_containers_add_up_to_the_single_gpu_container(world, cuda_device, world_size, sharding):
    """``BAMGenomeArray(..., shard=(rank, world))`` owns one position range (halo reads, range-only
    planes); with no process group the partial tables are summed here the way the all-reduce does."""
    w = world
    fac, sf = pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS), pb.SizeFilterFactory(22, 38)
    hb = w["hb"].pack()
    chains = masked_chains(w)
    table = ChainTable.from_chains(chains, w["layout"])
    whole = pb.BAMGenomeArray(hb, mapping=fac, device=cuda_device)
    whole.add_filter("size", sf)
    ref_s, ref_l = whole.count_chains(table, planes=True)
    seg = pb.GenomicSegment(w["chroms"][0], 16384 * 3 - 700, 16384 * 3 + 900, "+")
    ref_vec = whole[seg]
    acc_planes, acc_direct, acc_vec, total_bins, reads_held = 0, 0, 0, 0, 0
    for rank in range(world_size):
        ga = pb.BAMGenomeArray(hb, mapping=fac, device=cuda_device, shard=(rank, world_size), sharding=sharding)
        ga.add_filter("size", sf)
        lo, hi = ga.bin_range
        total_bins += hi - lo
        s_direct, l_direct = ga.count_chains(table, planes=False)
        s_planes, l_planes = ga.count_chains(table, planes=True)
        assert np.array_equal(l_direct, ref_l) and np.array_equal(l_planes, ref_l)
        assert ga._planes.planes["+"].numel() == max(hi - lo, 1)
        acc_direct = acc_direct + s_direct
        acc_planes = acc_planes + s_planes
        acc_vec = acc_vec + ga[seg]
        reads_held += len(ga._local)
    assert total_bins == w["layout"].total_bins and reads_held >= len(hb)
    assert np.array_equal(acc_planes, ref_s) and np.array_equal(acc_direct, ref_s) and np.array_equal(acc_vec, ref_vec)


def test_genome_array_get_returns_a_copy_documented_divergence(cuda_device):
    """The reference's ``GenomeArray.get`` returns a VIEW of the chromosome's numpy array (genome_array.py:1526), so
    writing into the result changes the array.  Here the planes live in HBM and ``get`` returns a fresh host vector
    (deliberate divergence): writes go through ``__setitem__``, as the reference's own scripts do."""
    ga = pb.GenomeArray({"chrA": 1000}, strands=("+", "-"), device=cuda_device)
    seg = pb.GenomicSegment("chrA", 100, 110, "+")
    ga[seg] = np.arange(10, dtype=float)
    got = ga.get(seg)
    assert (got == np.arange(10)).all()
    got[:] = 99.0                                        # a view would write through
    assert (ga.get(seg) == np.arange(10)).all()
    ga[seg] = got                                        # the supported way to write
    assert (ga[seg] == 99.0).all() and ga.sum() == 990.0


def test_vectorised_filters_are_asked_once_per_batch(world, cuda_device):
    """`add_filter` takes any predicate over reads (genome_array.py:697-722); one that also offers
    `batch_mask(batch)` is evaluated once for the whole batch — same planes as the per-read evaluation."""
    w = world

    class ForwardOnlyShort(object):
        calls = 0

        def __call__(self, read):
            ForwardOnlyShort.calls += 1
            return (not read.is_reverse) and len(read.positions) < 30

    class Vectorised(ForwardOnlyShort):
        def batch_mask(self, batch):
            return (~batch.is_reverse) & (batch.aligned_len < 30)

    small = pdist.shard_chromosomes(w["hb"], [0])                 # per-read python calls: keep it small
    a = pb.BAMGenomeArray(small, mapping=pb.FivePrimeMapFactory(3), device=cuda_device)
    a.add_filter("mine", ForwardOnlyShort())
    b = pb.BAMGenomeArray(small, mapping=pb.FivePrimeMapFactory(3), device=cuda_device)
    b.add_filter("mine", Vectorised())
    before = ForwardOnlyShort.calls
    pa, pbb = a.count_planes(("+", "-", ".")), b.count_planes(("+", "-", "."))
    import torch
    for s in "+-.":
        assert torch.equal(pa.planes[s], pbb.planes[s])
    assert ForwardOnlyShort.calls - before == len(small)          # only the first container called per read
    assert int(pa.planes["-"].sum()) == 0 and int(pa.planes["+"].sum()) > 0
