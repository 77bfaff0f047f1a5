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
