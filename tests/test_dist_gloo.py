"""World-size-2 gloo tests of the multi-GPU host logic (sharding plans + collectives) on CPU.  The
per-rank compute is stood in by the C oracle (checker), so the test pins that read-range sharding +
all-reduce and chromosome sharding + gather both reproduce the unsharded tables exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _world():
    import plastid_b200 as pb
    from plastid_b200 import synth
    chroms, lens = synth.yeast_like_genome(total=500_000, n_chrom=5)
    ann = synth.make_annotation(chroms, lens, 80, seed=2, exons=(1, 3), exon_len=(150, 500), intron_len=(40, 300))
    hb = synth.device_batch_to_host(synth.riboseq_reads(ann, 40_000, seed=6, device="cpu"), chroms, lens)
    spliced = synth.device_batch_to_host(synth.rnaseq_reads(chroms, lens, 8_000, seed=3, device="cpu",
                                                            intron=(50, 2000)), chroms, lens)
    return chroms, lens, ann, hb, spliced


def _region_table(hb, chains, offset=14):
    """Unmasked 5' region sums of `chains` over batch `hb` with the C oracle."""
    from oracle import coracle
    vec = {}
    out = np.zeros(len(chains))
    for i, ch in enumerate(chains):
        if ch.chrom not in hb.chroms:
            continue
        c = hb.chroms.index(ch.chrom)
        key = (c, ch.strand)
        if key not in vec:
            vec[key] = coracle.genome_vector(hb, c, ch.strand, rule="fiveprime", offset=offset, size_filter=(25, 100))[0]
        out[i] = sum(vec[key][s.start:s.end].sum() for s in ch)
    return out


def _worker(rank, world_size, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from plastid_b200 import dist as pd
        chroms, lens, ann, hb, spliced = _world()
        chains = ann.chains()
        full = _region_table(hb, chains)

        # read-range sharding: every rank a slice of the reads, tables all-reduced
        for batch in (hb, spliced):
            shard = pd.shard_reads(batch, rank, world_size)
            shard.check_sorted()
            assert shard.mapped == batch.mapped
            n = torch.tensor([len(shard)], dtype=torch.int64)
            pd.allreduce_sum(n)
            assert int(n.item()) == len(batch)
            if batch.blk is not None:
                for i in range(0, len(shard), 501):
                    a, _b = pd.read_range(len(batch), rank, world_size)
                    assert shard.positions_of(i) == batch.positions_of(a + i)
        local = torch.from_numpy(_region_table(pd.shard_reads(hb, rank, world_size), chains))
        pd.allreduce_sum(local)
        assert (local.numpy() == full).all()

        # position-range sharding: range-only vectors from the reads of the range + halo, region
        # tables clipped to the range and all-reduced
        import plastid_b200 as pb
        from plastid_b200.regions import ChainTable
        from oracle import coracle
        lay = pb.GenomeLayout(chroms, lens)
        table = ChainTable.from_chains(chains, lay, use_masks=False)
        for batch in (hb, spliced):
            cuts = pd.position_cuts(batch, lay, world_size)
            sub, lo, hi = pd.shard_positions(batch, lay, rank, world_size, cuts)
            sub.check_sorted()
            clipped = pd.clip_table(table, lo, hi)
            part = np.zeros(len(chains))
            whole = np.zeros(len(chains))
            vec_sub, vec_all = {}, {}
            for i, ch in enumerate(chains):
                c = chroms.index(ch.chrom)
                key = (c, ch.strand)
                if key not in vec_sub:
                    kw = dict(rule="fiveprime", offset=14, size_filter=(25, 100))
                    vec_sub[key] = coracle.genome_vector(sub, c, ch.strand, **kw)[0]
                    vec_all[key] = coracle.genome_vector(batch, c, ch.strand, **kw)[0]
                    base = int(lay.chrom_bin_off[c])
                    a, b = max(lo - base, 0), min(hi - base, int(lens[c]))
                    if a < b:      # inside the rank's range the shard reproduces the whole-batch vector
                        assert (vec_sub[key][a:b] == vec_all[key][a:b]).all()
                base = int(lay.chrom_bin_off[c])
                for j in range(int(clipped.chain_off[i]), int(clipped.chain_off[i + 1])):
                    part[i] += vec_sub[key][int(clipped.bstart[j]) - base:int(clipped.bend[j]) - base].sum()
                whole[i] = sum(vec_all[key][s.start:s.end].sum() for s in ch)
            t = torch.from_numpy(part)
            pd.allreduce_sum(t)
            assert (t.numpy() == whole).all()
            ln = torch.from_numpy(clipped.chain_len.copy())
            pd.allreduce_sum(ln)
            assert (ln.numpy() == table.chain_len).all()

        # chromosome sharding: disjoint ownership, tables gathered
        owned = pd.assign_chromosomes(hb, world_size)
        assert sorted(c for ids in owned for c in ids) == list(range(len(chroms)))
        loads = [sum(int(np.diff(hb.chrom_read_off)[c]) for c in ids) for ids in owned]
        assert max(loads) <= 0.75 * len(hb)
        mine = pd.shard_chromosomes(hb, owned[rank])
        mine.check_sorted()
        owner = pd.owner_of_chains(chains, hb, owned)
        my_chains = [ch for ch, o in zip(chains, owner) if o == rank]
        gathered = pd.gather_rows(torch.from_numpy(_region_table(mine, my_chains)), owner)
        assert (gathered.numpy() == full).all()

        # profiles: mean by all-reduce, median needs the gathered rows
        rng = np.random.default_rng(5)
        mat = rng.random((37, 20))
        mask = rng.random((37, 20)) < 0.2
        rows = np.array_split(np.arange(37), world_size)[rank]
        col_sum = torch.from_numpy(np.where(mask[rows], 0.0, mat[rows]).sum(0))
        cnt = torch.from_numpy((~mask[rows]).sum(0))
        prof = pd.mean_profile(col_sum, cnt)
        np.testing.assert_allclose(prof.numpy(), np.ma.MaskedArray(mat, mask=mask).mean(0).filled(np.nan), rtol=1e-12)
        allrows = pd.gather_matrix(torch.from_numpy(np.where(mask[rows], np.nan, mat[rows])))
        assert allrows.shape == (37, 20)
        med = np.nanmedian(allrows.numpy(), axis=0)
        np.testing.assert_allclose(med, np.ma.median(np.ma.MaskedArray(mat, mask=mask), axis=0).filled(np.nan), rtol=1e-12)
        # Center rule under position sharding: every rank ends up with the histogram of the whole batch
        from plastid_b200.batch import meta_length_hist
        lay = pb.GenomeLayout(chroms, lens)
        for batch in (hb, spliced):
            sub, lo, hi = pd.shard_positions(batch, lay, rank, world_size)
            assert (pd.global_length_hist(sub, lay, lo, hi) == meta_length_hist(batch.meta)).all()
        # the container shards itself when a process group exists (round 2): every rank of the group picks its own
        # position range — by reads or at chromosome boundaries — keeps the reads that can map into it (+ halo) in the
        # form the decoder emitted, and knows the length histogram of the WHOLE batch (Center tables)
        import plastid_b200 as pb2
        for sharding in ("positions", "chromosomes"):
            ga = pb2.BAMGenomeArray(hb.pack(), mapping=pb2.FivePrimeMapFactory(14), sharding=sharding)
            assert (ga._rank, ga._world) == (rank, world_size) and ga.is_sharded and ga._collective
            lo, hi = ga.bin_range
            spans = [torch.zeros(2, dtype=torch.int64) for _ in range(world_size)]
            dist.all_gather(spans, torch.tensor([lo, hi], dtype=torch.int64))
            spans = sorted((int(a), int(b)) for a, b in (x.tolist() for x in spans))
            assert spans[0][0] == 0 and spans[-1][1] == ga.layout.total_bins
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))          # the ranges tile the genome
            if sharding == "chromosomes":
                assert {lo, hi} <= set(int(x) for x in ga.layout.chrom_bin_off)
            assert ga._local.transfer is not None and ga.sum() == hb.mapped          # sum() is the whole file's
            held = torch.tensor([len(ga._local)])
            dist.all_reduce(held)
            assert int(held) >= len(hb)
            assert (ga._global_hist == meta_length_hist(hb.meta)).all()
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    port = _free_port()
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}


def test_sharding_plans_single_process():
    from plastid_b200 import dist as pd
    chroms, lens, ann, hb, spliced = _world()
    assert [pd.read_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    total = 0
    for r in range(3):
        s = pd.shard_reads(spliced, r, 3)
        total += len(s)
        assert s.chrom_read_off[-1] == len(s) and (np.diff(s.chrom_read_off) >= 0).all()
    assert total == len(spliced)
    owned = pd.assign_chromosomes(hb, 8)          # more ranks than chromosomes: some ranks idle
    assert sum(len(x) for x in owned) == len(chroms) and sum(1 for x in owned if not x) == 3
    import plastid_b200 as pb
    from plastid_b200.regions import ChainTable
    lay = pb.GenomeLayout(chroms, lens)
    for W in (1, 3, 8):
        cuts = pd.position_cuts(hb, lay, W)
        assert len(cuts) == W + 1 and cuts[0] == 0 and cuts[-1] == lay.total_bins and (cuts % 16384 == 0).all()
        covered = 0
        for r in range(W):
            sub, lo, hi = pd.shard_positions(spliced, lay, r, W, cuts)
            sub.check_sorted()
            c_of = np.searchsorted(sub.chrom_read_off, np.arange(len(sub)), side="right") - 1
            g = lay.chrom_bin_off[c_of] + sub.ref_start
            assert (g < hi).all() and (g >= lo - spliced.max_span).all()
            covered += int(((g >= lo) & (g < hi)).sum())
        assert covered == len(spliced)            # every read is "owned" by exactly one rank
    chains = ann.chains()
    masks = synth_masks = __import__("plastid_b200").synth.make_masks(ann, frac=0.3, seed=1)
    for ch, m in zip(chains, masks):
        if m:
            ch.add_masks(*m)
    table = ChainTable.from_chains(chains, lay)
    cuts = pd.position_cuts(hb, lay, 4)
    parts = [pd.clip_table(table, int(cuts[r]), int(cuts[r + 1])) for r in range(4)]
    assert (sum(p.chain_len for p in parts) == table.chain_len).all()
    bits = np.unpackbits(table.mask_bits, bitorder="little")
    n_masked = sum(int(np.unpackbits(p.mask_bits, bitorder="little")[:int(p.chain_len.sum())].sum()) for p in parts)
    assert n_masked == int(bits[:int(table.chain_len.sum())].sum()) > 0
    empty = pd.shard_chromosomes(hb, [])
    assert len(empty) == 0 and empty.chroms == []
    assert pd.world() == (0, 1)


def _exchange_worker(rank, world_size, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from plastid_b200 import dist as pd
        width = 11
        k = [5, 0, 7][rank % 3]
        rows = torch.arange(k * width, dtype=torch.float64).reshape(k, width) + 1000.0 * rank
        part, bounds = pd.exchange_column_slices(rows)
        every = [torch.arange([5, 0, 7][r % 3] * width, dtype=torch.float64).reshape(-1, width) + 1000.0 * r for r in range(world_size)]
        want = torch.cat(every, dim=0)[:, bounds[rank]:bounds[rank + 1]]
        assert bounds[0] == 0 and bounds[-1] == width and torch.equal(part, want)
        ranges = pd.all_ranges(100 * rank, 100 * (rank + 1))
        assert ranges.tolist() == [[100 * r, 100 * (r + 1)] for r in range(world_size)]
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_column_slice_exchange_under_gloo(world_size):
    """The exchange behind multi-GPU exact medians (dist.exchange_column_slices; gloo takes the all-gather route, NCCL
    an all-to-all): every rank ends with its column slice of the rows of ALL ranks, in rank order, uneven row counts
    and an empty rank included."""
    port = _free_port()
    results = mp.Manager().dict()
    mp.spawn(_exchange_worker, args=(world_size, port, results), nprocs=world_size, join=True)
    assert [results[r] for r in range(world_size)] == ["ok"] * world_size


def test_row_owners_of_a_window_table():
    """dist.row_owners: the owner of a row is the rank holding its first position on the genome; rows with positions
    on both sides of a cut are listed as straddlers; chains without a position on the genome belong to rank 0."""
    import plastid_b200 as pb
    from plastid_b200 import dist as pd
    from plastid_b200.regions import ChainTable
    layout = pb.GenomeLayout(["c1", "c2"], [40000, 30000])              # c2 starts at global bin 49152
    S = pb.GenomicSegment
    chains = [pb.SegmentChain(S("c1", 100, 400, "+")),                              # rank 0
              pb.SegmentChain(S("c1", 16000, 16300, "+"), S("c1", 16500, 16700, "+")),   # across the cut at 16384: straddler, owner 0
              pb.SegmentChain(S("c1", 20000, 20100, "-")),                          # rank 1
              pb.SegmentChain(S("c2", 10, 90, "+")),                                # rank 2 (bins 49162..)
              pb.SegmentChain(S("c1", -50, 30, "+")),                               # starts before the chromosome: first real position on rank 0
              pb.SegmentChain(),                                                     # no position
              pb.SegmentChain(S("nope", 5, 50, "+"))]                               # unknown chromosome
    t = ChainTable.from_chains(chains, layout)
    ranges = np.array([[0, 16384], [16384, 49152], [49152, layout.total_bins]])
    owner, strad = pd.row_owners(t, ranges)
    assert owner.tolist() == [0, 0, 1, 2, 0, 0, 0] and strad.tolist() == [1]
    owner1, strad1 = pd.row_owners(t, np.array([[0, layout.total_bins]]))
    assert not owner1.any() and len(strad1) == 0

