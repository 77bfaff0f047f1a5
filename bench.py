#!/usr/bin/env python
"""Benchmark of the read-to-coverage hot path (BASELINE.json metric: mapped reads/sec + region
counts/sec, % of the HBM roofline, reference CPU path timed beside it).

Workload (BASELINE.json configs[1], the metric's quoted configuration): VariableFivePrimeMapFactory
with per-read-length P-site offsets + size filter 25-100 over 200 M synthetic 25-35 nt ribo-seq reads
on a human-scale genome (hg38 chromosome lengths, 3.09 Gb) into dense '+' and '-' count vectors,
followed by masked-free region counts over 60 k CDS-like SegmentChains.  One "step" = that whole
pass over one batch.  At N > 1 every rank owns one such read shard (read-range sharding, weak
scaling) and the per-region count tables are summed with one NCCL all-reduce.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mapped_reads_per_sec"
UNIT = "reads/s"


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL (its version banner), torchrun children and other
    libraries write to file descriptor 1 behind Python's back, so descriptor 1 is pointed at stderr for the
    whole run and the JSON line alone is written to the original stdout (``emit``)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(text):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=200_000_000, help="reads per GPU")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="fraction of hg38 chromosome lengths")
    ap.add_argument("--regions", type=int, default=60_000)
    ap.add_argument("--e2e-format", default="delta3", choices=["delta3", "delta8", "wire16"],
                    help="host transfer format of the end-to-end leg (unspliced batches)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="upload chunks overlapped with mapping in the e2e leg")
    ap.add_argument("--e2e-weights", default="1,1,1,1,1,1,1,1",
                    help="relative read counts of the upload chunks of the e2e leg (delta3 / delta8), comma-separated")
    ap.add_argument("--e2e-sweep", default=None,
                    help="measurement aid: further chunk schedules (';'-separated weight lists) timed after the e2e leg, "
                         "one JSON line each on stderr")
    ap.add_argument("--sharding", default="reads", choices=["reads", "positions"],
                    help="multi-GPU mode: 'reads' (default, weak scaling: every GPU maps its own batch over the whole genome) "
                         "or 'positions' (strong scaling of ONE batch: every GPU owns a contiguous bin range, SURVEY 8e)")
    ap.add_argument("--cpu-sample-chroms", type=int, default=1, help="chromosomes in the cpu_baseline sample")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c2p", "c3", "c4", "c5", "peaks"],
                    help="BASELINE.json config: c2 is the metric's quoted configuration (default); the others "
                         "are side measurements (c1 counts_in_region yeast-scale, c3 CenterMapFactory(12) on "
                         "spliced 100-nt reads, c2p the psite pass over 60 k start windows x 11 read lengths of the c2 reads, "
                         "c4 metagene count over 60 k windows of the c2 planes, "
                         "c5 ThreePrimeMapFactory 500 M reads)")
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPU cores NVML reports as local to its GPU before any pinned host buffer is
    allocated, so that the e2e upload reads host memory of the GPU's own NUMA node (first touch).
    Returns a short description for the JSON line; silently a no-op where NVML or affinity is missing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + bit for w, word in enumerate(words) for bit in range(64) if (word >> bit) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to GPU %d" % (len(cpus), idx)
    except Exception as exc:      # measurement nicety only
        return "unbound (%s)" % type(exc).__name__
    return "unbound"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_world(args, rank, device):
    """Synthetic genome, annotation, device-resident read batch and mapping rule of the workload."""
    import torch
    import plastid_b200 as pb
    from plastid_b200 import synth
    wl = args.workload
    if wl in ("c4", "c2p"):
        wl = "c2"           # same reads, genome and mapping rule; the timed step is the metagene pass
    if wl == "c1":
        chroms, lens = synth.yeast_like_genome()
        n_reads = 2_000_000 if args.reads == 200_000_000 else args.reads
        ann = synth.make_annotation(chroms, lens, 6000, seed=0, exons=(1, 2), exon_len=(300, 700), intron_len=(80, 200))
        fac, sf, oracle_kw = pb.FivePrimeMapFactory(14), pb.SizeFilterFactory(25, 100), dict(rule="fiveprime", offset=14)
        name = "C1: counts_in_region, FivePrimeMapFactory(14)+size filter 25-100, yeast-scale 12 Mb genome"
    else:
        chroms, lens = synth.human_like_genome(args.genome_scale)
        n_tx = args.regions if wl != "c5" else 20000
        ann = synth.make_annotation(chroms, lens, n_tx, seed=0, exons=(1, 3), exon_len=(150, 600), intron_len=(100, 3000))
        if wl == "c2":
            n_reads = args.reads
            fac, sf = pb.VariableFivePrimeMapFactory(synth.RIBO_OFFSETS), pb.SizeFilterFactory(25, 100)
            oracle_kw = dict(rule="variable", luts=(fac.forward_offsets, fac.reverse_offsets))
            name = "C2: VariableFivePrimeMapFactory(p-site offsets)+size filter 25-100"
        elif wl == "c3":
            n_reads = 100_000_000 if args.reads == 200_000_000 else args.reads
            fac, sf, oracle_kw = pb.CenterMapFactory(12), None, dict(nibble=12)
            name = "C3: CenterMapFactory(nibble=12), 100-nt reads, 30% one N gap, 3% two"
        else:
            n_reads = 500_000_000 if args.reads == 200_000_000 else args.reads
            fac, sf, oracle_kw = pb.ThreePrimeMapFactory(0), pb.SizeFilterFactory(25, 100), dict(rule="threeprime", offset=0)
            name = "C5: cs count, ThreePrimeMapFactory(0)+size filter 25-100"
    layout = pb.GenomeLayout(chroms, lens)
    table = synth.annotation_table(ann, layout)
    seed = 100 if getattr(args, "sharding", "reads") == "positions" else 100 + rank     # positions: ONE batch, all ranks
    if wl == "c3":
        dbatch = synth.rnaseq_reads(chroms, lens, n_reads, seed=seed, device=device)
    else:
        dbatch = synth.riboseq_reads(ann, n_reads, seed=seed, device=device, frac_in=0.85 if wl != "c1" else 0.9)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return dict(chroms=chroms, lens=lens, ann=ann, layout=layout, table=table, dbatch=dbatch, fac=fac, sf=sf,
                oracle_kw=oracle_kw, name=name, center=(wl == "c3"))


def cpu_reference_pass(hb, lens, table, layout, oracle_kw, size_filter, chrom_ids, threads):
    """The oracle's restatement of the reference path on `chrom_ids`: per chromosome x strand the
    per-read loop of VariableFivePrimeMapFactory.__call__ over the whole chromosome segment, then the
    SegmentChain sums of the regions on those chromosomes.  Returns (seconds, reads, regions)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import coracle
    coracle.lib()

    def one(c):
        n_reg = 0
        for pidx, strand in enumerate(("+", "-")):
            vec = coracle.genome_vector(hb, c, strand, size_filter=size_filter, **oracle_kw)[0]
            base = int(layout.chrom_bin_off[c])
            sel = np.nonzero((table.chain_plane == pidx) & (table.bstart[table.chain_off[:-1]] >= base)
                             & (table.bstart[table.chain_off[:-1]] < int(layout.chrom_bin_off[c + 1])))[0]
            if len(sel):
                offs = [0]
                bs, be = [], []
                for i in sel:
                    a, b = table.chain_off[i], table.chain_off[i + 1]
                    bs.append(table.bstart[a:b] - base)
                    be.append(table.bend[a:b] - base)
                    offs.append(offs[-1] + b - a)
                coracle.region_sums(vec if vec.dtype == np.float64 else vec.astype(np.uint32),
                                    np.concatenate(bs), np.concatenate(be), offs)
            n_reg += len(sel)
        return int(hb.chrom_read_off[c + 1] - hb.chrom_read_off[c]), n_reg

    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(one, chrom_ids))
    else:
        res = [one(c) for c in chrom_ids]
    dt = time.perf_counter() - t0
    return dt, sum(r[0] for r in res), sum(r[1] for r in res)


def host_sample(dbatch, chroms, lens, chrom_ids):
    """Host copy of the reads of the chosen chromosomes only (bounded CPU sample)."""
    from plastid_b200 import dist as pdist
    from plastid_b200.synth import device_batch_to_host
    if len(chrom_ids) == len(chroms) or dbatch.blk_off is not None:
        hb = device_batch_to_host(dbatch, chroms, lens)
        if len(chrom_ids) == len(chroms):
            return hb
        sub = pdist.shard_chromosomes(hb, sorted(chrom_ids))
    else:
        from plastid_b200.batch import AlignmentBatch
        off = dbatch.chrom_read_off.cpu().numpy()
        starts, metas, new_off = [], [], [0]
        for c in sorted(chrom_ids):
            a, b = int(off[c]), int(off[c + 1])
            starts.append(dbatch.ref_start[a:b].cpu().numpy())
            metas.append(dbatch.meta[a:b].cpu().numpy().view(np.uint32))
            new_off.append(new_off[-1] + b - a)
        sub = AlignmentBatch([chroms[c] for c in sorted(chrom_ids)], np.asarray(lens)[sorted(chrom_ids)],
                             np.concatenate(starts), np.concatenate(metas), new_off, max_span=dbatch.max_span)
    # re-expand to the full chromosome list so chromosome indices keep their meaning
    full_off = np.zeros(len(chroms) + 1, dtype=np.int64)
    pos = 0
    for c in range(len(chroms)):
        if c in chrom_ids:
            j = sorted(chrom_ids).index(c)
            pos += int(sub.chrom_read_off[j + 1] - sub.chrom_read_off[j])
        full_off[c + 1] = pos
    from plastid_b200.batch import AlignmentBatch
    return AlignmentBatch(chroms, lens, sub.ref_start, sub.meta, full_off, sub.blk_off, sub.blk, max_span=sub.max_span)


def python_loop_estimate(hb, chrom_id, workload, n_sample=100_000):
    """"Reference as shipped" estimate: the per-read Python/Cython loop of the mapping rule
    restated in pure Python (oracle/pyoracle.py), timed on the first `n_sample` reads of one chromosome,
    read objects prebuilt (BAM decoding excluded).  One host core."""
    from oracle import pyoracle as po
    from plastid_b200 import synth
    a = int(hb.chrom_read_off[chrom_id])
    b = min(int(hb.chrom_read_off[chrom_id + 1]), a + n_sample)
    if b <= a:
        return None
    if hb.blk_off is None:
        reads = [po.Read(int(s), [(po.CMATCH, int(m & 0xFFFF))], bool((m >> 16) & 1))
                 for s, m in zip(hb.ref_start[a:b].tolist(), hb.meta[a:b].tolist())]
    else:
        reads = []
        for i in range(a, b):
            pos = hb.positions_of(i)
            ops, prev = [], None
            for p in pos:       # rebuild a CIGAR of M runs and N gaps from the aligned positions
                if prev is not None and p != prev + 1:
                    ops.append((po.CREF_SKIP, p - prev - 1))
                if ops and ops[-1][0] == po.CMATCH and (prev is None or p == prev + 1):
                    ops[-1] = (po.CMATCH, ops[-1][1] + 1)
                else:
                    ops.append((po.CMATCH, 1))
                prev = p
            reads.append(po.Read(int(hb.ref_start[i]), ops, bool((hb.meta[i] >> 16) & 1)))
    fn = {"c1": lambda: po.FivePrimeMap(14), "c2": lambda: po.VariableFivePrimeMap(dict(synth.RIBO_OFFSETS)),
          "c3": lambda: po.CenterMap(12), "c5": lambda: po.ThreePrimeMap(0)}[workload]()
    seg_end = max(r.reference_end for r in reads) + 1
    t0 = time.perf_counter()
    for strand in ("+", "-"):      # get_reads_and_counts: strand filter, then the rule (genome_array.py:811-823)
        mine = [r for r in reads if r.is_reverse is (strand == "-")]
        fn(mine, po.Seg(hb.chroms[chrom_id], int(reads[0].reference_start), seg_end, strand))
    dt = time.perf_counter() - t0
    return {"value": len(reads) / dt, "unit": UNIT, "cores": 1,
            "sample": "%d reads of %s through the pure-Python restatement of the rule's per-read loop in %.2f s"
                      % (len(reads), hb.chroms[chrom_id], dt)}


def run_peaks(args, device):
    """the L2 atomic peak next to hbm_gbs — 2^30 `red.global.add.u32` updates, (i) uniformly
    random over 24.8 GB of bins, (ii) coordinate-sorted with +-64 nt jitter over a 3.1 G-bin plane.  One JSON
    line; also written to gpurun_out/atomic_peaks.json (committed copy: profiles/atomic_peaks_r01.json)."""
    import torch
    from plastid_b200 import _lib
    L = _lib.lib()
    n_upd = 1 << 30
    out = {"metric": "atomic_updates_per_sec", "unit": "updates/s", "n_updates": n_upd, "op": "red.global.add.u32"}
    for name, mode, n_bins, jitter in (("uniform_24.8GB", 0, 6_200_000_000, 0), ("sorted_jitter64_3.1Gbins", 1, 3_100_000_000, 64)):
        bins = torch.zeros(n_bins, dtype=torch.int32, device=device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for it in range(5):
            ev0.record()
            _lib.check(L.pb_atomic_probe(_lib.ptr(bins), n_bins, n_upd, mode, jitter, _lib.stream_ptr()))
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            best = ms if best is None or (it > 0 and ms < best) else best
        total = int(bins.to(torch.int64).sum().item()) if n_bins <= 3_100_000_000 else None
        out[name] = {"ms": best, "updates_per_sec": n_upd / (best / 1000.0), "n_bins": n_bins, "jitter": jitter,
                     "updates_landed_check": total}
        del bins
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "atomic_peaks.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    emit(json.dumps(out))


def run_position_sharded(args, W, device, rank, world, dist):
    """Strong scaling of one C2 batch by position ranges: every rank builds the SAME batch,
    keeps the reads that start in its bin range (+ halo), allocates range-only planes, maps its range
    with pb_map_point_range and sums the clipped region table; one all-reduce per step."""
    import torch
    from plastid_b200 import dist as pdist
    from plastid_b200 import synth
    from plastid_b200.batch import DeviceBatch
    from plastid_b200.genome_array import map_batch, region_sums, CountPlanes, length_histogram
    from plastid_b200.map_factories import CenterMapFactory
    layout, table, fac, sf, dbatch = W["layout"], W["table"], W["fac"], W["sf"], W["dbatch"]
    n_total = dbatch.n_reads
    is_center = isinstance(fac, CenterMapFactory)
    # Center rule: every rank derives its slot tables from the histogram of the WHOLE batch (in a real run: one
    # 512 KB all-reduce), so that the sharded planes equal the unsharded ones bit for bit
    hist = length_histogram(dbatch, fac, sf) if is_center else None
    if dbatch.blk_off is None:
        sub, lo, hi, cuts = pdist.shard_positions_device(dbatch, layout, rank, world)
    else:                                       # spliced batches are sharded on the host (set-up, not timed)
        hb = synth.device_batch_to_host(dbatch, W["chroms"], W["lens"])
        h_sub, lo, hi = pdist.shard_positions(hb, layout, rank, world)
        sub = DeviceBatch.from_host(h_sub, device)
        del hb, h_sub
    del dbatch
    W["dbatch"] = None
    torch.cuda.empty_cache()
    clipped = pdist.clip_table(table, lo, hi)
    clipped.device(device)
    planes = CountPlanes(layout, "f64" if is_center else "u32", device, bin_range=(lo, hi))
    planes.alloc(("+", "-"))

    def step():
        map_batch(sub, layout, fac, sf, strands=("+", "-"), planes=planes, sync_stats=False, bin_range=(lo, hi),
                  length_hist=hist)
        sums, live = region_sums(planes, clipped)
        if world > 1:
            dist.all_reduce(sums)
        return sums

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        sums = step()
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=device)
    share = torch.tensor([float(sub.n_reads), float(hi - lo)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(share) for _ in range(world)]
        dist.all_gather(gathered, share)
    else:
        gathered = [share]
    if rank != 0:
        return
    ms = float(t.item())
    line = {"metric": METRIC, "value": n_total / (ms / 1000.0), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if is_center else "u32", "data": "synthetic",
            "config": {"workload": "%s, ONE batch of %d synthetic reads over %d bins sharded by position range over %d GPUs "
                                   "(range-only planes, halo reads, clipped region tables all-reduced)"
                                   % (W["name"], n_total, layout.total_bins, world),
                       "reads_per_rank_incl_halo": [int(g[0].item()) for g in gathered],
                       "bins_per_rank": [int(g[1].item()) for g in gathered]},
            "region_counts_per_sec": W["ann"].n_tx / (ms / 1000.0), "table_checksum": float(sums.sum().item())}
    emit(json.dumps(line))


def run_c4(args, W, device, rank, world, dist):
    """BASELINE config 4: metagene count over 60 k windows (-50/+300 nt, 5 % masked) of the C2 count
    planes: gather the window matrix, (N > 1: all-reduce it, counts are linear in the read shards),
    normalise, exact per-column median.  One step = that pass; the mapping itself is not timed."""
    import torch
    from plastid_b200 import synth
    from plastid_b200.genome_array import map_batch, gather_windows, window_normalize, column_profile
    layout, ann, dbatch = W["layout"], W["ann"], W["dbatch"]
    planes = map_batch(dbatch, layout, W["fac"], W["sf"], strands=("+", "-"))
    table, cols = synth.window_table(ann, layout, width=350)
    table.device(device)
    width, n = 350, table.n_chains
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    gather_ms = []

    def step(timed=False):
        if timed:
            ev[2].record()
        mat, mmask = gather_windows(planes, table, cols, width)
        if timed:
            ev[3].record()
        if world > 1:
            mat = torch.nan_to_num(mat, nan=0.0)
            dist.all_reduce(mat)
        denom, sel, norm, nmask = window_normalize(mat, mmask, 70, 100, 10)
        prof, nreg, csum = column_profile(norm, nmask, sel, "median")
        return prof, nreg

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.steps):
        prof, nreg = step(timed=True)
        torch.cuda.synchronize()
        gather_ms.append(ev[2].elapsed_time(ev[3]))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    ms = float(t.item())
    # `metagene generate` geometry at the same scale: landmark windows of 3 isoforms per gene and
    # the maximal spanning window of every gene (count + fill launches), tables resident on the device
    from plastid_b200.windows import landmark_windows, spanning_windows
    iso_table, grp_off, grp_tx = synth.isoform_table(ann, layout, n_iso=3)
    iso_table.device(device)
    gen_ms = []
    for i in range(3 + args.steps):
        torch.cuda.synchronize()
        ev[2].record()
        win, flags = landmark_windows(iso_table, 50, 300, device)
        ev[3].record()
        t0 = time.perf_counter()
        res = spanning_windows(iso_table, win, flags, grp_off, grp_tx, 50, 300, device)
        wall = (time.perf_counter() - t0) * 1e3
        if i >= 3:
            gen_ms.append((ev[2].elapsed_time(ev[3]), wall))
    generate = {"genes": int(len(grp_off) - 1), "transcripts": int(iso_table.n_tx), "window": 350,
                "landmark_kernel_ms": float(np.mean([a for a, _ in gen_ms])),
                "spanning_windows_wall_ms": float(np.mean([b for _, b in gen_ms])),
                "windows_found": int((res["status"] == 1).sum()), "window_blocks": int(res["n_blk"].sum()),
                "note": "wall time of count + scan + fill launches incl. the device->host copy of the result tables"}
    positions = int(table.chain_len.sum())
    alg = 4.0 * positions + positions / 8.0 + 16.0 * len(table.bstart) + 9.0 * n * width
    g_ms = float(np.mean(gather_ms))
    peak, peak_src = peaks()
    line = {"metric": "metagene_windows_per_sec", "value": n / (ms / 1000.0), "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: metagene count, %d windows x %d nt over the C2 planes, 5%% masked, "
                                   "norm window [20,50) from the landmark, min_counts 10, exact median profile" % (n, width),
                       "reads_per_gpu": dbatch.n_reads, "windows": n, "width": width},
            "roofline": {"bound": "hbm", "kernel": "pb_gather_windows_kernel", "achieved": alg / (g_ms / 1000.0) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (g_ms / 1000.0) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": g_ms,
                         "kernel_share_of_step": g_ms / ms},
            "generate": generate,
            "profile_checksum": float(torch.nan_to_num(prof).sum().item()), "regions_counted_max": int(nreg.max().item())}
    emit(json.dumps(line))


def run_c2p(args, W, device, rank, world, dist):
    """The psite pass of BASELINE config 2: 5' ends (offset 0, no size filter — psite.py:357-383) of the
    C2 reads counted per read length 25..35 over 60 k start-codon windows (350 nt) in ONE launch
    (pb_stratified_windows), then per length normalise + exact median profile."""
    import torch
    import plastid_b200 as pb
    from plastid_b200 import synth
    from plastid_b200.genome_array import stratified_windows, count_profiles
    layout, ann, dbatch = W["layout"], W["ann"], W["dbatch"]
    table, cols = synth.window_table(ann, layout, width=350)
    table.device(device)
    fac = pb.FivePrimeMapFactory(0)
    width, n, lo, hi = 350, table.n_chains, 25, 35
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    k_ms = []

    def step(timed=False):
        if timed:
            ev[2].record()
        strat, maskmat = stratified_windows(dbatch, layout, fac, None, table, cols, width, lo, hi)
        if timed:
            ev[3].record()
        if world > 1:
            dist.all_reduce(strat)
        # normalisation fused with key extraction, then one (length, column) median launch
        return count_profiles(strat, maskmat, 70, 100, 10, "median")[0]

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.steps):
        profs = step(timed=True)
        torch.cuda.synchronize()
        k_ms.append(ev[2].elapsed_time(ev[3]))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    ms = float(t.item())
    line = {"metric": "psite_window_profiles_per_sec", "value": n * (hi - lo + 1) / (ms / 1000.0), "unit": "window-lengths/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2 psite pass: FivePrimeMapFactory(0), %d reads/GPU, %d windows x %d nt x lengths %d-%d, "
                                   "median profiles" % (dbatch.n_reads, n, width, lo, hi)},
            "stratified_kernel_ms": float(np.mean(k_ms)), "kernel_share_of_step": float(np.mean(k_ms)) / ms,
            "profile_checksum": float(torch.nan_to_num(profs).sum().item())}
    emit(json.dumps(line))


def main():
    args = parse_args()
    claim_stdout()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    numa = bind_to_gpu_numa_node(local_rank) if args.impl != "reference" else "all host threads"
    if world > 1 and args.impl != "reference":
        dist.init_process_group("nccl", device_id=torch.device(device))

    import plastid_b200 as pb
    from plastid_b200 import synth, _lib
    from plastid_b200.genome_array import map_batch, region_sums, CountPlanes

    if args.workload == "peaks":
        if rank == 0:
            run_peaks(args, device)
        return 0
    W = build_world(args, rank, device)
    chroms, lens, ann, layout, table, dbatch = W["chroms"], W["lens"], W["ann"], W["layout"], W["table"], W["dbatch"]
    fac, sf, is_center = W["fac"], W["sf"], W["center"]
    sf_tuple = None if sf is None else (sf.min_, sf.max_)
    n_reads = dbatch.n_reads
    workload = ("%s, %d synthetic reads/GPU, %d chromosomes (%d bins, '+' and '-' planes), %d region counts"
                % (W["name"], n_reads, len(chroms), layout.total_bins, ann.n_tx))
    config = {"workload": workload, "reads_per_gpu": n_reads, "genome_bins": int(layout.total_bins),
              "regions": ann.n_tx, "host_affinity": numa, "sharding": "read-range per GPU; NCCL all-reduce of region tables" if world > 1
              else "single GPU", "l2": "inputs (%.2f GB) and outputs (%.1f GB) larger than the 126 MB L2"
              % (8 * n_reads / 1e9, (16 if is_center else 8) * layout.total_bins / 1e9)}

    if args.sharding == "positions" and args.impl != "reference":
        if args.workload not in ("c2", "c3", "c5"):
            raise SystemExit("--sharding positions is implemented for the mapping workloads c2, c3 and c5")
        run_position_sharded(args, W, device, rank, world, dist)
        if world > 1:
            dist.destroy_process_group()
        return 0
    if args.workload == "c2p" and args.impl != "reference":
        run_c2p(args, W, device, rank, world, dist)
        if world > 1:
            dist.destroy_process_group()
        return 0
    if args.workload == "c4" and args.impl != "reference":
        run_c4(args, W, device, rank, world, dist)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        threads = os.cpu_count() or 1
        # every chromosome of the workload, longest first, one chromosome per task over all host threads
        chrom_ids = [int(c) for c in np.argsort(-np.asarray(lens))]
        hb = host_sample(dbatch, chroms, lens, set(chrom_ids))
        del dbatch
        torch.cuda.empty_cache()
        times = []
        for it in range(args.warmup + args.steps):
            dt, nr, nreg = cpu_reference_pass(hb, lens, table, layout, W["oracle_kw"], sf_tuple, chrom_ids, threads)
            if it >= args.warmup:
                times.append(dt)
        ms = 1000.0 * float(np.mean(times))
        val = nr / (ms / 1000.0)
        sample = "the whole workload per step: %d chromosomes, %d reads, %d regions" % (len(chrom_ids), nr, nreg)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64" if is_center else "int64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "oracle port of map_factories.pyx/roitools.pyx loops (the Cython reference cannot be "
                        "built here: pysam absent); chromosome-parallel over %d host threads" % threads}
        emit(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ device-resident steps
    planes = CountPlanes(layout, "f64" if is_center else "u32", device)
    planes.alloc(("+", "-"))
    table.device(device)
    L = _lib.lib()

    def step():
        map_batch(dbatch, layout, fac, sf, strands=("+", "-"), planes=planes, sync_stats=False)
        sums, live = region_sums(planes, table)
        if world > 1:
            dist.all_reduce(sums)
        return sums, live

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs ~1 s to come up: start before the warm-up
    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    L.pb_enable_kernel_timing(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        sums, live = step()
    ev1.record()
    fence()
    total_ms = ev0.elapsed_time(ev1)
    kms, kn = C.c_float(0), C.c_int(0)
    _lib.check(L.pb_tiles_kernel_ms_total(C.byref(kms), C.byref(kn)))
    L.pb_enable_kernel_timing(0)
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_reads * world / (ms_per_step / 1000.0)
    region_rate = ann.n_tx * world / (ms_per_step / 1000.0)

    # ------------------------------------------------------------------ launch-bound workloads: CUDA graph replay
    graph_ms = None
    if args.workload == "c1" and world == 1:
        from plastid_b200.genome_array import GraphedCount
        gc_ = GraphedCount(dbatch, layout, fac, sf, table, strands=("+", "-"))
        for _ in range(3):
            gc_.replay()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps * 10):
            g_sums, g_live = gc_.replay()
        ev1.record()
        torch.cuda.synchronize()
        graph_ms = ev0.elapsed_time(ev1) / (args.steps * 10)
        assert torch.equal(g_sums, sums) and torch.equal(g_live, live), "graph replay differs from the eager pass"

    # ------------------------------------------------------------------ end-to-end (host buffers)
    # The caller holds the batch in pinned host memory in the transfer format the host decoder
    # emits: wire16 (4 B/read) for unspliced batches, the plain SoA otherwise.  Every step copies it
    # to the device, expands it, runs the same kernels and reads the region table back.
    from plastid_b200.batch import Wire16Batch, Wire16Receiver, Delta8Batch, Delta8Receiver, Delta3Batch, Delta3Receiver
    h_sums = torch.empty(ann.n_tx, dtype=torch.float64).pin_memory()
    h_live = torch.empty(ann.n_tx, dtype=torch.int64).pin_memory()
    d2h = h_sums.numel() * 8 + h_live.numel() * 8
    use_wire16 = dbatch.blk_off is None
    if use_wire16:
        WireBatch, WireReceiver = {"delta3": (Delta3Batch, Delta3Receiver), "delta8": (Delta8Batch, Delta8Receiver),
                                   "wire16": (Wire16Batch, Wire16Receiver)}[args.e2e_format]
        wire = WireBatch.from_batch(synth.device_batch_to_host(dbatch, chroms, lens))
        pinned = wire.pinned()
        receiver = WireReceiver(wire, device)
        h2d = wire.nbytes
        # small batches are launch-bound: fewer, larger chunks (about 8 M reads each at least)
        n_chunks = max(1, min(args.e2e_chunks, n_reads // 8_000_000))
        weights = [float(x) for x in args.e2e_weights.split(",")]
        if args.e2e_format == "wire16" or n_chunks < args.e2e_chunks or len(weights) < 2:
            chunks = WireReceiver.plan_chunks(wire, layout, n_chunks)
        else:
            chunks = WireReceiver.plan_chunks(wire, layout, len(weights), weights)
        copy_stream = torch.cuda.Stream(device=device)
        from plastid_b200.genome_array import map_wire16_streamed

        def e2e_step():
            # upload in chunks on a copy stream; each chunk's bins are mapped as soon as it has landed
            map_wire16_streamed(receiver, pinned, chunks, layout, fac, sf, ("+", "-"), planes, copy_stream)
            s, l = region_sums(planes, table)
            if world > 1:
                dist.all_reduce(s)
            h_sums.copy_(s, non_blocking=True)
            h_live.copy_(l, non_blocking=True)
            torch.cuda.synchronize()      # the caller holds the table before the next batch starts
    elif args.e2e_format == "delta3":
        # batches with multi-block reads: delta3 streams for (ref_start, meta) + one 4-byte block word per aligned
        # block; blk_off is rebuilt on the device (pb_unpack_blocks).  The Center / binned path needs the whole
        # batch before it can start, so the upload is not chunked.
        from plastid_b200.batch import Delta3SplicedBatch, Delta3SplicedReceiver
        swire = Delta3SplicedBatch.from_batch(synth.device_batch_to_host(dbatch, chroms, lens))
        spinned = swire.pinned()
        sreceiver = Delta3SplicedReceiver(swire, device)
        h2d = swire.nbytes
        resident = dbatch

        s_chunks = Delta3SplicedReceiver.plan_chunks(swire, layout, max(1, min(args.e2e_chunks, n_reads // 8_000_000)))
        s_copy = torch.cuda.Stream(device=device)
        from plastid_b200.genome_array import map_center_streamed

        def e2e_step():
            nonlocal dbatch
            if is_center:
                # chunked upload; every chunk's final bin range is mapped (pb_map_center_range) while later chunks land
                map_center_streamed(sreceiver, spinned, s_chunks, layout, fac, sf, ("+", "-"), planes, s_copy)
                s, l = region_sums(planes, table)
                if world > 1:
                    dist.all_reduce(s)
            else:
                dbatch = sreceiver.receive(spinned)
                s, l = step()
                dbatch = resident
            h_sums.copy_(s, non_blocking=True)
            h_live.copy_(l, non_blocking=True)
            torch.cuda.synchronize()
    else:
        h_start = torch.empty(n_reads, dtype=torch.int32).pin_memory()
        h_meta = torch.empty(n_reads, dtype=torch.int32).pin_memory()
        h_start.copy_(dbatch.ref_start)
        h_meta.copy_(dbatch.meta)
        h_blk_off = torch.empty_like(dbatch.blk_off, device="cpu").pin_memory()
        h_blk = torch.empty_like(dbatch.blk, device="cpu").pin_memory()
        h_blk_off.copy_(dbatch.blk_off)
        h_blk.copy_(dbatch.blk)
        h2d = (h_start.numel() + h_meta.numel() + h_blk_off.numel() + h_blk.numel()) * 4

        def e2e_step():
            dbatch.ref_start.copy_(h_start, non_blocking=True)
            dbatch.meta.copy_(h_meta, non_blocking=True)
            dbatch.blk_off.copy_(h_blk_off, non_blocking=True)
            dbatch.blk.copy_(h_blk, non_blocking=True)
            s, l = step()
            h_sums.copy_(s, non_blocking=True)
            h_live.copy_(l, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    fence()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    fence()
    e2e_ms = max(ev0.elapsed_time(ev1), 1000.0 * (time.perf_counter() - t0)) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_reads * world / (float(t.item()) / 1000.0)
    if args.e2e_sweep and use_wire16 and args.e2e_format != "wire16" and world == 1:
        for sched in args.e2e_sweep.split(";"):
            chunks = WireReceiver.plan_chunks(wire, layout, 0, [float(x) for x in sched.split(",")])
            for _ in range(2):
                e2e_step()
            fence()
            t1 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            fence()
            sys.stderr.write(json.dumps({"e2e_schedule": sched, "chunks": len(chunks),
                                         "ms_per_step": 1000.0 * (time.perf_counter() - t1) / e2e_steps}) + "\n")
    clocks = sampler.stop()              # samples cover warm-up, the timed steps and the e2e steps
    table_checksum = float(h_sums.sum().item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ roofline of the tiles kernel
    peak, peak_src = peaks()
    n_blk = 0 if dbatch.blk is None else dbatch.blk.shape[0]
    # SoA in once (+ block table for spliced reads) + every bin of both planes out once
    alg_bytes = 8.0 * n_reads + (4.0 * (n_reads + 1) + 8.0 * n_blk if n_blk else 0.0) \
        + (8.0 if is_center else 4.0) * layout.total_bins * 2
    k_ms = kms.value / max(kn.value, 1)
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    kernel_name = "pb_center_tiles_kernel" if is_center else "pb_point_tiles_kernel"
    traffic = None                    # DRAM bytes per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath) and args.genome_scale == 1.0:
        with open(tpath) as fh:
            traffic = json.load(fh).get("%s:%s" % (kernel_name, args.workload))
        if traffic is not None and abs(n_reads - {"c2": 200_000_000, "c3": 100_000_000}.get(args.workload, -1)) > 0:
            traffic = None            # captured at the default size only
    roofline = {"bound": "hbm", "kernel": "pb_center_tiles_kernel" if is_center else "pb_point_tiles_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "launches_timed": kn.value,
                "kernel_share_of_step": k_ms / ms_per_step}

    # ------------------------------------------------------------------ cpu_baseline (rank 0, N=1)
    cpu = None
    if world == 1:
        order = np.argsort(-np.asarray(lens))
        chrom_ids = sorted(int(c) for c in order[:args.cpu_sample_chroms])
        hb = host_sample(dbatch, chroms, lens, set(chrom_ids))
        dt, nr, nreg = cpu_reference_pass(hb, lens, table, layout, W["oracle_kw"], sf_tuple, chrom_ids, 1)
        cpu = {"value": nr / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%s: %d reads + %d region sums in %.2f s (oracle C port, one thread)"
                         % (",".join(chroms[c] for c in chrom_ids), nr, nreg, dt)}
        try:
            cpu["python_loop"] = python_loop_estimate(hb, chrom_ids[0], args.workload)
        except Exception as exc:      # an estimate beside the baseline, never a reason to lose the line
            cpu["python_loop"] = {"error": repr(exc)}

    binning = ["pb_bin_kernel(count)", "pb_scan_chunks_kernel", "pb_scan_top_kernel", "pb_scan_add_kernel",
               "pb_bin_kernel(fill)"] if dbatch.blk_off is not None else []
    if is_center:
        kernels_per_step = ["pb_length_hist_kernel", "pb_tile_index_kernel"] + binning + ["pb_center_tiles_kernel"]
    else:
        kernels_per_step = ["pb_tile_index_kernel"] + binning + ["pb_point_tiles_kernel", "pb_point_overflow_kernel"]
    kernels_per_step += ["pb_stats_finish_kernel", "pb_region_sums_kernel"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if is_center else "u32", "data": "synthetic",
            "config": config,
            "region_counts_per_sec": region_rate,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t.item()), "steps": e2e_steps,
                    "host_format": ("%s (%.2f B/read), %d-chunk upload overlapped with pb_map_point_range"
                                    % (args.e2e_format, h2d / max(n_reads, 1), len(chunks)))
                    if use_wire16 else (("delta3 + block words (%.2f B/read), %s" % (
                        h2d / max(n_reads, 1), ("%d-chunk upload overlapped with pb_map_center_range" % len(s_chunks)) if is_center
                        else "whole batch uploaded before the binned path starts"))
                        if args.e2e_format == "delta3" else "SoA (8 B/read + blocks)")},
            "gpu_launches": len(kernels_per_step) * args.steps, "kernels_per_step": kernels_per_step,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "table_checksum": table_checksum}
    if graph_ms is not None:
        # the same pass replayed as one CUDA graph launch (launch-latency bound workload)
        line["cuda_graph"] = {"ms_per_step": graph_ms, "value": n_reads / (graph_ms / 1000.0), "unit": UNIT,
                              "region_counts_per_sec": ann.n_tx / (graph_ms / 1000.0), "replays_timed": args.steps * 10}
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
