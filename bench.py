#!/usr/bin/env python
"""Benchmark of the read-to-coverage hot path (BASELINE.json metric: mapped reads/sec + region
counts/sec, % of the HBM roofline, reference CPU path timed beside it).

Workload (BASELINE.json configs[1], the metric's quoted configuration): VariableFivePrimeMapFactory
with per-read-length P-site offsets + size filter 25-100 over 200 M synthetic 25-35 nt ribo-seq reads
on a human-scale genome (hg38 chromosome lengths, 3.09 Gb) into dense '+' and '-' count vectors,
followed by masked-free region counts over 60 k CDS-like SegmentChains.  One "step" = that whole
pass over one batch.  At N > 1 (default `--sharding positions`, strong scaling) ONE such batch is cut into
contiguous position ranges of equal cost (reads streamed + plane bins written): every rank owns the planes of
its range and the reads that can reach it, and the per-region count tables are completed with one NCCL
all-reduce.  `--sharding reads` is round 1's weak-scaling mode (every rank its own batch over the whole genome).

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
    ap.add_argument("--sharding", default="positions", choices=["positions", "chromosomes", "reads"],
                    help="multi-GPU mode: 'positions' (default; strong scaling of ONE batch: every GPU owns a contiguous bin "
                         "range of equal cost = reads + plane bins, SURVEY 8e), 'chromosomes' (the same with cuts on chromosome boundaries, "
                         "BASELINE config 5) or 'reads' (weak scaling: every GPU maps its own batch over the whole genome)")
    ap.add_argument("--pileup", type=int, default=0,
                    help="c3 only: this many of the reads lie in the last 16.5 kb of the last chromosome (a chrM-like pile-up in the last tiles)")
    ap.add_argument("--c4-exchange", default="slices", choices=["slices", "matrix"],
                    help="c4 at N > 1: 'slices' (default) leaves the count matrix on its ranks (rows completed by their owner, "
                         "medians per column slice after one all-to-all); 'matrix' all-reduces the whole 168 MB matrix (round 2a)")
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


def bind_to_gpu_numa_node(local_rank, world=1):
    """Pin this rank to CPU cores NVML reports as local to its GPU before any pinned host buffer is allocated, so that
    the e2e upload reads host memory of the GPU's own NUMA node (first touch).  Several ranks whose GPUs share one
    node (all eight do on this pool's boxes) take DISJOINT slices of its cores, so that their encoder / copy threads
    do not pile onto the same ones.  Returns a short description; silently a no-op where NVML or affinity is missing."""
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
            mine = sorted(cpus)
            if world > 1 and len(mine) >= world:
                per = len(mine) // world
                mine = mine[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, set(mine))
            return "%d cpus (%d-%d) local to GPU %d" % (len(mine), mine[0], mine[-1], idx)
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
            name = "C3: CenterMapFactory(nibble=12), 100-nt reads, 30% one N gap, 3% two" + (
                ", %d of them piled up in the last 16.5 kb of the genome" % args.pileup if getattr(args, "pileup", 0) else "")
        else:
            n_reads = 500_000_000 if args.reads == 200_000_000 else args.reads
            fac, sf, oracle_kw = pb.ThreePrimeMapFactory(0), pb.SizeFilterFactory(25, 100), dict(rule="threeprime", offset=0)
            name = "C5: cs count, ThreePrimeMapFactory(0)+size filter 25-100"
    layout = pb.GenomeLayout(chroms, lens)
    table = synth.annotation_table(ann, layout)
    seed = 100 + rank if getattr(args, "sharding", "positions") == "reads" else 100     # position sharding: ONE batch on all ranks
    if wl == "c3":
        dbatch = synth.rnaseq_reads(chroms, lens, n_reads, seed=seed, device=device, pileup=getattr(args, "pileup", 0))
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
                coracle.region_sums(vec, np.concatenate(bs), np.concatenate(be), offs)     # int64 / float64 as the rule returned it
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
    # gather pattern: scattered segments of the size of an exon block / window out of a 12.4 GB plane
    plane = torch.zeros(3_088_465_920, dtype=torch.int32, device=device)
    out["gather_segments"] = {}
    for chunk, n_chunks in ((256, 480_000), (768, 120_000), (1536, 120_000), (4096, 60_000)):
        res = torch.empty(n_chunks, dtype=torch.int32, device=device)
        best = None
        for it in range(6):
            ev0.record()
            _lib.check(L.pb_gather_probe(_lib.ptr(plane), plane.numel(), chunk, n_chunks, _lib.ptr(res), _lib.stream_ptr()))
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            best = ms if best is None or (it > 0 and ms < best) else best
        nbytes = 4.0 * chunk * n_chunks
        out["gather_segments"]["%d_bins_x_%d" % (chunk, n_chunks)] = {"ms": best, "bytes": nbytes, "GBps": nbytes / (best / 1000.0) / 1e9}
    del plane
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "atomic_peaks.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    emit(json.dumps(out))


def run_c4(args, W, device, rank, world, dist):
    """BASELINE config 4: metagene count over 60 k windows (-50/+300 nt, 5 % masked) of the C2 count
    planes: gather the window matrix, (N > 1: all-reduce it, counts are linear in the read shards),
    normalise, exact per-column median.  One step = that pass; the mapping itself is not timed."""
    import torch
    from plastid_b200 import synth
    from plastid_b200.genome_array import map_batch, gather_windows, window_normalize, column_profile
    from plastid_b200.genome_array import CountPlanes
    layout, ann = W["layout"], W["ann"]
    # N > 1: ONE batch sharded by position range — every rank maps the planes of its own range, fills the window cells
    # of its own positions, and the count matrix (60 k x 350 float64 = 168 MB) is completed with one all-reduce
    dbatch, lo, hi, n_total = shard_world(args, W, device, rank, world)
    ranged = (lo, hi) != (0, int(layout.total_bins))
    planes = CountPlanes(layout, "u32", device, (lo, hi) if ranged else None)
    map_batch(dbatch, layout, W["fac"], W["sf"], strands=("+", "-"), planes=planes, bin_range=(lo, hi) if ranged else None)
    table, cols = synth.window_table(ann, layout, width=350)
    table.device(device)
    cols = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32)).to(device)       # the ROI table's offsets, resident
    width, n = 350, table.n_chains
    from plastid_b200 import dist as pdist
    ranges = pdist.all_ranges(lo, hi, device=device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    gather_ms = []

    def step(timed=False):
        if timed:
            ev[2].record()
        mat, mmask = gather_windows(planes, table, cols, width, touched_only=world > 1 and args.c4_exchange == "slices")
        if timed:
            ev[3].record()
        if world > 1 and args.c4_exchange == "matrix":
            dist.all_reduce(mat)         # cells of other ranks' positions are 0, cells without a position NaN everywhere
        elif world > 1:
            # the matrix stays put: rows completed by their owner, exact medians per column slice after one all-to-all
            prof, nreg, _d, _s = pdist.window_profile(mat, mmask, table, ranges, 70, 100, 10, "median", want_rows=False)
            return prof, nreg
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
            "scaling": "weak" if args.sharding == "reads" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: metagene count, %d windows x %d nt over the C2 planes, 5%% masked, "
                                   "norm window [20,50) from the landmark, min_counts 10, exact median profile" % (n, width),
                       "reads": n_total, "windows": n, "width": width},
            "sharding": "single GPU" if world == 1 else args.sharding,
            "exchange": None if world == 1 else {"slices": "matrix stays on its ranks: straddling rows all-reduced, rows normalised by their "
                                                 "owner, medians per column slice after one all-to-all (dist.window_profile)",
                                                 "matrix": "whole count matrix all-reduced"}[args.c4_exchange],
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
    layout, ann = W["layout"], W["ann"]
    dbatch, b_lo, b_hi, n_total = shard_world(args, W, device, rank, world)     # N > 1: sites are counted by the rank owning them
    table, cols = synth.window_table(ann, layout, width=350)
    table.device(device)
    fac = pb.FivePrimeMapFactory(0)
    width, n, lo, hi = 350, table.n_chains, 25, 35
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    k_ms = []

    def step(timed=False):
        if timed:
            ev[2].record()
        strat, maskmat = stratified_windows(dbatch, layout, fac, None, table, cols, width, lo, hi, bin_range=(b_lo, b_hi))
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
            "higher_is_better": True, "scaling": "weak" if args.sharding == "reads" else "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2 psite pass: FivePrimeMapFactory(0), %d reads, %d windows x %d nt x lengths %d-%d, "
                                   "median profiles" % (n_total, n, width, lo, hi)},
            "sharding": "single GPU" if world == 1 else args.sharding,
            "stratified_kernel_ms": float(np.mean(k_ms)), "kernel_share_of_step": float(np.mean(k_ms)) / ms,
            "profile_checksum": float(torch.nan_to_num(profs).sum().item())}
    emit(json.dumps(line))


def shard_world(args, W, device, rank, world):
    """The reads and bin range of this rank.  ``positions`` (default at N > 1; strong scaling): every rank
    built the SAME batch, keeps the reads that start in its bin range plus a halo, and owns the planes of that range
    only (``chromosomes``: cuts on chromosome boundaries, BASELINE config 5).  ``reads`` (weak scaling): every rank has
    its own batch and maps it over the whole genome."""
    import torch
    from plastid_b200 import dist as pdist
    from plastid_b200 import synth
    from plastid_b200.batch import DeviceBatch
    layout, dbatch = W["layout"], W["dbatch"]
    total = int(layout.total_bins)
    if world == 1 or args.sharding == "reads":
        return dbatch, 0, total, dbatch.n_reads * world
    n_total = dbatch.n_reads
    if dbatch.blk_off is None and args.sharding == "positions":
        sub, lo, hi, _cuts = pdist.shard_positions_device(dbatch, layout, rank, world)
    else:                                       # spliced batches / chromosome cuts are sharded on the host (set-up)
        hb = synth.device_batch_to_host(dbatch, W["chroms"], W["lens"])
        h_sub, lo, hi = pdist.shard_positions(hb, layout, rank, world,
                                              snap="chromosomes" if args.sharding == "chromosomes" else "bins")
        sub = DeviceBatch.from_host(h_sub, device)
        del hb, h_sub
    W["dbatch"] = None
    del dbatch
    torch.cuda.empty_cache()
    return sub, int(lo), int(hi), n_total


def _owned_reads(hb, layout, lo, hi):
    """Mask of the reads of a host shard that START in the rank's own bins (halo reads belong to the neighbour)."""
    c_of = np.searchsorted(hb.chrom_read_off, np.arange(len(hb)), side="right") - 1
    g = layout.chrom_bin_off[c_of] + hb.ref_start.astype(np.int64)
    return (g >= lo) & (g < hi)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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
    numa = bind_to_gpu_numa_node(local_rank, world) if args.impl != "reference" else "all host threads"
    if world > 1 and args.impl != "reference":
        dist.init_process_group("nccl", device_id=torch.device(device))

    import plastid_b200 as pb
    from plastid_b200 import synth, _lib
    from plastid_b200.genome_array import map_batch, region_sums, chain_counts, CountPlanes, length_histogram
    from plastid_b200.map_factories import CenterMapFactory

    if args.workload == "peaks":
        if rank == 0:
            run_peaks(args, device)
        return 0
    W = build_world(args, rank, device)
    chroms, lens, ann, layout, table, dbatch = W["chroms"], W["lens"], W["ann"], W["layout"], W["table"], W["dbatch"]
    fac, sf, is_center = W["fac"], W["sf"], W["center"]
    sf_tuple = None if sf is None else (sf.min_, sf.max_)
    n_batch_reads = dbatch.n_reads
    # `config` names the workload and nothing that differs between the two arms (the driver compares it)
    workload = ("%s, %d synthetic reads, %d chromosomes (%d bins, '+' and '-' planes), %d region counts"
                % (W["name"], n_batch_reads, len(chroms), layout.total_bins, ann.n_tx))
    config = {"workload": workload, "reads": n_batch_reads, "genome_bins": int(layout.total_bins), "regions": ann.n_tx,
              "l2": "inputs (%.2f GB) and outputs (%.1f GB) larger than the 126 MB L2"
                    % (8 * n_batch_reads / 1e9, (16 if is_center else 8) * layout.total_bins / 1e9)}

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
                "scaling": "strong" if args.sharding != "reads" else "weak", "vs_baseline": None,
                "dtype": "f64" if is_center else "int64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                 "cpu_model": cpu_model()},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "host": {"affinity": numa, "cpu_model": cpu_model(), "threads": threads},
                "note": "oracle port of map_factories.pyx/roitools.pyx loops (the Cython reference needs pysam, absent from "
                        "the GPU box); chromosome-parallel over %d host threads; region sums read the int64 vectors the "
                        "rules return" % threads}
        emit(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ this rank's share
    sub, lo, hi, n_total = shard_world(args, W, device, rank, world)
    ranged = (lo, hi) != (0, int(layout.total_bins))
    dbatch = None
    # Center rule: every rank derives its slot tables from the histogram of the WHOLE batch (in a real run: one
    # 512 KB all-reduce, BAMGenomeArray does it), so that sharded planes equal unsharded ones bit for bit
    hist = None
    if is_center and ranged:
        from plastid_b200.genome_array import _filtered_hist
        c_of = torch.bucketize(torch.arange(sub.n_reads, device=device), sub.chrom_read_off[1:], right=True)
        g = torch.from_numpy(layout.chrom_bin_off).to(device)[c_of] + sub.ref_start.to(torch.int64)
        keep = (g >= lo) & (g < hi) & (((sub.meta >> 17) & 1) == 0)       # reads are counted by the rank owning their start
        hist_t = torch.bincount((sub.meta[keep] & 0xFFFF).to(torch.int64), minlength=65536)
        dist.all_reduce(hist_t)
        hist = _filtered_hist(hist_t.cpu().numpy(), sf)
        del c_of, g, keep, hist_t
    planes = CountPlanes(layout, "f64" if is_center else "u32", device, (lo, hi) if ranged else None)
    planes.alloc(("+", "-"))
    table.device(device)
    L = _lib.lib()

    def step(events=None):
        if events is not None:
            events[0].record()
        map_batch(sub, layout, fac, sf, strands=("+", "-"), planes=planes, sync_stats=False,
                  bin_range=(lo, hi) if ranged else None, length_hist=hist)
        if events is not None:
            events[1].record()
        sums, live = region_sums(planes, table)
        if events is not None:
            events[2].record()
        if world > 1:
            dist.all_reduce(sums)
        if events is not None:
            events[3].record()
        return sums, live

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs ~1 s to come up: start before the warm-up
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        step()
    fence()
    L.pb_enable_kernel_timing(1)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for k in range(args.steps):
        sums, live = step(evs[k])
    fence()
    total_ms = evs[0][0].elapsed_time(evs[-1][3])
    kms, kn = C.c_float(0), C.c_int(0)
    _lib.check(L.pb_tiles_kernel_ms_total(C.byref(kms), C.byref(kn)))
    L.pb_enable_kernel_timing(0)
    step_ms = [evs[k][0].elapsed_time(evs[k + 1][0]) for k in range(args.steps - 1)] + [evs[-1][0].elapsed_time(evs[-1][3])]
    map_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    sums_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    coll_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_total / (ms_per_step / 1000.0)
    region_rate = ann.n_tx * (world if args.sharding == "reads" else 1) / (ms_per_step / 1000.0)
    table_checksum_dev = float(sums.sum().item())
    resident_table = sums.cpu().numpy().copy()

    # per-rank diagnostics: who holds what, where the step goes (a slow rank or a slow collective shows here)
    # tiles-kernel time per STEP: a spliced Center batch is mapped in several ranges (one tiles launch each, on two streams)
    k_ms = kms.value / max(args.steps, 1)
    mine = torch.tensor([float(sub.n_reads), float(hi - lo), k_ms, map_ms, sums_ms, coll_ms, float(np.median(step_ms)),
                         float(np.max(step_ms))], dtype=torch.float64, device=device)
    per_rank = [mine]
    if world > 1:
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
    names = ["reads_incl_halo", "bins", "tiles_kernel_ms", "map_ms", "region_sums_ms", "allreduce_wait_ms", "step_ms_median",
             "step_ms_max"]
    ranks_info = {n: [round(float(p[i].item()), 4) if i >= 2 else int(p[i].item()) for p in per_rank] for i, n in enumerate(names)}

    # a longer timed region for N > 1 (the K contract steps above stay the headline): a stall of one rank cannot
    # hide in, or dominate, a 100-step mean
    extended = None
    if world > 1:
        n_ext = max(100, args.steps)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        ev0.record()
        for _ in range(n_ext):
            step()
        ev1.record()
        fence()
        te = torch.tensor([ev0.elapsed_time(ev1) / n_ext], dtype=torch.float64, device=device)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        extended = {"steps": n_ext, "ms_per_step": float(te.item()), "value": n_total / (float(te.item()) / 1000.0)}

    # ------------------------------------------------------------------ launch-bound workloads: CUDA graph replay
    graph_ms = None
    if args.workload == "c1" and world == 1:
        from plastid_b200.genome_array import GraphedCount
        gc_ = GraphedCount(sub, layout, fac, sf, table, strands=("+", "-"))
        for _ in range(3):
            gc_.replay()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps * 10):
            g_sums, g_live = gc_.replay()
        ev1.record()
        torch.cuda.synchronize()
        graph_ms = ev0.elapsed_time(ev1) / (args.steps * 10)
        assert torch.equal(g_sums, sums) and torch.equal(g_live, live), "graph replay differs from the eager pass"
        del gc_

    # ------------------------------------------------------------------ table-only: plane-free region counts
    # counts_in_region / cs count need the region table, not the count vectors: pb_chain_counts maps every read's
    # site straight onto the chains.  Reported beside the dense figure (SURVEY 8d "touched-only"), never instead of it.
    table_only = None
    if not is_center:
        ref_sums = sums.clone()
        for _ in range(3):
            d_sums, d_live = chain_counts(sub, layout, fac, sf, table, (lo, hi))
            if world > 1:
                dist.all_reduce(d_sums)
        fence()
        n_to = max(args.steps, 10)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n_to):
            d_sums, d_live = chain_counts(sub, layout, fac, sf, table, (lo, hi))
            if world > 1:
                dist.all_reduce(d_sums)
        ev1.record()
        fence()
        tt = torch.tensor([ev0.elapsed_time(ev1) / n_to], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        to_ms = float(tt.item())
        n_blocks = len(table.bstart)
        # compulsory traffic: every read once (8 B) + the chain tables + the result table
        to_alg = 8.0 * sub.n_reads + 16.0 * n_blocks + 8.0 * (ann.n_tx + 1) + ann.n_tx + 16.0 * ann.n_tx
        table_only = {"kernel": "pb_chain_counts (pb_read_index, pb_chain_first_items, scan, pb_chain_items, pb_chain_totals)", "ms_per_step": to_ms, "value": n_total / (to_ms / 1000.0), "unit": UNIT,
                      "region_counts_per_sec": ann.n_tx / (to_ms / 1000.0), "steps": n_to,
                      "identical_to_plane_path": bool(torch.equal(d_sums, ref_sums) and torch.equal(d_live, live)),
                      "algorithmic_bytes_per_launch": to_alg, "plane_bytes_not_written": 4.0 * (hi - lo) * 2,
                      "speedup_vs_planes": ms_per_step / to_ms}

    # ------------------------------------------------------------------ end-to-end through the product API
    # Host buffers = the AlignmentBatch the decoder hands out (bam_io.batch_from_bam packs the transfer format while
    # decoding: delta3 streams + block words, in pinned memory).  Every step constructs a BAMGenomeArray on it and
    # asks for the planes and the region table: upload (chunked, overlapped with mapping), expansion, mapping, region
    # sums, the all-reduce at N > 1 and the device->host copy of the table are all inside the clock.
    del planes
    torch.cuda.empty_cache()
    hb = synth.device_batch_to_host(sub, chroms, lens, mapped=n_total)
    t_pack = time.perf_counter()
    hb.pack()
    hb.transfer_pinned()
    pack_s = time.perf_counter() - t_pack
    h2d = hb.transfer.nbytes
    d2h = ann.n_tx * 16 + _lib.PB_NSTATS * 8

    raw_hist = None
    if is_center and ranged:                 # the whole batch's histogram, known to whoever decoded and sharded it
        from plastid_b200.batch import meta_length_hist
        raw_hist = torch.from_numpy(meta_length_hist(hb.meta[_owned_reads(hb, layout, lo, hi)])).to(device)
        dist.all_reduce(raw_hist)
        raw_hist = raw_hist.cpu().numpy()

    def api_step(batch, use_planes=True):
        ga = pb.BAMGenomeArray(batch, mapping=fac, device=device, shard=None, bin_range=(lo, hi) if ranged else None,
                               length_hist=raw_hist)
        if sf is not None:
            ga.add_filter("size", sf)
        if use_planes:
            ga.count_planes(("+", "-"))
        return ga.count_chains(table, planes=use_planes)

    def timed(fn, n_steps, n_warm=2):
        for _ in range(n_warm):
            out = fn()
        fence()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            out = fn()
        fence()
        ms = 1000.0 * (time.perf_counter() - t0) / n_steps
        tm = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item()), out

    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, (h_sums, h_live) = timed(lambda: api_step(hb), e2e_steps)
    e2e_value = n_total / (e2e_ms / 1000.0)
    api_checksum = float(np.sum(h_sums))
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
           "steps": e2e_steps, "api": "BAMGenomeArray(batch).count_planes(('+','-')); .count_chains(table)",
           "host_format": "%s (%.2f B/read): what bam_io.batch_from_bam emits; uploaded in chunks, expanded on the device, "
                          "each chunk's bin range mapped while the next is on the wire"
                          % (type(hb.transfer).__name__, h2d / max(len(hb), 1)),
           "table_equals_device_resident_leg": bool(np.array_equal(h_sums, resident_table)),
           "pack_seconds_outside_clock": round(pack_s, 3),
           "pack_note": "the transfer format is written once per batch by the decoder (host, %d threads here); a plain SoA "
                        "batch is timed below with nothing outside the clock" % _lib.host_threads()}
    if table_only is not None:
        to_ms_e2e, (t_sums, _tl) = timed(lambda: api_step(hb, use_planes=False), e2e_steps)
        e2e["table_only"] = {"value": n_total / (to_ms_e2e / 1000.0), "ms_per_step": to_ms_e2e,
                             "api": "BAMGenomeArray(batch).count_chains(table)  (no planes: pb_chain_counts)",
                             "table_equals_plane_path": bool(np.array_equal(t_sums, h_sums))}
    # the same call on a plain SoA batch (8 B/read in pageable numpy arrays, no transfer format): every host
    # transform and the whole upload inside the clock
    from plastid_b200.batch import AlignmentBatch
    plain = AlignmentBatch(hb.chroms, hb.chrom_len, hb.ref_start, hb.meta, hb.chrom_read_off, hb.blk_off, hb.blk,
                           max_span=hb.max_span, mapped=hb.mapped)
    soa_steps = 3

    def soa_step():
        plain._dev.clear()                       # a fresh batch every step: no cached device copy
        return api_step(plain)
    soa_ms, (s_sums, _sl) = timed(soa_step, soa_steps, n_warm=1)
    soa_bytes = 8 * len(hb) + (0 if hb.blk is None else 4 * (len(hb) + 1) + 8 * len(hb.blk))
    e2e["soa_input"] = {"value": n_total / (soa_ms / 1000.0), "ms_per_step": soa_ms, "h2d_bytes_per_step": soa_bytes,
                        "steps": soa_steps, "host_format": "plain SoA in pageable numpy arrays, nothing precomputed",
                        "table_equals_packed_path": bool(np.array_equal(s_sums, h_sums))}
    clocks = sampler.stop()              # samples cover warm-up, the timed steps and the e2e steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ rooflines
    peak, peak_src = peaks()
    n_blk = 0 if sub.blk is None else sub.blk.shape[0]
    # tiles kernel: SoA in once (+ block table for spliced reads) + every bin of both planes out once — per rank
    alg_bytes = 8.0 * sub.n_reads + (4.0 * (sub.n_reads + 1) + 8.0 * n_blk if n_blk else 0.0) \
        + (8.0 if is_center else 4.0) * (hi - lo) * 2
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    kernel_name = "pb_center_tiles_kernel" if is_center else "pb_point_tiles_kernel"
    traffic = None                    # DRAM bytes per launch from the committed ncu --set full capture
    for tname in ("traffic_r01.json", "traffic_r02.json"):            # the later capture wins
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and args.genome_scale == 1.0 and world == 1:
            with open(tpath) as fh:
                traffic = json.load(fh).get("%s:%s" % (kernel_name, args.workload), traffic)
    if traffic is not None and abs(n_batch_reads - {"c2": 200_000_000, "c3": 100_000_000}.get(args.workload, -1)) > 0:
        traffic = None                # captured at the default size only
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "launches_timed": kn.value,
                "launches_per_step": kn.value / max(args.steps, 1),
                "kernel_share_of_step": k_ms / ms_per_step}
    # region sums: every chain position of this rank's bins read once (4 or 8 B) + tables + results
    own = np.clip(np.minimum(table.bend, hi) - np.maximum(table.bstart, lo), 0, None).sum()
    g_alg = (8.0 if is_center else 4.0) * float(own) + 16.0 * len(table.bstart) + 8.0 * (ann.n_tx + 1) + ann.n_tx + 16.0 * ann.n_tx
    g_traffic = None
    gpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(gpath) and args.genome_scale == 1.0 and world == 1:
        with open(gpath) as fh:
            g_traffic = json.load(fh).get("pb_region_sums:%s" % args.workload)
    roofline_gather = {"bound": "hbm", "kernel": "pb_block_sums_kernel + pb_chain_totals_kernel (pb_region_sums)", "achieved": g_alg / (sums_ms / 1000.0) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": g_alg / (sums_ms / 1000.0) / 1e9 / peak, "traffic": g_traffic,
                       "algorithmic_bytes_per_launch": g_alg, "kernel_ms": sums_ms, "positions": int(own), "chains": ann.n_tx,
                       "note": "CUDA events around the launch inside the timed steps"}
    if table_only is not None:
        table_only["roofline"] = {"bound": "hbm", "achieved": table_only["algorithmic_bytes_per_launch"] / (table_only["ms_per_step"] / 1000.0) / 1e9,
                                  "peak": peak, "unit": "GB/s",
                                  "frac": table_only["algorithmic_bytes_per_launch"] / (table_only["ms_per_step"] / 1000.0) / 1e9 / peak,
                                  "note": "touched-only figure of SURVEY 8(d): reads in + chain tables, no planes"}

    # ------------------------------------------------------------------ cpu_baseline (rank 0, N=1)
    cpu = None
    if world == 1:
        order = np.argsort(-np.asarray(lens))
        chrom_ids = sorted(int(c) for c in order[:args.cpu_sample_chroms])
        hs = host_sample(sub, chroms, lens, set(chrom_ids))
        dt, nr, nreg = cpu_reference_pass(hs, lens, table, layout, W["oracle_kw"], sf_tuple, chrom_ids, 1)
        cpu = {"value": nr / dt, "unit": UNIT, "cores": 1, "kind": "port", "cpu_model": cpu_model(),
               "sample": "%s: %d reads + %d region sums in %.2f s (oracle C port, one thread)"
                         % (",".join(chroms[c] for c in chrom_ids), nr, nreg, dt)}
        try:
            cpu["python_loop"] = python_loop_estimate(hs, chrom_ids[0], args.workload)
        except Exception as exc:      # an estimate beside the baseline, never a reason to lose the line
            cpu["python_loop"] = {"error": repr(exc)}

    binning = ["pb_bin_kernel(count)", "pb_scan_chunks_kernel", "pb_scan_top_kernel", "pb_scan_add_kernel",
               "pb_bin_kernel(fill)"] if sub.blk_off is not None else []
    if is_center:
        kernels_per_step = ["pb_tile_index_kernel"] + binning + ["pb_center_tiles_kernel"]
    else:
        kernels_per_step = ["pb_tile_index_kernel"] + binning + ["pb_point_tiles_kernel", "pb_point_overflow_kernel"]
    kernels_per_step += ["pb_stats_finish_kernel", "pb_block_sums_kernel", "pb_chain_totals_kernel"]
    sharding_text = {"reads": "read-range: every GPU maps its own batch over the whole genome (weak scaling)",
                     "positions": "ONE batch sharded by position range: range-only planes, halo reads, region tables all-reduced "
                                  "(strong scaling, SURVEY 8e)",
                     "chromosomes": "ONE batch sharded by runs of whole chromosomes (strong scaling, BASELINE config 5)"}[args.sharding]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if args.sharding == "reads" else "strong", "vs_baseline": None,
            "dtype": "f64" if is_center else "u32", "data": "synthetic", "config": config,
            "region_counts_per_sec": region_rate,
            "sharding": "single GPU" if world == 1 else sharding_text, "per_rank": ranks_info,
            "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(np.max(step_ms)),
            "host": {"affinity": numa, "cpu_model": cpu_model(), "host_threads": _lib.host_threads()},
            "e2e": e2e, "gpu_launches": len(kernels_per_step) * args.steps, "kernels_per_step": kernels_per_step,
            "roofline": roofline, "roofline_gather": roofline_gather, "table_only": table_only, "cpu_baseline": cpu,
            "clocks": clocks, "table_checksum": api_checksum}
    if extended is not None:
        line["extended"] = extended
    if graph_ms is not None:
        # the same pass replayed as one CUDA graph launch (launch-latency bound workload)
        line["cuda_graph"] = {"ms_per_step": graph_ms, "value": n_total / (graph_ms / 1000.0), "unit": UNIT,
                              "region_counts_per_sec": ann.n_tx / (graph_ms / 1000.0), "replays_timed": args.steps * 10}
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
