"""Shared test helpers: the reference's own test fixtures, transcribed, plus oracle glue."""
import numpy as np

from oracle import pyoracle as po


def kat_reads():
    """test_map_factories.py:19-28: per strand, 15 reads starting at 0 with '<L>M', L = 25..39."""
    return {s: [po.Read(0, [(po.CMATCH, L)], s == "-") for L in range(25, 40)] for s in ("+", "-")}


def kat_expected():
    """test_map_factories.py:29-49 closed-form expectations."""
    mn, mx = 25, 40
    exp = {}
    for mapping in ("fiveprime", "threeprime", "center"):
        for param in (0, 10):
            for strand in ("+", "-"):
                exp[(mapping, param, strand)] = np.zeros(2000)
    exp[("fiveprime", 0, "+")][0] = mx - mn
    exp[("fiveprime", 10, "+")][10] = mx - mn
    exp[("fiveprime", 0, "-")][mn - 1:mx - 1] = 1
    exp[("fiveprime", 10, "-")][mn - 11:mx - 11] = 1
    exp[("threeprime", 0, "-")][0] = mx - mn
    exp[("threeprime", 10, "-")][10] = mx - mn
    exp[("threeprime", 0, "+")][mn - 1:mx - 1] = 1
    exp[("threeprime", 10, "+")][mn - 11:mx - 11] = 1
    for L in range(mn, mx):
        exp[("center", 0, "+")][:L] += 1.0 / L
        exp[("center", 0, "-")][:L] += 1.0 / L
        exp[("center", 10, "+")][10:L - 10] += 1.0 / (L - 20)
        exp[("center", 10, "-")][10:L - 10] += 1.0 / (L - 20)
    return exp


def random_cigar_reads(rng, n, chrom_len, max_start=None):
    """Reads with every CIGAR op (M I D N S H P = X), for oracle cross-checks."""
    reads = []
    for _ in range(n):
        ops = []
        if rng.random() < 0.2:
            ops.append((po.CHARD_CLIP, int(rng.integers(1, 5))))
        if rng.random() < 0.3:
            ops.append((po.CSOFT_CLIP, int(rng.integers(1, 6))))
        nseg = int(rng.integers(1, 5))
        for k in range(nseg):
            ops.append((int(rng.choice([po.CMATCH, po.CEQUAL, po.CDIFF])), int(rng.integers(1, 30))))
            if k < nseg - 1:
                mid = int(rng.choice([po.CINS, po.CDEL, po.CREF_SKIP, po.CPAD, po.CMATCH]))
                ops.append((mid, int(rng.integers(1, 40 if mid != po.CREF_SKIP else 400))))
        if rng.random() < 0.3:
            ops.append((po.CSOFT_CLIP, int(rng.integers(1, 6))))
        span = sum(n_ for op, n_ in ops if op in (po.CMATCH, po.CEQUAL, po.CDIFF, po.CDEL, po.CREF_SKIP))
        hi = (max_start if max_start is not None else chrom_len - span - 1)
        start = int(rng.integers(0, max(hi, 1)))
        reads.append(po.Read(start, ops, bool(rng.integers(0, 2))))
    return reads


def oracle_planes(hb, kind, strand, **kw):
    """Concatenated whole-genome vector (one per chromosome) from the C oracle."""
    from oracle import coracle
    out = []
    dropped = 0
    for c in range(len(hb.chroms)):
        vec, _kept, d, _l = coracle.genome_vector(hb, c, strand, **kw)
        out.append(vec)
        dropped += d
    return out, dropped


def delta8_decode(w):
    """Plain-numpy statement of the delta8 format (include/plastid_b200.h, pb_unpack_delta8): test
    checker for the host encoder, independent of the device kernel."""
    n, K = w.n_reads, 128
    start = np.zeros(n, dtype=np.int64)
    meta = np.zeros(n, dtype=np.uint32)
    for B in range(len(w.blk_base)):
        cur, e = int(w.blk_base[B]), int(w.blk_exc_off[B])
        for i in range(B * K, min((B + 1) * K, n)):
            if w.dstart[i] == 255:
                cur, meta[i] = int(w.exc_start[e]), w.exc_meta[e]
                e += 1
            else:
                cur += int(w.dstart[i]) if i > B * K else 0
                meta[i] = w.meta_dict[w.code[i]]
            start[i] = cur
        assert e == int(w.blk_exc_off[B + 1])
    return start, meta


def genome_array_recipe(seed=0, n_regions=24, reads_per_region=150, read_length=30, chrom_len=60_000):
    """The reference's own golden-vector recipe (plastid/test/unit/genomics/test_genome_array.py:1832-1866),
    with our seed and synthetic regions instead of its BED file: per region (1-3 exons, either strand)
    ``reads_per_region`` reads of ``read_length`` nt are placed at random TRANSCRIPT offsets; the expected
    count vectors are written down directly in transcript coordinates — 5'/3' ends at offset 0 and 15,
    every position 1/30 (center 0) and the inner six positions 1/6 (center 12) — without any CIGAR
    walking.  Reads that span an exon junction are spliced (one N gap per junction).

    Returns ``(reads, vectors)``: ``reads`` = list of oracle Read objects (coordinate-sorted),
    ``vectors[(rule, param, strand)]`` = float64[chrom_len]."""
    rng = np.random.default_rng(seed)
    vectors = {(rule, par, st): np.zeros(chrom_len) for rule, pars in (("fiveprime", (0, 15)), ("threeprime", (0, 15)),
                                                                      ("center", (0, 12)))
               for par in pars for st in "+-"}
    reads = []
    cursor = 500
    for r in range(n_regions):
        n_ex = int(rng.integers(1, 4))
        exons = []
        for _ in range(n_ex):
            ln = int(rng.integers(40, 400))
            exons.append((cursor, cursor + ln))
            cursor += ln + int(rng.integers(30, 900))
        cursor += 300
        strand = "+-"[r % 2]
        genomic = np.concatenate([np.arange(a, b) for a, b in exons])          # ascending genomic positions
        tx = genomic if strand == "+" else genomic[::-1]                       # transcript order, 5' -> 3'
        for loc in rng.integers(0, len(tx) - read_length + 1, size=reads_per_region):
            for offset in (0, 15):                                             # test_genome_array.py:1851-1855
                vectors[("fiveprime", offset, strand)][tx[loc + offset]] += 1
                vectors[("threeprime", offset, strand)][tx[loc + read_length - offset - 1]] += 1
            pos = np.sort(tx[loc:loc + read_length])                           # :1858 get_subchain(...).get_position_list()
            for p_ in pos:
                vectors[("center", 0, strand)][p_] += 1.0 / len(pos)           # :1860-1861
            for p_ in pos[12:-12]:
                vectors[("center", 12, strand)][p_] += 1.0 / (len(pos) - 24)   # :1862-1866
            ops, run, prev = [], 0, None
            for p_ in pos.tolist():                                            # the alignment a spliced aligner reports
                if prev is not None and p_ != prev + 1:
                    ops += [(po.CMATCH, run), (po.CREF_SKIP, p_ - prev - 1)]
                    run = 0
                run += 1
                prev = p_
            ops.append((po.CMATCH, run))
            reads.append(po.Read(int(pos[0]), ops, strand == "-"))
    reads.sort(key=lambda x: x.reference_start)
    assert cursor < chrom_len
    return reads, vectors


def delta3_decode(w):
    """Plain-numpy statement of the delta3 format (include/plastid_b200.h, pb_unpack_delta3)."""
    n, K = w.n_reads, 128
    start = np.zeros(n, dtype=np.int64)
    meta = np.zeros(n, dtype=np.uint32)
    for B in range(len(w.blk_base)):
        cur, wi, e = int(w.blk_base[B]), int(w.blk_wide_off[B]), int(w.blk_exc_off[B])
        for i in range(B * K, min((B + 1) * K, n)):
            d, code = int(w.packed[i]) & 7, int(w.packed[i]) >> 3
            is_exc = False
            if d == 7:
                wv = int(w.wide[wi])
                wi += 1
                is_exc = wv == 255
                d = 7 + wv
            if is_exc:
                cur, meta[i] = int(w.exc_start[e]), w.exc_meta[e]
                e += 1
            else:
                cur += d if i > B * K else 0
                meta[i] = w.meta_dict[code]
            start[i] = cur
        assert wi == int(w.blk_wide_off[B + 1]) and e == int(w.blk_exc_off[B + 1])
    return start, meta


# ---------------------------------------------------------------------------------------------
# metagene generate golden data (tests/golden/metagene_generate.json, from the reference's
# plastid/test/unit/bin/test_metagene.py via tests/golden/make_metagene_generate_golden.py)
# ---------------------------------------------------------------------------------------------
def metagene_generate_golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metagene_generate.json")) as fh:
        return json.load(fh)


def gff3_transcript_records(text):
    """Minimal GFF3 transcript assembly for the golden GFF: what ``GFF3_TranscriptAssembler``
    (plastid/readers/gff.py:1440-1560) yields for it — segments = exon + CDS features of a parent
    transcript (1-based inclusive -> 0-based half-open), ``cds_genome_start/end`` = first CDS start /
    last CDS end, ``gene_id`` = the mRNA's sorted, comma-joined ``Parent``.  Returns a dict
    name -> dict(chrom, strand, segments, cds_genome_start, cds_genome_end, gene_id)."""
    tx_parent, exons, cds = {}, {}, {}
    for line in text.splitlines():
        if not line.strip() or line.startswith("#"):
            continue
        f = line.split("\t")
        chrom, ftype, start, end, strand = f[0], f[2].strip(), int(f[3]) - 1, int(f[4]), f[6].strip()
        attr = dict(kv.split("=", 1) for kv in f[8].strip().split(";") if "=" in kv)
        if ftype == "mRNA":
            tx_parent[attr["ID"]] = ",".join(sorted(attr.get("Parent", attr["ID"]).split(",")))
        elif ftype in ("exon", "CDS"):
            for parent in attr["Parent"].split(","):
                (exons if ftype == "exon" else cds).setdefault(parent, []).append((chrom, start, end, strand))
    out = {}
    for name in sorted(set(exons) | set(cds)):
        feats = exons.get(name, []) + cds.get(name, [])
        rec = dict(chrom=feats[0][0], strand=feats[0][3], segments=[(s, e) for _, s, e, _ in feats],
                   cds_genome_start=None, cds_genome_end=None, gene_id=tx_parent.get(name, name))
        if name in cds:
            ordered = sorted(cds[name], key=lambda x: (x[1], x[2]))
            rec["cds_genome_start"], rec["cds_genome_end"] = ordered[0][1], ordered[-1][2]
        out[name] = rec
    return out


def random_gene_models(rng, n_genes, chroms=("chrA", "chrB"), spacing=6000):
    """Seeded multi-isoform gene models for the maximal-spanning-window tests: per gene 1-4 transcripts
    derived from one exon chain by the variations the reference's own cases cover (shared start codon
    with different 5' UTRs, alternative start, skipped / shifted downstream exons, non-coding isoform,
    CDS start at a splice junction or at the transcript's first base).  Returns records shaped like
    :func:`gff3_transcript_records` (name -> dict), in a deterministic order."""
    out = {}
    for g in range(n_genes):
        chrom = chroms[g % len(chroms)]
        strand = "+-"[int(rng.integers(0, 2))]
        pos = 1000 + (g // len(chroms)) * spacing + int(rng.integers(0, 500))
        n_ex = int(rng.integers(1, 6))
        exons = []
        for _ in range(n_ex):
            ln = int(rng.integers(20, 400))
            exons.append((pos, pos + ln))
            pos += ln + int(rng.integers(30, 300))
        total = sum(e - s for s, e in exons)

        def genomic(chain, x):                       # unstranded chain coordinate -> genomic position
            for s, e in chain:
                if x < e - s:
                    return s + x
                x -= e - s
            raise IndexError(x)

        # coding region in unstranded chain coordinates [a, b)
        a = int(rng.integers(0, max(total // 2, 1)))
        b = int(rng.integers(min(a + 3, total), total + 1))
        if rng.random() < 0.15:
            a = 0
        if rng.random() < 0.15 and n_ex > 1:
            a = exons[0][1] - exons[0][0]            # first base of the second exon
            b = max(b, min(a + 3, total))
        if b <= a:
            b = min(a + 3, total)
        cds = (genomic(exons, a), genomic(exons, b - 1) + 1) if b > a else None
        n_iso = int(rng.integers(1, 5))
        for k in range(n_iso):
            segs = list(exons)
            this_cds = cds
            kind = int(rng.integers(0, 7)) if k else 0
            if kind == 1:                            # longer / shorter first exon (left end moves)
                s, e = segs[0]
                segs[0] = (max(s + int(rng.integers(-200, 15)), 0), e)
            elif kind == 2:                          # last exon's right end moves
                s, e = segs[-1]
                segs[-1] = (s, e + int(rng.integers(-15, 200)))
            elif kind == 3 and len(segs) > 2:        # skip an internal exon
                del segs[int(rng.integers(1, len(segs) - 1))]
            elif kind == 4:                          # non-coding isoform
                this_cds = None
            elif kind == 5 and cds is not None:      # alternative start / stop
                this_cds = (cds[0] + 3, cds[1]) if cds[1] - cds[0] > 6 else cds
            elif kind == 6:                          # extra upstream exon
                segs = [(max(segs[0][0] - 400, 0), max(segs[0][0] - 300, 1))] + segs
            segs = [(s, e) for s, e in segs if e > s]
            covered = set(p for s, e in segs for p in range(s, e))
            if this_cds is not None and not (this_cds[0] in covered and this_cds[1] - 1 in covered):
                this_cds = None                      # the assembler would reject it; keep it non-coding
            out["g%04d.t%d" % (g, k)] = dict(chrom=chrom, strand=strand, segments=segs,
                                             cds_genome_start=None if this_cds is None else this_cds[0],
                                             cds_genome_end=None if this_cds is None else this_cds[1],
                                             gene_id="g%04d" % g)
    return out


def add_shared_exon_genes(records, rng):
    """Extends :func:`random_gene_models` records with the cases ``cs generate`` merges or masks on:
    a second gene id owning a copy of another gene's transcript (shared exons -> merged gene, also a
    chain of three genes merged transitively), and two genes with identical position sets but no common
    exon (not merged, and not masked against each other, cs.py:364-366)."""
    out = dict(records)
    names = list(records)
    for i, name in enumerate(names[::7]):
        r = records[name]
        copy = dict(r, gene_id=r["gene_id"] + "m", segments=list(r["segments"]))
        if (i % 2 and copy["cds_genome_start"] is not None and copy["cds_genome_end"] - copy["cds_genome_start"] > 9
                and any(s <= copy["cds_genome_start"] + 3 < e for s, e in copy["segments"])):
            copy["cds_genome_start"] += 3
        out[name + ".m"] = copy
        if i % 3 == 0:                                               # third gene sharing only the LAST exon with the copy
            s, e = r["segments"][-1]
            out[name + ".mm"] = dict(chrom=r["chrom"], strand=r["strand"], segments=[(s, e), (e + 50, e + 90)],
                                     cds_genome_start=None, cds_genome_end=None, gene_id=r["gene_id"] + "mm")
    far = 10_000_000
    out["twin.a1"] = dict(chrom="chrA", strand="+", segments=[(far, far + 10)], cds_genome_start=None, cds_genome_end=None, gene_id="twinA")
    out["twin.a2"] = dict(chrom="chrA", strand="+", segments=[(far + 10, far + 20)], cds_genome_start=None, cds_genome_end=None, gene_id="twinA")
    out["twin.b"] = dict(chrom="chrA", strand="+", segments=[(far, far + 20)], cds_genome_start=far + 3, cds_genome_end=far + 12, gene_id="twinB")
    out["twin.c"] = dict(chrom="chrA", strand="+", segments=[(far + 15, far + 40)], cds_genome_start=None, cds_genome_end=None, gene_id="twinC")
    return out


def cs_generate_hand_case():
    """A hand-derived known answer for ``cs generate`` (the reference ships none in-tree): gene A with two
    coding isoforms, a non-coding gene B overlapping A's 3' end, one mask.  Derivation (cs.py:313-470):
    A's positions 100-200^300-450; B covers 380-450 of them and the mask 120-130 -> masked; pooled
    utr5 = 100-180, cds = 150-200^300-350, utr3 = 350-450; every class loses the other classes' pooled
    positions and the masked ones."""
    records = {
        "A1": dict(chrom="c", strand="+", segments=[(100, 200), (300, 400)], cds_genome_start=150, cds_genome_end=350, gene_id="A"),
        "A2": dict(chrom="c", strand="+", segments=[(100, 200), (300, 450)], cds_genome_start=180, cds_genome_end=350, gene_id="A"),
        "B1": dict(chrom="c", strand="+", segments=[(380, 500)], cds_genome_start=None, cds_genome_end=None, gene_id="B"),
    }
    masks = [("c", 120, 130, "+")]
    genes = {
        "A": dict(transcript_ids="A1,A2", exon_unmasked="c:100-200^300-450(+)", masked="c:120-130^380-450(+)",
                  exon="c:100-120^130-200^300-380(+)", utr5="c:100-120^130-150(+)", cds="c:180-200^300-350(+)",
                  utr3="c:350-380(+)"),
        "B": dict(transcript_ids="B1", exon_unmasked="c:380-500(+)", masked="c:380-450(+)", exon="c:450-500(+)",
                  utr5="na", cds="na", utr3="na"),
    }
    transcripts = {
        "A1": dict(exon="c:100-120^130-200^300-380(+)", utr5="c:100-120^130-150(+)", cds="c:180-200^300-350(+)",
                   utr3="c:350-380(+)", masked="c:120-130^380-450(+)", exon_unmasked="c:100-200^300-400(+)"),
        "A2": dict(exon="c:100-120^130-200^300-380(+)", utr5="c:100-120^130-150(+)", cds="c:180-200^300-350(+)",
                   utr3="c:350-380(+)", masked="c:120-130^380-450(+)", exon_unmasked="c:100-200^300-450(+)"),
        "B1": dict(exon="c:450-500(+)", utr5="na", cds="na", utr3="na", masked="c:380-450(+)", exon_unmasked="c:380-500(+)"),
    }
    return records, masks, genes, transcripts


def merge_segments_known_answers():
    """plastid/test/unit/genomics/test_roitools.py:356-398 (test_merge_segments), transcribed:
    list of (input intervals, expected merged intervals) on one chromosome strand."""
    seg = [(0, 100), (50, 200), (250, 300), (300, 330), (500, 550), (525, 575), (570, 590), (575, 580)]
    s01, s23, s456 = (0, 200), (250, 330), (500, 590)
    pick = lambda *idx: [seg[i] for i in idx]                           # noqa: E731
    return [(pick(0, 1), [s01]), (pick(1, 0), [s01]),                   # two overlapping
            (pick(0, 2), [seg[0], seg[2]]), (pick(2, 0), [seg[0], seg[2]]),   # two non-overlapping
            (pick(2, 3), [s23]), (pick(3, 2), [s23]),                   # two adjacent
            (pick(0, 1, 2, 3, 4), [s01, s23, seg[4]]), (pick(4, 1, 3, 2, 0), [s01, s23, seg[4]]),
            (pick(4, 5, 6), [s456]), (pick(4, 6, 5), [s456]),           # three overlapping
            (pick(6, 7), [seg[6]]), (pick(7, 6), [seg[6]]),             # one internal
            (pick(4, 5, 7, 6), [s456]), (pick(7, 4, 6, 5), [s456]),
            (pick(0), [seg[0]]), (pick(0, 0, 0), [seg[0]]), ([], [])]


def positions_to_segments_known_answers():
    """plastid/test/unit/genomics/test_roitools.py:212-268 (test_positions_to_segments), transcribed:
    list of (positions, expected intervals)."""
    return [([], []), (list(range(100)), [(0, 100)]),
            (list(range(100)) + list(range(150, 200)), [(0, 100), (150, 200)]),
            (list(range(100)) + list(range(150, 200)) + list(range(195, 205)), [(0, 100), (150, 205)])]
