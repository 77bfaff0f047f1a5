"""``cs``: ``generate`` (plastid/bin/cs.py:190-664: merge genes that share exons, classify and mask their
positions) and ``count`` (:667-729: per-gene exon / utr5 / cds / utr3 counts and RPKM)."""
import argparse
import sys

import numpy as np

from . import _cli
from ..chains import ChainSet, chain_binary, chain_union
from ..masks import GenomeHash
from ..roitools import GenomicSegment, SegmentChain
from ..windows import layout_for_features

KEYS = ("exon", "utr5", "cds", "utr3")
CLASSES = ("utr5", "cds", "utr3")


# ---------------------------------------------------------------------------------------------
# generate
# ---------------------------------------------------------------------------------------------
def merge_genes(tx_ivcs):
    """plastid/bin/cs.py:190-239: genes whose transcripts share an exon (same chromosome, strand, start and
    end) are merged, transitively -> dict raw gene name -> merged name (sorted, comma-joined)."""
    parent = {}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    owner = {}
    for txid, chain in tx_ivcs.items():
        if chain.strand not in ("+", "-"):
            raise KeyError(chain.strand)           # the reference keeps '+' and '-' exon tables only (cs.py:203)
        gene = chain.get_gene()
        parent.setdefault(gene, gene)
        for iv in chain:
            key = (chain.strand, chain.chrom, iv.start, iv.end)
            if key in owner:
                a, b = find(owner[key]), find(gene)
                if a != b:
                    parent[b] = a
            else:
                owner[key] = gene
    groups = {}
    for gene in parent:
        groups.setdefault(find(gene), []).append(gene)
    dout = {}
    for members in groups.values():
        name = ",".join(sorted(members))
        for gene in members:
            dout[gene] = name
    return dout


def _chains_from(bstart, bend, off, i, layout, chrom, strand, **attr):
    base = int(layout.chrom_bin_off[layout.index[chrom]])
    segs = [GenomicSegment(chrom, int(a) - base, int(b) - base, strand) for a, b in zip(bstart[off[i]:off[i + 1]], bend[off[i]:off[i + 1]])]
    return SegmentChain(*segs, **attr)


def process_partial_group(transcripts, mask_hash=None, printer=None, device="cuda"):
    """plastid/bin/cs.py:242-496 for a dict ``{transcript id: Transcript}`` -> (gene table, transcript
    table, merged_genes) with the reference's columns (``pandas.DataFrame`` s sorted by ``region``).

    The reference pools, intersects and subtracts python sets of positions gene by gene; here every
    step is one batched chain operation over all genes / transcripts (``plastid_b200.chains``):
    E = union of a gene's transcripts; P = transcript AND {5' of CDS, CDS, 3' of CDS}; U = union of P per
    gene and class; M = E AND (union of the other genes overlapping E, plus the mask features);
    gene class = U minus the other two classes' U and M; transcript class = P minus the same."""
    import pandas as pd
    mask_hash = GenomeHash([]) if mask_hash is None else mask_hash
    txids = list(transcripts)
    txs = [transcripts[t] for t in txids]
    merged_genes = merge_genes(transcripts)
    merged_gene_tx = {}
    for t, tx in zip(txids, txs):
        merged_gene_tx.setdefault(merged_genes[tx.get_gene()], []).append(t)
    gene_ids = list(merged_gene_tx)
    # A gene whose transcripts lie on several chromosomes or strands: the reference prints "Skipping gene ..." and
    # then pools the bare position numbers of all of them on the FIRST transcript's chromosome and strand
    # (cs.py:324-343); every chain of the gene and of its transcripts is written there (:405-406, :444-460), while
    # 5' / 3' of a transcript's coding region follow the transcript's own strand.  Reproduced: such transcripts are
    # lowered into the bins of their gene's place.
    place = {}
    for gene_id, members in merged_gene_tx.items():
        chroms_seen, strands_seen = [], []
        for t in members:
            chroms_seen.append(transcripts[t].chrom)
            strands_seen.append(transcripts[t].strand)
            if printer is not None and len(set(chroms_seen)) > 1:
                printer.write("Skipping gene %s which contains multiple chromosomes: %s" % (gene_id, ",".join(chroms_seen)))
            if printer is not None and len(set(strands_seen)) > 1:
                printer.write("Skipping gene %s which contains multiple strands: %s" % (gene_id, ",".join(strands_seen)))
        place[gene_id] = (chroms_seen[0], strands_seen[0])
    gene_index = {g: i for i, g in enumerate(gene_ids)}
    gene_of_tx = np.asarray([gene_index[merged_genes[tx.get_gene()]] for tx in txs], dtype=np.int64)
    n_tx, n_gene = len(txs), len(gene_ids)
    cols = ["region", "transcript_ids", "exon_unmasked", "exon", "masked", "utr5", "cds", "utr3",
            "exon_bed", "utr5_bed", "cds_bed", "utr3_bed", "masked_bed"]
    if n_tx == 0:
        return pd.DataFrame({c: [] for c in cols}), pd.DataFrame({c: [] for c in cols}), merged_genes
    tx_place = [place[merged_genes[tx.get_gene()]] for tx in txs]
    moved = [SegmentChain(*[GenomicSegment(c, seg.start, seg.end, st) for seg in tx])
             for tx, (c, st) in zip(txs, tx_place) if (c, st) != (tx.chrom, tx.strand) and len(tx)]
    layout = layout_for_features(txs, mask_hash.features, moved)

    # T: transcripts; R: the three genomic ranges of each transcript (5' of the CDS, CDS, 3' of it)
    bs, be, off, rs, re_, roff = [], [], [0], [], [], [0]
    for tx, (p_chrom, _p_strand) in zip(txs, tx_place):
        base = int(layout.chrom_bin_off[layout.index[p_chrom]])
        top = int(layout.chrom_bin_off[layout.index[p_chrom] + 1])
        for seg in tx:
            bs.append(base + seg.start)
            be.append(base + seg.end)
        off.append(len(bs))
        if getattr(tx, "cds_genome_start", None) is not None and tx.cds_genome_end is not None:
            gs, ge = base + int(tx.cds_genome_start), base + int(tx.cds_genome_end)
            ranges = [(base, gs), (gs, ge), (ge, top)]
            if tx.strand == "-":
                ranges = ranges[::-1]
        else:
            ranges = [(0, 0)] * 3                                      # get_cds / get_utr5 / get_utr3 are empty
        for a, b in ranges:
            if b > a:
                rs.append(a)
                re_.append(b)
            roff.append(len(rs))
    T = ChainSet.from_numpy(bs, be, off, device)
    R = ChainSet.from_numpy(rs, re_, roff, device)
    tx_ids3 = np.repeat(np.arange(n_tx, dtype=np.int64), 3)
    P = chain_binary("and", T, tx_ids3, R, np.arange(3 * n_tx, dtype=np.int64))          # P[3t+k]

    # gene-level pools: E[g] = union of transcripts, U[3g+k] = union of P[3t+k] over the gene's transcripts
    order = np.argsort(gene_of_tx, kind="stable")
    g_off = np.zeros(n_gene + 1, dtype=np.int64)
    np.cumsum(np.bincount(gene_of_tx, minlength=n_gene), out=g_off[1:])
    E = chain_union(T, g_off, order)
    u_off = np.zeros(3 * n_gene + 1, dtype=np.int64)
    np.cumsum(np.repeat(np.diff(g_off), 3), out=u_off[1:])
    u_mem = np.concatenate([3 * order[g_off[g]:g_off[g + 1]] + k for g in range(n_gene) for k in range(3)]) if n_gene else np.zeros(0, np.int64)
    U = chain_union(P, u_off, u_mem)

    # other genes overlapping a gene (same chromosome and strand; genes with the identical position set
    # are not masked against each other, cs.py:364-366)
    e_bs, e_be, e_off = E.numpy()
    first_tx = [txs[order[g_off[g]]] for g in range(n_gene)]
    span_lo = np.asarray([e_bs[e_off[g]] if e_off[g + 1] > e_off[g] else 0 for g in range(n_gene)], dtype=np.int64)
    span_hi = np.asarray([e_be[e_off[g + 1] - 1] if e_off[g + 1] > e_off[g] else 0 for g in range(n_gene)], dtype=np.int64)
    strand_cls = np.asarray([1 if tx.strand == "-" else 0 for tx in first_tx], dtype=np.int64)
    def blocks_key(g):
        return e_bs[e_off[g]:e_off[g + 1]].tobytes() + e_be[e_off[g]:e_off[g + 1]].tobytes()

    cand_of = [[] for _ in range(n_gene)]
    for cls in (0, 1):
        idx = np.flatnonzero(strand_cls == cls)
        if len(idx) == 0:
            continue
        idx = idx[np.argsort(span_lo[idx], kind="stable")]
        los, his = span_lo[idx], span_hi[idx]
        run_hi = np.maximum.accumulate(his)
        upper = np.searchsorted(los, his, side="left")         # sorted genes that start before this gene's end ...
        lower = np.searchsorted(run_hi, los, side="right")     # ... from the first whose running max end passes its start
        for j, g in enumerate(idx):
            if upper[j] - lower[j] <= 1:
                continue                                       # only the gene itself
            key = blocks_key(g)
            for h in idx[lower[j]:upper[j]]:
                if h != g and span_hi[h] > los[j] and blocks_key(h) != key:
                    cand_of[g].append(int(h))
    members = np.asarray([h for cand in cand_of for h in cand], dtype=np.int64)
    n_off = np.zeros(n_gene + 1, dtype=np.int64)
    np.cumsum([len(cand) for cand in cand_of], out=n_off[1:])
    N = chain_union(E, n_off, members)
    gene_ids_arr = np.arange(n_gene, dtype=np.int64)
    F = chain_binary("and", E, gene_ids_arr, N, gene_ids_arr)
    parts = [F]
    if len(mask_hash):
        mi = mask_hash.mask_index(layout)
        MK = ChainSet.from_numpy(mi.mask_start, mi.mask_end, mi.class_off, device)
        parts.append(chain_binary("and", E, gene_ids_arr, MK, strand_cls))
    if len(parts) == 1:
        M = F
    else:
        both = ChainSet.cat(parts)
        M = chain_union(both, np.arange(n_gene + 1, dtype=np.int64) * 2,
                        np.stack([gene_ids_arr, gene_ids_arr + n_gene], axis=1).reshape(-1))

    # B[3g+k] = everything class k must not contain: the other two pooled classes and the masked positions
    UM = ChainSet.cat([U, M])
    b_mem = np.concatenate([[3 * g + j for j in range(3) if j != k] + [3 * n_gene + g] for g in range(n_gene) for k in range(3)])
    B = chain_union(UM, np.arange(3 * n_gene + 1, dtype=np.int64) * 3, b_mem.astype(np.int64))
    gene_exon = chain_binary("sub", E, gene_ids_arr, M, gene_ids_arr)
    u_ids = np.arange(3 * n_gene, dtype=np.int64)
    gene_cls = chain_binary("sub", U, u_ids, B, u_ids)
    p_ids = np.arange(3 * n_tx, dtype=np.int64)
    tx_cls = chain_binary("sub", P, p_ids, B, (3 * gene_of_tx[:, None] + np.arange(3)[None, :]).reshape(-1))
    tx_exon = chain_binary("sub", T, np.arange(n_tx, dtype=np.int64), M, gene_of_tx)

    # results -> chains -> the reference's tables
    h_m, h_ge, h_gc, h_tc, h_te = M.numpy(), gene_exon.numpy(), gene_cls.numpy(), tx_cls.numpy(), tx_exon.numpy()
    gene_table, transcript_table = {c: [] for c in cols}, {c: [] for c in cols}
    for g, gene_id in enumerate(gene_ids):
        chrom, strand = first_tx[g].chrom, first_tx[g].strand
        raw = _chains_from(e_bs, e_be, e_off, g, layout, chrom, strand)
        masked = _chains_from(*h_m, g, layout, chrom, strand, ID=gene_id)
        exon = _chains_from(*h_ge, g, layout, chrom, strand, ID=gene_id)
        gene_table["region"].append(gene_id)
        gene_table["transcript_ids"].append(",".join(sorted(merged_gene_tx[gene_id])))
        gene_table["exon_unmasked"].append(str(raw))
        gene_table["masked"].append(str(masked))
        gene_table["masked_bed"].append(masked.as_bed())
        gene_table["exon"].append(str(exon))
        gene_table["exon_bed"].append(exon.as_bed())
        for k, key in enumerate(CLASSES):
            ch = _chains_from(*h_gc, 3 * g + k, layout, chrom, strand, ID=gene_id)
            gene_table[key].append(str(ch))
            gene_table["%s_bed" % key].append(ch.as_bed())
    for t, txid in enumerate(txids):
        tx, g = txs[t], int(gene_of_tx[t])
        t_chrom, t_strand = tx_place[t]
        masked = _chains_from(*h_m, g, layout, t_chrom, t_strand, ID=txid)
        exon = _chains_from(*h_te, t, layout, t_chrom, t_strand, ID=txid)
        transcript_table["region"].append(txid)
        transcript_table["exon"].append(str(exon))
        transcript_table["exon_bed"].append(exon.as_bed())
        for k, key in enumerate(CLASSES):
            ch = _chains_from(*h_tc, 3 * t + k, layout, t_chrom, t_strand, ID=txid)
            transcript_table[key].append(str(ch))
            transcript_table["%s_bed" % key].append(ch.as_bed())
        transcript_table["masked"].append(str(masked))
        transcript_table["masked_bed"].append(masked.as_bed())
        transcript_table["exon_unmasked"].append(str(tx))
        transcript_table["transcript_ids"].append(txid)
    gene_df = pd.DataFrame(gene_table)
    gene_df.sort_values(["region"], inplace=True)
    transcript_df = pd.DataFrame(transcript_table)
    transcript_df.sort_values(["region"], inplace=True)
    return gene_df, transcript_df, merged_genes


def write_output_files(table, title, outbase):
    """plastid/bin/cs.py:121-187: ``OUTBASE_<title>_<key>.bed`` and ``OUTBASE_<title>.positions``."""
    for k in ("utr5", "utr3", "cds", "masked", "exon"):
        with open("%s_%s_%s.bed" % (outbase, title, k), "w") as fh:
            for line in table["%s_bed" % k]:
                fh.write(line)
    table.to_csv("%s_%s.positions" % (outbase, title), sep="\t", header=True, index=False, na_rep="nan",
                 float_format="%.8f",
                 columns=["region", "exon", "utr5", "cds", "utr3", "masked", "exon_unmasked", "transcript_ids"])


def do_generate(transcripts, mask_hash=None, outbase=None, device="cuda"):
    """plastid/bin/cs.py:498-664 for an iterable of transcripts -> (gene table, transcript table,
    merged_genes); writes the reference's output files when ``outbase`` is given."""
    tx_dict = {tx.get_name(): tx for tx in transcripts}
    gene_table, transcript_table, merged_genes = process_partial_group(tx_dict, mask_hash, device=device)
    if outbase is not None:
        with open("%s_merged.txt" % outbase, "w") as fout:
            for gene, merged_name in sorted(merged_genes.items()):
                fout.write("%s\t%s\n" % (gene, merged_name))
        write_output_files(gene_table, "gene", outbase)
        write_output_files(transcript_table, "transcript", outbase)
    return gene_table, transcript_table, merged_genes


# ---------------------------------------------------------------------------------------------
# count
# ---------------------------------------------------------------------------------------------


def do_count(ga, gene_positions):
    """``gene_positions``: dict with ``region`` + one chain-string column per key in KEYS (the
    positions table ``cs generate`` writes).  Returns (column_order, columns dict)."""
    normconst = 1000.0 * 1e6 / ga.sum()
    names = list(gene_positions["region"])
    chains = []
    for k in KEYS:
        chains.extend(SegmentChain.from_str(s) for s in gene_positions[k])
    sums, _live = ga.count_chains(chains, use_masks=False)      # cs applies no masks at count time
    n = len(names)
    cols, order = {"region": names}, ["region"]
    for j, k in enumerate(KEYS):
        total = sums[j * n:(j + 1) * n]
        length = np.asarray([ch.length for ch in chains[j * n:(j + 1) * n]], dtype=np.int64)
        with np.errstate(all="ignore"):
            rpkm = np.where(length > 0, normconst * total / np.maximum(length, 1), np.nan)
        cols["%s_reads" % k] = [0 if L == 0 else t for t, L in zip(total, length)]   # sum([]) == 0 (int)
        cols["%s_length" % k] = list(length)
        cols["%s_rpkm" % k] = list(rpkm)
        order += ["%s_reads" % k, "%s_length" % k, "%s_rpkm" % k]
    return order, cols


def _fmt(v):
    if isinstance(v, (float, np.floating)):
        return "nan" if np.isnan(v) else "%.8f" % v
    return str(v)


def write_table(fout, order, cols):
    """The reference writes a pandas table with ``float_format="%.8f"`` (cs.py:716-727): a column holding any float
    is a float column, and its integer zeros (``sum([])`` of an empty chain) print as ``0.00000000`` too."""
    fout.write("\t".join(order) + "\n")
    is_float = {c: any(isinstance(v, (float, np.floating)) for v in cols[c]) for c in order}
    for i in range(len(cols["region"])):
        fout.write("\t".join(_fmt(float(cols[c][i]) if is_float[c] else cols[c][i]) for c in order) + "\n")


def main(argv=sys.argv[1:]):
    """``cs generate OUTBASE --annotation_files ...`` / ``cs count POSITION_FILE OUTBASE --count_files ...``
    (plastid/bin/cs.py:1259-1380).  ``count`` under ``torchrun``: every rank counts its own genome range."""
    parser = argparse.ArgumentParser(description=__doc__)
    sub = parser.add_subparsers(dest="program")
    gp = sub.add_parser("generate")
    _cli.add_base_args(gp)
    _cli.add_annotation_args(gp)
    _cli.add_mask_args(gp)
    gp.add_argument("--device", default="cuda")
    gp.add_argument("outbase")
    cp = sub.add_parser("count")
    _cli.add_base_args(cp)
    _cli.add_alignment_args(cp, disabled=("normalize",))
    cp.add_argument("position_file")
    cp.add_argument("outbase")
    args = parser.parse_args(argv)
    if args.program == "generate":
        transcripts = _cli.chains_from_args(args, as_transcripts=True)
        masks = _cli.chains_from_args(args, prefix="mask_")
        do_generate(transcripts, GenomeHash(masks), args.outbase, args.device)
        return
    if args.program != "count":
        parser.error("the `generate` and `count` sub-programs are on the GPU path")
    ga = _cli.genome_array_from_args(args, disabled=("normalize",))
    order, cols = do_count(ga, _cli.read_pl_table(args.position_file))
    if _cli.is_writer():
        with open("%s.txt" % args.outbase, "w") as fout:
            write_table(fout, order, cols)
    _cli.finish_distributed()


if __name__ == "__main__":
    main()
