"""Access to the committed golden files produced by the REFERENCE'S OWN programs
(tests/golden/make_script_goldens.py ran plastid's unmodified ``main()`` functions in the build container):
inputs under tests/golden/ref_scripts/in, outputs under tests/golden/ref_scripts/out."""
import gzip
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IN = os.path.join(HERE, "golden", "ref_scripts", "in")
OUT = os.path.join(HERE, "golden", "ref_scripts", "out")

_CIGAR_OPS = "MIDNSHP=X"
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")

OFFSETS = {25: 12, 26: 12, 27: 13, 28: 13, 29: 14, 30: 14, 31: 14, 32: 14, 33: 15, 34: 15, 35: 15, "default": 14}


def inp(name):
    return os.path.join(IN, name)


def out(name):
    return os.path.join(OUT, name)


def read_alignments():
    """-> (ordered {chrom: length}, [(chrom, start, is_reverse, cigartuples), ...] in file order)."""
    chrom_lengths, reads = {}, []
    with gzip.open(inp("reads.aln.gz"), "rt") as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if f[0] == "@SQ":
                chrom_lengths[f[1]] = int(f[2])
            elif line.strip():
                reads.append((f[0], int(f[1]), f[2] == "-", [(_CIGAR_OPS.index(op), int(n)) for n, op in _CIGAR_RE.findall(f[3])]))
    return chrom_lengths, reads


def bed_rows(name):
    """Rows of a BED file as lists of str."""
    with open(inp(name)) as fh:
        return [ln.rstrip("\n").split("\t") for ln in fh if ln.strip() and not ln.startswith(("#", "track", "browser"))]


def table(path, comment="##"):
    """Tab-delimited table with a header row -> (header, list of row lists); '##' lines skipped."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        lines = [ln.rstrip("\n") for ln in fh if not ln.startswith(comment) and ln.strip()]
    return lines[0].split("\t"), [ln.split("\t") for ln in lines[1:]]


def columns(path):
    head, rows = table(path)
    return {h: [r[i] for r in rows] for i, h in enumerate(head)}


def matrix(path):
    """numpy.savetxt output (tab-delimited, trailing tab tolerated) -> float64 2-D array."""
    with gzip.open(path, "rt") as fh:
        return np.array([[float(x) for x in ln.split("\t") if x.strip()] for ln in fh if ln.strip()], dtype=np.float64)


def track_lines(path):
    """Lines of a wiggle / bedGraph track without the `track` line (it names the output path)."""
    with gzip.open(path, "rt") as fh:
        return [ln.rstrip("\n") for ln in fh if not ln.startswith("track")]


def count_vectors():
    """get_count_vectors output bundle -> {file name: float64 vector}."""
    vecs = {}
    with gzip.open(out("count_vectors.txt.gz"), "rt") as fh:
        for ln in fh:
            name, vals = ln.rstrip("\n").split("\t")
            vecs[name] = np.array([float(x) for x in vals.split()], dtype=np.float64)
    return vecs


def fnum(text):
    return float("nan") if text in ("nan", "NaN", "") else float(text)


def assert_float_columns_equal(got_rows, want_rows, float_cols, rtol=0.0, label=""):
    """Row lists of str: non-float columns equal as text, float columns equal as numbers (nan == nan)."""
    assert len(got_rows) == len(want_rows), "%s: %d rows, reference has %d" % (label, len(got_rows), len(want_rows))
    for i, (g, w) in enumerate(zip(got_rows, want_rows)):
        assert len(g) == len(w), "%s row %d: column count" % (label, i)
        for j, (a, b) in enumerate(zip(g, w)):
            if j in float_cols:
                fa, fb = fnum(a), fnum(b)
                if np.isnan(fa) or np.isnan(fb):
                    assert np.isnan(fa) and np.isnan(fb), "%s row %d col %d: %s vs %s" % (label, i, j, a, b)
                elif rtol:
                    assert abs(fa - fb) <= rtol * max(abs(fa), abs(fb)), "%s row %d col %d: %s vs %s" % (label, i, j, a, b)
                else:
                    assert a == b or fa == fb, "%s row %d col %d: %s vs %s" % (label, i, j, a, b)
            else:
                assert a == b, "%s row %d col %d: %r vs %r" % (label, i, j, a, b)


def rules():
    """tests/golden/ref_rules.json.gz: the reference's own map factories / BAMGenomeArray on seeded reads with every
    CIGAR op (tests/golden/make_rule_goldens.py)."""
    import json
    with gzip.open(os.path.join(HERE, "golden", "ref_rules.json.gz"), "rt") as fh:
        d = json.load(fh)
    d["offsets_default"] = {(k if k == "default" else int(k)): v for k, v in d["offsets_default"].items()}
    d["offsets_plain"] = {int(k): v for k, v in d["offsets_plain"].items()}
    d["read_tuples"] = [(s, strand == "-", [(_CIGAR_OPS.index(op), int(n)) for n, op in _CIGAR_RE.findall(c)]) for s, strand, c in d["reads"]]
    return d
