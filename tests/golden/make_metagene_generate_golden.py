#!/usr/bin/env python
"""Regenerates tests/golden/metagene_generate.json from the REFERENCE's own unit test data for
``metagene generate`` (plastid/test/unit/bin/test_metagene.py: the GFF3 text at :336-449, the mask
list :330-334, the query lists and expected windows / offsets / reference points :451-781).

The reference test module cannot be imported here (it imports plastid, whose Cython needs pysam), so
the module is parsed with ``ast`` and only its data literals are evaluated.  The expected values are
DATA of the reference's tests, transcribed unchanged; they pin oracle/generate.py and the CUDA path
(pb_spanning_windows) for SURVEY 8f-4.  Needs /root/reference; the JSON output is committed."""
import ast
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/plastid/test/unit/bin/test_metagene.py"

WANTED = ["_FLANKS", "_MASKS", "_TRANSCRIPTS_GFF", "_CDS_START_QUERIES", "_CDS_START_RESULTS",
          "_CDS_STOP_QUERIES", "_CDS_STOP_RESULTS", "_CDS_STOP_WITH_DELTA_RESULTS",
          "_DO_GENERATE_MAX_WINDOW", "_DO_GENERATE_MAX_WINDOW_RESULTS",
          "_DO_GENERATE_MAX_WINDOW_RESULTS_MASKED", "_DO_GENERATE_MULTI_GENE",
          "_DO_GENERATE_MULTI_GENE_RESULTS"]


class _FromStr(object):
    """``SegmentChain.from_str("...")`` in the mask list evaluates to the string itself."""
    @staticmethod
    def from_str(text):
        return text


def _jsonable(x):
    if isinstance(x, float) and math.isnan(x):
        return None                       # numpy.nan in the reference's tables
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    return x


def main():
    tree = ast.parse(open(SRC).read())
    env = {"nan": float("nan"), "SegmentChain": _FromStr, "numpy": type("np", (), {"nan": float("nan")})}
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name in WANTED:
                value = eval(compile(ast.Expression(node.value), SRC, "eval"), env)
                env[name] = value
                out[name.lstrip("_").lower()] = _jsonable(value)
    missing = [w for w in WANTED if w.lstrip("_").lower() not in out]
    assert not missing, missing
    out["source"] = "plastid/test/unit/bin/test_metagene.py (plastid v0.6.1), data literals only"
    with open(os.path.join(HERE, "metagene_generate.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote metagene_generate.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
