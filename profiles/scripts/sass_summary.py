"""cuobjdump -sass of the shipped library -> mnemonic counts per kernel (proof of the TMA bulk copies / reductions, the
256-bit stores, the 16-byte loads; and of the absence of tensor-core instructions on this path).
  python profiles/scripts/sass_summary.py > profiles/sass_summary_r02c.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
COLS = ["UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG", "STG.E.ENL2.256", "LDG.E.128", "LDS", "ATOMS", "REDG", "ATOMG", "UTCHMMA", "HMMA",
        "MATCH", "SHFL", "REDUX", "MEMBAR"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "plastid_b200", "libplastid_b200.so")],
                         capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = counts.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for c in COLS:
                if op.startswith(c):
                    cur[c] += 1
    print("# cuobjdump -sass plastid_b200/libplastid_b200.so (sm_100a), instruction mnemonic counts per kernel (prefix match)")
    print("# UBLKCP = cp.async.bulk (TMA bulk copy), UBLKRED = cp.reduce.async.bulk, STG.E.ENL2.256 = 256-bit global stores,")
    print("# LDG.E.128 = 16-byte loads, LDS = shared loads (site tables, tiles), ATOMS = shared atomics, REDG = fire-and-forget global")
    print("# reductions (count pass of the binning), ATOMG = returning global atomics; no UTC*MMA / HMMA: no tensor-core work on this path")
    print("kernel\t" + "\t".join(COLS))
    for name, c in counts.items():
        print(name + "\t" + "\t".join(str(c[k]) for k in COLS))


if __name__ == "__main__":
    sys.exit(main())
