#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the text summaries committed under profiles/.

  python profiles/summarize.py rep   gpurun_out/prof_x.ncu-rep  profiles/ncu_x_r01.txt  "<how it was captured>"
  python profiles/summarize.py list  gpurun_out/launches.csv    profiles/launches_x_r01_summary.txt  <n_steps> "<how>"
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = re.compile(r"^(dram__bytes|gpu__dram_throughput|gpu__time_duration|l1tex__data_pipe_lsu_wavefronts|l1tex__throughput|"
                  r"lts__throughput|lts__t_sectors_op_(read|write|red|atom)\.sum$|launch__|sm__inst_executed\.sum|sm__throughput|"
                  r"sm__warps_active|smsp__average_warps_issue_stalled|sm__inst_executed_pipe_(tma|lsu|alu|fp64)|"
                  r"smsp__inst_executed\.sum$|sm__cycles_elapsed\.max)")


def rep(path, out, how):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write("# %s\n" % how)
        for r in rows[2:]:
            name = r[head.index("Kernel Name")]
            fh.write("## %s\n" % name[:160])
            for h, u, v in sorted(zip(head, units, r)):
                if KEEP.match(h):
                    fh.write("%s\t%s\t%s\n" % (h, u, v))


def launches(path, out, n_steps, how, only="pb_"):
    """`only`: keep the library's own kernels (the bench generates its synthetic reads with torch kernels first)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = rows[0]
    ik, iv = head.index("Kernel Name"), head.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if only and only not in r[ik]:
            continue
        k = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")
        agg.setdefault(k, []).append(float(r[iv].replace(",", "")) / 1000.0)
    total = sum(sum(v) for v in agg.values())
    with open(out, "w") as fh:
        fh.write("# %s\n# per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's event-timed\n"
                 "# roofline.kernel_share_of_step, not absolutes\nkernel\tlaunches\tavg_us\tshare_of_listed\n" % how)
        for k, v in agg.items():
            fh.write("%s\t%d\t%.1f\t%.3f\n" % (k, len(v), sum(v) / len(v), sum(v) / total))


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
