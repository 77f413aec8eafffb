#!/usr/bin/env python3
"""Print (and optionally write as markdown) the per-kernel rows of an `ncu --csv` launch list with several metrics.
Usage: launch_list.py launches.csv [min_id]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_id = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > iv:
        d.setdefault((int(r[iid]), r[ik]), {})[r[im]] = r[iv]
ms = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
      "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
print("| id | kernel | us | warp-instr (M) | issue % | warps active % | fp64 pipe % |")
print("|---|---|---|---|---|---|---|")
for (i, k), v in d.items():
    if i < min_id or "fit_kernel<25>" in k or "fit_kernel<(int)25>" in k:
        continue
    name = k.split("(")[0].replace("npswf::", "").replace("void ", "")
    if "<" in k.split("(")[0]:
        name = k[:k.index(">") + 1].replace("npswf::", "").replace("void ", "").replace("(int)", "")
    f = lambda m, s=1.0: ("%.1f" % (float(v[m].replace(",", "")) * s)) if m in v else "-"
    print("| %d | %s | %s | %s | %s | %s | %s |" % (i, name, f(ms[0], 1e-3), f(ms[1], 1e-6), f(ms[2]), f(ms[3]), f(ms[4])))
