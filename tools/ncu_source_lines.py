#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS addresses, executed
instructions, stall samples) with the line table of the built object (`cuobjdump -xelf` + `nvdisasm -g`), because
ncu's own CUDA-source view needs the sources at the path they had on the GPU box.

Usage: ncu_source_lines.py report.ncu-rep object.o kernel_substring [units=1] [top=40] [kernel_id_in_report=1]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    kid = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
    off2line, cur = {}, None
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::%d" % kid], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    marks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    rows = rows[marks[0]:marks[1]] if len(marks) > 1 else rows[marks[0]:]
    hdr = rows[1]
    ia, ii, isrc, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    byline, bysamp, fp64, byop = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in rows[2:]:
        n, s = int(r[ii]), int(r[ist])
        ln = off2line.get(int(r[ia], 16) - base)
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        byline[ln] += n; bysamp[ln] += s; tot += n; tots += s
        byop[op.split(".")[0]] += n
        if op.startswith("D"):
            fp64[ln] += n
    print("kernel %s: %d warp-instructions, %.1f per unit (%g units), %d stall samples" % (rows[0][1][:60], tot, tot / units, units, tots))
    print("%7s %10s %9s %8s  line" % ("instr%", "per unit", "fp64/unit", "stall%"))
    for ln, n in byline.most_common(top):
        print("%6.2f%% %10.1f %9.1f %7.2f%%  %s" % (100.0 * n / tot, n / units, fp64[ln] / units, 100.0 * bysamp[ln] / max(1, tots), ln))
    print("opcode mix (per unit):")
    for op, n in byop.most_common(24):
        print("  %-10s %9.1f %6.2f%%" % (op, n / units, 100.0 * n / tot))


if __name__ == "__main__":
    main()
