#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: opcode mix by executed instructions and the
instructions with the most stall samples.  Usage: ncu_sass_summary.py file.csv [top_n]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ops, samp = collections.Counter(), collections.Counter()
    tot = totS = 0
    lines = []
    for r in rows:
        if len(r) <= iE or r is hdr:
            continue
        try:
            e, s = int(r[iE] or 0), int(r[iSm] or 0)
        except ValueError:
            continue
        src = r[iS].strip()
        toks = src.split()
        m = toks[1] if toks[0].startswith("@") else toks[0]
        m = m.split(".")[0]
        ops[m] += e; samp[m] += s; tot += e; totS += s
        lines.append((s, e, src, r))
    print("total warp-instructions", tot, "stall samples", totS)
    for m, c in ops.most_common(top_n):
        print("%-10s inst %12d %5.1f%%   samples %5.1f%%" % (m, c, 100.0 * c / tot, 100.0 * samp[m] / max(1, totS)))
    print("instructions with most stall samples:")
    lines.sort(key=lambda x: -x[0])
    for s, e, src, r in lines[:top_n]:
        st = {h: int(r[i] or 0) for h, i in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("%6d %10d %-58s %s" % (s, e, src[:58], top))


if __name__ == "__main__":
    main()
