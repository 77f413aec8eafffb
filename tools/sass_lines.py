#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (line table via nvdisasm -g): where the code size goes.

Usage: sass_lines.py object.o kernel_substring [top=40]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
    cnt, byfile, cur, tot = collections.Counter(), collections.Counter(), None, 0
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", l):
            cnt[cur] += 1
            byfile[cur[0] if cur else None] += 1
            tot += 1
    print("%s: %d instructions (%.1f KB)" % (kern, tot, tot * 16 / 1024))
    for f, n in byfile.most_common():
        print("  %6d  %s" % (n, f))
    for ln, n in cnt.most_common(top):
        print("%6d  %s" % (n, ln))


if __name__ == "__main__":
    main()
