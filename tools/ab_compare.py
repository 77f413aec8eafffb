#!/usr/bin/env python3
"""A/B aid: compare the dumps of tools/ab_variant.py bit for bit against the first tag.  Usage: ab_compare.py base tag..."""
import sys
import numpy as np

base = sys.argv[1]
for tag in sys.argv[2:]:
    for cfg in (2, 3):
        a = np.load("/tmp/ab_%s_cfg%d.npz" % (base, cfg))
        b = np.load("/tmp/ab_%s_cfg%d.npz" % (tag, cfg))
        res = []
        for k in a.files:
            x, y = a[k], b[k]
            same = np.array_equal(x.view(np.uint8), y.view(np.uint8))
            res.append("%s %s" % (k, "==" if same else "DIFF(%d)" % int((x != y).sum())))
        print("CMP %s vs %s cfg%d: %s" % (tag, base, cfg, ", ".join(res)), flush=True)
