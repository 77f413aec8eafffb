#!/usr/bin/env python3
"""Generates tests/golden/*.npz: seeded synthetic inputs + the ORACLE's outputs for them.

The reference ships no golden vectors and cannot be run here (needs ROOT), so
these fixtures pin the oracle against itself over time (regression) and give the GPU tests a
fixed target.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402


def main():
    cal = synth.make_calibration()
    out_dir = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out_dir, "calib.npz"), interpY=cal["interpY"], timeref=cal["timeref"],
                        cortime=cal["cortime"], preswf=cal["preswf"], kappa=cal["kappa"])
    for timerefacc in (0.0, -5.0):
        o = oracle.Oracle(cal, timerefacc=timerefacc)
        spl = o.spline_coeffs()
        for cfg, n, absent in ((1, 1, 0.0), (2, 2, 0.05), (3, 2, 0.03)):
            if timerefacc != 0.0 and cfg != 2:
                continue
            ev = synth.generate_host(synth.config_params(cfg, absent_frac=absent), spl, cal, 7000 + cfg, n,
                                     n_threads=4, counts=True)
            assert np.array_equal(ev["counts"] * synth.LSB, ev["signal"])
            r = o.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=8)
            hist = np.zeros((n, 1080, 110), np.float32)
            for e in range(n):
                for b in range(0, 1080, 7):
                    if ev["pres"][e, b]:
                        hist[e, b] = o.matched_filter(b, ev["signal"][e])[1]
            name = "cfg%d_acc%s.npz" % (cfg, "0" if timerefacc == 0 else "m5")
            np.savez_compressed(os.path.join(out_dir, name), counts=ev["counts"], pres=ev["pres"],
                                corr=ev["corr_time_HMS"], timerefacc=timerefacc, mfhist_every7=hist[:, ::7],
                                **{k: r[k] for k in ("wfnpulse", "wftime", "wfampl", "chi2", "timewf", "amplwf",
                                                     "status")})
            print(name, "pulses", int(r["wfnpulse"].sum()), "fits ok", int(((r["status"] & 12) > 0).sum()),
                  "fallback", int(((r["status"] & 16) > 0).sum()))


if __name__ == "__main__":
    main()
