#!/usr/bin/env python3
"""Agreement of the CUDA path with the CPU oracle (Migrad restatement) on seeded synthetic sets, per configuration
and per pulse multiplicity, with the chi2 of the disagreeing fits compared.  Test infrastructure: it runs the
oracle, so it is a checker (like tests/), never a product path.

Usage: python tools/agreement_report.py [events_cfg1=24] [events_cfg2=24] [events_cfg3=12]
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")

TOL_T_BIN, TOL_A_REL, TOL_CHI2_REL = 0.01, 1e-3, 1e-3


def main():
    n_ev = {1: int(sys.argv[1]) if len(sys.argv) > 1 else 24, 2: int(sys.argv[2]) if len(sys.argv) > 2 else 24,
            3: int(sys.argv[3]) if len(sys.argv) > 3 else 12}
    cal = synth.make_calibration()
    orc = oracle.Oracle(cal)
    gpu = pkg.NpsWf(cal)
    spl = orc.spline_coeffs()
    threads = os.cpu_count() or 1
    for cfg in (1, 2, 3):
        ev = synth.generate_host(synth.config_params(cfg), spl, cal, 5_000_000, n_ev[cfg], n_threads=threads)
        ref = orc.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=threads)
        got = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
        exact = (np.array_equal(got["wfnpulse"], ref["wfnpulse"]) and np.array_equal(got["status"] & 3, ref["status"] & 3))
        n = ref["wfnpulse"]
        valid = np.arange(12)[None, None, :] < n[..., None]
        seeds_equal = True
        nofit = (ref["status"] & 28) == 0
        for k in ("wftime", "wfampl", "chi2"):
            seeds_equal &= bool(np.array_equal(got[k][nofit], ref[k][nofit]))
        r_ok, g_ok = (ref["status"] & 12) > 0, (got["status"] & 12) > 0
        both = r_ok & g_ok
        d_t = np.where(valid, np.abs(ref["wftime"] - got["wftime"]) / 4.0, 0.0).max(axis=-1)
        d_a = np.where(valid, np.abs(ref["wfampl"] - got["wfampl"]) / np.maximum(np.abs(ref["wfampl"]), 1e-300), 0.0).max(axis=-1)
        d_c = np.abs(ref["chi2"] - got["chi2"]) / np.maximum(np.abs(ref["chi2"]), 1e-300)
        good = (d_t <= TOL_T_BIN) & (d_a <= TOL_A_REL) & (d_c <= TOL_CHI2_REL)
        fitted = (ref["status"] & 28) > 0
        print("config %d: %d events, %d block-waveforms fitted; peak count/position, threshold decision and all "
              "non-fitted outputs exact: %s" % (cfg, n_ev[cfg], int(fitted.sum()), exact and seeds_equal))
        print("  verdicts: both converge %d | oracle only %d | GPU only %d | neither %d" % (
            int(both.sum()), int((r_ok & ~g_ok & fitted).sum()), int((g_ok & ~r_ok & fitted).sum()),
            int((fitted & ~r_ok & ~g_ok).sum())))
        print("  both converge: within tolerance %.4f %%" % (100.0 * good[both].mean() if both.any() else 100.0))
        print("  %3s %9s %10s %10s | of the disagreeing: GPU chi2 lower / equal(1e-6) / higher" % ("N", "both", "within", "frac"))
        for N in range(1, 13):
            m = both & (n == N)
            if not m.any():
                continue
            bad = m & ~good
            lo = int((got["chi2"][bad] < ref["chi2"][bad] * (1 - 1e-6)).sum())
            hi = int((got["chi2"][bad] > ref["chi2"][bad] * (1 + 1e-6)).sum())
            eq = int(bad.sum()) - lo - hi
            print("  %3d %9d %10d %10.4f | %d / %d / %d" % (N, int(m.sum()), int((m & good).sum()),
                                                           float(good[m].mean()), lo, eq, hi))
        bad = both & ~good
        if bad.any():
            rel = (got["chi2"][bad] - ref["chi2"][bad]) / ref["chi2"][bad]
            print("  disagreeing fits: %d; relative chi2 difference (GPU - oracle)/oracle: median %.3g, 10%% %.3g, 90%% %.3g; "
                  "median |dt| %.3g bin" % (int(bad.sum()), float(np.median(rel)), float(np.quantile(rel, 0.1)),
                                            float(np.quantile(rel, 0.9)), float(np.median(d_t[bad]))))


if __name__ == "__main__":
    main()
