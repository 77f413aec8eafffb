#!/usr/bin/env python3
"""MIGRAD fit mode of the CUDA path against the CPU oracle's Migrad restatement: every output compared bit for bit
(wfnpulse, wftime, wfampl, chi2, timewf, amplwf, status), per configuration, with the fit-stage rate.
Test infrastructure (runs the oracle as the checker).

Usage: python tools/migrad_check.py [events_cfg1=8] [events_cfg2=8] [events_cfg3=4]
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")


def main():
    n_ev = {1: int(sys.argv[1]) if len(sys.argv) > 1 else 8, 2: int(sys.argv[2]) if len(sys.argv) > 2 else 8,
            3: int(sys.argv[3]) if len(sys.argv) > 3 else 4}
    cal = synth.make_calibration()
    orc = oracle.Oracle(cal)
    gpu = pkg.NpsWf(cal, fit_mode=pkg.FIT_MIGRAD)
    spl = orc.spline_coeffs()
    print("spline coefficients bitwise equal to the oracle's:", np.array_equal(gpu.spline_coeffs(), spl))
    threads = os.cpu_count() or 1
    for cfg in (1, 2, 3):
        if n_ev[cfg] <= 0:
            continue
        ev = synth.generate_host(synth.config_params(cfg), spl, cal, 7_000_000 + cfg, n_ev[cfg], n_threads=threads)
        t0 = time.time()
        ref = orc.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=threads)
        t_or = time.time() - t0
        gpu.analyze(ev["signal"][:1], ev["pres"][:1], ev["corr_time_HMS"][:1])
        gpu.reset_counters()
        t0 = time.time()
        got = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
        t_gpu = time.time() - t0
        c = gpu.counters()
        fitted = (ref["status"] & 28) > 0
        line = []
        for k in ("wfnpulse", "status", "wftime", "wfampl", "chi2", "timewf", "amplwf"):
            line.append("%s %s" % (k, "==" if np.array_equal(got[k], ref[k]) else "DIFF"))
        print("config %d: %d events, %d fits | %s" % (cfg, n_ev[cfg], int(fitted.sum()), ", ".join(line)))
        if not np.array_equal(got["status"], ref["status"]):
            d = got["status"] != ref["status"]
            print("   status differs on %d blocks; e.g. gpu %s oracle %s" % (int(d.sum()), got["status"][d][:8], ref["status"][d][:8]))
        for k in ("wftime", "wfampl", "chi2"):
            if not np.array_equal(got[k], ref[k]):
                d = got[k] != ref[k]
                rel = np.abs(got[k][d] - ref[k][d]) / np.maximum(np.abs(ref[k][d]), 1e-300)
                print("   %s differs on %d values, max rel %.3g, median rel %.3g" % (k, int(d.sum()), rel.max(), np.median(rel)))
        print("   chi2 evaluations: gpu %d oracle %d | verdicts ok1 %d ok2 %d fallback %d | oracle %.2f s (%d threads), gpu call %.3f s "
              "(%.3g fits/s end to end)" % (c["n_fit_evals"], int(ref["ncalls"].sum()), c["n_fit_ok_first"], c["n_fit_ok_retry"],
                                            c["n_fallback"], t_or, threads, t_gpu, fitted.sum() / t_gpu))


if __name__ == "__main__":
    main()
