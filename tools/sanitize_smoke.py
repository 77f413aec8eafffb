#!/usr/bin/env python3
"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the library once.
Usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")
cal = synth.make_calibration()
spl = oracle.Oracle(cal).spline_coeffs()
h = pkg.NpsWf(cal, chunk_events=148)
for cfg in (2, 3):
    ev = synth.generate_host(synth.config_params(cfg, absent_frac=0.05), spl, cal, 7, 3, n_threads=4, counts=True)
    a = h.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    b = h.analyze_i16(ev["counts"], synth.LSB, ev["pres"], ev["corr_time_HMS"])
    samp, offs = synth.pack_events(ev["signal"], ev["pres"], seed=1)
    c = h.analyze_packed(samp, offs, ev["corr_time_HMS"])
    n, t, amp = h.FindPulsesMF(ev["signal"], ev["pres"])
    ok = h.PassClusterThreshold(ev["signal"], ev["pres"])
    r = h.Fitwf(ev["signal"], ev["corr_time_HMS"], ok & (ev["pres"] == 1), n, t, amp)
    d = h.event_diagnostics(ev["signal"])
    print("cfg", cfg, "fitted", int(((a["status"] & 28) > 0).sum()), "equal i16", all(np.array_equal(a[k], b[k]) for k in a))
hist = np.abs(np.random.default_rng(0).normal(0, 1, (40, 110))).astype(np.float32)
hist[:, :5] = 0; hist[:, 105:] = 0
h.tspectrum_debug(hist)
print("exact ops", h.debug_exact_ops(1_000_000))
print("done")
