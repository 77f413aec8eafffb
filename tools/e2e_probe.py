#!/usr/bin/env python3
"""Host-buffer (end-to-end) timing probe: f64 and int16 ABIs, a few chunk sizes.  Usage: e2e_probe.py [events]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
import oracle  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
    cal = synth.make_calibration()
    spl = oracle.Oracle(cal).spline_coeffs()
    base = synth.generate_host(synth.config_params(2), spl, cal, 0, 296, n_threads=16, counts=True)
    reps = (E + 295) // 296
    hs = pkg.pinned_empty((E, 1080, 110), np.float64)
    hk = pkg.pinned_empty((E, 1080, 110), np.int16)
    hp = pkg.pinned_empty((E, 1080), np.int32)
    hc = pkg.pinned_empty((E,), np.float64)
    for r in range(reps):
        n = min(296, E - 296 * r)
        hs[296 * r:296 * r + n] = base["signal"][:n]; hk[296 * r:296 * r + n] = base["counts"][:n]
        hp[296 * r:296 * r + n] = base["pres"][:n]; hc[296 * r:296 * r + n] = base["corr_time_HMS"][:n]
    for chunk in (296, 444, 592):
        h = pkg.NpsWf(cal, chunk_events=chunk)
        ho = h.alloc_outputs(E, pinned=True)
        for name, fn in (("f64", lambda: h.analyze(hs, hp, hc, out=ho)), ("i16", lambda: h.analyze_i16(hk, synth.LSB, hp, hc, out=ho))):
            fn(); fn()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            dt = (time.perf_counter() - t0) / 3
            print("chunk cap %4d %s: %.1f ms/call -> %.1f M block-wf/s" % (chunk, name, dt * 1e3, E * 1080 / dt / 1e6), flush=True)
            if os.environ.get("E2E_STAGES"):
                h.set_profiling(True); h.stage_times(reset=True)
                t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
                print("    profiled call %.1f ms, stage sums %s" % (dt * 1e3, h.stage_times(reset=True)), flush=True)
                h.set_profiling(False)
        del h


if __name__ == "__main__":
    main()
