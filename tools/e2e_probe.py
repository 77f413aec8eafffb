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
    print("host cores", os.cpu_count(), flush=True)
    hpage = np.array(hs[:1184])   # pageable copy of a quarter of the batch
    hppage = np.array(hp[:1184]); hcpage = np.array(hc[:1184])
    for chunk in (1184,):
        h = pkg.NpsWf(cal, chunk_events=chunk)
        ho = h.alloc_outputs(E, pinned=True)
        ho_q = h.alloc_outputs(1184, pinned=True)
        variants = [("f64 raw", 0, 0, lambda: h.analyze(hs, hp, hc, out=ho))]
        for nt in (4, 8, 16, 32):
            variants.append(("f64 packed x%d" % nt, 2, nt, lambda: h.analyze(hs, hp, hc, out=ho)))
        variants.append(("f64 auto x4", 1, 4, lambda: h.analyze(hs, hp, hc, out=ho)))
        variants.append(("f64 auto x8", 1, 8, lambda: h.analyze(hs, hp, hc, out=ho)))
        variants.append(("f64 auto", 1, 16, lambda: h.analyze(hs, hp, hc, out=ho)))
        variants.append(("i16", 0, 0, lambda: h.analyze_i16(hk, synth.LSB, hp, hc, out=ho)))
        variants.append(("pageable f64 raw (1184 ev)", 0, 0, lambda: h.analyze(hpage, hppage, hcpage, out=ho_q)))
        variants.append(("pageable f64 packed x16 (1184 ev)", 2, 16, lambda: h.analyze(hpage, hppage, hcpage, out=ho_q)))
        for name, mode, nt, fn in variants:
            h.set_host_packing(mode, n_threads=nt)
            fn(); fn()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            dt = (time.perf_counter() - t0) / 3
            ne = 1184 if "1184 ev" in name else E
            print("chunk cap %4d %-34s: %.1f ms/call -> %.1f M block-wf/s  %s" % (chunk, name, dt * 1e3, ne * 1080 / dt / 1e6,
                                                                        h.host_packing_stats()), flush=True)
            if os.environ.get("E2E_STAGES"):
                h.set_profiling(True); h.stage_times(reset=True)
                t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
                print("    profiled call %.1f ms, stage sums %s" % (dt * 1e3, h.stage_times(reset=True)), flush=True)
                h.set_profiling(False)
        del h


if __name__ == "__main__":
    main()
