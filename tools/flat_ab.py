"""A/B probe: npswf_analyze_batch (padded outputs) against npswf_analyze_batch_flat (pulses packed on the device), f64 host input."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import synth, oracle
pkg = importlib.import_module("nps-waveform-analysis_b200")
E = 4736
cal = synth.make_calibration(); spl = oracle.Oracle(cal).spline_coeffs()
base = synth.generate_host(synth.config_params(2), spl, cal, 0, 296, n_threads=16)
hs = pkg.pinned_empty((E, 1080, 110), np.float64); hp = pkg.pinned_empty((E, 1080), np.int32); hc = pkg.pinned_empty((E,), np.float64)
for r in range(E // 296):
    hs[296 * r:296 * (r + 1)] = base["signal"]; hp[296 * r:296 * (r + 1)] = base["pres"]; hc[296 * r:296 * (r + 1)] = base["corr_time_HMS"]
h = pkg.NpsWf(cal); ho = h.alloc_outputs(E, pinned=True); hf = h.alloc_flat_outputs(E, E * 1080 * 4, pinned=True)
for name, fn in (("padded", lambda: h.analyze(hs, hp, hc, out=ho)), ("flat", lambda: h.analyze_flat(hs, hp, hc, out=hf))):
    fn(); fn(); ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print("%-7s median %.1f ms -> %.1f M/s" % (name, sorted(ts)[2] * 1e3, E * 1080 / sorted(ts)[2] / 1e6), flush=True)
