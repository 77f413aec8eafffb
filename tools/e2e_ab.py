"""A/B probe of the host-buffer entry points (f64 with / without the int16 transport, int16 ABI): best and median of 5 calls."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import synth, oracle
pkg = importlib.import_module("nps-waveform-analysis_b200")
E = 4736
cal = synth.make_calibration()
spl = oracle.Oracle(cal).spline_coeffs()
base = synth.generate_host(synth.config_params(2), spl, cal, 0, 296, n_threads=16, counts=True)
hs = pkg.pinned_empty((E, 1080, 110), np.float64); hk = pkg.pinned_empty((E, 1080, 110), np.int16)
hp = pkg.pinned_empty((E, 1080), np.int32); hc = pkg.pinned_empty((E,), np.float64)
for r in range((E + 295) // 296):
    n = min(296, E - 296 * r)
    hs[296 * r:296 * r + n] = base["signal"][:n]; hk[296 * r:296 * r + n] = base["counts"][:n]
    hp[296 * r:296 * r + n] = base["pres"][:n]; hc[296 * r:296 * r + n] = base["corr_time_HMS"][:n]
h = pkg.NpsWf(cal)
ho = h.alloc_outputs(E, pinned=True)
variants = (("f64 auto", 1, lambda: h.analyze(hs, hp, hc, out=ho)), ("f64 raw", 0, lambda: h.analyze(hs, hp, hc, out=ho)),
                       ("i16", 0, lambda: h.analyze_i16(hk, synth.LSB, hp, hc, out=ho)))
if os.environ.get("E2E_ONLY_AUTO"):
    variants = variants[:1]
for name, mode, fn in variants:
    h.set_host_packing(mode)
    fn(); fn()
    ts = []
    for _ in range(9 if os.environ.get("E2E_ONLY_AUTO") else 5):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(os.environ.get("NPSWF_LIB", "default").split("/")[-1], "ramp=%s %-9s best %.1f ms median %.1f ms -> %.1f M/s" % (os.environ.get("NPSWF_CHUNK_RAMP", "1"), name, min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3, E * 1080 / sorted(ts)[len(ts) // 2] / 1e6), h.host_packing_stats(), flush=True)
