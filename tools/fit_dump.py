#!/usr/bin/env python3
"""Dump, per fitted block-waveform, the seeds, the CUDA path's fit and the CPU oracle's (Migrad restatement) fit of
seeded synthetic sets into a compressed .npz for offline study of where the two disagree (which fits the Migrad
arbiter of the FAST mode has to take).  Test infrastructure: runs the oracle as the checker.

Usage: python tools/fit_dump.py out.npz [events_cfg2=100] [events_cfg3=60] [seed=9100000]
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
NP = 6


def main():
    out = sys.argv[1]
    n_ev = {2: int(sys.argv[2]) if len(sys.argv) > 2 else 100, 3: int(sys.argv[3]) if len(sys.argv) > 3 else 60}
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 9_100_000
    cal = synth.make_calibration()
    orc = oracle.Oracle(cal)
    gpu = pkg.NpsWf(cal)
    spl = orc.spline_coeffs()
    threads = os.cpu_count() or 1
    rows = {k: [] for k in ("cfg", "event", "bn", "N", "seed_t", "seed_a", "g_t", "g_a", "g_chi2", "g_st", "o_t", "o_a",
                            "o_chi2", "o_st", "o_ncalls", "corr", "trace")}
    for cfg in (2, 3):
        if n_ev[cfg] <= 0:
            continue
        ev = synth.generate_host(synth.config_params(cfg), spl, cal, seed + cfg, n_ev[cfg], n_threads=threads)
        ref = orc.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=threads)
        got = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
        n0, t0, a0 = gpu.FindPulsesMF(ev["signal"], ev["pres"])
        assert np.array_equal(n0, ref["wfnpulse"])
        fitted = ((ref["status"] & 28) > 0) & (ref["wfnpulse"] <= NP)
        e, b = np.nonzero(fitted)
        rows["cfg"].append(np.full(e.size, cfg, np.int8)); rows["event"].append(e.astype(np.int32)); rows["bn"].append(b.astype(np.int16))
        rows["N"].append(ref["wfnpulse"][e, b].astype(np.int8))
        rows["seed_t"].append(t0[e, b, :NP]); rows["seed_a"].append(a0[e, b, :NP])
        rows["g_t"].append(got["wftime"][e, b, :NP]); rows["g_a"].append(got["wfampl"][e, b, :NP])
        rows["g_chi2"].append(got["chi2"][e, b]); rows["g_st"].append(got["status"][e, b])
        rows["o_t"].append(ref["wftime"][e, b, :NP]); rows["o_a"].append(ref["wfampl"][e, b, :NP])
        rows["o_chi2"].append(ref["chi2"][e, b]); rows["o_st"].append(ref["status"][e, b]); rows["o_ncalls"].append(ref["ncalls"][e, b])
        rows["corr"].append(np.asarray(ev["corr_time_HMS"])[e])
        sig = np.asarray(ev["signal"]).reshape(-1, 1080, 110)
        rows["trace"].append(np.round(sig[e, b] * 4.096).astype(np.int16))
        print("config %d: %d events, %d fits dumped" % (cfg, n_ev[cfg], e.size), flush=True)
    np.savez_compressed(out, timeref=np.asarray(cal["timeref"]), cortime=np.asarray(cal["cortime"]),
                        **{k: np.concatenate(v) for k, v in rows.items()})
    print("wrote", out, os.path.getsize(out) / 1e6, "MB")


if __name__ == "__main__":
    main()
