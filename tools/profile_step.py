#!/usr/bin/env python3
"""Minimal resident-data step for ncu: W warm-up + K timed steps of the full pipeline on one batch
(config 2, generated on device).  Usage: python tools/profile_step.py [events] [steps] [config] [chunk_events]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    cal = synth.make_calibration()
    h = pkg.NpsWf(cal, chunk_events=chunk)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_spl = torch.from_numpy(h.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)
    sig = torch.empty((E, 1080, 110), dtype=torch.float64, device=dev)
    pres = torch.empty((E, 1080), dtype=torch.int32, device=dev)
    corr = torch.empty((E,), dtype=torch.float64, device=dev)
    synth.generate_device(synth.config_params(cfg), d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), 0, E,
                          sig.data_ptr(), 0, pres.data_ptr(), corr.data_ptr(), st)
    o = dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev),
             wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             status=torch.empty((E, 1080), dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()
    h.set_profiling(True)
    for i in range(1 + K):
        if i == 1:
            h.sync_device(stream=st)
            h.stage_times(reset=True)
            h.reset_counters()
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                         o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(),
                         o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)
    h.sync_device(stream=st)
    t = h.stage_times()
    c = h.counters()
    if os.environ.get("NPSWF_FIT_MODE") == "2":
        r = h.vm_reasons(); print("vm hand-offs by reason [limit/inexact, g2<=0, edm<0, above edm, not descent]:", r[:5], "| other tallies:", r[5:])
    import time
    h.set_profiling(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                         o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(),
                         o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)
    h.sync_device(stream=st)
    wall = (time.perf_counter() - t0) / K * 1e3
    print("wall %.3f ms/step | " % wall, end="")
    print("events %d steps %d cfg %d: per step front %.3f search %.3f fit %.3f ms | fits %d iters/fit %.2f evals/fit %.2f retry %d fb %d" % (
        E, K, cfg, t["front_ms"] / K, t["search_ms"] / K, t["fit_ms"] / K, c["n_fit_attempted"] // K,
        c["n_fit_iterations"] / max(1, c["n_fit_attempted"]), c["n_fit_evals"] / max(1, c["n_fit_attempted"]),
        c["n_fit_ok_retry"], c["n_fallback"]))


if __name__ == "__main__":
    main()
