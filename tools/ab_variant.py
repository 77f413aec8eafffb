#!/usr/bin/env python3
"""A/B aid: one library variant (NPSWF_LIB) per process.  Dumps the outputs of a config-2 and a config-3 batch
(device-generated, FAST mode) to /tmp/ab_<tag>_cfg<k>.npz and prints stage times (profiling hooks on) and the
overlapped wall time per step of 9 472 config-2 events.  tools/ab_compare.py compares the dumps bit for bit.
Usage: NPSWF_LIB=.../libnpswf_<tag>.so python tools/ab_variant.py <tag> [steps]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")


def main():
    tag = sys.argv[1]
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cal = synth.make_calibration()
    h = pkg.NpsWf(cal)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_spl = torch.from_numpy(h.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)

    def batch(E, cfg):
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
        return sig, pres, corr, o

    def run(E, sig, pres, corr, o):
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                         o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(),
                         o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)

    for cfg, E in ((2, 592), (3, 296)):
        sig, pres, corr, o = batch(E, cfg)
        run(E, sig, pres, corr, o)
        h.sync_device(stream=st)
        torch.cuda.synchronize()
        np.savez("/tmp/ab_%s_cfg%d.npz" % (tag, cfg), **{k: v.cpu().numpy() for k, v in o.items()})
        del sig, pres, corr, o
    print("AB %-8s inlined division / square-root chains vs IEEE (mismatches, must be 0 0):" % tag, h.debug_exact_ops(400_000_000, seed=11), flush=True)
    E = 9472
    sig, pres, corr, o = batch(E, 2)
    h.set_profiling(True)
    for i in range(1 + K):
        if i == 1:
            h.sync_device(stream=st)
            h.stage_times(reset=True)
            h.reset_counters()
        run(E, sig, pres, corr, o)
    h.sync_device(stream=st)
    t = h.stage_times()
    c = h.counters()
    fz = h.search_fused() if hasattr(h, "search_fused") else (0, 0)
    h.set_profiling(False)
    torch.cuda.synchronize()
    walls = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            run(E, sig, pres, corr, o)
        e1.record(stream)
        h.sync_device(stream=st)
        torch.cuda.synchronize()
        walls.append(e0.elapsed_time(e1) / K)
    print("AB %-8s cfg2 x %d: front %.3f search %.3f fit %.3f ms | overlapped step %.3f ms (best of 3: %s) | evals/fit %.2f | search fused %d redone %d" % (
        tag, E, t["front_ms"] / K, t["search_ms"] / K, t["fit_ms"] / K, min(walls), " ".join("%.2f" % w for w in walls),
        c["n_fit_evals"] / max(1, c["n_fit_attempted"]), fz[0], fz[1]), flush=True)


if __name__ == "__main__":
    main()
