#!/usr/bin/env python3
"""How many fits converge within K accepted LM steps (first attempt only)?  Usage: iter_hist.py [events] [config]"""
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
    cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cal = synth.make_calibration()
    dev = torch.device("cuda:0")
    for K in (4, 6, 8, 10, 12, 16, 20, 24, 32, 48, 60):
        h = pkg.NpsWf(cal, fit_max_iter=K, fit_retry_max_iter=1)
        d_spl = torch.from_numpy(h.spline_coeffs()).to(dev)
        d_tref = torch.from_numpy(cal["timeref"]).to(dev)
        d_kap = torch.from_numpy(cal["kappa"]).to(dev)
        sig = torch.empty((E, 1080, 110), dtype=torch.float64, device=dev)
        pres = torch.empty((E, 1080), dtype=torch.int32, device=dev)
        corr = torch.empty((E,), dtype=torch.float64, device=dev)
        synth.generate_device(synth.config_params(cfg), d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), 0, E,
                              sig.data_ptr(), 0, pres.data_ptr(), corr.data_ptr(), 0)
        o = [torch.empty((E, 1080), dtype=torch.int32, device=dev), torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             torch.empty((E, 1080, 12), dtype=torch.float64, device=dev), torch.empty((E, 1080), dtype=torch.float64, device=dev),
             torch.empty((E, 1080), dtype=torch.float64, device=dev), torch.empty((E, 1080), dtype=torch.float64, device=dev),
             torch.empty((E, 1080), dtype=torch.uint8, device=dev)]
        torch.cuda.synchronize()
        h.reset_counters()
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), *[t.data_ptr() for t in o], stream=0)
        h.sync_device(stream=0)
        c = h.counters()
        n = c["n_fit_attempted"]
        print("K=%2d: ok within K %.4f  (not converged %d of %d), accepted steps/fit %.2f" % (
            K, c["n_fit_ok_first"] / n, n - c["n_fit_ok_first"], n, c["n_fit_iterations"] / n))
        del h


if __name__ == "__main__":
    main()
