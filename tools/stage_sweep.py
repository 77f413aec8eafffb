#!/usr/bin/env python3
"""BASELINE configs[4] (SURVEY.md §8d "config 5"): per-stage sweep over the resident batch size.

For each batch size the three stages run serialised (stage profiling on, CUDA events around each stage on its
launching stream) on a device-generated batch; the streaming stages (matched filter + cluster threshold = the
front kernel) are quoted against the HBM roofline, the search and fit stages against the FP64 pipe.

Usage: python tools/stage_sweep.py [config=2] [sizes=64,256,1024,2368,4096,16384] [steps=3]
Prints one JSON line per batch size and a closing table.
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")

B, T, MAXP = 1080, 110, 12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = 6454.9
    try:
        d = json.load(open(p))
        hbm = float(d.get("hbm_gbs", hbm))
    except Exception:
        pass
    return hbm


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "64,256,1024,2368,4096,16384").split(",")]
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    cal = synth.make_calibration()
    h = pkg.NpsWf(cal)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_spl = torch.from_numpy(h.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)
    fp64_peak = h.fp64_peak_gflops()
    hbm_peak = peaks()
    rows = []
    for E in sizes:
        sig = torch.empty((E, B, T), dtype=torch.float64, device=dev)
        pres = torch.empty((E, B), dtype=torch.int32, device=dev)
        corr = torch.empty((E,), dtype=torch.float64, device=dev)
        synth.generate_device(synth.config_params(cfg), d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), 0, E,
                              sig.data_ptr(), 0, pres.data_ptr(), corr.data_ptr(), st)
        o = dict(n=torch.empty((E, B), dtype=torch.int32, device=dev),
                 t=torch.empty((E, B, MAXP), dtype=torch.float64, device=dev),
                 a=torch.empty((E, B, MAXP), dtype=torch.float64, device=dev),
                 c=torch.empty((E, B), dtype=torch.float64, device=dev),
                 tw=torch.empty((E, B), dtype=torch.float64, device=dev),
                 aw=torch.empty((E, B), dtype=torch.float64, device=dev),
                 s=torch.empty((E, B), dtype=torch.uint8, device=dev))

        def step():
            h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["n"].data_ptr(), o["t"].data_ptr(),
                             o["a"].data_ptr(), o["c"].data_ptr(), o["tw"].data_ptr(), o["aw"].data_ptr(),
                             o["s"].data_ptr(), stream=st)

        h.set_profiling(True)
        step()
        h.sync_device(stream=st)
        h.stage_times(reset=True)
        h.reset_counters()
        for _ in range(K):
            step()
        h.sync_device(stream=st)
        tms = h.stage_times()
        c = h.counters()
        h.set_profiling(False)
        step()
        h.sync_device(stream=st)
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        h.sync_device(stream=st)
        wall = (time.perf_counter() - t0) / K * 1e3
        nb = E * B
        front = tms["front_ms"] / K
        search = tms["search_ms"] / K
        fit = tms["fit_ms"] / K
        fits = c["n_fit_attempted"] / K
        # front: 880 B read once per block-waveform + 1 B decision + 4 B minsig/flags
        front_gbs = nb * 885.0 / (front * 1e-3) / 1e9
        row = {"events": E, "front_ms": front, "search_ms": search, "fit_ms": fit, "overlapped_wall_ms": wall,
               "front_block_wf_per_s": nb / (front * 1e-3), "front_hbm_gbs": front_gbs, "front_hbm_frac": front_gbs / hbm_peak,
               "search_block_wf_per_s": nb / (search * 1e-3),
               "search_fp64_frac": (nb * 39000.0 / (search * 1e-3) / 1e9) / (fp64_peak / 2.0),
               "fit_fits_per_s": fits / (fit * 1e-3), "fit_evals_per_fit": c["n_fit_evals"] / max(1, c["n_fit_attempted"]),
               "pipeline_block_wf_per_s": fits / (wall * 1e-3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del sig, pres, corr, o
        torch.cuda.empty_cache()
    print("\nconfig %d, %d steps per size; HBM peak %.1f GB/s, FP64 peak %.0f GFLOP/s (FMA=2)" % (cfg, K, hbm_peak, fp64_peak))
    print("%8s %10s %10s %10s %10s | %12s %8s | %12s %8s | %12s | %12s" % (
        "events", "front ms", "search ms", "fit ms", "wall ms", "front bwf/s", "HBM frac", "search bwf/s", "FP64 fr",
        "fits/s", "pipeline/s"))
    for r in rows:
        print("%8d %10.3f %10.3f %10.3f %10.3f | %12.4g %8.3f | %12.4g %8.3f | %12.4g | %12.4g" % (
            r["events"], r["front_ms"], r["search_ms"], r["fit_ms"], r["overlapped_wall_ms"], r["front_block_wf_per_s"],
            r["front_hbm_frac"], r["search_block_wf_per_s"], r["search_fp64_frac"], r["fit_fits_per_s"],
            r["pipeline_block_wf_per_s"]))


if __name__ == "__main__":
    main()
