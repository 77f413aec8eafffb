#!/usr/bin/env python3
"""Verbose stage-by-stage GPU-vs-oracle comparison + first timings (development aid, run under gpurun)."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")


def main():
    cal = synth.make_calibration()
    orc = oracle.Oracle(cal)
    spl = orc.spline_coeffs()
    h = pkg.NpsWf(cal)
    print("devices:", pkg.lib().npswf_device_count())
    x = np.random.default_rng(2).uniform(-3, 3, 100000)
    print("det_exp bit-exact:", np.array_equal(h.debug_exp(x), oracle.det_exp(x)))
    for cfg in (1, 2, 3):
        ev = synth.generate_host(synth.config_params(cfg, absent_frac=0.03), spl, cal, 100 * cfg, 2, n_threads=4)
        sig, pres, corr = ev["signal"], ev["pres"], ev["corr_time_HMS"]
        mf = h.matched_filter(sig, pres)
        nbad = 0
        hists = []
        for b in range(1080):
            if pres[0, b]:
                ref = orc.matched_filter(b, sig[0])[1]
                hists.append(ref)
                if not np.array_equal(mf[0, b], ref):
                    nbad += 1
                    if nbad <= 3:
                        i = np.nonzero(mf[0, b] != ref)[0]
                        print("  MF mismatch cfg", cfg, "b", b, "idx", i[:5], mf[0, b][i[:5]], ref[i[:5]])
        print("cfg%d MF mismatching blocks: %d" % (cfg, nbad))
        hists = np.array(hists[:200], np.float32)
        npk, px, sm, de = h.tspectrum_debug(hists)
        bad = [0, 0, 0, 0]
        for i, hh in enumerate(hists):
            n, pos, s, d = oracle.search_highres(hh.astype(np.float64))
            bad[0] += npk[i] != n
            bad[1] += not np.array_equal(sm[i], s)
            bad[2] += not np.array_equal(de[i], d)
            bad[3] += not (npk[i] == n and np.array_equal(px[i, :n], pos))
            if (not np.array_equal(sm[i], s)) and bad[1] <= 2:
                j = np.nonzero(sm[i] != s)[0]
                print("  smoothed mismatch i", i, "first idx", j[:5], sm[i][j[:3]], s[j[:3]])
            if (not np.array_equal(de[i], d)) and bad[2] <= 2:
                j = np.nonzero(de[i] != d)[0]
                print("  decon mismatch i", i, "first idx", j[:5], de[i][j[:3]], d[j[:3]])
        print("cfg%d tspectrum mismatches [npeaks, smoothed, decon, pos]: %s of %d" % (cfg, bad, len(hists)))
        ok = h.PassClusterThreshold(sig, pres)
        ref_ok = np.array([[orc.pass_cluster_threshold(b, sig[e], pres[e]) for b in range(1080)] for e in range(2)])
        print("cfg%d threshold mismatches: %d (pass frac %.3f)" % (cfg, int((ok != ref_ok).sum()), ref_ok.mean()))
        t0 = time.time()
        ref = orc.analyze_batch(sig, pres, corr, n_threads=os.cpu_count())
        t1 = time.time()
        got = h.analyze(sig, pres, corr)
        t2 = time.time()
        print("cfg%d oracle %.2fs gpu(host api, cold) %.3fs" % (cfg, t1 - t0, t2 - t1))
        print("  wfnpulse equal:", np.array_equal(got["wfnpulse"], ref["wfnpulse"]),
              " status&3 equal:", np.array_equal(got["status"] & 3, ref["status"] & 3))
        if not np.array_equal(got["wfnpulse"], ref["wfnpulse"]):
            i = np.argwhere(got["wfnpulse"] != ref["wfnpulse"])
            print("  first npulse mismatches", i[:5], got["wfnpulse"][tuple(i[0])], ref["wfnpulse"][tuple(i[0])])
        both = ((ref["status"] & 12) > 0) & ((got["status"] & 12) > 0)
        valid = np.arange(12)[None, None, :] < ref["wfnpulse"][..., None]
        d_t = np.where(valid, np.abs(ref["wftime"] - got["wftime"]) / 4.0, 0.0).max(-1)
        d_a = np.where(valid, np.abs(ref["wfampl"] - got["wfampl"]) / np.maximum(np.abs(ref["wfampl"]), 1e-300), 0).max(-1)
        d_c = np.abs(ref["chi2"] - got["chi2"]) / np.maximum(np.abs(ref["chi2"]), 1e-300)
        good = (d_t <= 0.01) & (d_a <= 1e-3) & (d_c <= 1e-3)
        print("  fits: both ok %d | within tol %.5f | gpu status hist ok1/ok2/fb %d/%d/%d | oracle %d/%d/%d" % (
            both.sum(), good[both].mean() if both.any() else 1,
            ((got["status"] & 4) > 0).sum(), ((got["status"] & 8) > 0).sum(), ((got["status"] & 16) > 0).sum(),
            ((ref["status"] & 4) > 0).sum(), ((ref["status"] & 8) > 0).sum(), ((ref["status"] & 16) > 0).sum()))
        if both.any():
            print("  max dt %.3g bin, max dA/A %.3g, max dchi2 %.3g (over both-ok)" % (d_t[both].max(), d_a[both].max(), d_c[both].max()))
        print("  counters", h.counters())
    # ---- timing on resident data (config 2)
    import torch
    dev = torch.device("cuda:0")
    E = 592
    p = synth.config_params(2)
    d_spl = torch.from_numpy(h.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)
    sig = torch.empty((E, 1080, 110), dtype=torch.float64, device=dev)
    pres = torch.empty((E, 1080), dtype=torch.int32, device=dev)
    corr = torch.empty((E,), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    synth.generate_device(p, d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), 0, E, sig.data_ptr(), 0,
                          pres.data_ptr(), corr.data_ptr(), st)
    torch.cuda.synchronize()
    o = dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev),
             wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             status=torch.empty((E, 1080), dtype=torch.uint8, device=dev))

    def run():
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                         o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(),
                         o["timewf"].data_ptr(), o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    h.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    h.sync_device(stream=st)
    print("stage times (5 steps):", h.stage_times())
    nfit = int(((o["status"] & 28) > 0).sum().item())
    print("resident cfg2: %d events %.2f ms/step -> %.3g block-wf/s, %.3g fitted/s (nfit %d, npulse hist %s)" % (
        E, ms, E * 1080 / ms * 1e3, nfit / ms * 1e3, nfit,
        torch.bincount(o["wfnpulse"].flatten().long(), minlength=13).tolist()))
    # device-generated data vs oracle (D2H copy of 2 events)
    sig_h = sig[:2].cpu().numpy(); pres_h = pres[:2].cpu().numpy(); corr_h = corr[:2].cpu().numpy()
    ref = orc.analyze_batch(sig_h, pres_h, corr_h, n_threads=os.cpu_count())
    print("device-generated events: wfnpulse equal", np.array_equal(o["wfnpulse"][:2].cpu().numpy(), ref["wfnpulse"]),
          "lattice ok", np.array_equal(np.round(sig_h / synth.LSB) * synth.LSB, sig_h))


if __name__ == "__main__":
    main()
