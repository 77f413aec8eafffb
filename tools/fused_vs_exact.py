#!/usr/bin/env python3
"""Peak search, fused first pass vs the reference's arithmetic only, at scale: device-generated events of the three
BASELINE configurations run through two handles (NPSWF_SEARCH_FUSED = 1, the default, and 0) and every output of
`analyze` compared bit for bit on the device.  Prints one line per configuration and the repeat counters.
Usage: python tools/fused_vs_exact.py [events_cfg1 events_cfg2 events_cfg3] [batch]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import synth  # noqa: E402

pkg = importlib.import_module("nps-waveform-analysis_b200")


def main():
    n_ev = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else [47360, 189440, 47360]
    E = int(sys.argv[4]) if len(sys.argv) > 4 else 4736
    cal = synth.make_calibration()
    os.environ["NPSWF_SEARCH_FUSED"] = "1"
    h1 = pkg.NpsWf(cal)
    os.environ["NPSWF_SEARCH_FUSED"] = "0"
    h0 = pkg.NpsWf(cal)
    del os.environ["NPSWF_SEARCH_FUSED"]
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_spl = torch.from_numpy(h1.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(cal["timeref"]).to(dev)
    d_kap = torch.from_numpy(cal["kappa"]).to(dev)
    sig = torch.empty((E, 1080, 110), dtype=torch.float64, device=dev)
    pres = torch.empty((E, 1080), dtype=torch.int32, device=dev)
    corr = torch.empty((E,), dtype=torch.float64, device=dev)

    def outs():
        return dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev),
                    wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
                    wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
                    chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
                    timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
                    amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
                    status=torch.empty((E, 1080), dtype=torch.uint8, device=dev))

    oa, ob = outs(), outs()

    def run(h, o):
        h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(), o["wftime"].data_ptr(),
                         o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(), o["amplwf"].data_ptr(),
                         o["status"].data_ptr(), stream=st)
        h.sync_device(stream=st)

    h1.search_fused(reset=True)
    for cfg, total in zip((1, 2, 3), n_ev):
        spectra = pulses = differing = 0
        for e0 in range(0, total, E):
            synth.generate_device(synth.config_params(cfg), d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), e0, E,
                                  sig.data_ptr(), 0, pres.data_ptr(), corr.data_ptr(), st)
            run(h1, oa)
            run(h0, ob)
            torch.cuda.synchronize()
            for k in oa:
                va, vb = oa[k].view(torch.uint8), ob[k].view(torch.uint8)
                if not torch.equal(va, vb):
                    differing += int((va != vb).sum().item())
            spectra += E * 1080
            pulses += int(oa["wfnpulse"].sum().item())
        fused, redone = h1.search_fused(reset=True)
        print("config %d: %d events, %d block-waveforms, %d pulses: %d differing output bytes between the fused first pass and the "
              "exact arithmetic; fused pass on %d spectra, %d repeated exactly (%.4f %%)" % (
                  cfg, total, spectra, pulses, differing, fused, redone, 100.0 * redone / max(1, fused)), flush=True)


if __name__ == "__main__":
    main()
