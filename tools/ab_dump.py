"""A/B aid: dump status / chi2 / wfnpulse of 592 device-generated config-3 events for the library selected by NPSWF_LIB."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch, synth
pkg = importlib.import_module("nps-waveform-analysis_b200")
tag = sys.argv[1]; E = 592
cal = synth.make_calibration(); h = pkg.NpsWf(cal)
dev = torch.device("cuda:0"); stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
d_spl = torch.from_numpy(h.spline_coeffs()).to(dev); d_tref = torch.from_numpy(cal["timeref"]).to(dev); d_kap = torch.from_numpy(cal["kappa"]).to(dev)
sig = torch.empty((E, 1080, 110), dtype=torch.float64, device=dev); pres = torch.empty((E, 1080), dtype=torch.int32, device=dev); corr = torch.empty((E,), dtype=torch.float64, device=dev)
synth.generate_device(synth.config_params(3), d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), 0, E, sig.data_ptr(), 0, pres.data_ptr(), corr.data_ptr(), st)
o = dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev), wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
         wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev), chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
         timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev), amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
         status=torch.empty((E, 1080), dtype=torch.uint8, device=dev))
h.reset_counters()
h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(), o["wftime"].data_ptr(), o["wfampl"].data_ptr(),
                 o["chi2"].data_ptr(), o["timewf"].data_ptr(), o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)
h.sync_device(stream=st); torch.cuda.synchronize()
print(tag, {k: v for k, v in h.counters().items() if "fit" in k or "fallback" in k})
np.savez("gpurun_out/ab_%s.npz" % tag, status=o["status"].cpu().numpy(), chi2=o["chi2"].cpu().numpy(), n=o["wfnpulse"].cpu().numpy())
