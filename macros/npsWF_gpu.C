// npsWF_gpu.C — ROOT-side shim: how npsWF.C / TEST_2.C would call libnpswf.so.
// NOT compiled or tested in this repo's environment (ROOT is not installed); kept deliberately small.
// It replaces only T2:1305-1387 (Define("tuple", analyze) ... Snapshot): events are read in batches,
// unpacked exactly as T2:851-889 does (or handed over packed: npswf::Analyzer::analyze_packed unpacks on the
// device), analysed by ONE call per batch, and written to the WF tree with the reference's 17 columns (T2:1387):
//   chi2 ampl amplwf wfnpulse Sampampl Samptime timewf enertot integtot pres corr_time_HMS h1time h2time runnum evt
//   wfampl wftime
// Everything before (chain, calibration loading T2:360-469) and after (BuildIndex("runnum","evt") / CloneTree,
// T2:1395-1432) stays as in the reference, so plotstats.C reads the output unchanged.
//
//   root -l -b -q 'npsWF_gpu.C+(run, seg)'      with   gSystem->Load("libnpswf.so") beforehand
#include <vector>
#include "TFile.h"
#include "TTree.h"
#include "TTreeReader.h"
#include "TTreeReaderArray.h"
#include "TTreeReaderValue.h"
#include "../include/npswf_host.hpp"

// the reference's globals (T2:74-85), filled by the unchanged loader code of T2:360-469
extern std::vector<std::vector<double>> interpX, interpY;
extern Double_t timeref[1080];
extern Float_t cortime[1080];
extern Int_t preswf[1080];
extern Double_t timerefacc;

// tdcoffset: T2:368-375; timemean2: T2:526-529; fit_mode: NPSWF_FIT_MIGRAD reproduces Minuit2's path bit for bit,
// NPSWF_FIT_VM follows it with analytic derivatives (~4x faster), NPSWF_FIT_FAST (Levenberg-Marquardt) is ~8x faster
void npsWF_gpu_event_loop(TTree *T, TTree *WF, const Float_t *tdcoffset, const Float_t *timemean2, int batch_events = 592,
                          int fit_mode = NPSWF_FIT_MIGRAD)
{
    const int B = NPSWF_NBLOCKS, NT = NPSWF_NTIME, nslots = 1104;
    std::vector<double> X(B * NT), Y(B * NT);
    for (int b = 0; b < B; b++)
        for (int it = 0; it < NT; it++) {
            X[b * NT + it] = preswf[b] ? interpX[b][it] : it;
            Y[b * NT + it] = preswf[b] ? interpY[b][it] : 0.;
        }
    NpsWfConfig cfg = npswf::Analyzer::defaults();
    cfg.timerefacc = timerefacc;
    cfg.fit_mode = fit_mode;
    NpsWfCalib cal{X.data(), Y.data(), timeref, cortime, preswf};
    npswf::Analyzer gpu(cfg, cal);

    TTreeReader rd(T);
    TTreeReaderValue<Int_t> NSamp(rd, "Ndata.NPS.cal.fly.adcSampWaveform");
    TTreeReaderArray<Double_t> Samp(rd, "NPS.cal.fly.adcSampWaveform");
    TTreeReaderArray<Double_t> adcCounter(rd, "NPS.cal.fly.adcCounter");
    TTreeReaderArray<Double_t> pulseTime(rd, "NPS.cal.fly.adcSampPulseTime");
    TTreeReaderArray<Double_t> pulseTimeRaw(rd, "NPS.cal.fly.adcSampPulseTimeRaw");
    TTreeReaderArray<Double_t> pulseAmp(rd, "NPS.cal.fly.adcSampPulseAmp");
    TTreeReaderValue<Double_t> evnum(rd, "g.evnum");
    TTreeReaderValue<Double_t> runnumIn(rd, "g.runnum");

    std::vector<double> signal((size_t)batch_events * B * NT), corr(batch_events), evt(batch_events), rn(batch_events);
    std::vector<double> amplAll((size_t)batch_events * B), enerAll(batch_events), integAll(batch_events);
    std::vector<int32_t> presAll((size_t)batch_events * B);
    std::vector<npswf::HcanaPulses> hc(batch_events);
    // the 17 branches of T2:1387
    std::vector<double> chi2, ampl, amplwf, Sampampl, Samptime, timewf, h1time, h2time, wfampl, wftime;
    std::vector<Int_t> wfnpulse, pres;
    Double_t enertot, integtot, corrOut, runnum, evtOut;
    WF->Branch("chi2", &chi2); WF->Branch("ampl", &ampl); WF->Branch("amplwf", &amplwf); WF->Branch("wfnpulse", &wfnpulse);
    WF->Branch("Sampampl", &Sampampl); WF->Branch("Samptime", &Samptime); WF->Branch("timewf", &timewf);
    WF->Branch("enertot", &enertot); WF->Branch("integtot", &integtot); WF->Branch("pres", &pres);
    WF->Branch("corr_time_HMS", &corrOut); WF->Branch("h1time", &h1time); WF->Branch("h2time", &h2time);
    WF->Branch("runnum", &runnum); WF->Branch("evt", &evtOut); WF->Branch("wfampl", &wfampl); WF->Branch("wftime", &wftime);

    auto flush = [&](int n) {
        auto res = gpu.analyze(n, signal.data(), presAll.data(), corr.data());                     // T2:942-1022
        gpu.diagnostics(n, signal.data(), amplAll.data(), enerAll.data(), integAll.data());        // T2:1026-1056
        for (int e = 0; e < n; e++) {
            chi2 = res[e].chi2; timewf = res[e].timewf; amplwf = res[e].amplwf;
            wfnpulse.assign(res[e].wfnpulse.begin(), res[e].wfnpulse.end());
            wfampl = res[e].wfampl; wftime = res[e].wftime; h1time = res[e].h1time; h2time = res[e].h2time;
            ampl.assign(amplAll.begin() + (size_t)e * B, amplAll.begin() + (size_t)(e + 1) * B);
            enertot = enerAll[e]; integtot = integAll[e];
            pres.assign(presAll.begin() + (size_t)e * B, presAll.begin() + (size_t)(e + 1) * B);
            Sampampl = hc[e].Sampampl; Samptime = hc[e].Samptime; corrOut = corr[e];
            runnum = rn[e]; evtOut = evt[e];
            WF->Fill();
        }
    };
    int n = 0;
    while (rd.Next()) {
        double *sig = &signal[(size_t)n * B * NT];
        int32_t *pr = &presAll[(size_t)n * B];
        std::fill(sig, sig + B * NT, 0.);
        std::fill(pr, pr + B, 0);
        if (*NSamp <= nslots * (NT + 2)) {                    // T2:830-836: longer events are not analysed
            for (int ns = 0; ns + 1 < *NSamp;) {              // unpack, T2:855-889
                int bloc = (int)Samp[ns++], nsamp = (int)Samp[ns++];
                if (bloc == 2000) bloc = 1080;
                if (bloc == 2001) bloc = 1081;
                if (bloc < 0 || bloc > nslots - 0.5) break;
                if (bloc < B) pr[bloc] = 1;                    // (the reference also writes pres[] out of bounds here)
                for (int it = 0; it < nsamp; it++, ns++)
                    if (bloc < B && it < NT && ns < *NSamp) sig[bloc * NT + it] = Samp[ns];
            }
        }
        // HMS time correction + hcana pulse closest to the expected time per block, T2:893-939
        const bool any = adcCounter.GetSize() > 0;
        hc[n] = npswf::hcana_pulses((int32_t)adcCounter.GetSize(), any ? &adcCounter[0] : nullptr, any ? &pulseTime[0] : nullptr,
                                    any ? &pulseTimeRaw[0] : nullptr, any ? &pulseAmp[0] : nullptr,
                                    tdcoffset, timemean2);
        corr[n] = hc[n].corr_time_HMS;
        evt[n] = *evnum;
        rn[n] = *runnumIn;
        if (++n == batch_events) { flush(n); n = 0; }
    }
    if (n) flush(n);
}
