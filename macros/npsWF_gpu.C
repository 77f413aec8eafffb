// npsWF_gpu.C — ROOT-side shim: how npsWF.C / TEST_2.C would call libnpswf.so.
// NOT compiled or tested in this repo's environment (ROOT is not installed); kept deliberately small.
// It replaces only T2:1305-1387 (Define("tuple", analyze) ... Snapshot): events are read in batches,
// unpacked exactly as T2:851-889 does (or handed over packed: npswf::Analyzer::analyze_packed unpacks on the
// device), analysed by ONE call per batch, and written to the WF tree
// with the reference's branch names.  Everything before (chain, calibration loading T2:360-469) and
// after (BuildIndex / CloneTree, T2:1395-1432) stays as in the reference.
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

void npsWF_gpu_event_loop(TTree *T, TTree *WF, const Float_t *tdcoffset, int batch_events = 592)
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
    NpsWfCalib cal{X.data(), Y.data(), timeref, cortime, preswf};
    npswf::Analyzer gpu(cfg, cal);

    TTreeReader rd(T);
    TTreeReaderValue<Int_t> NSamp(rd, "Ndata.NPS.cal.fly.adcSampWaveform");
    TTreeReaderArray<Double_t> Samp(rd, "NPS.cal.fly.adcSampWaveform");
    TTreeReaderArray<Double_t> adcCounter(rd, "NPS.cal.fly.adcCounter");
    TTreeReaderArray<Double_t> pulseTime(rd, "NPS.cal.fly.adcSampPulseTime");
    TTreeReaderArray<Double_t> pulseTimeRaw(rd, "NPS.cal.fly.adcSampPulseTimeRaw");
    TTreeReaderValue<Double_t> evnum(rd, "g.evnum");

    std::vector<double> signal((size_t)batch_events * B * NT), corr(batch_events), evt(batch_events);
    std::vector<int32_t> pres((size_t)batch_events * B);
    std::vector<double> chi2, timewf, amplwf, wfampl, wftime, h2time;
    std::vector<Int_t> wfnpulse;
    Double_t evtOut, corrOut;
    WF->Branch("chi2", &chi2); WF->Branch("amplwf", &amplwf); WF->Branch("timewf", &timewf);
    WF->Branch("wfnpulse", &wfnpulse); WF->Branch("wfampl", &wfampl); WF->Branch("wftime", &wftime);
    WF->Branch("evt", &evtOut); WF->Branch("corr_time_HMS", &corrOut); WF->Branch("h2time", &h2time);

    auto flush = [&](int n) {
        auto res = gpu.analyze(n, signal.data(), pres.data(), corr.data());
        for (int e = 0; e < n; e++) {
            chi2 = res[e].chi2; timewf = res[e].timewf; amplwf = res[e].amplwf;
            wfnpulse.assign(res[e].wfnpulse.begin(), res[e].wfnpulse.end());
            wfampl = res[e].wfampl; wftime = res[e].wftime; h2time = res[e].h2time;
            evtOut = evt[e]; corrOut = corr[e];
            WF->Fill();
        }
    };
    int n = 0;
    while (rd.Next()) {
        double *sig = &signal[(size_t)n * B * NT];
        int32_t *pr = &pres[(size_t)n * B];
        std::fill(sig, sig + B * NT, 0.);
        std::fill(pr, pr + B, 0);
        for (int ns = 0; ns < *NSamp;) {                      // unpack, T2:855-889
            int bloc = (int)Samp[ns++], nsamp = (int)Samp[ns++];
            if (bloc == 2000) bloc = 1080;
            if (bloc == 2001) bloc = 1081;
            if (bloc < 0 || bloc > nslots - 0.5) break;
            if (bloc < B) pr[bloc] = 1;                        // (the reference also writes pres[] out of bounds here)
            for (int it = 0; it < nsamp; it++, ns++)
                if (bloc < B && it < NT) sig[bloc * NT + it] = Samp[ns];
        }
        int c0 = adcCounter.GetSize() ? (int)adcCounter[0] : 0;
        if (c0 == 2000) c0 = 1080;                            // scintillator renumbering, T2:894-897
        if (c0 == 2001) c0 = 1081;
        corr[n] = adcCounter.GetSize() ? pulseTime[0] - pulseTimeRaw[0] / 16. - tdcoffset[c0] : 0.;  // T2:903
        evt[n] = *evnum;
        if (++n == batch_events) { flush(n); n = 0; }
    }
    if (n) flush(n);
}
