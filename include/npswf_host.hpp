// npswf_host.hpp — C++ host mirror of the reference's hot-path interface over the C ABI (npswf.h).
//
// The reference (/root/reference/TEST_2.C, "T2") calls analyze(event) once per RDataFrame row
// (T2:1305).  Per-event synchronous calls defeat batching, so the drop-in replaces the
// Define("tuple", analyze)...Snapshot section (T2:1305-1387) with: read + unpack a batch of events,
// one call of npswf::Analyzer::analyze(), fill the WF tree.  Names and sentinels follow the
// reference: wfnpulse, wftime/wfampl (flattened with blockOffset like T2:1294-1295), chi2 = -100,
// timewf/amplwf = -100.  Header-only; link with -lnpswf.  No exceptions cross the C boundary; this
// wrapper turns error codes into std::runtime_error for C++ callers.
#ifndef NPSWF_HOST_HPP
#define NPSWF_HOST_HPP
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "npswf.h"

namespace npswf {

struct EventResult {                      // what analyze() returns per event (T2:1289-1296)
    std::vector<double> chi2, timewf, amplwf;   // [1080]
    std::vector<int32_t> wfnpulse;              // [1080]
    std::vector<double> wfampl, wftime;         // flattened: blocks with pulses only, block order
    std::vector<int32_t> blockOffset;           // [1081]
    std::vector<double> h1time, h2time;         // per pulse with wfampl > 20 of the fitted blocks, block order (T2:988-996)
};

// hcana pulse variables of one event (T2:893-939): what analyze() computes between the unpack and the block loop
struct HcanaPulses {
    double corr_time_HMS = 0.;                  // T2:903
    std::vector<double> Sampampl, Samptime;     // [1080], -100 where the block has no hcana pulse
};
inline HcanaPulses hcana_pulses(int32_t NadcCounter, const double *adcCounter, const double *adcSampPulseTime,
                                const double *adcSampPulseTimeRaw, const double *adcSampPulseAmp, const float *tdcoffset,
                                const float *timemean2)
{
    HcanaPulses r;
    r.Sampampl.resize(NPSWF_NBLOCKS);
    r.Samptime.resize(NPSWF_NBLOCKS);
    if (npswf_hcana_pulses(NadcCounter, adcCounter, adcSampPulseTime, adcSampPulseTimeRaw, adcSampPulseAmp, tdcoffset, timemean2,
                           &r.corr_time_HMS, r.Sampampl.data(), r.Samptime.data()))
        throw std::runtime_error("npswf_hcana_pulses: bad arguments");
    return r;
}

class Analyzer {
public:
    Analyzer(const NpsWfConfig &cfg, const NpsWfCalib &calib) : cortime_(calib.cortime, calib.cortime + NPSWF_NBLOCKS), dt_(cfg.dt)
    {
        if (int rc = npswf_create(&cfg, &calib, &h_)) throw std::runtime_error(std::string("npswf_create: ") + npswf_last_error(nullptr) + " (" + std::to_string(rc) + ")");
    }
    ~Analyzer() { npswf_destroy(h_); }
    Analyzer(const Analyzer &) = delete;
    Analyzer &operator=(const Analyzer &) = delete;

    static NpsWfConfig defaults()
    {
        NpsWfConfig c;
        npswf_default_config(&c);
        return c;
    }

    // analyze() for n_events events.  signal: [E][1080][110] (the `signal` vector of T2:549 per event),
    // pres: [E][1080] (T2:548), corr_time_HMS: [E] (T2:903).
    std::vector<EventResult> analyze(int64_t n_events, const double *signal, const int32_t *pres,
                                     const double *corr_time_HMS)
    {
        // wfampl / wftime arrive already packed the reference's way (npswf_analyze_batch_flat).  The pools hold the
        // 12-pulses-per-block maximum but are left uninitialised, so only the pages the pulses land in are ever touched:
        // one call, no retry with a larger pool (which would run the whole batch -- and count its fits -- twice).
        const size_t nb = (size_t)n_events * NPSWF_NBLOCKS;
        std::vector<int32_t> n(nb), cnt((size_t)n_events);
        std::vector<int64_t> off((size_t)n_events);
        std::vector<double> c(nb), tw(nb), aw(nb);
        std::vector<uint8_t> st(nb);
        const size_t cap = nb * (size_t)NPSWF_MAXWFPULSES;
        std::unique_ptr<double[]> pt(new double[cap ? cap : 1]), pa(new double[cap ? cap : 1]);
        check(npswf_analyze_batch_flat(h_, n_events, signal, pres, corr_time_HMS, n.data(), off.data(), cnt.data(), pt.get(),
                                       pa.get(), (int64_t)cap, c.data(), tw.data(), aw.data(), st.data(), nullptr));
        std::vector<EventResult> out((size_t)n_events);
        for (int64_t e = 0; e < n_events; e++) {
            EventResult &r = out[(size_t)e];
            const size_t o = (size_t)e * NPSWF_NBLOCKS;
            r.chi2.assign(c.begin() + o, c.begin() + o + NPSWF_NBLOCKS);
            r.timewf.assign(tw.begin() + o, tw.begin() + o + NPSWF_NBLOCKS);
            r.amplwf.assign(aw.begin() + o, aw.begin() + o + NPSWF_NBLOCKS);
            r.wfnpulse.assign(n.begin() + o, n.begin() + o + NPSWF_NBLOCKS);
            r.blockOffset.resize(NPSWF_NBLOCKS + 1);
            int32_t run = 0;
            for (int b = 0; b < NPSWF_NBLOCKS; b++) {   // T2:959-961
                r.blockOffset[b] = run;
                run += r.wfnpulse[b];
            }
            r.blockOffset[NPSWF_NBLOCKS] = run;          // T2:1022
            r.wftime.assign(pt.get() + off[e], pt.get() + off[e] + cnt[e]);
            r.wfampl.assign(pa.get() + off[e], pa.get() + off[e] + cnt[e]);
            // T2:988-996 runs for the blocks that passed the cluster threshold (the others `continue` at T2:984):
            // h2time = wftime, h1time = fit parameter - timerefacc + corr_time_HMS/dt = (wftime + cortime) / dt
            // (npswf_event_times), one entry per pulse with wfampl > 20
            for (int b = 0; b < NPSWF_NBLOCKS; b++) {
                if (!(st[o + b] & NPSWF_ST_OKTOFIT)) continue;
                for (int p = 0; p < r.wfnpulse[b]; p++) {
                    const size_t k = (size_t)r.blockOffset[b] + p;
                    if (r.wfampl[k] > 20) {
                        r.h2time.push_back(r.wftime[k]);
                        r.h1time.push_back((r.wftime[k] + (double)cortime_[b]) / dt_);
                    }
                }
            }
        }
        return out;
    }

    // The same, fed with what the reference's analyze receives (T2:540): the packed branch
    // NPS.cal.fly.adcSampWaveform of the events, concatenated, event e = samp[offsets[e] .. offsets[e+1])
    // (offsets = prefix sums of Ndata.NPS.cal.fly.adcSampWaveform).  The unpack of T2:851-889 runs on the device.
    std::vector<EventResult> analyze_packed(int64_t n_events, const double *samp, const int64_t *offsets,
                                            const double *corr_time_HMS)
    {
        Padded p(n_events);
        check(npswf_analyze_batch_packed(h_, n_events, samp, offsets, corr_time_HMS, p.n.data(), p.t.data(), p.a.data(),
                                         p.c.data(), p.tw.data(), p.aw.data(), p.st.data()));
        return flatten(n_events, p);
    }

    // WF-tree diagnostics (T2:1026-1056): ampl[E][1080], enertot[E], integtot[E]
    void diagnostics(int64_t n_events, const double *signal, double *ampl, double *enertot, double *integtot)
    {
        check(npswf_event_diagnostics_batch(h_, n_events, signal, ampl, enertot, integtot));
    }

    // Stage-level mirrors (batched): FindPulsesMF (T2:124), PassClusterThreshold (T2:218), Fitwf (T2:601)
    void FindPulsesMF(int64_t n_events, const double *signal, const int32_t *pres, int32_t *wfnpulse, double *wftime,
                      double *wfampl)
    {
        check(npswf_find_pulses_mf_batch(h_, n_events, signal, pres, wfnpulse, wftime, wfampl));
    }
    void PassClusterThreshold(int64_t n_events, const double *signal, const int32_t *pres, uint8_t *ok)
    {
        check(npswf_pass_cluster_threshold_batch(h_, n_events, signal, pres, ok));
    }
    void Fitwf(int64_t n_events, const double *signal, const double *corr_time_HMS, const uint8_t *fit_mask,
               const int32_t *wfnpulse, double *wftime, double *wfampl, double *chi2, uint8_t *status)
    {
        check(npswf_fitwf_batch(h_, n_events, signal, corr_time_HMS, fit_mask, wfnpulse, wftime, wfampl, chi2, status));
    }
    NpsWfCounters counters()
    {
        NpsWfCounters c;
        check(npswf_get_counters(h_, &c));
        return c;
    }
    // Transport of analyze()'s binary64 traces (npswf_set_host_packing): 0 doubles, 1 automatic, 2 counts whenever lossless.
    void set_host_packing(int mode, int n_threads = 0, double lsb_mV = 0.0) { check(npswf_set_host_packing(h_, mode, n_threads, lsb_mV)); }
    npswf_handle *raw() { return h_; }

private:
    struct Padded {
        std::vector<int32_t> n;
        std::vector<double> t, a, c, tw, aw;
        std::vector<uint8_t> st;
        explicit Padded(int64_t E)
            : n((size_t)E * NPSWF_NBLOCKS), t(n.size() * NPSWF_MAXWFPULSES), a(n.size() * NPSWF_MAXWFPULSES), c(n.size()),
              tw(n.size()), aw(n.size()), st(n.size())
        {
        }
    };
    // padded [E][1080][12] arrays -> the reference's flattened per-event vectors (T2:1289-1296)
    std::vector<EventResult> flatten(int64_t n_events, const Padded &p) const
    {
        std::vector<EventResult> out((size_t)n_events);
        for (int64_t e = 0; e < n_events; e++) {
            EventResult &r = out[(size_t)e];
            const size_t o = (size_t)e * NPSWF_NBLOCKS;
            r.chi2.assign(p.c.begin() + o, p.c.begin() + o + NPSWF_NBLOCKS);
            r.timewf.assign(p.tw.begin() + o, p.tw.begin() + o + NPSWF_NBLOCKS);
            r.amplwf.assign(p.aw.begin() + o, p.aw.begin() + o + NPSWF_NBLOCKS);
            r.wfnpulse.assign(p.n.begin() + o, p.n.begin() + o + NPSWF_NBLOCKS);
            r.blockOffset.resize(NPSWF_NBLOCKS + 1);
            r.wftime.resize((size_t)NPSWF_NBLOCKS * NPSWF_MAXWFPULSES);
            r.wfampl.resize((size_t)NPSWF_NBLOCKS * NPSWF_MAXWFPULSES);
            const int64_t tot = npswf_flatten_event(&p.n[o], &p.t[o * NPSWF_MAXWFPULSES], &p.a[o * NPSWF_MAXWFPULSES],
                                                    r.wftime.data(), r.wfampl.data(), r.blockOffset.data());
            r.wftime.resize((size_t)tot);
            r.wfampl.resize((size_t)tot);
            std::vector<double> h1((size_t)tot + 1), h2((size_t)tot + 1);
            const int64_t nt = npswf_event_times(&p.n[o], &p.t[o * NPSWF_MAXWFPULSES], &p.a[o * NPSWF_MAXWFPULSES], &p.st[o],
                                                 cortime_.data(), dt_, h1.data(), h2.data());     // T2:988-996
            r.h1time.assign(h1.begin(), h1.begin() + (nt > 0 ? nt : 0));
            r.h2time.assign(h2.begin(), h2.begin() + (nt > 0 ? nt : 0));
        }
        return out;
    }
    void check(int rc)
    {
        if (rc) throw std::runtime_error(std::string("npswf: ") + npswf_last_error(h_) + " (" + std::to_string(rc) + ")");
    }
    npswf_handle *h_ = nullptr;
    std::vector<float> cortime_;
    double dt_ = 4.0;
};

}  // namespace npswf
#endif
