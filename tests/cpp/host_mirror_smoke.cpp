// Compiles include/npswf_host.hpp against libnpswf.so and exercises the C++ mirror without a GPU:
// handle creation + derived calibration work on the host, compute calls must fail loudly.
#include <cmath>
#include <cstdio>
#include <vector>
#include "npswf_host.hpp"

int main()
{
    const int B = NPSWF_NBLOCKS, T = NPSWF_NTIME;
    std::vector<double> X(B * T), Y(B * T), tref(B, 35.0);
    std::vector<float> cort(B, 0.5f);
    std::vector<int32_t> preswf(B, 1);
    for (int b = 0; b < B; b++)
        for (int it = 0; it < T; it++) {
            X[b * T + it] = it;
            const double u = it - 30.0;
            Y[b * T + it] = u > 0 ? u * u * std::exp(-u / 3.0) / 4.872 : 0.0;
        }
    for (int b = 0; b < B; b++) {  // timeref = x of the maximum sample (T2:434-438)
        int im = 0;
        for (int it = 0; it < T; it++) if (Y[b * T + it] > Y[b * T + im]) im = it;
        tref[b] = im;
    }
    NpsWfConfig cfg = npswf::Analyzer::defaults();
    NpsWfCalib cal{X.data(), Y.data(), tref.data(), cort.data(), preswf.data()};
    npswf::Analyzer an(cfg, cal);
    std::vector<double> mfy(B * NPSWF_MFWIDTH), mfi(B);
    if (npswf_get_mf_calib(an.raw(), mfy.data(), mfi.data()) != 0 || !(mfi[0] > 0)) return 2;
    if (npswf_device_count() > 0) {
        // GPU present: the mirror's analyze() (flat outputs packed on the device) must equal the padded C entry point
        // followed by the host-side packing npswf_flatten_event, event by event
        const int E = 5;
        const double lsb = 1000.0 / 4096;
        std::vector<double> sig((size_t)E * B * T), corr(E);
        std::vector<int32_t> pres((size_t)E * B, 1);
        for (int e = 0; e < E; e++) {
            corr[e] = 0.25 * e;
            for (int b = 0; b < B; b++)
                for (int it = 0; it < T; it++) {
                    const int sh = (b * 7 + e * 3) % 11 - 5;
                    const int src = it - sh;
                    double v = 0.4 * (((b * 131 + it * 17 + e * 29) % 7) - 3);                      // pedestal wiggle
                    if ((b + e) % 3 != 2 && src >= 0 && src < T) v += (15.0 + (b % 23)) * Y[b * T + src];  // one pulse
                    if ((b + e) % 5 == 0 && src + 14 >= 0 && src + 14 < T) v += 9.0 * Y[b * T + src + 14]; // pile-up
                    sig[((size_t)e * B + b) * T + it] = std::nearbyint(v / lsb) * lsb;
                }
        }
        auto res = an.analyze(E, sig.data(), pres.data(), corr.data());
        const size_t nb = (size_t)E * B;
        std::vector<int32_t> n(nb);
        std::vector<double> t(nb * NPSWF_MAXWFPULSES), a(nb * NPSWF_MAXWFPULSES), c(nb), tw(nb), aw(nb);
        std::vector<uint8_t> st(nb);
        if (npswf_analyze_batch(an.raw(), E, sig.data(), pres.data(), corr.data(), n.data(), t.data(), a.data(), c.data(),
                                tw.data(), aw.data(), st.data()) != 0) return 4;
        long long pulses = 0;
        for (int e = 0; e < E; e++) {
            std::vector<double> ft((size_t)B * NPSWF_MAXWFPULSES), fa(ft.size());
            std::vector<int32_t> bo(B + 1);
            const size_t o = (size_t)e * B;
            const int64_t tot = npswf_flatten_event(&n[o], &t[o * NPSWF_MAXWFPULSES], &a[o * NPSWF_MAXWFPULSES], ft.data(),
                                                    fa.data(), bo.data());
            const npswf::EventResult &r = res[e];
            if ((int64_t)r.wftime.size() != tot || (int64_t)r.wfampl.size() != tot) return 5;
            for (int64_t i = 0; i < tot; i++)
                if (r.wftime[i] != ft[i] || r.wfampl[i] != fa[i]) return 6;
            for (int b = 0; b <= B; b++)
                if (r.blockOffset[b] != bo[b]) return 7;
            for (int b = 0; b < B; b++)
                if (r.wfnpulse[b] != n[o + b] || r.chi2[b] != c[o + b] || r.timewf[b] != tw[o + b] || r.amplwf[b] != aw[o + b]) return 8;
            // h2time: times of the fitted blocks' pulses with amplitude > 20 (T2:987-992)
            std::vector<double> h2;
            for (int b = 0; b < B; b++)
                if (st[o + b] & NPSWF_ST_OKTOFIT)
                    for (int p = 0; p < n[o + b]; p++)
                        if (a[(o + b) * NPSWF_MAXWFPULSES + p] > 20) h2.push_back(t[(o + b) * NPSWF_MAXWFPULSES + p]);
            if (h2 != r.h2time || h2.empty()) return 10;
            pulses += tot;
        }
        std::printf("gpu present: mirror analyze() == C entry point + flatten on %d events, %lld pulses\n", E, pulses);
        return pulses > 1000 ? 0 : 9;
    }
    std::vector<double> sig((size_t)B * T, 0.0), corr(1, 0.0);
    std::vector<int32_t> pres(B, 1);
    int refused = 0;
    try {
        an.analyze(1, sig.data(), pres.data(), corr.data());
    } catch (const std::runtime_error &e) {
        std::printf("expected failure: %s\n", e.what());
        refused++;
    }
    std::vector<double> samp = {3, 110};
    samp.resize(112, 1.0);
    const int64_t offs[2] = {0, 112};
    try {
        an.analyze_packed(1, samp.data(), offs, corr.data());
    } catch (const std::runtime_error &) {
        refused++;
    }
    std::vector<double> ampl(B), et(1), it(1);
    try {
        an.diagnostics(1, sig.data(), ampl.data(), et.data(), it.data());
    } catch (const std::runtime_error &) {
        refused++;
    }
    return refused == 3 ? 0 : 3;  // a CPU fallback would be a bug
}
