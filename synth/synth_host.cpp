// Host build of the synthetic event generator (test + bench infrastructure). See npswf_synth.h.
#include "npswf_synth.h"
#include <atomic>
#include <thread>
#include <vector>

static void gen_event(const SynthParams *p, const SynthCalibView *cal, uint64_t ev, double *sig, int16_t *cnt,
                      int32_t *pres, double *corr, int32_t *tn, double *tpos, double *tamp, double *tped)
{
    uint32_t r[4];
    sy_philox(0u, 0xFFFFFFFFu, (uint32_t)ev, ((uint32_t)(ev >> 32) << 8) | 2u, p->seed, r);
    if (corr) *corr = -5.0 + 10.0 * sy_u01(r[0]);  // corr_time_HMS ~ U[-5,5] ns
    for (int b = 0; b < SY_NBLOCKS; b++) {
        SyBlockTruth t;
        sy_block_truth(p, cal, ev, b, &t);
        sy_block_trace(p, cal, ev, b, &t, sig ? sig + (size_t)b * SY_NTIME : nullptr,
                       cnt ? cnt + (size_t)b * SY_NTIME : nullptr);
        if (pres) pres[b] = t.present;
        if (tn) tn[b] = t.present ? t.npulse : 0;
        if (tped) tped[b] = t.ped;
        for (int n = 0; n < SY_MAXPULSES; n++) {
            if (tpos) tpos[(size_t)b * SY_MAXPULSES + n] = (n < t.npulse) ? t.pos[n] : -999.0;
            if (tamp) tamp[(size_t)b * SY_MAXPULSES + n] = (n < t.npulse) ? t.amp[n] : -999.0;
        }
    }
}

extern "C" int synth_generate_host(const SynthParams *p, const double *spline, const double *timeref,
                                   const double *kappa, int64_t event0, int64_t n_events, double *signal,
                                   int16_t *counts, int32_t *pres, double *corr_time_HMS, int32_t *truth_n,
                                   double *truth_pos, double *truth_amp, double *truth_ped, int n_threads)
{
    SynthCalibView cal{spline, timeref, kappa};
    if (n_threads < 1) n_threads = 1;
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t e = next.fetch_add(1);
            if (e >= n_events) break;
            const size_t eb = (size_t)e * SY_NBLOCKS;
            gen_event(p, &cal, (uint64_t)(event0 + e), signal ? signal + eb * SY_NTIME : nullptr,
                      counts ? counts + eb * SY_NTIME : nullptr, pres ? pres + eb : nullptr,
                      corr_time_HMS ? corr_time_HMS + e : nullptr, truth_n ? truth_n + eb : nullptr,
                      truth_pos ? truth_pos + eb * SY_MAXPULSES : nullptr,
                      truth_amp ? truth_amp + eb * SY_MAXPULSES : nullptr, truth_ped ? truth_ped + eb : nullptr);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    return 0;
}
