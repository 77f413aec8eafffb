// Host-side callers on either side of the hot path: scalar per-event work of analyze() that
// the reference does between the waveform unpack and the block loop, and after it.  Plain C++, no device code.
//   npswf_hcana_pulses  -- T2:893-939: HMS time correction + the hcana pulse (amplitude, time) closest to the
//                          expected time per block
//   npswf_event_times   -- T2:988-996: the h1time / h2time vectors of an event from the analysis outputs
#include <cmath>
#include <cstdint>
#include "../../include/npswf.h"

extern "C" {

int npswf_hcana_pulses(int32_t n_adc, const double *adcCounter, const double *adcSampPulseTime, const double *adcSampPulseTimeRaw,
                       const double *adcSampPulseAmp, const float *tdcoffset, const float *timemean2, double *corr_time_HMS,
                       double *Sampampl, double *Samptime)
{
    if (n_adc < 0 || (n_adc > 0 && (!adcCounter || !adcSampPulseTime || !adcSampPulseTimeRaw || !adcSampPulseAmp)) || !tdcoffset ||
        !timemean2 || !corr_time_HMS)
        return NPSWF_ERR_ARG;
    const int nblocks = NPSWF_NBLOCKS;
    double corr = 0.;                                              // T2:557
    int32_t npulse[NPSWF_NBLOCKS];
    for (int i = 0; i < nblocks; i++) {
        npulse[i] = 0;                                             // T2:848
        if (Sampampl) Sampampl[i] = -100;                          // T2:569, 571
        if (Samptime) Samptime[i] = -100;
    }
    for (int32_t k = 0; k < n_adc; k++) {
        double counter = adcCounter[k];
        if (counter == 2000) counter = 1080;                       // T2:895-898: the two scintillator channels
        if (counter == 2001) counter = 1081;
        if (k == 0) {
            // T2:903.  The reference indexes tdcoffset[1080] with the counter as it is; a first pulse from a
            // scintillator (1080 / 1081) or from a corrupt counter reads past that array there.  Not replicated:
            // the offset of such a channel is taken as 0.
            const int c = (int)counter;
            const double off = (c >= 0 && c < nblocks) ? (double)tdcoffset[c] : 0.0;
            corr = adcSampPulseTime[k] - (adcSampPulseTimeRaw[k] / 16.) - off;
        }
        if (counter >= 0 && counter < nblocks) {                   // T2:917
            const int c = (int)counter;
            npulse[c] += 1;
            bool take = npulse[c] == 1;                            // T2:921-927
            if (!take && Samptime)                                 // T2:928-937: a later pulse closer to the expected time wins
                take = std::fabs(Samptime[c] - (double)timemean2[c]) > std::fabs(adcSampPulseTime[k] - (double)timemean2[c]);
            if (take) {
                if (Sampampl) Sampampl[c] = adcSampPulseAmp[k];
                if (Samptime) Samptime[c] = adcSampPulseTime[k];
            }
        }
    }
    *corr_time_HMS = corr;
    return 0;
}

int64_t npswf_event_times(const int32_t *wfnpulse, const double *wftime_padded, const double *wfampl_padded, const uint8_t *status,
                          const float *cortime, double dt, double *h1time, double *h2time)
{
    if (!wfnpulse || !wftime_padded || !wfampl_padded || !status || !cortime || !(dt != 0)) return NPSWF_ERR_ARG;
    int64_t n = 0;
    for (int b = 0; b < NPSWF_NBLOCKS; b++) {
        if (!(status[b] & NPSWF_ST_OKTOFIT)) continue;             // T2:980-986: blocks below the cluster threshold are skipped
        const int np = wfnpulse[b] < NPSWF_MAXWFPULSES ? wfnpulse[b] : NPSWF_MAXWFPULSES;
        for (int p = 0; p < np; p++) {
            const double a = wfampl_padded[(size_t)b * NPSWF_MAXWFPULSES + p], t = wftime_padded[(size_t)b * NPSWF_MAXWFPULSES + p];
            if (a > 20) {                                          // T2:990
                if (h2time) h2time[n] = t;                         // T2:992
                // T2:993: finter->GetParameter(1+2p) - timerefacc + corr_time_HMS/dt, where the parameter is the
                // pulse's bin offset (fitted, or the seed when the fit failed).  The corrected time is
                // offset*dt + corr - cortime - timerefacc*dt in both cases (T2:779-790, 812-815), so this is
                // (wftime + cortime) / dt -- the same number up to the rounding of that round trip (~1e-15 relative).
                if (h1time) h1time[n] = (t + (double)cortime[b]) / dt;
                n++;
            }
        }
    }
    return n;
}

}  // extern "C"
