// Device build of the synthetic event generator (bench infrastructure). See npswf_synth.h.
// One warp per (event, block): lanes split the 110 samples in groups of 4 (28 groups).
#include "npswf_synth.h"
#include <cuda_runtime.h>

__global__ void synth_kernel(SynthParams p, SynthCalibView cal, int64_t event0, int64_t n_events, double *signal,
                             int16_t *counts, int32_t *pres, double *corr)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_events * SY_NBLOCKS) return;
    const int64_t e = gw / SY_NBLOCKS;
    const int b = (int)(gw % SY_NBLOCKS);
    const uint64_t ev = (uint64_t)(event0 + e);
    SyBlockTruth t;
    sy_block_truth(&p, &cal, ev, b, &t);
    if (lane == 0) {
        if (pres) pres[gw] = t.present;
        if (b == 0 && corr) {
            uint32_t r[4];
            sy_philox(0u, 0xFFFFFFFFu, (uint32_t)ev, ((uint32_t)(ev >> 32) << 8) | 2u, p.seed, r);
            corr[e] = -5.0 + 10.0 * sy_u01(r[0]);
        }
    }
    const double *spl = cal.spline + (size_t)b * (SY_NTIME - 1) * 4;
    const double tref = cal.timeref[b];
    const uint32_t elo = (uint32_t)ev, ehi = (uint32_t)(ev >> 32) << 8;
    if (lane < 28) {
        const int it0 = lane * 4;
        uint32_t r[4];
        sy_philox((uint32_t)lane, (uint32_t)b, elo, ehi | 1u, p.seed, r);
        double g[4];
        for (int h = 0; h < 2; h++) {
            const double u1 = sy_u01(r[2 * h]), u2 = sy_u01(r[2 * h + 1]);
            const double rad = sqrt(-2.0 * log(u1));
            double s, c;
            sincos(6.283185307179586476925 * u2, &s, &c);
            g[2 * h] = rad * c;
            g[2 * h + 1] = rad * s;
        }
        for (int j = 0; j < 4 && it0 + j < SY_NTIME; j++) {
            const int it = it0 + j;
            double v = 0.0;
            if (t.present) {
                v = t.ped + p.noise_sigma * g[j];
                for (int n = 0; n < t.npulse; n++) v += t.amp[n] * sy_spline(spl, (double)it - (t.pos[n] - tref));
            }
            const double k = rint(v / SY_LSB);
            if (signal) signal[(size_t)gw * SY_NTIME + it] = k * SY_LSB;
            if (counts) counts[(size_t)gw * SY_NTIME + it] = (int16_t)k;
        }
    }
}

// All pointers are DEVICE pointers; `stream` is a cudaStream_t (0 = default).
extern "C" int synth_generate_device(const SynthParams *p, const double *d_spline, const double *d_timeref,
                                     const double *d_kappa, int64_t event0, int64_t n_events, double *d_signal,
                                     int16_t *d_counts, int32_t *d_pres, double *d_corr, void *stream)
{
    SynthCalibView cal{d_spline, d_timeref, d_kappa};
    const int64_t warps = n_events * SY_NBLOCKS;
    const int threads = 256;
    const int64_t blocks = (warps * 32 + threads - 1) / threads;
    if (blocks <= 0) return 0;
    synth_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*p, cal, event0, n_events, d_signal, d_counts,
                                                                       d_pres, d_corr);
    return (int)cudaGetLastError();
}
