/* Synthetic NPS event generator (test + bench infrastructure; neither oracle nor product).
 *
 * The JLab replay files the reference reads (/root/reference/TEST_2.C:290-301) are not
 * available, so inputs are synthesised as BASELINE.json prescribes:
 *   s[it] = Q( ped + sum_n A_n * S_b(it - tau_n) + N(0, sigma^2) )
 * with S_b the natural cubic spline through the block's reference waveform (the same curve the
 * fit model of T2:621-635 evaluates), Q = rounding to the 12-bit ADC lattice 1000/4096 mV
 * (T2:357), counter-based Philox4x32-10 keyed by (seed, event, block) so shards are independent.
 *
 * One header, compiled twice: g++ -> synth/libnpswf_synth.so (host, for CPU tests and the CPU
 * baseline), nvcc -> synth/libnpswf_synth_cuda.so (device kernel for resident bench batches).
 * Host and device streams are statistically identical, not bit-identical (libm vs CUDA exp/log).
 */
#ifndef NPSWF_SYNTH_H
#define NPSWF_SYNTH_H
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SY_HD __host__ __device__ __forceinline__
#else
#define SY_HD static inline
#endif

#define SY_NTIME 110
#define SY_NBLOCKS 1080
#define SY_MAXPULSES 12
#define SY_LSB (1000.0 / 4096.0) /* ADCtomV, T2:357 */

typedef struct SynthParams {
    uint64_t seed;
    int32_t nmin, nmax;      /* pulses per block drawn uniformly in [nmin, nmax] */
    int32_t amp_mode;        /* 0: A ~ logU[a_lo, a_hi] mV;  1: A = U[a_lo, a_hi] / kappa_b (MF-height targeted) */
    double a_lo, a_hi;
    double tau_lo, tau_hi;   /* pulse peak position relative to timeref[b], bins */
    double min_sep;          /* minimal pulse separation, bins */
    double noise_sigma;      /* mV */
    double ped_lo, ped_hi;   /* mV */
    double absent_frac;      /* fraction of blocks with pres = 0 (signal all zero) */
} SynthParams;

typedef struct SynthCalibView {
    const double *spline;   /* [B][109][4] = y, b, c, d per unit interval */
    const double *timeref;  /* [B] */
    const double *kappa;    /* [B] matched-filter gain for a unit pulse (amp_mode 1) */
} SynthCalibView;

typedef struct SyPhilox { uint32_t c[4]; uint32_t k[2]; } SyPhilox;

SY_HD void sy_philox_round(uint32_t *c, const uint32_t *k)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

/* Philox4x32-10: out[4] = f(counter, key) */
SY_HD void sy_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t key, uint32_t *out)
{
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    for (int r = 0; r < 10; r++) {
        sy_philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

SY_HD double sy_u01(uint32_t a) { return ((double)a + 0.5) * (1.0 / 4294967296.0); } /* (0,1) */

typedef struct SyBlockTruth {
    int32_t present;
    int32_t npulse;
    double ped;
    double amp[SY_MAXPULSES];
    double pos[SY_MAXPULSES]; /* absolute peak position in bins (timeref + tau), ascending */
} SyBlockTruth;

SY_HD double sy_spline(const double *spl_b, double x)
{
    if (!(x >= 0.0) || x > (double)(SY_NTIME - 1)) return 0.0;
    int i = (int)x;
    if (i > SY_NTIME - 2) i = SY_NTIME - 2;
    const double *q = spl_b + 4 * i;
    const double d = x - (double)i;
    return q[0] + d * (q[1] + d * (q[2] + d * q[3]));
}

/* Draw the per-(event, block) truth.  Counter layout: (draw, block, event_lo, stream | event_hi<<8) */
SY_HD void sy_block_truth(const SynthParams *p, const SynthCalibView *cal, uint64_t event, int block, SyBlockTruth *t)
{
    uint32_t r[4];
    const uint32_t elo = (uint32_t)event, ehi = (uint32_t)(event >> 32) << 8;
    sy_philox(0u, (uint32_t)block, elo, ehi | 0u, p->seed, r);
    t->present = (sy_u01(r[3]) >= p->absent_frac) ? 1 : 0;
    t->ped = p->ped_lo + (p->ped_hi - p->ped_lo) * sy_u01(r[0]);
    int n = p->nmin + (int)(sy_u01(r[1]) * (double)(p->nmax - p->nmin + 1));
    if (n > p->nmax) n = p->nmax;
    if (n > SY_MAXPULSES) n = SY_MAXPULSES;
    t->npulse = n;
    /* positions: n sorted uniforms on the shrunk interval, then spread by min_sep */
    double span = (p->tau_hi - p->tau_lo) - (double)(n > 0 ? n - 1 : 0) * p->min_sep;
    if (span < 0) span = 0;
    for (int i = 0; i < n; i += 2) {
        sy_philox(1u + (uint32_t)(i >> 1), (uint32_t)block, elo, ehi | 0u, p->seed, r);
        for (int j = 0; j < 2 && i + j < n; j++) {
            t->pos[i + j] = sy_u01(r[2 * j]) * span;
            const double u = sy_u01(r[2 * j + 1]);
            if (p->amp_mode == 0) t->amp[i + j] = p->a_lo * exp(u * log(p->a_hi / p->a_lo));
            else t->amp[i + j] = (p->a_lo + (p->a_hi - p->a_lo) * u) / cal->kappa[block];
        }
    }
    for (int i = 1; i < n; i++) { /* insertion sort by position, amplitudes follow */
        double ps = t->pos[i], am = t->amp[i];
        int j = i - 1;
        while (j >= 0 && t->pos[j] > ps) { t->pos[j + 1] = t->pos[j]; t->amp[j + 1] = t->amp[j]; j--; }
        t->pos[j + 1] = ps; t->amp[j + 1] = am;
    }
    for (int i = 0; i < n; i++) t->pos[i] += cal->timeref[block] + p->tau_lo + (double)i * p->min_sep;
}

/* One trace (110 samples) of one (event, block), quantised to the ADC lattice; counts (int16) optional. */
SY_HD void sy_block_trace(const SynthParams *p, const SynthCalibView *cal, uint64_t event, int block,
                          const SyBlockTruth *t, double *out_mv, int16_t *out_counts)
{
    const double *spl = cal->spline + (size_t)block * (SY_NTIME - 1) * 4;
    const double tref = cal->timeref[block];
    const uint32_t elo = (uint32_t)event, ehi = (uint32_t)(event >> 32) << 8;
    for (int it0 = 0; it0 < SY_NTIME; it0 += 4) {
        uint32_t r[4];
        sy_philox((uint32_t)(it0 >> 2), (uint32_t)block, elo, ehi | 1u, p->seed, r);
        double g[4];
        for (int h = 0; h < 2; h++) { /* Box-Muller: two normals per uniform pair */
            const double u1 = sy_u01(r[2 * h]), u2 = sy_u01(r[2 * h + 1]);
            const double rad = sqrt(-2.0 * log(u1));
            const double ang = 6.283185307179586476925 * u2;
            g[2 * h] = rad * cos(ang);
            g[2 * h + 1] = rad * sin(ang);
        }
        for (int j = 0; j < 4 && it0 + j < SY_NTIME; j++) {
            const int it = it0 + j;
            double v = 0.0;
            if (t->present) {
                v = t->ped + p->noise_sigma * g[j];
                /* a pulse peaking at pos has shape S(it - (pos - timeref)) */
                for (int n = 0; n < t->npulse; n++) v += t->amp[n] * sy_spline(spl, (double)it - (t->pos[n] - tref));
            }
            const double k = rint(v / SY_LSB);
            if (out_mv) out_mv[it] = k * SY_LSB;
            if (out_counts) out_counts[it] = (int16_t)k;
        }
    }
}

#endif
