// fit_migrad_kernel: Fitwf (T2:601-828) with the reference's own minimiser -- Minuit2 Migrad on numerical
// gradients, strategy 1, retry with strategy 2 from the same seeds, fall back to the TSpectrum values -- instead of
// the Levenberg-Marquardt solver of the other fit kernels.  Same chi2 (ROOT::Fit::Chi2FCN over the BinData of
// T2:680-688 with Err of T2:946-956), same operation order as a scalar x86-64 build without FMA contraction: this
// translation unit is compiled with -fmad=false, divisions and square roots are IEEE, and the 90 chi2 terms are
// summed serially in sample order, so the minimiser sees the function values the reference's Chi2FCN produces and
// takes the same path (checked bit for bit against the CPU oracle's Migrad restatement in tests/).
//
// One warp per fit.  Migrad is sequential, branchy scalar code with a small state (parameters, gradient, P x P
// inverse-Hessian estimate): lane 0 runs it (migrad_core.hpp) on a shared-memory workspace that no other lane
// touches; lanes 1..31 sit in a service loop and, for every chi2 evaluation lane 0 asks for, compute the terms of
// their three samples (spline coefficients of consecutive intervals: coalesced reads); lane 0 adds
// the 90 terms in order.  The two sides meet at __syncwarp() -- a warp barrier does not need its participants to
// arrive from the same instruction.  Jobs are claimed one at a time from the per-multiplicity job list.
#pragma once
#include "common.cuh"
#include "migrad_core.hpp"

namespace npswf {

constexpr int MG_WARPS = 8;
constexpr int MG_THREADS = MG_WARPS * 32;
// list entries: item | N << MG_NSHIFT (the kernel is instantiated per workspace size, not per multiplicity)
constexpr int MG_NSHIFT = 27;

template <int PMAX>
struct alignas(16) MgSmem {
    alignas(16) double tk[96];   // the 90 chi2 terms of the evaluation in flight (read back as 16-byte pairs)
    mg::Work<PMAX> W;
    double pe[PMAX];        // parameters of the evaluation in flight
    double start[PMAX], werr[PMAX];
    volatile int cmd;       // 1: evaluate at pe, 0: the fit is done
    int pad;
};

// Service lanes 1..31 own the 90 chi2 terms, three each: sample k = (lane - 1) + 31 * kk.  Lane 0 computes none, so
// that the divergent halves of the warp never issue the same work twice.
__device__ __forceinline__ int mg_sample(int lane, int kk) { return (lane - 1) + 31 * kk; }

// a service lane's terms of chi2(pe) -> tk
template <int PMAX>
__device__ __forceinline__ void mg_points(MgSmem<PMAX> *sm, int N, int lane, const double *__restrict__ spl,
                                          const double *__restrict__ knots, const double (&y)[3], const double (&w)[3])
{
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
        const int k = mg_sample(lane, kk);
        if (k < mg::FIT_NPT) sm->tk[k] = mg::chi2_term(k, sm->pe, N, spl, knots, y[kk], w[kk]);
    }
}

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// the chi2 functor lane 0 hands to the minimiser
template <int PMAX>
struct MgWarpFcn {
    MgSmem<PMAX> *sm;
    int P;
    int ncalls;
    __device__ __noinline__ double operator()(const double *x)
    {
        ncalls++;
        for (int i = 0; i < P; i++) sm->pe[i] = x[i];
        sm->cmd = 1;
        __syncwarp();                                   // release the service lanes
        __syncwarp();                                   // all 90 terms are in tk
        const uint32_t a = smem_u32(sm->tk);
        double chi2 = 0;
#pragma unroll
        for (int k = 0; k < mg::FIT_NPT / 2; k++) {     // serial sum in sample order, as FitUtil::EvaluateChi2 does
            const double2 v = lds_f64x2(a + 16u * (uint32_t)k);
            chi2 += v.x;
            chi2 += v.y;
        }
        return chi2;
    }
    MG_DEFAULT_PAIR(__device__ __forceinline__)
};

template <int PMAX>
__global__ void __launch_bounds__(MG_THREADS, (PMAX <= 13 ? 3 : 1))
fit_migrad_kernel(const int *__restrict__ job_list, const int *__restrict__ job_count, int *__restrict__ job_next, int list_N,
                  const double *__restrict__ signal, const double *__restrict__ corr_time_HMS, DevCalib cal, KParams kp,
                  double *__restrict__ wftime, double *__restrict__ wfampl, double *__restrict__ chi2_out,
                  double *__restrict__ timewf, double *__restrict__ amplwf, uint8_t *__restrict__ status,
                  DeviceCounters *__restrict__ ctr)
{
    // list_N > 0: every entry of the list is a fit with list_N pulses; list_N == 0: N is packed into the entry
    extern __shared__ __align__(16) unsigned char mg_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    MgSmem<PMAX> *sm = reinterpret_cast<MgSmem<PMAX> *>(mg_smem_raw) + warp;
    const int njobs = *job_count;
    unsigned long long c_ok1 = 0, c_ok2 = 0, c_fb = 0, c_calls = 0, c_att = 0;

    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(job_next, 1);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= njobs) break;
        const int raw = job_list[job];
        const int N = list_N > 0 ? list_N : (raw >> MG_NSHIFT) & 15;
        const long long item = list_N > 0 ? (long long)(raw & (FIT_CONT_RESTART - 1)) : (long long)(raw & ((1 << MG_NSHIFT) - 1));
        const int P = 2 * N + 1;
        const long long e = item / B;
        const int bn = (int)(item % B);
        const double *sig = signal + (size_t)item * T;
        const double *spl = cal.spline + (size_t)bn * (T - 1) * 4;
        const double *knots = cal.knots_x ? cal.knots_x + (size_t)bn * T : nullptr;
        const double tref = cal.timeref[bn];

        MgWarpFcn<PMAX> fcn;
        fcn.sm = sm; fcn.P = P; fcn.ncalls = 0;
        double y[3], w[3];
#pragma unroll
        for (int kk = 0; kk < 3; kk++) {   // BinData (T2:680-688): sample and inverse error
            const int k = mg_sample(lane, kk);
            y[kk] = 0; w[kk] = 0;
            if (lane > 0 && k < mg::FIT_NPT) {
                y[kk] = sig[mg::FIT_X0 + k];
                w[kk] = mg::inv_err(y[kk]);
            }
        }
        const double *seed_t = wftime + (size_t)item * MAXP, *seed_a = wfampl + (size_t)item * MAXP;
        mg::FitOutcome out{0, 0.0, 0};
        __syncwarp();
        if (lane == 0) {
            mg::fit_seeds(sig, tref, seed_t, seed_a, N, sm->start);
            out = mg::fitwf_minimise<PMAX>(fcn, sm->W, N, sm->start, sm->werr);
            sm->cmd = 0;
            __syncwarp();
        } else {
            for (;;) {
                __syncwarp();
                if (sm->cmd == 0) break;
                mg_points<PMAX>(sm, N, lane, spl, knots, y, w);
                __syncwarp();
            }
        }
        __syncwarp();
        const int st = __shfl_sync(0xffffffffu, out.status, 0);
        const double fmin = __shfl_sync(0xffffffffu, out.fmin, 0);
        const int ncalls = __shfl_sync(0xffffffffu, out.ncalls, 0);
        // write-back (T2:774-827)
        const double corr = corr_time_HMS ? corr_time_HMS[e] : 0.0;
        const double cort = (double)cal.cortime[bn];
        const double accdt = kp.timerefacc * kp.dt;
        double out_t = 0, out_a = 0;
        if (lane < N) {
            if (st == NPSWF_ST_FALLBACK) {   // TSpectrum values, time converted to corrected ns (T2:779-790)
                out_t = (seed_t[lane] - tref) * kp.dt + corr - cort - accdt;
                out_a = seed_a[lane];
            } else {                         // T2:796-817
                out_t = sm->W.x[1 + 2 * lane] * kp.dt + corr - cort - accdt;
                out_a = sm->W.x[2 + 2 * lane];
            }
        }
        __syncwarp();
        if (lane < N) {
            wftime[(size_t)item * MAXP + lane] = out_t;
            wfampl[(size_t)item * MAXP + lane] = out_a;
        }
        double bt = out_t, ba = out_a;   // timewf / amplwf: the pulse with the smallest |wftime| (T2:999-1016)
        for (int p = 1; p < N; p++) {
            const double tp = __shfl_sync(0xffffffffu, out_t, p), ap = __shfl_sync(0xffffffffu, out_a, p);
            if (fabs(tp) < fabs(bt)) { bt = tp; ba = ap; }
        }
        if (lane == 0) {
            chi2_out[item] = (st == NPSWF_ST_FALLBACK) ? -100. : fmin / (double)(mg::FIT_NPT - P);   // T2:824-827
            if (timewf) timewf[item] = bt;
            if (amplwf) amplwf[item] = ba;
            if (status) status[item] |= (uint8_t)st;
        }
        if (st == NPSWF_ST_FIT_OK1) c_ok1++;
        else if (st == NPSWF_ST_FIT_OK2) c_ok2++;
        else c_fb++;
        c_calls += (unsigned long long)ncalls;
        c_att++;
        __syncwarp();
    }
    if (ctr && lane == 0) {
        if (c_att) atomicAdd(&ctr->n_fit_attempted, c_att);
        if (c_ok1) atomicAdd(&ctr->n_fit_ok_first, c_ok1);
        if (c_ok2) atomicAdd(&ctr->n_fit_ok_retry, c_ok2);
        if (c_fb) atomicAdd(&ctr->n_fallback, c_fb);
        if (c_calls) atomicAdd(&ctr->n_fit_evals, c_calls);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// fit_migrad_thread_kernel<N>: the same minimisation, ONE THREAD PER FIT, for N = 1..3 pulses on traces that sit on
// the ADC lattice (sample = integer count * lsb, T2:357) -- the bulk of the fits.
//
// Migrad is scalar code; with a warp per fit 31 lanes idle through it and through the serial chi2 sum.  Here every
// lane runs its own minimisation (migrad_core.hpp on a thread-private workspace) and evaluates its own chi2 serially
// in sample order -- the reference's own summation order costs nothing extra -- so a warp advances 32 fits at once
// wherever their control flow agrees: the seed gradient, the first line-search points, MnHesse.  Where it does not
// (line-search tails, different iteration counts) lanes wait for one another at the structured join points.
// Data per thread: the 90 samples as int16 counts in shared memory, transposed [sample][thread]; the sample value is
// count * lsb (exact), its inverse error comes from a table indexed by |count| built with the same inv_err() at
// npswf_create.  A trace with a sample off the lattice (or beyond the table) is handed, untouched, to the
// warp-per-fit kernel above, which takes any doubles.  Both kernels evaluate identical expressions, so which one runs
// a fit cannot be seen in the result.
#ifndef NPSWF_MT_THREADS
#define NPSWF_MT_THREADS 128
#endif
#ifndef NPSWF_MT_MINBLOCKS
#define NPSWF_MT_MINBLOCKS 4
#endif
#ifdef NPSWF_MT_FCN_INLINE
#define MT_FCN_ATTR __forceinline__
#else
#define MT_FCN_ATTR __noinline__
#endif
constexpr int MT_THREADS = NPSWF_MT_THREADS;
constexpr int MT_WTAB = 8192;        // |count| < MT_WTAB: a 12-bit ADC minus its pedestal stays far inside

// the four coefficients of one spline interval (32 bytes, 32-byte aligned) in ONE request: sm_100's 256-bit load.  The
// 32 lanes of a warp sit in different intervals, so every request is its own sector -- two 128-bit loads would put
// twice as many sector requests through L1, which is what bounds this kernel.
__device__ __forceinline__ void ldg_quad(const double *p, double &y, double &b, double &c, double &d)
{
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(y), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ int lds_s16(uint32_t addr)
{
    short v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (int)v;
}

// chi2 of one fit at x, all 90 samples in order, one thread.  Out of line (Migrad calls it from ~20 places) with
// everything it needs in argument registers; tile = shared-space address of this thread's column of counts.
template <int N>
__device__ __noinline__ double mt_chi2(uint32_t tile, const double *__restrict__ wtab, const double *__restrict__ spl, double lsb,
                                       const double *__restrict__ x, int wlow, double w0)
{
    const double p0 = x[0];
    double t[N], A[N];
#pragma unroll
    for (int n = 0; n < N; n++) { t[n] = x[1 + 2 * n]; A[n] = x[2 + 2 * n]; }
    double chi2 = 0;
#pragma unroll 2
    for (int k = 0; k < mg::FIT_NPT; k++) {
        const int c = lds_s16(tile + (uint32_t)k * (2u * MT_THREADS));
        const double y = (double)c * lsb;
        // |count| < wlow: Err is the constant floor sqrt(2.048)/4.096 (T2:952-954), the table entry is w0 -- most samples
        // of a trace are pedestal, and the gather of 32 different table entries is 32 sector requests saved
        const int ac = abs(c);
        const double w = ac < wlow ? w0 : __ldg(wtab + ac);
        const double xk = (double)(mg::FIT_X0 + k);
        double val = p0;
#pragma unroll
        for (int n = 0; n < N; n++) {
            const double dt0 = xk - t[n];
            if (dt0 > 1 && dt0 < mg::FIT_T - 1) {   // T2:629
                const int i = (int)dt0;
                const double delx = dt0 - (double)i;
                double q0, q1, q2, q3;
                ldg_quad(spl + 4 * i, q0, q1, q2, q3);
                val += A[n] * (q0 + delx * (q1 + delx * (q2 + delx * q3)));
            }
        }
        const double tmp = (y - val) * w;
        chi2 += tmp * tmp;
    }
    return chi2;
}

// chi2 at x with x[i] = vp and with x[i] = vm in ONE pass over the samples (every central difference of Migrad and
// MnHesse): sample, weight and the spline terms of the untouched pulses are shared, each of the two sums is built
// with exactly the operations of mt_chi2 in the same order, so the two values carry the same bits as two calls.
template <int N>
__device__ __noinline__ double2 mt_chi2_pair(uint32_t tile, const double *__restrict__ wtab, const double *__restrict__ spl,
                                             double lsb, const double *__restrict__ x, int i, double vp, double vm, int wlow, double w0)
{
    double t[N], A[N];
#pragma unroll
    for (int n = 0; n < N; n++) { t[n] = x[1 + 2 * n]; A[n] = x[2 + 2 * n]; }
    const bool ped = i == 0, is_t = (i & 1) != 0;
    const int m = (i - 1) >> 1;                       // pulse of parameter i (i > 0): t_m = x[1+2m], A_m = x[2+2m]
    const double p1 = ped ? vp : x[0], p2 = ped ? vm : x[0];
    double c1 = 0, c2 = 0;
#pragma unroll 2
    for (int k = 0; k < mg::FIT_NPT; k++) {
        const int c = lds_s16(tile + (uint32_t)k * (2u * MT_THREADS));
        const double y = (double)c * lsb;
        // |count| < wlow: Err is the constant floor sqrt(2.048)/4.096 (T2:952-954), the table entry is w0 -- most samples
        // of a trace are pedestal, and the gather of 32 different table entries is 32 sector requests saved
        const int ac = abs(c);
        const double w = ac < wlow ? w0 : __ldg(wtab + ac);
        const double xk = (double)(mg::FIT_X0 + k);
        double v1 = p1, v2 = p2;
#pragma unroll
        for (int n = 0; n < N; n++) {
            if (is_t && n == m) {                     // the pulse whose time is varied: two spline evaluations
                const double d1 = xk - vp, d2 = xk - vm;
                if (d1 > 1 && d1 < mg::FIT_T - 1) {
                    const int j = (int)d1;
                    const double delx = d1 - (double)j;
                    double q0, q1, q2, q3;
                    ldg_quad(spl + 4 * j, q0, q1, q2, q3);
                    v1 += A[n] * (q0 + delx * (q1 + delx * (q2 + delx * q3)));
                }
                if (d2 > 1 && d2 < mg::FIT_T - 1) {
                    const int j = (int)d2;
                    const double delx = d2 - (double)j;
                    double q0, q1, q2, q3;
                    ldg_quad(spl + 4 * j, q0, q1, q2, q3);
                    v2 += A[n] * (q0 + delx * (q1 + delx * (q2 + delx * q3)));
                }
            } else {
                const double dt0 = xk - t[n];
                if (dt0 > 1 && dt0 < mg::FIT_T - 1) {
                    const int j = (int)dt0;
                    const double delx = dt0 - (double)j;
                    double q0, q1, q2, q3;
                    ldg_quad(spl + 4 * j, q0, q1, q2, q3);
                    const double sv = q0 + delx * (q1 + delx * (q2 + delx * q3));
                    const bool amp = !ped && !is_t && n == m;     // the pulse whose amplitude is varied
                    v1 += (amp ? vp : A[n]) * sv;
                    v2 += (amp ? vm : A[n]) * sv;
                }
            }
        }
        const double r1 = (y - v1) * w, r2 = (y - v2) * w;
        c1 += r1 * r1;
        c2 += r2 * r2;
    }
    return make_double2(c1, c2);
}

template <int N>
struct MgThreadFcn {
    uint32_t tile;          // shared-space address of this thread's column of the count tile: sample k at tile + k * 2 * MT_THREADS
    const double *wtab;     // inverse error by |count|
    const double *spl;
    double lsb, w0;
    int wlow;
    int ncalls;
    // A thread's fit is as slow as 32 fits: one that needs far more evaluations than usual (call limit, the strategy-2
    // retry) would hold its warp -- and, at the end of a short job list, the whole kernel -- for tens of milliseconds.
    // Past `cap` evaluations the functor answers NaN, which ends every loop of the minimiser at once, and the fit is
    // handed to the warp-per-fit kernel, which runs it again from its seeds (same result, a fraction of the latency).
    int cap;
    bool aborted;
    __device__ __forceinline__ double operator()(const double *x)
    {
        ncalls++;
        if (ncalls > cap) { aborted = true; return __longlong_as_double(0x7ff8000000000000LL); }
        return mt_chi2<N>(tile, wtab, spl, lsb, x, wlow, w0);
    }
    __device__ __forceinline__ void pair(double *x, int i, double vp, double vm, double &f1, double &f2)
    {
        ncalls += 2;
        if (ncalls > cap) { aborted = true; f1 = f2 = __longlong_as_double(0x7ff8000000000000LL); return; }
        const double2 r = mt_chi2_pair<N>(tile, wtab, spl, lsb, x, i, vp, vm, wlow, w0);
        f1 = r.x;
        f2 = r.y;
    }
};

__global__ void mg_wtab_kernel(double *wtab, int n, double lsb)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) wtab[c] = mg::inv_err((double)c * lsb);
}

template <int N>
__global__ void __launch_bounds__(MT_THREADS, NPSWF_MT_MINBLOCKS)
fit_migrad_thread_kernel(const int *__restrict__ job_list, const int *__restrict__ job_count, int *__restrict__ job_next,
                         const double *__restrict__ signal, const double *__restrict__ corr_time_HMS, DevCalib cal, KParams kp,
                         double *__restrict__ wftime, double *__restrict__ wfampl, double *__restrict__ chi2_out,
                         double *__restrict__ timewf, double *__restrict__ amplwf, uint8_t *__restrict__ status,
                         DeviceCounters *__restrict__ ctr, const double *__restrict__ wtab, double lsb, int wlow,
                         int *__restrict__ ho_count, int *__restrict__ ho_list)
{
    constexpr int P = 2 * N + 1;
    __shared__ int16_t tile[mg::FIT_NPT * MT_THREADS];
    const int lane = threadIdx.x & 31;
    const int njobs = *job_count;
    unsigned long long c_ok1 = 0, c_ok2 = 0, c_fb = 0, c_calls = 0, c_att = 0;
    mg::Work<P> W;
    double start[P], werr[P];

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(job_next, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= njobs) break;
        const int job = base + lane;
        if (job < njobs) {
            bool continue_next = false;
            const int raw = job_list[job];
            const long long item = (long long)(raw & (FIT_CONT_RESTART - 1));
            const long long e = item / B;
            const int bn = (int)(item % B);
            const double *sig = signal + (size_t)item * T;
            // BinData (T2:680-688) as counts; off the lattice -> the warp-per-fit kernel takes the fit
            bool lattice = true;
            for (int k = 0; k < mg::FIT_NPT; k++) {
                const double y = sig[mg::FIT_X0 + k];
                const double q = rint(y / lsb);
                lattice = lattice && (fabs(q) < (double)MT_WTAB) && (q * lsb == y);
                tile[k * MT_THREADS + threadIdx.x] = (int16_t)(int)q;
            }
            if (!lattice) {
                ho_list[atomicAdd(ho_count, 1)] = raw;
            } else {
                MgThreadFcn<N> fcn;
                fcn.tile = smem_u32(tile + threadIdx.x); fcn.wtab = wtab; fcn.spl = cal.spline + (size_t)bn * (T - 1) * 4;
                fcn.lsb = lsb; fcn.ncalls = 0; fcn.wlow = wlow; fcn.w0 = __ldg(wtab);
                fcn.cap = 60 + 150 * N; fcn.aborted = false;     // ~3x the usual 64 / 123 / 192 evaluations of a first attempt
                const double tref = cal.timeref[bn];
                double *wt = wftime + (size_t)item * MAXP, *wa = wfampl + (size_t)item * MAXP;
                mg::fit_seeds(sig, tref, wt, wa, N, start);
                // first attempt only (Migrad strategy 1, T2:755): a fit that fails it, or runs long, leaves for the warp kernel
#pragma unroll
                for (int i = 0; i < P; i++) werr[i] = (start[i] == 0) ? 0.3 : 0.3 * fabs(start[i]);
                const mg::Result r1 = mg::migrad<P>(fcn, W, P, start, werr, 1, 1000u + 100u * P + 5u * P * P, 0.01);
                if (fcn.aborted || !r1.valid) {
                    ho_list[atomicAdd(ho_count, 1)] = raw;
                    continue_next = true;
                }
                const mg::FitOutcome out{NPSWF_ST_FIT_OK1, r1.fval, r1.ncalls};
                if (!continue_next) {
                // write-back (T2:774-827)
                const double corr = corr_time_HMS ? corr_time_HMS[e] : 0.0;
                const double cort = (double)cal.cortime[bn];
                const double accdt = kp.timerefacc * kp.dt;
                double bt = 0, ba = 0;
#pragma unroll
                for (int p = 0; p < N; p++) {
                    double ot, oa;
                    if (out.status == NPSWF_ST_FALLBACK) {   // TSpectrum values, time converted to corrected ns (T2:779-790)
                        ot = (wt[p] - tref) * kp.dt + corr - cort - accdt;
                        oa = wa[p];
                    } else {                                 // T2:796-817
                        ot = W.x[1 + 2 * p] * kp.dt + corr - cort - accdt;
                        oa = W.x[2 + 2 * p];
                    }
                    wt[p] = ot;
                    wa[p] = oa;
                    if (p == 0 || fabs(ot) < fabs(bt)) { bt = ot; ba = oa; }   // T2:999-1016
                }
                chi2_out[item] = (out.status == NPSWF_ST_FALLBACK) ? -100. : out.fmin / (double)(mg::FIT_NPT - P);   // T2:824-827
                if (timewf) timewf[item] = bt;
                if (amplwf) amplwf[item] = ba;
                if (status) status[item] |= (uint8_t)out.status;
                if (out.status == NPSWF_ST_FIT_OK1) c_ok1++;
                else if (out.status == NPSWF_ST_FIT_OK2) c_ok2++;
                else c_fb++;
                c_calls += (unsigned long long)out.ncalls;
                c_att++;
                }
            }
        }
        __syncwarp();
    }
    // per-warp totals -> one atomic each
    c_att = warp_sum_u64(c_att); c_ok1 = warp_sum_u64(c_ok1); c_ok2 = warp_sum_u64(c_ok2); c_fb = warp_sum_u64(c_fb);
    c_calls = warp_sum_u64(c_calls);
    if (ctr && lane == 0) {
        if (c_att) atomicAdd(&ctr->n_fit_attempted, c_att);
        if (c_ok1) atomicAdd(&ctr->n_fit_ok_first, c_ok1);
        if (c_ok2) atomicAdd(&ctr->n_fit_ok_retry, c_ok2);
        if (c_fb) atomicAdd(&ctr->n_fallback, c_fb);
        if (c_calls) atomicAdd(&ctr->n_fit_evals, c_calls);
    }
}

}  // namespace npswf
