// Fit kernel: Fitwf (T2:601-828) as a batched damped Gauss-Newton / Levenberg-Marquardt solver.
//
// Model (T2:621-635):  f(x) = p0 + sum_n [1 < x - t_n < 109] * A_n * S_b(x - t_n),  x = 10..99,
// S_b = natural cubic spline of the block's reference waveform (coefficients precomputed once per
// block at npswf_create; the reference rebuilds them on every call, T2:612-619).  Objective: the
// chi2 of ROOT::Fit::Chi2FCN over the BinData of T2:680-688 with Err[] of T2:946-956.
// The reference minimises with Minuit2 Migrad on numerical gradients; here the Jacobian is analytic
// (spline value + derivative), the (2N+1)x(2N+1) normal equations are solved by Cholesky, and the
// reference's policy is kept: attempt -> tougher retry from the same seeds -> fall back to the
// TSpectrum values with chi2 = -100 (T2:755-791).
//
// One warp per fit job; lanes own the 90 points (3 each), weighted Jacobian rows are staged in
// shared memory, normal-matrix entries are spread over lanes.  Two instantiations: PMAX = 7
// (N <= 3 pulses, 6.6 KB smem per warp) and PMAX = 25 (N <= 12).  FMA contraction is allowed here:
// agreement with the reference is by tolerance (|dt| <= 0.01 bin, |dA|/A <= 1e-3, chi2 rel 1e-3).
#pragma once
#include "common.cuh"

namespace npswf {

constexpr int FIT_THREADS = 128;
constexpr int FIT_WARPS = FIT_THREADS / 32;
constexpr int FIT_ROWS = 96;  // 90 points padded to 3 per lane

template <int PMAX>
struct FitSmem {
    double J[FIT_ROWS * PMAX];       // weighted Jacobian rows
    double r[FIT_ROWS];              // weighted residuals
    double JtJ[PMAX * (PMAX + 1) / 2];
    double A[PMAX * (PMAX + 1) / 2]; // damped copy -> Cholesky factor (row-packed lower triangle)
    double Jtr[PMAX];
    double dp[PMAX];
    double par[PMAX];
    double trial[PMAX];
    double s2[PMAX / 2];             // second-order term of the exact half-Hessian on (t_n, t_n): -sum r w A_n S''
};

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b

// weighted residual sum of squares at `par` for this lane's points; optional Jacobian rows
template <int PMAX, bool WITH_J>
__device__ __forceinline__ double eval_points(const double *__restrict__ par, int N, int P, int lane,
                                              const double *y, const double *w, const double *__restrict__ spl,
                                              FitSmem<PMAX> *sm, const double *__restrict__ knots = nullptr)
{
    // knots != nullptr: the spline's abscissae are not the sample indices (general interpX, T2:432): the interval
    // comes from a bisection over the block's 110 knots (gsl_interp_bsearch), the offset from the knot itself
    double ss = 0;
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
        const int k = lane + 32 * kk;
        const bool valid = k < NFIT;
        const double x = (double)(MFSTART + k);
        double val = par[0];
        if (WITH_J) sm->J[k * P] = valid ? w[kk] : 0.0;
        for (int n = 0; n < N; n++) {
            const double tn = par[1 + 2 * n], an = par[2 + 2 * n];
            const double d = x - tn;
            double s = 0, ds = 0;
            if (d > 1.0 && d < (double)(T - 1)) {  // T2:629
                int i = (int)d;
                double f = d - (double)i;
                if (knots) {
                    int lo = 0, hi = T - 1;
                    while (hi > lo + 1) {
                        const int m = (hi + lo) >> 1;
                        if (knots[m] > d) hi = m;
                        else lo = m;
                    }
                    i = lo;
                    f = d - knots[lo];
                }
                const double4 q = *reinterpret_cast<const double4 *>(spl + 4 * i);
                s = q.x + f * (q.y + f * (q.z + f * q.w));
                if (WITH_J) ds = q.y + f * (2.0 * q.z + 3.0 * f * q.w);
            }
            val += an * s;
            if (WITH_J) {
                sm->J[k * P + 1 + 2 * n] = valid ? -an * ds * w[kk] : 0.0;
                sm->J[k * P + 2 + 2 * n] = valid ? s * w[kk] : 0.0;
            }
        }
        const double res = valid ? (y[kk] - val) * w[kk] : 0.0;
        if (WITH_J) sm->r[k] = res;
        ss += res * res;
    }
    return warp_sum(ss);
}

// Cholesky solve of (JtJ + lambda*diag) dp = Jtr, executed redundantly by every lane on the
// warp's shared copy (identical values, benign same-value stores).  Returns false if not PD.
// With `newton` the matrix is the exact half-Hessian (see NormalEq in kernel_fit_small.cuh): the (t_n, t_n) term
// comes from sm->s2, the (A_n, t_n) term sum r w S' equals -Jtr[t_n] / A_n.
template <int PMAX>
__device__ __forceinline__ bool damped_solve(FitSmem<PMAX> *sm, int P, double lambda, bool newton)
{
    double *A = sm->A;
    for (int a = 0; a < P; a++)
        for (int b = 0; b <= a; b++) {
            double v = sm->JtJ[tri(a, b)];
            if (newton) {
                if (a == b && (a & 1)) v += sm->s2[(a - 1) / 2];
                if (a == b + 1 && (b & 1)) {
                    const double An = sm->par[a];
                    if (An != 0.0) v -= sm->Jtr[b] / An;
                }
            }
            if (a == b) v += lambda * (fabs(v) + 1e-12);
            A[tri(a, b)] = v;
        }
    __syncwarp();
    for (int a = 0; a < P; a++) {
        for (int b = 0; b <= a; b++) {
            double s = A[tri(a, b)];
            for (int k = 0; k < b; k++) s -= A[tri(a, k)] * A[tri(b, k)];
            if (a == b) {
                if (!(s > 0)) return false;
                A[tri(a, a)] = sqrt(s);
            } else {
                A[tri(a, b)] = s / A[tri(b, b)];
            }
            __syncwarp();
        }
    }
    double *dp = sm->dp;
    for (int a = 0; a < P; a++) {
        double s = sm->Jtr[a];
        for (int k = 0; k < a; k++) s -= A[tri(a, k)] * dp[k];
        dp[a] = s / A[tri(a, a)];
        __syncwarp();
    }
    for (int a = P - 1; a >= 0; a--) {
        double s = dp[a];
        for (int k = a + 1; k < P; k++) s -= A[tri(k, a)] * dp[k];
        dp[a] = s / A[tri(a, a)];
        __syncwarp();
    }
    return true;
}

struct LmOutcome { bool ok; double chi2; int iters; };

// Levenberg-Marquardt from the parameters currently in sm->par (same schedule as the CPU
// prototype in oracle/npswf_oracle.cpp lm_minimise): accept if chi2 does not increase, lambda *0.2
// on accept / *10 on reject, converged when the relative decrease drops below rel_tol or no
// descent step exists any more.
template <int PMAX>
__device__ __forceinline__ LmOutcome lm_warp(FitSmem<PMAX> *sm, int N, int P, int lane, const double *y, const double *w,
                                             const double *__restrict__ spl, int max_iter, double lambda0, double rel_tol,
                                             const double *__restrict__ knots = nullptr)
{
    double lambda = lambda0;
    double chi2 = eval_points<PMAX, false>(sm->par, N, P, lane, y, w, spl, sm, knots);
    bool converged = false, newton = false;   // exact-Hessian steps once an accepted step gains < 5 %
    int it = 0;
    const int ntri = P * (P + 1) / 2;
    for (; it < max_iter; it++) {
        __syncwarp();
        eval_points<PMAX, true>(sm->par, N, P, lane, y, w, spl, sm, knots);
        __syncwarp();
        // normal equations: entries spread over lanes
        for (int idx = lane; idx < ntri + P + N; idx += 32) {
            double s = 0;
            if (idx >= ntri + P) {
                // s2[n] = -A_n sum_k r_k w_k S''(x_k - t_n): the weights are column 0 of J
                const int n = idx - ntri - P;
                const double tn = sm->par[1 + 2 * n], an = sm->par[2 + 2 * n];
                for (int k = 0; k < NFIT; k++) {
                    const double d = (double)(MFSTART + k) - tn;
                    if (d > 1.0 && d < (double)(T - 1)) {
                        int i = (int)d;
                        double f = d - (double)i;
                        if (knots) {
                            int lo = 0, hi = T - 1;
                            while (hi > lo + 1) {
                                const int m = (hi + lo) >> 1;
                                if (knots[m] > d) hi = m;
                                else lo = m;
                            }
                            i = lo;
                            f = d - knots[lo];
                        }
                        const double2 q23 = *reinterpret_cast<const double2 *>(spl + 4 * i + 2);
                        s += sm->r[k] * sm->J[k * P] * (2.0 * q23.x + 6.0 * f * q23.y);
                    }
                }
                sm->s2[n] = -an * s;
            } else if (idx < ntri) {
                int a = 0;
                while (tri(a + 1, 0) <= idx) a++;
                const int b = idx - tri(a, 0);
                for (int k = 0; k < NFIT; k++) s += sm->J[k * P + a] * sm->J[k * P + b];
                sm->JtJ[idx] = s;
            } else {
                const int a = idx - ntri;
                for (int k = 0; k < NFIT; k++) s += sm->J[k * P + a] * sm->r[k];
                sm->Jtr[a] = s;
            }
        }
        __syncwarp();
        bool accepted = false;
        for (int tries = 0; tries < 30 && !accepted; tries++) {
            const bool pd = damped_solve<PMAX>(sm, P, lambda, newton);
            __syncwarp();
            if (!pd) { lambda = fmax(lambda * 10, 1e-6); continue; }
            if (lambda <= 1e-2) {   // predicted-decrease stop (see fit_thread_kernel)
                double pred2 = 0;
                for (int a = 0; a < P; a++) pred2 += sm->Jtr[a] * sm->dp[a];
                if (2.0 * pred2 < rel_tol * (fabs(chi2) + 1e-30)) { converged = true; accepted = true; break; }
            }
            if (lane < P) sm->trial[lane] = sm->par[lane] + sm->dp[lane];
            __syncwarp();
            const double c2 = eval_points<PMAX, false>(sm->trial, N, P, lane, y, w, spl, sm, knots);
            if (c2 <= chi2) {
                const double rel = (chi2 - c2) / (fabs(chi2) + 1e-30);
                if (lane < P) sm->par[lane] = sm->trial[lane];
                chi2 = c2;
                lambda = fmax(lambda * 0.2, 1e-12);
                accepted = true;
                if (rel < 0.05) newton = true;
                if (rel < rel_tol) converged = true;
            } else {
                lambda = fmax(lambda * 10, 1e-6);
            }
            __syncwarp();
        }
        if (!accepted) { converged = isfinite(chi2); break; }   // no descent step left; a chi2 that is not a number is a failed attempt
        if (converged) break;
    }
    return {converged, chi2, it + 1};
}

// grid-stride (by warp) over the fit jobs of one pulse multiplicity N.
template <int PMAX>
__global__ void __launch_bounds__(FIT_THREADS)
fit_kernel(const int *__restrict__ job_list, const int *__restrict__ job_count, int N, const double *__restrict__ signal,
           const double *__restrict__ corr_time_HMS, DevCalib cal, KParams kp, double *__restrict__ wftime,
           double *__restrict__ wfampl, double *__restrict__ chi2_out, double *__restrict__ timewf,
           double *__restrict__ amplwf, uint8_t *__restrict__ status, DeviceCounters *__restrict__ ctr,
           int first_attempt_done = 0, int first_attempt_iters = 0, int *__restrict__ job_next = nullptr)
{
    // job_next: a zeroed cursor -> jobs are claimed one at a time (the lists this kernel gets are short and their
    // fits very unequal, so a fixed deal leaves most warps waiting for a few); null -> fixed deal by warp index
    // first_attempt_done: the list holds fits whose first attempt was run (and exhausted) by fit_thread_kernel;
    // only the retry is left to do here
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FitSmem<PMAX> *sm = reinterpret_cast<FitSmem<PMAX> *>(smem_raw) + warp;
    const int njobs = *job_count;
    const int P = 2 * N + 1;
    const int warps_total = gridDim.x * FIT_WARPS;
    unsigned long long c_ok1 = 0, c_ok2 = 0, c_fb = 0, c_it = 0, c_att = 0;

    for (int job = blockIdx.x * FIT_WARPS + warp;; job += warps_total) {
        if (job_next) {
            int j = 0;
            if (lane == 0) j = atomicAdd(job_next, 1);
            job = __shfl_sync(0xffffffffu, j, 0);
        }
        if (job >= njobs) break;
        const int raw = job_list[job];
        const long long item = raw & (FIT_CONT_RESTART - 1);
        const bool first_done = first_attempt_done && !(raw & FIT_CONT_RESTART);
        const long long e = item / B;
        const int bn = (int)(item % B);
        const double *sig = signal + (size_t)item * T;
        const double *spl = cal.spline + (size_t)bn * (T - 1) * 4;
        const double *knots = cal.knots_x ? cal.knots_x + (size_t)bn * T : nullptr;
        // BinData (T2:680-688) with Err of T2:946-956; stored as inverse error like ROOT::Fit::BinData
        double y[3], w[3];
#pragma unroll
        for (int kk = 0; kk < 3; kk++) {
            const int k = lane + 32 * kk;
            y[kk] = 0; w[kk] = 0;
            if (k < NFIT) {
                const double v = sig[MFSTART + k];
                double er = dsqrt(fabs(dmul(v, 4.096) / 2.)) / 4.096;
                if (er < 1.) er = dsqrt(fabs(1.0 * 4.096 / 2.)) / 4.096;
                y[kk] = v;
                w[kk] = 1.0 / er;
            }
        }
        // seeds (T2:656-677)
        double ped = 0;
        for (int i = 0; i < 20; i++) ped = dadd(ped, sig[i]);
        ped = ped / 20;
        const double tref = cal.timeref[bn];
        double seed_t = 0, seed_a = 0;
        if (lane < N) {
            seed_t = wftime[(size_t)item * MAXP + lane];
            seed_a = wfampl[(size_t)item * MAXP + lane];
        }
        __syncwarp();
        if (lane == 0) sm->par[0] = ped;
        if (lane < N) {
            sm->par[1 + 2 * lane] = dsub(seed_t, tref);
            sm->par[2 + 2 * lane] = seed_a;
        }
        __syncwarp();
        LmOutcome r = {false, 0.0, first_attempt_iters};
        if (!first_done) r = lm_warp<PMAX>(sm, N, P, lane, y, w, spl, kp.fit_max_iter, 1e-3, FIT_REL_TOL, knots);
        int st = 0;
        int iters = r.iters;
        if (r.ok) st = NPSWF_ST_FIT_OK1;
        else {  // retry from the same seeds with a tougher configuration (T2:761-768)
            __syncwarp();
            if (lane == 0) sm->par[0] = ped;
            if (lane < N) {
                sm->par[1 + 2 * lane] = dsub(seed_t, tref);
                sm->par[2 + 2 * lane] = seed_a;
            }
            __syncwarp();
            r = lm_warp<PMAX>(sm, N, P, lane, y, w, spl, kp.fit_retry_max_iter, 1.0, FIT_REL_TOL, knots);
            iters += r.iters;
            if (r.ok) st = NPSWF_ST_FIT_OK2;
        }
        __syncwarp();
        const double corr = corr_time_HMS ? corr_time_HMS[e] : 0.0;
        const double cort = (double)cal.cortime[bn];
        const double accdt = dmul(kp.timerefacc, kp.dt);
        double out_t = 0, out_a = 0;
        if (st == 0) {  // fallback: TSpectrum values, time converted to corrected ns (T2:779-790)
            st = NPSWF_ST_FALLBACK;
            if (lane < N) {
                out_t = dsub(dsub(dadd(dmul(dsub(seed_t, tref), kp.dt), corr), cort), accdt);
                out_a = seed_a;
            }
            if (lane == 0) chi2_out[item] = -100.;
            c_fb++;
        } else {  // T2:796-827
            if (lane < N) {
                const double binOff = sm->par[1 + 2 * lane];
                out_a = sm->par[2 + 2 * lane];
                out_t = dsub(dsub(dadd(dmul(binOff, kp.dt), corr), cort), accdt);
            }
            if (lane == 0) chi2_out[item] = r.chi2 / (double)(NFIT - P);
            if (st == NPSWF_ST_FIT_OK1) c_ok1++;
            else c_ok2++;
        }
        if (lane < N) {
            wftime[(size_t)item * MAXP + lane] = out_t;
            wfampl[(size_t)item * MAXP + lane] = out_a;
        }
        // timewf / amplwf: the pulse with the smallest |wftime| (T2:999-1016)
        double bt = out_t, ba = out_a;
        for (int p = 1; p < N; p++) {
            const double tp = __shfl_sync(0xffffffffu, out_t, p), ap = __shfl_sync(0xffffffffu, out_a, p);
            if (fabs(tp) < fabs(bt)) { bt = tp; ba = ap; }
        }
        if (lane == 0) {
            if (timewf) timewf[item] = bt;
            if (amplwf) amplwf[item] = ba;
            if (status) status[item] |= (uint8_t)st;
        }
        c_it += iters;
        c_att++;
        __syncwarp();
    }
    if (ctr && lane == 0) {
        if (c_att) atomicAdd(&ctr->n_fit_attempted, c_att);
        if (c_ok1) atomicAdd(&ctr->n_fit_ok_first, c_ok1);
        if (c_ok2) atomicAdd(&ctr->n_fit_ok_retry, c_ok2);
        if (c_fb) atomicAdd(&ctr->n_fallback, c_fb);
        if (c_it) atomicAdd(&ctr->n_fit_iterations, c_it);
    }
}

}  // namespace npswf
