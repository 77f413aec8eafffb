// Register-resident fit kernel for N = 1..3 pulses (P = 3, 5, 7 parameters): the common case.
//
// Same objective / policy as kernel_fit.cuh (Fitwf, T2:601-828), different mapping: a sub-warp group
// of GROUP = 8 (N <= 2) or 16 (N = 3) lanes owns one fit (4 or 2 fits per warp).  Lane g of a group owns
// points k = g + GROUP j (90 points), so the group's loads of the trace are contiguous; y and 1/err stay
// in registers for the whole fit; each lane accumulates its partial normal equations
// (P(P+1)/2 + P + 1 doubles) in registers, a 3-step xor-shuffle butterfly sums them inside the group,
// and every lane solves the tiny damped system redundantly (fully unrolled Cholesky, no shared
// memory, no barriers).  With SPEC the trial point is evaluated together with its Jacobian, so an
// accepted step costs one pass over the points instead of two.
#pragma once
#include "common.cuh"

namespace npswf {

constexpr int FS_THREADS = 128;
constexpr int FS_MINB1 = 3, FS_MINB2 = 2, FS_MINB3 = 2;  // min resident CTAs per SM (register budget)

// Normal equations of one fit.  H = J^T J is the Gauss-Newton matrix; s1 / s2 are the second-order terms of the
// exact half-Hessian of chi2 that involve the pulse times (the model is linear in p0 and A):
//   H_tt(n)  -=  sum r w A_n S''(x - t_n)   -> s2[n]        H_tA(n)  +=  sum r w S'(x - t_n)   -> s1[n]
// (r = weighted residual).  They are always accumulated and only used once a fit has entered its slowly
// converging phase (see solve_damped / the `newton` flag): with a large residual -- fewer pulses found than are
// present -- Gauss-Newton converges linearly and needs 20-60 steps, the exact Hessian needs 2-4.
template <int P>
struct NormalEq {
    double H[P * (P + 1) / 2];  // lower triangle, row-packed
    double g[P];
    double c2;
    double s1[(P - 1) / 2], s2[(P - 1) / 2];
};

template <int P, int GROUP>
__device__ __forceinline__ void group_reduce(NormalEq<P> &ne)
{
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < P * (P + 1) / 2; i++) ne.H[i] += __shfl_xor_sync(0xffffffffu, ne.H[i], o);
#pragma unroll
        for (int i = 0; i < P; i++) ne.g[i] += __shfl_xor_sync(0xffffffffu, ne.g[i], o);
        ne.c2 += __shfl_xor_sync(0xffffffffu, ne.c2, o);
#pragma unroll
        for (int i = 0; i < (P - 1) / 2; i++) {
            ne.s1[i] += __shfl_xor_sync(0xffffffffu, ne.s1[i], o);
            ne.s2[i] += __shfl_xor_sync(0xffffffffu, ne.s2[i], o);
        }
    }
}

// One pass over this lane's points at parameters `par`: chi2 and (WITH_J) the normal equations.
template <int N, int GROUP, bool WITH_J>
__device__ __forceinline__ void eval_group(const double (&par)[2 * N + 1], const double (&y)[96 / GROUP],
                                           const double (&w)[96 / GROUP], int g, const double4 *__restrict__ spl4,
                                           NormalEq<2 * N + 1> &ne)
{
    constexpr int P = 2 * N + 1;
    constexpr int PTS = 96 / GROUP;
#pragma unroll
    for (int i = 0; i < P * (P + 1) / 2; i++) ne.H[i] = 0;
#pragma unroll
    for (int i = 0; i < P; i++) ne.g[i] = 0;
    ne.c2 = 0;
#pragma unroll
    for (int n = 0; n < N; n++) { ne.s1[n] = 0; ne.s2[n] = 0; }
#pragma unroll
    for (int j = 0; j < PTS; j++) {
        const int k = g + GROUP * j;
        const double x = (double)(MFSTART + k);
        const double wk = w[j];  // 0 for the padding points k >= 90
        double val = par[0];
        double J[P];
        double dsw[N], d2w[N];
        J[0] = wk;
#pragma unroll
        for (int n = 0; n < N; n++) {
            const double d = x - par[1 + 2 * n];
            const bool in = d > 1.0 && d < (double)(T - 1);  // T2:629
            int i = (int)d;
            i = i < 0 ? 0 : (i > T - 2 ? T - 2 : i);
            const double f = d - (double)i;
            const double2 q01 = __ldg(reinterpret_cast<const double2 *>(spl4 + i));
            const double2 q23 = __ldg(reinterpret_cast<const double2 *>(spl4 + i) + 1);
            double s = q01.x + f * (q01.y + f * (q23.x + f * q23.y));
            s = in ? s : 0.0;
            val += par[2 + 2 * n] * s;
            if (WITH_J) {
                double ds = q01.y + f * (2.0 * q23.x + 3.0 * f * q23.y);
                ds = in ? ds : 0.0;
                double d2 = 2.0 * q23.x + 6.0 * f * q23.y;
                d2 = in ? d2 : 0.0;
                dsw[n] = ds * wk;
                d2w[n] = d2 * wk;
                J[1 + 2 * n] = -par[2 + 2 * n] * dsw[n];
                J[2 + 2 * n] = s * wk;
            }
        }
        const double r = (y[j] - val) * wk;
        ne.c2 += r * r;
        if (WITH_J) {
#pragma unroll
            for (int n = 0; n < N; n++) {
                ne.s1[n] += r * dsw[n];
                ne.s2[n] += r * d2w[n];
            }
#pragma unroll
            for (int a = 0; a < P; a++) {
                ne.g[a] += J[a] * r;
#pragma unroll
                for (int b = 0; b <= a; b++) ne.H[a * (a + 1) / 2 + b] += J[a] * J[b];
            }
        }
    }
#pragma unroll
    for (int n = 0; n < N; n++) ne.s2[n] *= -par[2 + 2 * n];
    group_reduce<P, GROUP>(ne);
}

// (H + lambda |diag(H)|) dp = g by Cholesky, all in registers; one rsqrt per pivot, no divisions.
// Returns false if the damped matrix is not positive definite.  With `newton` H is the exact half-Hessian
// (Gauss-Newton matrix + the second-order terms s1, s2), which may be indefinite away from a minimum: the
// caller then raises lambda exactly as for a rejected step.
template <int P>
__device__ __forceinline__ bool solve_damped(const NormalEq<P> &ne, double lambda, double (&dp)[P], bool newton = false)
{
    double L[P * (P + 1) / 2];  // strict lower part; the diagonal slot holds 1 / L_aa
    bool pd = true;
    const double nf = newton ? 1.0 : 0.0;
#pragma unroll
    for (int a = 0; a < P; a++) {
#pragma unroll
        for (int b = 0; b <= a; b++) {
            double s = ne.H[a * (a + 1) / 2 + b];
            if (a == b && (a & 1)) s = fma(nf, ne.s2[(a - 1) / 2], s);                   // (t_n, t_n)
            if (a == b + 1 && (b & 1)) s = fma(nf, ne.s1[(b - 1) / 2], s);               // (A_n, t_n)
            if (a == b) s += lambda * (fabs(s) + 1e-12);
#pragma unroll
            for (int k = 0; k < b; k++) s -= L[a * (a + 1) / 2 + k] * L[b * (b + 1) / 2 + k];
            if (a == b) {
                pd = pd && (s > 0);
                L[a * (a + 1) / 2 + a] = rsqrt(s > 0 ? s : 1.0);
            } else {
                L[a * (a + 1) / 2 + b] = s * L[b * (b + 1) / 2 + b];
            }
        }
    }
#pragma unroll
    for (int a = 0; a < P; a++) {
        double s = ne.g[a];
#pragma unroll
        for (int k = 0; k < a; k++) s -= L[a * (a + 1) / 2 + k] * dp[k];
        dp[a] = s * L[a * (a + 1) / 2 + a];
    }
#pragma unroll
    for (int a = P - 1; a >= 0; a--) {
        double s = dp[a];
#pragma unroll
        for (int k = a + 1; k < P; k++) s -= L[k * (k + 1) / 2 + a] * dp[k];
        dp[a] = s * L[a * (a + 1) / 2 + a];
    }
    return pd;
}

// upper bound 2 g.dp of the chi2 decrease a Gauss-Newton step can achieve (the quadratic model gives
// 2 g.dp - dp.H.dp and H is positive semi-definite)
template <int P>
__device__ __forceinline__ double lm_pred2(const NormalEq<P> &ne, const double (&dp)[P])
{
    double s = 0;
#pragma unroll
    for (int a = 0; a < P; a++) s = fma(ne.g[a], dp[a], s);
    return 2.0 * s;
}

// cp.async (LDGSTS): global -> shared without staging registers; completion is awaited only at use time
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int FS_STAGE_DOUBLES = 136;  // per group: 110 samples | wftime[12] | wfampl[12] (+2 pad)

// stage C of the job pipeline: asynchronously copy the next job's trace and seeds into the group's
// shared-memory staging buffer (880 B = 55 x 16 B, 16-byte aligned because 880 = 16 * 55)
template <int N, int GROUP>
__device__ __forceinline__ void stage_job(double *buf, long long item, int g, const double *__restrict__ signal,
                                          const double *__restrict__ wftime, const double *__restrict__ wfampl)
{
    const double *src = signal + (size_t)item * T;
#pragma unroll
    for (int c = g; c < T / 2; c += GROUP) cp_async16(buf + 2 * c, src + 2 * c);
    if (g < N) {
        cp_async8(buf + T + g, wftime + (size_t)item * MAXP + g);
        cp_async8(buf + T + MAXP + g, wfampl + (size_t)item * MAXP + g);
    }
    cp_async_commit();
}

// 1 / Err[ib] with Err of T2:946-956: e = sqrt(|y*4.096/2|)/4.096, replaced by sqrt(2.048)/4.096 when
// e < 1.  The comparison is exact: every operation in e(|y|) is a monotone rounded function, and
// bisection on the rounded expression gives  e < 1  <=>  |y| < 0x1.0624dd2f1a9fcp+3 (= 8.192).
__device__ __forceinline__ double inv_err(double v)
{
    const double a = fabs(v);
    if (a < 0x1.0624dd2f1a9fcp+3) return 0x1.6e5b7d16657e1p+1;  // 1 / (sqrt(2.048)/4.096)
    return 4.096 * rsqrt(dmul(a, 4.096) / 2.);
}

// N pulses, GROUP lanes per fit (32/GROUP fits per warp), persistent groups: a group that finishes its
// fit immediately claims the next job from a global counter, so the lanes of a warp never wait for the
// slowest fit of a batch.  The groups of a warp still execute in lock-step (the shuffles need the whole
// warp); every state update is predicated per group.  Both attempts of the reference's policy run in
// the same loop: if the first attempt (lambda0 = 1e-3, fit_max_iter accepted steps) does not
// converge the group restarts from the same seeds with the tougher configuration (lambda0 = 1,
// fit_retry_max_iter steps) (T2:761-768); if that fails too the TSpectrum values are kept (T2:774-791).
// A fresh (or restarted) fit enters the loop with its seeds as the trial point, so its first
// evaluation shares the pass the other groups use for their trial steps.
template <int N, int GROUP, int MINB>
__global__ void __launch_bounds__(FS_THREADS, MINB)
fit_small_kernel(const int *__restrict__ job_list, const int *__restrict__ job_count, int *__restrict__ job_next,
                 const double *__restrict__ signal, const double *__restrict__ corr_time_HMS, DevCalib cal, KParams kp,
                 double *__restrict__ wftime, double *__restrict__ wfampl, double *__restrict__ chi2_out,
                 double *__restrict__ timewf, double *__restrict__ amplwf, uint8_t *__restrict__ status,
                 DeviceCounters *__restrict__ ctr, const double *__restrict__ cont_state = nullptr)
{
    constexpr int P = 2 * N + 1;
    constexpr int PTS = 96 / GROUP;
    constexpr double REL_TOL = FIT_REL_TOL;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int g = lane & (GROUP - 1);
    const int leader = lane & ~(GROUP - 1);
    const int njobs = *job_count;
    unsigned long long c_ok1 = 0, c_ok2 = 0, c_fb = 0, c_it = 0, c_att = 0, c_ev = 0;

    bool has_job = false, exhausted = false, fresh = false;
    long long item = 0;
    int bn = 0;
    const double4 *spl4 = reinterpret_cast<const double4 *>(cal.spline);
    double y[PTS], w[PTS], par[P], seed[P];
    NormalEq<P> cur;
    double lambda = 1e-3;
    int attempt = 1, max_iter = kp.fit_max_iter, iters = 0, it_total = 0, rejects = 0;
    bool newton = false;   // exact-Hessian steps: switched on by the first accepted step that gains less than 5 %
#pragma unroll
    for (int j = 0; j < PTS; j++) { y[j] = 0; w[j] = 0; }
#pragma unroll
    for (int i = 0; i < P; i++) { par[i] = 0; seed[i] = 0; }
    cur.c2 = 0;
    const unsigned group_mask = (GROUP == 32) ? FULL : (((1u << GROUP) - 1u) << leader);
    // first claim: resolved synchronously (once per kernel)
    int next_job = 0;
    if (g == 0) next_job = atomicAdd(job_next, 1);
    next_job = __shfl_sync(FULL, next_job, leader);
    long long next_item = (next_job < njobs) ? (long long)job_list[next_job] : -1;
    __shared__ __align__(16) double s_stage[(FS_THREADS / GROUP) * FS_STAGE_DOUBLES];
    double *buf = s_stage + (threadIdx.x / GROUP) * FS_STAGE_DOUBLES;
    if (next_item >= 0) stage_job<N, GROUP>(buf, next_item, g, signal, wftime, wfampl);
    int stage = 0;  // 0: next_item staged; 1: next_job claimed, list read pending; 2: staging copy pending

    for (;;) {
        // ---- job pipeline.  Every group keeps one job claimed AHEAD of the one it is fitting, and the three
        // dependent memory round trips behind a claim (atomic cursor -> job list -> trace) are spread over
        // three loop iterations, so none of them is waited for: stage A (at a refill) bumps the cursor,
        // stage B (next iteration) reads the job list, stage C (the one after) prefetches the trace into L1.
        if (stage == 2) {
            if (next_item >= 0) stage_job<N, GROUP>(buf, next_item, g, signal, wftime, wfampl);
            stage = 0;
        }
        if (stage == 1) {
            next_item = (next_job < njobs) ? (long long)job_list[next_job] : -1;
            stage = 2;
        }
        const bool need = !has_job && !exhausted;
        if (__any_sync(FULL, need)) {
            if (need) {
                if (stage == 1 || stage == 2) {  // the claim-ahead has not been resolved yet (very short fit)
                    if (stage == 1) next_item = (next_job < njobs) ? (long long)job_list[next_job] : -1;
                    if (next_item >= 0) stage_job<N, GROUP>(buf, next_item, g, signal, wftime, wfampl);
                    stage = 0;
                }
                const long long it_now = next_item;
                const int job_now = next_job;
                int j2 = 0;
                if (g == 0) j2 = atomicAdd(job_next, 1);  // stage A: result first used next iteration
                next_job = __shfl_sync(group_mask, j2, leader);
                stage = 1;
                if (it_now >= 0) {
                    item = it_now;
                    bn = (int)(item % B);
                    spl4 = reinterpret_cast<const double4 *>(cal.spline) + (size_t)bn * (T - 1);
                    cp_async_wait_all();
                    __syncwarp(group_mask);  // every lane's copies have landed and are visible to the group
#pragma unroll
                    for (int jj = 0; jj < PTS; jj++) {
                        const int k = g + GROUP * jj;
                        const double v = (k < NFIT) ? buf[MFSTART + k] : 0.0;
                        y[jj] = v;
                        w[jj] = (k < NFIT) ? inv_err(v) : 0.0;
                    }
                    // pedestal seed = mean of the first 20 samples (T2:671-677); summed as a tree inside the
                    // group (a seed: its last bit does not matter for the tolerance-based fit outputs)
                    double ped = 0;
#pragma unroll
                    for (int i = g; i < 20; i += GROUP) ped += buf[i];
#pragma unroll
                    for (int o = GROUP / 2; o > 0; o >>= 1) ped += __shfl_xor_sync(group_mask, ped, o);
                    seed[0] = ped / 20;
                    const double tref = cal.timeref[bn];
#pragma unroll
                    for (int n = 0; n < N; n++) {
                        seed[1 + 2 * n] = dsub(buf[T + n], tref);   // wftime - timeref   T2:662
                        seed[2 + 2 * n] = buf[T + MAXP + n];        // wfampl             T2:663
                    }
#pragma unroll
                    for (int i = 0; i < P; i++) par[i] = seed[i];
                    __syncwarp(group_mask);  // the buffer may be refilled from here on
                    has_job = true; fresh = true;
                    lambda = 1e-3; attempt = 1; max_iter = kp.fit_max_iter; iters = 0; it_total = 0; rejects = 0;
                    newton = false;
                    if (cont_state) {   // continuation of a fit started by fit_thread_kernel: its parameters, damping, counters
                        const double *cs = cont_state + (size_t)job_now * 10;
#pragma unroll
                        for (int i = 0; i < P; i++) par[i] = cs[i];
                        lambda = cs[P];
                        const int pk = (int)cs[P + 1];
                        iters = pk >> 7; rejects = (pk >> 1) & 63; newton = pk & 1;
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(FULL, has_job)) break;

        // ---- one LM try (or the first evaluation of a fresh fit)
        double dp[P], trial[P];
        const bool pd = fresh ? true : solve_damped<P>(cur, lambda, dp, newton);
#pragma unroll
        for (int i = 0; i < P; i++) trial[i] = par[i] + ((pd && !fresh) ? dp[i] : 0.0);
        // predicted-decrease stop (see fit_thread_kernel): same rule, so a continued fit behaves as it would have there
        const bool pre_done = has_job && !fresh && pd && lambda <= 1e-2 && lm_pred2<P>(cur, dp) < REL_TOL * (fabs(cur.c2) + 1e-30);
        NormalEq<P> nxt;
        eval_group<N, GROUP, true>(trial, y, w, g, spl4, nxt);
        bool finished = pre_done, converged = pre_done;
        if (has_job && !pre_done) {
            if (g == 0) c_ev++;
            if (fresh) {
                cur = nxt;
                fresh = false;
            } else if (pd && nxt.c2 <= cur.c2) {
                const double rel = (cur.c2 - nxt.c2) / (fabs(cur.c2) + 1e-30);
#pragma unroll
                for (int i = 0; i < P; i++) par[i] = trial[i];
                cur = nxt;
                lambda = fmax(lambda * 0.2, 1e-12);
                rejects = 0;
                iters++;
                if (rel < 0.05) newton = true;
                if (rel < REL_TOL) { converged = true; finished = true; }
                else if (iters >= max_iter) {
                    if (attempt == 1) {  // retry from the same seeds, tougher configuration
#pragma unroll
                        for (int i = 0; i < P; i++) par[i] = seed[i];
                        fresh = true; lambda = 1.0; attempt = 2; max_iter = kp.fit_retry_max_iter; newton = false;
                        it_total += iters; iters = 0; rejects = 0;
                    } else {
                        finished = true;
                    }
                }
            } else {
                lambda = fmax(lambda * 10, 1e-6);
                rejects++;
                if (rejects >= 30) {   // no descent step left: a minimum to machine precision -- unless chi2 is not a number
                    if (isfinite(cur.c2)) { converged = true; finished = true; iters++; }
                    else if (attempt == 1) {   // (NaN / Inf sample, Cholesky failing on every try): failed attempt, as Migrad's !ok (T2:755-768)
#pragma unroll
                        for (int i = 0; i < P; i++) par[i] = seed[i];
                        fresh = true; lambda = 1.0; attempt = 2; max_iter = kp.fit_retry_max_iter; newton = false;
                        it_total += iters; iters = 0; rejects = 0;
                    } else {
                        finished = true;
                    }
                }
            }
        }
        // ---- write back finished fits
        if (finished) {
            it_total += iters;
            if (g == 0) {
                const long long e = item / B;
                const double corr = corr_time_HMS ? corr_time_HMS[e] : 0.0;
                const double cort = (double)cal.cortime[bn];
                const double accdt = dmul(kp.timerefacc, kp.dt);
                double out_t[N], out_a[N];
                int st;
                if (!converged) {  // fallback: TSpectrum values, time converted to corrected ns (T2:779-790)
                    st = NPSWF_ST_FALLBACK;
#pragma unroll
                    for (int n = 0; n < N; n++) {
                        out_t[n] = dsub(dsub(dadd(dmul(seed[1 + 2 * n], kp.dt), corr), cort), accdt);
                        out_a[n] = seed[2 + 2 * n];
                    }
                    chi2_out[item] = -100.;
                    c_fb++;
                } else {  // T2:796-827
                    st = attempt == 1 ? NPSWF_ST_FIT_OK1 : NPSWF_ST_FIT_OK2;
#pragma unroll
                    for (int n = 0; n < N; n++) {
                        out_a[n] = par[2 + 2 * n];
                        out_t[n] = dsub(dsub(dadd(dmul(par[1 + 2 * n], kp.dt), corr), cort), accdt);
                    }
                    chi2_out[item] = cur.c2 / (double)(NFIT - P);
                    if (attempt == 1) c_ok1++;
                    else c_ok2++;
                }
                double bt = out_t[0], ba = out_a[0];  // T2:999-1016
#pragma unroll
                for (int n = 0; n < N; n++) {
                    wftime[(size_t)item * MAXP + n] = out_t[n];
                    wfampl[(size_t)item * MAXP + n] = out_a[n];
                    if (n > 0 && fabs(out_t[n]) < fabs(bt)) { bt = out_t[n]; ba = out_a[n]; }
                }
                if (timewf) timewf[item] = bt;
                if (amplwf) amplwf[item] = ba;
                if (status) status[item] = (uint8_t)(NPSWF_ST_PRESENT | NPSWF_ST_OKTOFIT | st);  // fit jobs are present && okToFit
                c_it += it_total;
                c_att++;
            }
            has_job = false;
        }
    }
    if (ctr && g == 0) {
        if (c_att) atomicAdd(&ctr->n_fit_attempted, c_att);
        if (c_ok1) atomicAdd(&ctr->n_fit_ok_first, c_ok1);
        if (c_ok2) atomicAdd(&ctr->n_fit_ok_retry, c_ok2);
        if (c_fb) atomicAdd(&ctr->n_fallback, c_fb);
        if (c_it) atomicAdd(&ctr->n_fit_iterations, c_it);
        if (c_ev) atomicAdd(&ctr->n_fit_evals, c_ev);
    }
}

}  // namespace npswf
