// Thread-per-fit kernel for N = 1, 2 pulses (P = 3, 5 parameters): the bulk of the fits.
//
// Same objective / policy as kernel_fit.cuh and kernel_fit_small.cuh (Fitwf, T2:601-828): chi2 of
// p0 + sum A_n S(x - t_n) over the 90 points x = 10..99, damped Gauss-Newton with analytic spline
// derivatives, attempt -> retry from the same seeds -> fall back to the TSpectrum values.
//
// Mapping: ONE THREAD owns one fit.  Its 90 trace samples and weights 1/Err sit in shared memory as float2,
// transposed ([point][lane], leading dimension 33) so that the 32 fits of a warp read point j conflict-free
// (ADC samples are multiples of 1000/4096 mV and therefore exact in binary32; a trace that is not exactly
// representable is handed to the sub-warp kernel untouched, so nothing is ever rounded); the
// normal equations, the Cholesky factor and the LM state are plain registers.  No shuffles, no
// redundant solves, no partial sums.  Because the knots of the reference shape are the integers
// 0..109, every point of a pulse has the same fractional offset f = frac(10 - t): the spline is a
// 4-term polynomial in f with coefficient quads read at consecutive indices, and the in-range test
// 1 < x - t < 109 (T2:629) becomes an integer interval of point indices, computed once per try.
// Lanes are persistent: a thread whose fit is finished takes the next job from the warp's queue (32 job
// ids claimed with one atomic, traces prefetched into L2 on claim), and the warp copies that trace -- and
// the per-point weights 1/Err -- into the thread's shared-memory column.
// A thread's latency per LM try is ~30x that of a sub-warp group, so this kernel only runs the first
// `fit_thread_tries` tries of the first attempt (97 % of the fits converge inside them); a fit that needs
// more hands its state (parameters, damping, step counters) to a continuation list that
// fit_small_kernel finishes, including the retry / fall-back policy.
#pragma once
#include "common.cuh"
#include "kernel_fit_small.cuh"

namespace npswf {

constexpr int FT_WARPS = 4;   // 2 CTAs per SM = 2 warps per scheduler: what a 16 K-register sub-partition holds at 255 registers
constexpr int FT_THREADS = FT_WARPS * 32;
constexpr int FT_LD = 33;
constexpr int FT_WARP_BYTES = NFIT * FT_LD * (int)sizeof(float2);   // (y, 1/err) per point
constexpr size_t FT_SMEM = (size_t)FT_WARPS * FT_WARP_BYTES;         // 95 040 B -> 2 CTAs per SM
// Instances with N <= NPSWF_FT_YONLY_MAXN keep the samples alone in the tile (47 520 B per CTA) and recompute the
// weight 1 / Err per point on the FP32 / MUFU pipes -- bit for bit what inv_err_f32 stores, so the results do not
// depend on the layout -- which lets NPSWF_FT_MINB<N> CTAs share an SM.  Measured (9 472 config-2 events, fit stage):
// N = 1 at 3 CTAs / 168 registers 27.1 -> 26.5 ms (adopted); N = 2 as well: 27.1 (its spills land outside the point
// loop, no gain); N = 1 at 4 CTAs / 128 registers with one point per loop body: 27.7; N = 3 at 3 CTAs: 29.9 (spills
// inside the point loop).  The thread-per-fit kernels are bound by their FP64 instruction stream, not by latency.
#ifndef NPSWF_FT_YONLY_MAXN
#define NPSWF_FT_YONLY_MAXN 1
#endif
#ifndef NPSWF_FT_MINB1
#define NPSWF_FT_MINB1 3
#endif
#ifndef NPSWF_FT_MINB2
#define NPSWF_FT_MINB2 2
#endif
#ifndef NPSWF_FT_MINB3
#define NPSWF_FT_MINB3 2
#endif
#ifndef NPSWF_FT_U1
#define NPSWF_FT_U1 5
#endif
__host__ __device__ constexpr bool ft_yonly(int N) { return N <= NPSWF_FT_YONLY_MAXN; }
__host__ __device__ constexpr int ft_warp_bytes(int N) { return ft_yonly(N) ? (NFIT * FT_LD * (int)sizeof(float) + 15) / 16 * 16 : FT_WARP_BYTES; }
__host__ __device__ constexpr size_t ft_smem(int N) { return (size_t)FT_WARPS * ft_warp_bytes(N); }
__host__ __device__ constexpr int ft_minblocks(int N) { return N == 1 ? NPSWF_FT_MINB1 : N == 2 ? NPSWF_FT_MINB2 : N == 3 ? NPSWF_FT_MINB3 : 2; }
template <bool YONLY> struct FtTile { typedef float2 type; };
template <> struct FtTile<true> { typedef float type; };
constexpr int FT_CONT_STRIDE = 10;  // doubles per continuation record: par[P] | lambda | iters << 7 | rejects << 1 | newton

// 1 / Err (T2:946-956) as the binary32 weight the kernel stores: the same branch point as inv_err(), the
// square root through the FP32 MUFU (relative error 2^-22, the storage format itself rounds at 2^-24; both
// are far inside the chi2 tolerance of 1e-3)
__device__ __forceinline__ float inv_err_f32(double v)
{
    const double a = fabs(v);
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)a));   // 1/sqrt(a)
    const float wv = r * 0x1.6e5b7ep+1f;                             // 4.096 / sqrt(2.048 a) = (1/sqrt a) * 4.096/sqrt(2.048)
    return (a < 0x1.0624dd2f1a9fcp+3) ? 0x1.6e5b7ep+1f : wv;
}

// sample and weight of one tile entry: stored pair, or the weight recomputed from the (binary32-exact) sample
__device__ __forceinline__ void ft_point(const float2 e, double &y, double &w) { y = (double)e.x; w = (double)e.y; }
__device__ __forceinline__ void ft_point(const float e, double &y, double &w)
{
    const float af = fabsf(e);
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(af));
    const float wf = (af < 0x1.0624dep+3f) ? 0x1.6e5b7ep+1f : rs * 0x1.6e5b7ep+1f;   // = inv_err_f32((double)e)
    y = (double)e;
    w = (double)wf;
}
__device__ __forceinline__ void ft_store(float2 *dst, float yf, double v) { *dst = make_float2(yf, inv_err_f32(v)); }
__device__ __forceinline__ void ft_store(float *dst, float yf, double) { *dst = yf; }

// chi2 and normal equations of one fit at parameters p, all 90 points, one thread.
// kn points at knot 0 of the block's zero-padded knot array.  U points per loop body; the loads of the next
// body are issued before the arithmetic of the current one.
// DIAG: only the diagonal of the normal matrix is accumulated (the variable-metric kernel needs chi2, its gradient and
// its second derivatives along the axes, not the full Gauss-Newton matrix).
template <int N, int U, bool DIAG = false, typename TY = float2>
__device__ __forceinline__ void eval_thread(const double (&p)[2 * N + 1], const TY *__restrict__ ywcol,
                                            const double2 *__restrict__ kn, NormalEq<2 * N + 1> &ne)
{
    constexpr int P = 2 * N + 1;
    static_assert(NFIT % (2 * U) == 0, "2 U must divide the number of fit points");
    const double2 *kp[N];
    double2 k0[N];
    double wa[N], wb[N], wc[N], wd[N], dc[N], dd[N], e0[N], e1[N], nA[N];
    int jlo[N];
    unsigned span[N];
#pragma unroll
    for (int n = 0; n < N; n++) {
        // d_j = (10 + j) - t = u + j;  knot index floor(d_j) = floor(u) + j, fraction f = u - floor(u)
        double u = (double)MFSTART - p[1 + 2 * n];
        u = fmin(fmax(u, -120.0), 120.0);   // a wild trial step stays inside the zero padding of the knot array
        const double fl = floor(u);
        const int i0 = (int)fl;
        const double f = u - fl, g = 1.0 - f;
        // S = g y0 + f y1 + ((g^3 - g) c0 + (f^3 - f) c1) / 3,  S' = y1 - y0 + ((1 - 3 g^2) c0 + (3 f^2 - 1) c1) / 3
        wa[n] = g; wb[n] = f;
        wc[n] = (g * g * g - g) * (1.0 / 3.0); wd[n] = (f * f * f - f) * (1.0 / 3.0);
        dc[n] = (1.0 - 3.0 * g * g) * (1.0 / 3.0); dd[n] = (3.0 * f * f - 1.0) * (1.0 / 3.0);
        e0[n] = 2.0 * g; e1[n] = 2.0 * f;   // S'' = 2 (g c0 + f c1)
        nA[n] = -p[2 + 2 * n];
        // 1 < u + j < 109  (T2:629)  <=>  jl <= j <= jh
        int jl = (int)floor(1.0 - u) + 1, jh = (int)ceil((double)(T - 1) - u) - 1;
        jl = max(jl, 0);
        jh = min(jh, NFIT - 1);
        if (jh < jl) { jl = 1 << 20; jh = jl; }
        jlo[n] = jl; span[n] = (unsigned)(jh - jl);
        kp[n] = kn + i0;
        k0[n] = __ldg(kp[n]);
    }
#pragma unroll
    for (int i = 0; i < P * (P + 1) / 2; i++) ne.H[i] = 0;
#pragma unroll
    for (int i = 0; i < P; i++) ne.g[i] = 0;
    ne.c2 = 0;
#pragma unroll
    for (int n = 0; n < N; n++) { ne.s1[n] = 0; ne.s2[n] = 0; }
    const double p0 = p[0];
    // two register buffers of U points each: while one is consumed the other is being loaded (no copies)
    TY ywA[U], ywB[U];
    double2 kA[N][U], kB[N][U];
    auto load = [&](int jb, TY (&yw)[U], double2 (&kk)[N][U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            yw[u] = ywcol[(jb + u) * FT_LD];
#pragma unroll
            for (int n = 0; n < N; n++) kk[n][u] = __ldg(kp[n] + jb + u + 1);
        }
    };
    auto consume = [&](int jb, const TY (&yw)[U], const double2 (&kk)[N][U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + u;
            double wk, yv;
            ft_point(yw[u], yv, wk);
            double r = (yv - p0) * wk;
            double J[P];
            double dsw[N], d2w[N];
            J[0] = wk;
#pragma unroll
            for (int n = 0; n < N; n++) {
                const double2 k1 = kk[n][u];
                const double wkm = ((unsigned)(j - jlo[n]) <= span[n]) ? wk : 0.0;   // in range: 1 < x - t < 109
                const double s = fma(wd[n], k1.y, fma(wc[n], k0[n].y, fma(wb[n], k1.x, wa[n] * k0[n].x)));
                const double ds = fma(dd[n], k1.y, fma(dc[n], k0[n].y, k1.x - k0[n].x));
                const double d2 = fma(e1[n], k1.y, e0[n] * k0[n].y);
                k0[n] = k1;
                dsw[n] = ds * wkm;
                d2w[n] = d2 * wkm;
                J[2 + 2 * n] = s * wkm;
                J[1 + 2 * n] = dsw[n];   // without its factor -A_n: applied once to the sums after the loop
                r = fma(nA[n], J[2 + 2 * n], r);
            }
            ne.c2 = fma(r, r, ne.c2);
#pragma unroll
            for (int n = 0; n < N; n++) ne.s2[n] = fma(r, d2w[n], ne.s2[n]);
#pragma unroll
            for (int a = 0; a < P; a++) {
                ne.g[a] = fma(J[a], r, ne.g[a]);
#pragma unroll
                for (int b = 0; b <= a; b++)
                    if (!DIAG || b == a) ne.H[a * (a + 1) / 2 + b] = fma(J[a], J[b], ne.H[a * (a + 1) / 2 + b]);
            }
        }
    };
    load(0, ywA, kA);
#pragma unroll 1
    for (int j0 = 0; j0 < NFIT; j0 += 2 * U) {
        load(j0 + U, ywB, kB);
        consume(j0, ywA, kA);
        if (j0 + 2 * U < NFIT) load(j0 + 2 * U, ywA, kA);
        consume(j0 + U, ywB, kB);
    }
    // the time columns were accumulated without their factor -A_n: sum r w S' is the second-order term s1 as it stands,
    // and the gradient / normal-matrix entries get the factor(s) now
#pragma unroll
    for (int n = 0; n < N; n++) {
        ne.s1[n] = ne.g[1 + 2 * n];
        ne.s2[n] *= nA[n];
    }
#pragma unroll
    for (int a = 0; a < P; a++) {
        const bool ta = a >= 1 && ((a - 1) & 1) == 0;
        if (ta) ne.g[a] *= nA[(a - 1) >> 1];
#pragma unroll
        for (int b = 0; b <= a; b++) {
            if (DIAG && b != a) continue;
            const bool tb = b >= 1 && ((b - 1) & 1) == 0;
            if (ta && tb) ne.H[a * (a + 1) / 2 + b] *= nA[(a - 1) >> 1] * nA[(b - 1) >> 1];
            else if (ta) ne.H[a * (a + 1) / 2 + b] *= nA[(a - 1) >> 1];
            else if (tb) ne.H[a * (a + 1) / 2 + b] *= nA[(b - 1) >> 1];
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int N>
__global__ void __launch_bounds__(FT_THREADS, ft_minblocks(N))
fit_thread_kernel(const int *__restrict__ job_list, const int *__restrict__ job_count, int *__restrict__ job_next,
                  const double *__restrict__ signal, const double *__restrict__ corr_time_HMS, DevCalib cal, KParams kp,
                  double *__restrict__ wftime, double *__restrict__ wfampl, double *__restrict__ chi2_out,
                  double *__restrict__ timewf, double *__restrict__ amplwf, uint8_t *__restrict__ status,
                  DeviceCounters *__restrict__ ctr, int *__restrict__ cont_count, int *__restrict__ cont_list,
                  double *__restrict__ cont_state)
{
    constexpr int P = 2 * N + 1;
    constexpr int U = (N == 1) ? NPSWF_FT_U1 : 1;
    constexpr double REL_TOL = FIT_REL_TOL;
    typedef typename FtTile<ft_yonly(N)>::type TY;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char ft_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TY *ywwarp = reinterpret_cast<TY *>(ft_smem + (size_t)warp * ft_warp_bytes(N));   // [90][33] (sample, 1/err) or samples
    const TY *ywcol = ywwarp + lane;
    const int njobs = *job_count;
    // N >= 4 has no sub-warp continuation kernel that could take over the LM state: a fit that is not done after
    // fit_thread_tries + 10 tries (a thread's try takes ~30x as long as a warp's, and these kernels have few jobs, so
    // such a fit would be the tail of the whole stage) is handed to the warp-per-fit kernel, which runs its first
    // attempt again from the seeds and then the retry
    const int max_tries = (N <= 3) ? kp.fit_thread_tries : kp.fit_thread_tries + 10;
    unsigned long long c_ok1 = 0, c_it = 0, c_att = 0, c_ev = 0;

    bool has_job = false, exhausted = false, fresh = false;
    long long item = 0;
    int bn = 0;
    const double2 *kn = cal.knots + KN_LO;
    double par[P];
    NormalEq<P> cur;
    double lambda = 1e-3;
    int iters = 0, rejects = 0, tries = 0;
    bool newton = false;   // exact-Hessian steps (see NormalEq)
#pragma unroll
    for (int i = 0; i < P; i++) par[i] = 0;
#pragma unroll
    for (int i = 0; i < P * (P + 1) / 2; i++) cur.H[i] = 0;
#pragma unroll
    for (int i = 0; i < P; i++) cur.g[i] = 0;
    cur.c2 = 0;

    // warp job queue: lane i holds the i-th job id of the current batch of 32
    int q_item = -1, qpos = 32;
    bool drained = false;   // the cursor has passed the end of the list

    for (;;) {
        // ---- hand new jobs to the lanes without one (up to 4 traces in flight per round)
        // (a round costs a few hundred issue slots whatever the number of idle lanes, so it waits for four of them
        // unless the warp has nothing else to do)
        unsigned m = __ballot_sync(FULL, !has_job && !exhausted);
        if (__popc(m) < 4 && __any_sync(FULL, has_job)) m = 0;
        bool got = false, inexact = false;
        double ped = 0;
        while (m) {
            int ls[4];
            long long its[4];
            double v[4][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                ls[k] = -1; its[k] = -1;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) v[k][c4] = 0;
                if (m) {
                    if (qpos == 32 && !drained) {   // claim the next 32 job ids, start pulling their traces into L2
                        int base = 0;
                        if (lane == 0) base = atomicAdd(job_next, 32);
                        base = __shfl_sync(FULL, base, 0);
                        const int j = base + lane;
                        q_item = (j < njobs) ? job_list[j] : -1;
                        qpos = 0;
                        drained = base + 32 >= njobs;
                        if (q_item >= 0) {
                            const char *pt = reinterpret_cast<const char *>(signal + (size_t)q_item * T);
#pragma unroll
                            for (int c = 0; c < 7; c++) prefetch_l2(pt + 128 * c);
                            prefetch_l2(pt + T * 8 - 8);
                        }
                    }
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int it = (qpos < 32) ? __shfl_sync(FULL, q_item, qpos) : -1;
                    if (qpos < 32) qpos++;
                    ls[k] = l;
                    its[k] = it;
                    if (it >= 0) {
                        const double *src = signal + (size_t)it * T;
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            const int c = 32 * c4 + lane;
                            if (c < T) v[k][c4] = src[c];
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (ls[k] >= 0) {   // warp-uniform
                    // pedestal seed = mean of the first 20 samples (T2:671-677); a seed: its last bit does not matter
                    const double sum = warp_sum(lane < 20 ? v[k][0] : 0.0);
                    bool exact = true;
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        const int c = 32 * c4 + lane;
                        if (c >= MFSTART && c < MFEND && its[k] >= 0) {
                            const float yf = (float)v[k][c4];
                            exact = exact && ((double)yf == v[k][c4]) && (fabsf(yf) <= 3.0e38f);   // NaN fails the first test, +-Inf the second
                            ft_store(&ywwarp[(c - MFSTART) * FT_LD + ls[k]], yf, v[k][c4]);
                        }
                    }
                    exact = __all_sync(FULL, exact);
                    if (lane == ls[k]) {
                        if (its[k] >= 0) { item = its[k]; ped = sum / 20; got = true; inexact = !exact; }
                        else exhausted = true;
                    }
                }
            }
        }
        __syncwarp();
        if (got) {
            bn = (int)(item % B);
            kn = cal.knots + (size_t)bn * KN_LEN + KN_LO;
            const double tref = cal.timeref[bn];
            par[0] = ped;
#pragma unroll
            for (int n = 0; n < N; n++) {
                par[1 + 2 * n] = dsub(wftime[(size_t)item * MAXP + n], tref);   // wftime - timeref   T2:662
                par[2 + 2 * n] = wfampl[(size_t)item * MAXP + n];               // wfampl             T2:663
            }
            has_job = true; fresh = true;
            lambda = 1e-3; iters = 0; rejects = 0; newton = false;
            tries = inexact ? (1 << 20) : 0;   // samples not exact in binary32: hand the fit over after the seed evaluation
        }
        if (!__any_sync(FULL, has_job)) break;

        // ---- one LM try (or the first evaluation of a fresh fit)
        double dp[P], trial[P];
        const bool pd = fresh ? true : solve_damped<P>(cur, lambda, dp, newton);
#pragma unroll
        for (int i = 0; i < P; i++) trial[i] = par[i] + ((pd && !fresh) ? dp[i] : 0.0);
        // predicted-decrease stop: in the Gauss-Newton regime chi2 cannot drop by more than 2 g.dp, so a step
        // whose bound is below the tolerance is not worth its evaluation
        const bool pre_done = has_job && !fresh && pd && lambda <= 1e-2 && lm_pred2<P>(cur, dp) < REL_TOL * (fabs(cur.c2) + 1e-30);
        NormalEq<P> nxt;
        eval_thread<N, U, false, TY>(trial, ywcol, kn, nxt);
        bool finished = pre_done, handoff = false;
        if (has_job && !pre_done) {
            tries++;
            c_ev++;
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
                if (rel < REL_TOL) finished = true;
                else if (iters >= kp.fit_max_iter) handoff = true;   // the retry policy lives in fit_small_kernel
            } else {
                lambda = fmax(lambda * 10, 1e-6);
                rejects++;
                if (rejects >= 30) { finished = true; iters++; }  // no descent step left (chi2 is finite here: a trace with a NaN / Inf sample never stays in this kernel)
            }
            bool restart = false;
            if (!finished && !handoff && tries >= max_tries) { handoff = true; restart = N >= 4; }
            if (handoff && restart) item |= FIT_CONT_RESTART;
        }
        if (handoff) {   // continuation record: fit_small_kernel re-evaluates at par and carries on
            const int idx = atomicAdd(cont_count, 1);
            cont_list[idx] = (int)item;
            if (cont_state) {   // (N >= 4: the warp-per-fit kernel restarts the fit from its seeds)
                double *cs = cont_state + (size_t)idx * FT_CONT_STRIDE;
#pragma unroll
                for (int i = 0; i < P; i++) cs[i < FT_CONT_STRIDE - 2 ? i : 0] = par[i];
                cs[P < FT_CONT_STRIDE - 2 ? P : 0] = lambda;
                cs[P + 1 < FT_CONT_STRIDE ? P + 1 : 0] = (double)((iters << 7) | (rejects << 1) | (newton ? 1 : 0));
            }
            has_job = false;
        }
        // ---- write back converged fits (T2:796-827)
        if (finished) {
            const long long e = item / B;
            const double corr = corr_time_HMS ? corr_time_HMS[e] : 0.0;
            const double cort = (double)cal.cortime[bn];
            const double accdt = dmul(kp.timerefacc, kp.dt);
            double bt = 0, ba = 0;  // T2:999-1016
#pragma unroll
            for (int n = 0; n < N; n++) {
                const double oa = par[2 + 2 * n];
                const double ot = dsub(dsub(dadd(dmul(par[1 + 2 * n], kp.dt), corr), cort), accdt);
                wftime[(size_t)item * MAXP + n] = ot;
                wfampl[(size_t)item * MAXP + n] = oa;
                if (n == 0 || fabs(ot) < fabs(bt)) { bt = ot; ba = oa; }
            }
            chi2_out[item] = cur.c2 / (double)(NFIT - P);
            if (timewf) timewf[item] = bt;
            if (amplwf) amplwf[item] = ba;
            if (status) status[item] = (uint8_t)(NPSWF_ST_PRESENT | NPSWF_ST_OKTOFIT | NPSWF_ST_FIT_OK1);
            c_ok1++;
            c_it += iters;
            c_att++;
            has_job = false;
        }
    }
    if (ctr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_att += __shfl_xor_sync(FULL, c_att, o);
            c_ok1 += __shfl_xor_sync(FULL, c_ok1, o);
            c_it += __shfl_xor_sync(FULL, c_it, o);
            c_ev += __shfl_xor_sync(FULL, c_ev, o);
        }
        if (lane == 0) {
            if (c_ev) atomicAdd(&ctr->n_fit_evals, c_ev);
            if (c_att) atomicAdd(&ctr->n_fit_attempted, c_att);
            if (c_ok1) atomicAdd(&ctr->n_fit_ok_first, c_ok1);
            if (c_it) atomicAdd(&ctr->n_fit_iterations, c_it);
        }
    }
}

}  // namespace npswf
