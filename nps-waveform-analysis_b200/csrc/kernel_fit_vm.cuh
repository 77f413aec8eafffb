// fit_vm_thread_kernel<N>: Fitwf (T2:601-828) minimised along MIGRAD'S OWN PATH at the cost of a few evaluations.
//
// The reference's Minuit2 Migrad is a variable-metric method: V0 = diag(1 / d2chi2/dp^2) at the seeds, then per
// iteration a line search (MnLineSearch: parabolic, <= 12 points) along -V g and a Davidon rank-two update of V,
// until the estimated distance to the minimum EDM = g^T V g / 2 (x (1 + 3 dcovar)) is below 2e-5.  Where the chi2 has
// several local minima (unresolved pile-up), WHICH of them a minimiser ends in is decided by those first steps -- a
// Levenberg-Marquardt solver starts with a full Gauss-Newton step and ends elsewhere in 0.14 % (1-3 pulses) to 6 %
// (up to 12 pulses near threshold) of the fits.  Migrad spends 2P chi2 evaluations per iteration on NUMERICAL
// gradients; its final steps are ~0.008 bin, so those gradients agree with the analytic ones to ~1e-5.  This kernel
// runs the same recursion -- seed, MnLineSearch, Davidon update, EDM stop (migrad_core.hpp spells them out for the
// exact kernels) -- with analytic gradients and second derivatives from the pass that evaluates chi2 anyway, and no
// MnHesse at the end (it does not move the parameters): ~11 evaluations per single-pulse fit instead of 65, and the
// same minimum as the CPU oracle's Migrad on 99.99 % (1-3 pulses) / 99.5 % (up to 12 near threshold) of the fits
// (tests/test_gpu_migrad.py::test_vm_mode_follows_migrad).
// It stops where Migrad stops (it does not converge further: the reference's numbers are Migrad's stopping points).
// NegativeG2LineSearch (a non-positive second derivative at the seeds) and MnPosDef (a direction that is not a descent)
// are followed in the kernel, out of line.  What leaves the common path for good -- EDM above the limit when the
// iterations end, too many evaluations, an invalid seed state, a trace that is not exact in binary32 -- is handed,
// untouched, to the exact Migrad kernels (fit_migrad_*), which run it from its seeds with the reference's retry /
// fall-back policy.
//
// Same machinery as fit_thread_kernel (kernel_fit_thread.cuh): one thread per fit (1-6 pulses), persistent lanes fed
// from the warp's job queue, one chi2 pass per trip of the main loop, the minimiser a per-lane state machine around
// that single evaluation site.  Its own: the tile holds the samples only (binary32; the weight is recomputed per
// point), the evaluation accumulates chi2, gradient and the diagonal second derivatives only, and everything is
// written for a SMALL loop body -- the kernel was bound by instruction fetches before it was bound by anything else
// (rare branches out of line, a single-copy trace loader, one point per loop body, the warps of a CTA in lockstep).
#pragma once
#include "kernel_fit_thread.cuh"
#include "migrad_core.hpp"


namespace npswf {

enum { VM_FRESH = 0, VM_LS_A = 1, VM_LS_B = 2, VM_LS_C = 3 };   // what the evaluation in flight is
enum { VM_EVAL = 0, VM_DONE = 1, VM_HANDOFF = 2 };              // what vm_advance asks for next
#ifndef NPSWF_VM_MAX_EVALS
#define NPSWF_VM_MAX_EVALS 48
#endif
// beyond this many evaluations the exact kernels take the fit: 48 for 1-3 pulses (100 / 250 would save 2-4 % of the
// stage for 0.07-0.1 point of agreement); 96 / 128 / 160 for 4 / 5 / 6 pulses, whose fits need more iterations and whose
// hand-overs run one warp per fit for milliseconds each
__host__ __device__ constexpr int vm_max_evals(int P) { return P <= 7 ? NPSWF_VM_MAX_EVALS : 16 * (P - 3); }
// why fits left the kernel (diagnostics; read by npswf_debug_vm_reasons): 0 evaluation limit / inexact trace, 1 second
// derivative <= 0 at the seeds, 2 EDM negative or not a number, 3 above the EDM limit, 4 not a descent direction
__device__ unsigned long long g_vm_reason[8];
#define VM_REASON(i) atomicAdd(&g_vm_reason[i], 1ULL)

template <int P>
struct VmState {
    double x0[P], g0[P], dir[P], gb[P];      // accepted point, its gradient, search direction, gradient at the best line-search point
    double g20[P], g2b[P], gs[P];            // NegativeG2LineSearch: second derivatives at x0 / at the best point, Minuit's gradient steps
    double V[P * (P + 1) / 2];               // inverse-Hessian estimate, lower triangle row-packed
    double f0, gdel, dcovar, edm;
    // MnLineSearch
    double slamin, overal, undral, toler8, slamax, flast, slam, xvmin, fvmin, p0x, p0y, p1x, p1y, p2x, p2y;
    int niter, phase, nev, iters, ng_iter;
    bool in_ng;                              // the line search in flight belongs to NegativeG2LineSearch
};

// The step Numerical2PGradientCalculator would leave in gstep for one parameter (value x, previous step gs) after its
// cycles at a point with chi2 = f: cycle 0 sizes the step from the PREVIOUS second derivative g2a, cycle 1 from the new
// one g2b_ (what the first cycle measures; the analytic value stands in for it), and stops once the step changes by
// less than 30 % -- or, after cycle 0, when the gradient moved by less than 5 % (gprev -> gnew).  NegativeG2LineSearch
// steps along an axis in units of this.
__device__ __noinline__ double vm_gstep(double x, double f, double gprev, double gnew, double g2a, double g2b_, double gs)
{
    constexpr double EPS = mg::EPS, EPS2 = mg::EPS2;
    const double dfmin = 8. * EPS2 * (fabs(f) + 1.0), vrysml = 8. * EPS * EPS;
    const double epspri = EPS2 + fabs(gprev * EPS2);
    const double stpmin = fmax(vrysml, 8. * fabs(EPS2 * x));
    double step = fmax(sqrt(dfmin / (fabs(g2a) + epspri)), fabs(0.1 * gs));
    step = fmin(step, 10. * fabs(gs));
    step = fmax(step, stpmin);
    gs = step;                                            // cycle 0 always runs ((step - 0) / step = 1)
    if (fabs(gprev - gnew) / (fabs(gnew) + dfmin / step) < 0.05) return gs;
    double step1 = fmax(sqrt(dfmin / (fabs(g2b_) + epspri)), fabs(0.1 * gs));
    step1 = fmin(step1, 10. * fabs(gs));
    step1 = fmax(step1, stpmin);
    if (fabs((step1 - step) / step1) < 0.3) return gs;
    return step1;
}

// One copy of the binary64 division sequence for the whole kernel (inlined, each of the ~20 quotients of the common path
// costs ~20 instructions of a loop body that has to stay inside the instruction cache)
#if NPSWF_VM_DIV_CALL
__device__ __noinline__ double vm_div(double a, double b) { return a / b; }
#else
__device__ __forceinline__ double vm_div(double a, double b) { return a / b; }
#endif

template <int P>
__device__ __forceinline__ double vm_edm(const double (&V)[P * (P + 1) / 2], const double (&g)[P])
{
    double s = 0;
#pragma unroll
    for (int a = 0; a < P; a++) {
        double r = 0;
#pragma unroll
        for (int b = 0; b < P; b++) r = fma(V[a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a], g[b], r);
        s = fma(g[a], r, s);
    }
    return 0.5 * s;
}

// MnPosDef on the packed estimate (rare: V has lost positive definiteness through a Davidon update): unpacked into
// thread-local arrays and handed to the scalar routine of the exact kernels (eigenvalues by Jacobi rotations).
// (Out of line on purpose: an inlined version with unrolled copies costs the main path 50 registers -> spills, +6 %.)
template <int P>
__device__ __noinline__ void vm_posdef(double (&V)[P * (P + 1) / 2])
{
    double A[P * P], Bm[P * P], sv[P];
    for (int a = 0; a < P; a++)
        for (int b = 0; b < P; b++) A[a * P + b] = V[a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a];
    int cov = 0;
    mg::make_posdef(A, P, Bm, sv, &cov);
    for (int a = 0; a < P; a++)
        for (int b = 0; b <= a; b++) V[a * (a + 1) / 2 + b] = A[a * P + b];
}

// NegativeG2LineSearch, the two pieces around its line search (rare: a second derivative <= 0 at the seeds; out of line
// to keep the common path small -- the loop body of this kernel is an instruction-cache problem first).
// Accept the best point along the axis and take the derivatives there:
template <int P>
__device__ __noinline__ void vm_ng_accept(VmState<P> &S)
{
    S.in_ng = false;
    if (S.xvmin != 0.) {
#pragma unroll 1
        for (int a = 0; a < P; a++) {
            S.x0[a] += S.xvmin * S.dir[a];
            S.gs[a] = vm_gstep(S.x0[a], S.fvmin, S.g0[a], S.gb[a], S.g20[a], S.g2b[a], S.gs[a]);
            S.g0[a] = S.gb[a];
            S.g20[a] = S.g2b[a];
        }
        S.f0 = S.fvmin;
    }
}

// ... and start the search along parameter ia, in units of the step Minuit's gradient calculator would hold for it
template <int P>
__device__ __noinline__ void vm_ng_begin(VmState<P> &S, int ia)
{
    constexpr double EPS = mg::EPS, EPS2 = mg::EPS2;
    if (S.ng_iter == 0) {
        // InitialGradientCalculator from the parameter step 0.3 |x| (0.3 for x = 0), then the seed's gradient cycles
#pragma unroll 1
        for (int a = 0; a < P; a++) {
            const double gsmin = 8. * EPS2 * (fabs(S.x0[a]) + EPS2);
            const double dirin = fmax(S.x0[a] == 0 ? 0.3 : 0.3 * fabs(S.x0[a]), gsmin);
            const double g2i = 2.0 / (dirin * dirin);
            S.gs[a] = vm_gstep(S.x0[a], S.f0, g2i * dirin, S.g0[a], g2i, S.g20[a], fmax(gsmin, 0.1 * dirin));
        }
    }
    S.ng_iter++;
    double gdel = 0, slamin = 0;
#pragma unroll 1
    for (int a = 0; a < P; a++) {
        S.dir[a] = 0.0;
        if (a == ia) {
            S.dir[a] = (S.g0[a] < 0) ? S.gs[a] : -S.gs[a];
            gdel = S.dir[a] * S.g0[a];
            if (S.dir[a] != 0) slamin = fabs(S.x0[a] / S.dir[a]);
        }
    }
    S.gdel = gdel;
    if (fabs(slamin) < EPS) slamin = EPS;
    S.slamin = slamin * EPS2;
    S.overal = 1000.; S.undral = -100.;
    S.slam = 1.0;
    S.in_ng = true;
    S.phase = VM_LS_A;
}

// Next step of the minimisation after the evaluation in flight returned (f, g, g2) at x0 + slam * dir (at the seeds
// when fresh).  Mirrors VariableMetricBuilder / MnLineSearch / DavidonErrorUpdator as migrad_core.hpp states them.
template <int P>
__device__ __forceinline__ int vm_advance(VmState<P> &S, double f, const double (&g)[P], const double (&g2)[P])
{
    constexpr double EPS = mg::EPS, EPS2 = mg::EPS2;
    constexpr double toler = 0.05, slambg = 5., alpha = 2., edmval = 0.002 * 0.01;
    constexpr int maxiter = 12;
    S.nev++;
    if (S.nev > vm_max_evals(P)) { VM_REASON(0); return VM_HANDOFF; }
    bool ls_done = false;
    int go = 0;   // 1: begin an iteration, 2: first loop of the line search, 3: its second loop, 4: its inner check

    if (S.phase == VM_FRESH) {
        // ---- MnSeedGenerator: gradient at the seeds, then V0 = diag(1 / g2), dcovar = 1 -- after NegativeG2LineSearch
        // has moved the point to where every second derivative is positive, if one is not
        if (!(f == f)) { VM_REASON(1); return VM_HANDOFF; }       // a chi2 that is not a number
        S.f0 = f;
        S.ng_iter = 0;
        S.in_ng = false;
#pragma unroll
        for (int a = 0; a < P; a++) {
            S.g0[a] = g[a];
            S.g20[a] = g2[a];
        }
        go = 5;
    } else if (S.phase == VM_LS_A) {
        S.niter = 2;
        S.fvmin = S.f0; S.xvmin = 0.0;
        if (f < S.f0) {
            S.fvmin = f; S.xvmin = 1.0;
#pragma unroll
            for (int a = 0; a < P; a++) { S.gb[a] = g[a]; S.g2b[a] = g2[a]; }
        }
        S.toler8 = toler; S.slamax = slambg; S.flast = f; S.slam = 1.0;
        S.p0x = 0.0; S.p0y = S.f0; S.p1x = 1.0; S.p1y = f;
        go = 2;
    } else if (S.phase == VM_LS_B) {
        S.niter++;
        if (f < S.fvmin) {
            S.fvmin = f; S.xvmin = S.slam;
#pragma unroll
            for (int a = 0; a < P; a++) { S.gb[a] = g[a]; S.g2b[a] = g2[a]; }
        }
        if (fabs(S.p0y - S.fvmin) < fabs(S.fvmin) * EPS) {
            S.flast = f;
            S.toler8 = toler * S.slam;
            S.overal = S.slam - S.toler8;
            S.slamax = S.overal;
            S.p1x = S.slam; S.p1y = S.flast;
            if (S.niter < maxiter) go = 2;
            else ls_done = true;
        } else if (S.niter >= maxiter) {
            ls_done = true;
        } else {
            S.p2x = S.slam; S.p2y = f;
            go = 3;
        }
    } else {   // VM_LS_C
        if (f > S.p0y && f > S.p1y && f > S.p2y) {
            if (S.slam > S.xvmin) S.overal = fmin(S.overal, S.slam - S.toler8);
            if (S.slam < S.xvmin) S.undral = fmax(S.undral, S.slam + S.toler8);
            S.slam = 0.5 * (S.slam + S.xvmin);
            S.niter++;
            if (S.niter < maxiter) go = 4;
            else ls_done = true;
        } else {
            if (S.p0y > S.p1y && S.p0y > S.p2y) { S.p0x = S.slam; S.p0y = f; }
            else if (S.p1y > S.p0y && S.p1y > S.p2y) { S.p1x = S.slam; S.p1y = f; }
            else { S.p2x = S.slam; S.p2y = f; }
            if (f < S.fvmin) {
                S.fvmin = f; S.xvmin = S.slam;
#pragma unroll
                for (int a = 0; a < P; a++) { S.gb[a] = g[a]; S.g2b[a] = g2[a]; }
            } else {
                if (S.slam > S.xvmin) S.overal = fmin(S.overal, S.slam - S.toler8);
                if (S.slam < S.xvmin) S.undral = fmax(S.undral, S.slam + S.toler8);
            }
            S.niter++;
            if (S.niter < maxiter) go = 3;
            else ls_done = true;
        }
    }

    // ---- the straight-line pieces between evaluations
    if (go == 2) {   // first loop of MnLineSearch: the point where the parabola through f0, gdel and the last value has its minimum
        double denom = vm_div(2. * (S.flast - S.f0 - S.gdel * S.slam), S.slam * S.slam);
        if (denom != 0) S.slam = vm_div(-S.gdel, denom);
        else S.slam = 1.;
        if (S.slam < 0.) S.slam = S.slamax;
        if (S.slam > S.slamax) S.slam = S.slamax;
        if (S.slam < S.toler8) S.slam = S.toler8;
        if (S.slam < S.slamin) ls_done = true;
        else if (fabs(S.slam - 1.) < S.toler8 && S.p1y < S.p0y) ls_done = true;
        else {
            if (fabs(S.slam - 1.) < S.toler8) S.slam = 1. + S.toler8;
            S.phase = VM_LS_B;
            return VM_EVAL;
        }
    }
    if (go == 3) {   // second loop: parabola through the three best points
        S.slamax = fmax(S.slamax, alpha * fabs(S.xvmin));
        double x1 = S.p0x, x2 = S.p1x, x3 = S.p2x;
        const double dx12 = x1 - x2, dx13 = x1 - x3, dx23 = x2 - x3;
        const double xm = (x1 + x2 + x3) * (1.0 / 3.0);
        x1 -= xm; x2 -= xm; x3 -= xm;
        // (three reciprocals instead of Minuit's six quotients: the last bit of lambda is far below what the analytic
        // gradients already differ by)
        const double q0 = vm_div(S.p0y, dx12 * dx13), q1 = vm_div(S.p1y, dx12 * dx23), q2 = vm_div(S.p2y, dx13 * dx23);
        const double pa = q0 - q1 + q2;
        double pb = -q0 * (x2 + x3) + q1 * (x1 + x3) - q2 * (x1 + x2);
        pb -= 2. * xm * pa;
        if (pa < EPS2) {
            const double slopem = 2. * pa * S.xvmin + pb;
            S.slam = (slopem < 0.) ? S.xvmin + S.slamax : S.xvmin - S.slamax;
        } else {
            S.slam = vm_div(-pb, 2. * pa);
            if (S.slam > S.xvmin + S.slamax) S.slam = S.xvmin + S.slamax;
            if (S.slam < S.xvmin - S.slamax) S.slam = S.xvmin - S.slamax;
        }
        if (S.slam > 0.) { if (S.slam > S.overal) S.slam = S.overal; }
        else { if (S.slam < S.undral) S.slam = S.undral; }
        go = 4;
    }
    if (go == 4) {   // a point too close to one already known ends the search
        const double toler9 = fmax(S.toler8, fabs(S.toler8 * S.slam));
        if (fabs(S.p0x - S.slam) < toler9 || fabs(S.p1x - S.slam) < toler9 || fabs(S.p2x - S.slam) < toler9) ls_done = true;
        else {
            S.phase = VM_LS_C;
            return VM_EVAL;
        }
    }

    double edm_s = 0;   // EDM x (1 + 3 dcovar), what the loop condition of VariableMetricBuilder looks at
    bool stop = false;
    if (ls_done && S.in_ng) {
        vm_ng_accept<P>(S);
        ls_done = false;
        go = 5;
    }
    if (go == 5) {
        int ia = -1;
        if (S.ng_iter <= 2 * P) {
#pragma unroll
            for (int a = P - 1; a >= 0; a--)
                if (S.g20[a] <= 0 && !(fabs(S.g0[a]) < EPS && fabs(S.g20[a]) < EPS)) ia = a;   // the first such parameter
        }
        if (ia < 0) {   // every second derivative is positive (or the search has run its 2 P + 1 rounds): the seed state
#pragma unroll
            for (int i = 0; i < P * (P + 1) / 2; i++) S.V[i] = 0.0;
#pragma unroll
            for (int a = 0; a < P; a++) S.V[a * (a + 1) / 2 + a] = (fabs(S.g20[a]) > EPS2) ? vm_div(1.0, S.g20[a]) : 1.0;
            S.dcovar = 1.0;
            S.edm = vm_edm<P>(S.V, S.g0);
            if (!(S.edm >= 0.)) { VM_REASON(1); return VM_HANDOFF; }   // still not positive: Migrad's seed is invalid, strategy 2 follows
            go = 1;
        } else {
            vm_ng_begin<P>(S, ia);
            return VM_EVAL;
        }
    }
    if (ls_done) {
        if (fabs(S.fvmin - S.f0) <= fabs(S.f0) * EPS) {   // no improvement along the line: the iteration loop ends
            edm_s = S.edm * (1. + 3. * S.dcovar);
            stop = true;
        } else {
            double gn[P], dx[P], dg[P], vg[P];
#pragma unroll
            for (int a = 0; a < P; a++) {
                gn[a] = S.gb[a];
                dx[a] = S.xvmin * S.dir[a];
                dg[a] = gn[a] - S.g0[a];
            }
            double edm = vm_edm<P>(S.V, gn);
            if (!(edm == edm)) { VM_REASON(2); return VM_HANDOFF; }
            if (edm < 0.) {   // "Matrix not pos.def., try to make pos.def."
                vm_posdef<P>(S.V);
                edm = vm_edm<P>(S.V, gn);
                if (edm < 0.) {   // Migrad returns the state it had: valid, the final EDM check decides
                    if (S.edm > 10. * edmval) { VM_REASON(2); return VM_HANDOFF; }
                    return VM_DONE;
                }
            }
            // DavidonErrorUpdator
            double delgam = 0, gvg = 0;
#pragma unroll
            for (int a = 0; a < P; a++) {
                double r = 0;
#pragma unroll
                for (int b = 0; b < P; b++) r = fma(S.V[a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a], dg[b], r);
                vg[a] = r;
                delgam = fma(dx[a], dg[a], delgam);
            }
#pragma unroll
            for (int a = 0; a < P; a++) gvg = fma(dg[a], vg[a], gvg);
            if (delgam != 0 && gvg > 0) {
                const double rd = vm_div(1.0, delgam), rg = vm_div(1.0, gvg);
                const bool rank2 = delgam > gvg;
                double sum_upd = 0, sum_v = 0;
#pragma unroll
                for (int a = 0; a < P; a++)
#pragma unroll
                    for (int b = 0; b <= a; b++) {
                        double u = dx[a] * dx[b] * rd - vg[a] * vg[b] * rg;
                        if (rank2) u += gvg * (dx[a] * rd - vg[a] * rg) * (dx[b] * rd - vg[b] * rg);
                        sum_upd += fabs(u);
                        const double v = S.V[a * (a + 1) / 2 + b] + u;
                        S.V[a * (a + 1) / 2 + b] = v;
                        sum_v += fabs(v);
                    }
                S.dcovar = 0.5 * (S.dcovar + vm_div(sum_upd, sum_v));
            }
#pragma unroll
            for (int a = 0; a < P; a++) { S.x0[a] += dx[a]; S.g0[a] = gn[a]; }
            S.f0 = S.fvmin;
            S.edm = edm;
            S.iters++;
            edm_s = edm * (1. + 3. * S.dcovar);
            if (edm_s > edmval) go = 1;
            else stop = true;
        }
    }
    if (stop) {   // VariableMetricBuilder's verdict on leaving the loop
        if (edm_s > edmval && !(edm_s < 10. * edmval) && !(edm_s < fabs(EPS2 * S.f0))) { VM_REASON(3); return VM_HANDOFF; }   // above the EDM limit
        return VM_DONE;
    }
    // ---- go == 1: a new iteration: direction -V g, slope, the first point of its line search (lambda = 1)
    double gdel = 0;
#pragma unroll
    for (int a = 0; a < P; a++) {
        double r = 0;
#pragma unroll
        for (int b = 0; b < P; b++) r = fma(S.V[a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a], S.g0[b], r);
        S.dir[a] = -r;
        gdel = fma(-r, S.g0[a], gdel);
    }
    if (gdel > 0.) {   // not a descent direction: MnPosDef, then the direction again
        vm_posdef<P>(S.V);
        gdel = 0;
#pragma unroll
        for (int a = 0; a < P; a++) {
            double r = 0;
#pragma unroll
            for (int b = 0; b < P; b++) r = fma(S.V[a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a], S.g0[b], r);
            S.dir[a] = -r;
            gdel = fma(-r, S.g0[a], gdel);
        }
        if (gdel > 0.) {   // Migrad gives up and returns the state it has (valid; the final EDM check decides)
            if (S.edm > 10. * edmval) { VM_REASON(4); return VM_HANDOFF; }
            return VM_DONE;
        }
    }
    S.gdel = gdel;
    double slamin = 0.;
#pragma unroll
    for (int a = 0; a < P; a++) {
        if (S.dir[a] != 0) {
            const double ratio = fabs(vm_div(S.x0[a], S.dir[a]));
            if (slamin == 0 || ratio < slamin) slamin = ratio;
        }
    }
    if (fabs(slamin) < EPS) slamin = EPS;
    S.slamin = slamin * EPS2;
    S.overal = 1000.; S.undral = -100.;
    S.slam = 1.0;
    S.phase = VM_LS_A;
    return VM_EVAL;
}

constexpr int VM_TILE_BYTES = (NFIT * FT_LD * (int)sizeof(float) + 15) / 16 * 16;   // the samples only: the weight is a function of the sample
constexpr int VM_WARP_BYTES = VM_TILE_BYTES + 20 * (int)sizeof(double);   // + the 20 pedestal samples of the trace being loaded
#ifndef NPSWF_VM_WARPS
#define NPSWF_VM_WARPS 4
#endif
#ifndef NPSWF_VM_MINBLOCKS
#define NPSWF_VM_MINBLOCKS 3
#endif
#ifndef NPSWF_VM_MINBLOCKS4   // 4-6 pulses: 9-13 parameters, the evaluation's sums alone are 40-56 registers
#define NPSWF_VM_MINBLOCKS4 2
#endif
#ifndef NPSWF_VM_MAXN         // multiplicities with an analytic-path kernel; above: the exact Migrad kernels
#define NPSWF_VM_MAXN 6
#endif
#ifndef NPSWF_VM_MINBLOCKS1   // single-pulse instance: 114 registers as it stands; a fourth CTA measures the same
#define NPSWF_VM_MINBLOCKS1 3
#endif
constexpr int VM_THREADS = NPSWF_VM_WARPS * 32;
constexpr size_t VM_SMEM = (size_t)NPSWF_VM_WARPS * VM_WARP_BYTES;   // 11 880 B per warp -> registers, not shared memory, set the residency
#ifndef NPSWF_VM_DIV_CALL
#define NPSWF_VM_DIV_CALL 1
#endif
#ifndef NPSWF_VM_LOCKSTEP
#define NPSWF_VM_LOCKSTEP 1
#endif

// chi2, d chi2 / d p_a and d2 chi2 / d p_a^2 of one fit at parameters p, all 90 points, one thread (the arithmetic of
// eval_thread, kernel_fit_thread.cuh, without the off-diagonal normal matrix).  The tile holds the samples as binary32
// (exact: other traces do not come here); the weight 1 / Err (T2:946-956) is recomputed per point on the FP32 / MUFU
// pipes, which this kernel leaves idle -- bit for bit what inv_err_f32 stores for the Levenberg-Marquardt kernel.
template <int N, int U>
__device__ __forceinline__ void eval_vm(const double (&p)[2 * N + 1], const float *__restrict__ ycol, const double2 *__restrict__ kn,
                                        double &c2_out, double (&g_out)[2 * N + 1], double (&g2_out)[2 * N + 1])
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
        double u = (double)MFSTART - p[1 + 2 * n];
        u = fmin(fmax(u, -120.0), 120.0);   // a wild trial step stays inside the zero padding of the knot array
        const double fl = floor(u);
        const int i0 = (int)fl;
        const double f = u - fl, g = 1.0 - f;
        wa[n] = g; wb[n] = f;
        wc[n] = (g * g * g - g) * (1.0 / 3.0); wd[n] = (f * f * f - f) * (1.0 / 3.0);
        dc[n] = (1.0 - 3.0 * g * g) * (1.0 / 3.0); dd[n] = (3.0 * f * f - 1.0) * (1.0 / 3.0);
        e0[n] = 2.0 * g; e1[n] = 2.0 * f;
        nA[n] = -p[2 + 2 * n];
        int jl = (int)floor(1.0 - u) + 1, jh = (int)ceil((double)(T - 1) - u) - 1;   // 1 < u + j < 109  (T2:629)
        jl = max(jl, 0);
        jh = min(jh, NFIT - 1);
        if (jh < jl) { jl = 1 << 20; jh = jl; }
        jlo[n] = jl; span[n] = (unsigned)(jh - jl);
        kp[n] = kn + i0;
        k0[n] = __ldg(kp[n]);
    }
    double c2 = 0, gs[P], hs[P], s2[N];
#pragma unroll
    for (int i = 0; i < P; i++) { gs[i] = 0; hs[i] = 0; }
#pragma unroll
    for (int n = 0; n < N; n++) s2[n] = 0;
    const double p0 = p[0];
    float yA[U], yB[U];
    double2 kA[N][U], kB[N][U];
    auto load = [&](int jb, float (&yy)[U], double2 (&kk)[N][U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            yy[u] = ycol[(jb + u) * FT_LD];
#pragma unroll
            for (int n = 0; n < N; n++) kk[n][u] = __ldg(kp[n] + jb + u + 1);
        }
    };
    auto consume = [&](int jb, const float (&yy)[U], const double2 (&kk)[N][U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + u;
            const float af = fabsf(yy[u]);
            float rs;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(af));
            const float wf = (af < 0x1.0624dep+3f) ? 0x1.6e5b7ep+1f : rs * 0x1.6e5b7ep+1f;   // = inv_err_f32((double)y)
            const double wk = (double)wf;
            double r = ((double)yy[u] - p0) * wk;
            double J[P], d2w[N];
            J[0] = wk;
#pragma unroll
            for (int n = 0; n < N; n++) {
                const double2 k1 = kk[n][u];
                const double wkm = ((unsigned)(j - jlo[n]) <= span[n]) ? wk : 0.0;
                const double sp = fma(wd[n], k1.y, fma(wc[n], k0[n].y, fma(wb[n], k1.x, wa[n] * k0[n].x)));
                const double ds = fma(dd[n], k1.y, fma(dc[n], k0[n].y, k1.x - k0[n].x));
                const double d2 = fma(e1[n], k1.y, e0[n] * k0[n].y);
                k0[n] = k1;
                d2w[n] = d2 * wkm;
                J[2 + 2 * n] = sp * wkm;
                J[1 + 2 * n] = ds * wkm;   // without its factor -A_n: applied once to the sums after the loop
                r = fma(nA[n], J[2 + 2 * n], r);
            }
            c2 = fma(r, r, c2);
#pragma unroll
            for (int n = 0; n < N; n++) s2[n] = fma(r, d2w[n], s2[n]);
#pragma unroll
            for (int a = 0; a < P; a++) {
                gs[a] = fma(J[a], r, gs[a]);
                hs[a] = fma(J[a], J[a], hs[a]);
            }
        }
    };
    load(0, yA, kA);
#pragma unroll 1
    for (int j0 = 0; j0 < NFIT; j0 += 2 * U) {
        load(j0 + U, yB, kB);
        consume(j0, yA, kA);
        if (j0 + 2 * U < NFIT) load(j0 + 2 * U, yA, kA);
        consume(j0 + U, yB, kB);
    }
    // d chi2 / d p_a = -2 sum r w df/dp_a;  d2 chi2 / d p_a^2 = 2 sum (w df/dp_a)^2 - 2 sum r w d2f/dp_a^2: the second term
    // is zero for the parameters the model is linear in; the time columns get their factor(s) -A_n here
    c2_out = c2;
    g_out[0] = -2.0 * gs[0];
    g2_out[0] = 2.0 * hs[0];
#pragma unroll
    for (int n = 0; n < N; n++) {
        g_out[1 + 2 * n] = -2.0 * gs[1 + 2 * n] * nA[n];
        g2_out[1 + 2 * n] = 2.0 * hs[1 + 2 * n] * (nA[n] * nA[n]) + 2.0 * (s2[n] * nA[n]);
        g_out[2 + 2 * n] = -2.0 * gs[2 + 2 * n];
        g2_out[2 + 2 * n] = 2.0 * hs[2 + 2 * n];
    }
}

template <int N>
__global__ void __launch_bounds__(VM_THREADS, (N == 1) ? NPSWF_VM_MINBLOCKS1 : (N <= 3) ? NPSWF_VM_MINBLOCKS : NPSWF_VM_MINBLOCKS4)
fit_vm_thread_kernel(const int *__restrict__ job_list, const int *__restrict__ job_count, int *__restrict__ job_next,
                  const double *__restrict__ signal, const double *__restrict__ corr_time_HMS, DevCalib cal, KParams kp,
                  double *__restrict__ wftime, double *__restrict__ wfampl, double *__restrict__ chi2_out,
                  double *__restrict__ timewf, double *__restrict__ amplwf, uint8_t *__restrict__ status,
                  DeviceCounters *__restrict__ ctr, int *__restrict__ cont_count, int *__restrict__ cont_list)
{
    constexpr int P = 2 * N + 1;
#ifndef NPSWF_VM_U1   // points per loop body of the evaluation: 1 measures best (code size before instruction-level parallelism)
#define NPSWF_VM_U1 1
#define NPSWF_VM_U2 1
#define NPSWF_VM_U3 1
#endif
    constexpr int U = (N == 1) ? NPSWF_VM_U1 : (N == 2) ? NPSWF_VM_U2 : (N == 3) ? NPSWF_VM_U3 : 1;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char ft_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ywarp = reinterpret_cast<float *>(ft_smem + (size_t)warp * VM_WARP_BYTES);   // [90][33] samples
    const float *ycol = ywarp + lane;
    double *pedbuf = reinterpret_cast<double *>(ft_smem + (size_t)warp * VM_WARP_BYTES + VM_TILE_BYTES);
    const int njobs = *job_count;
    unsigned long long c_ok1 = 0, c_it = 0, c_att = 0, c_ev = 0;

    bool has_job = false, exhausted = false;
    long long item = 0;
    int bn = 0;
    const double2 *kn = cal.knots + KN_LO;
    VmState<P> S;
    S.phase = VM_FRESH; S.nev = 0; S.iters = 0; S.slam = 0;
#pragma unroll
    for (int i = 0; i < P; i++) { S.x0[i] = 0; S.dir[i] = 0; }

    // warp job queue: lane i holds the i-th job id of the current batch of 32
    int q_item = -1, qpos = 32;
    bool drained = false;   // the cursor has passed the end of the list

    for (;;) {
        // ---- hand new jobs to the lanes without one: one trace per turn of the loop, the next one's loads in flight
        // meanwhile (a round costs a few hundred issue slots whatever the number of idle lanes, so it waits for four
        // of them unless the warp has nothing else to do; ONE copy of the code: the loop body of the kernel has to
        // stay small, see the note on lockstep below)
        unsigned m = __ballot_sync(FULL, !has_job && !exhausted);
        if (__popc(m) < 4 && __any_sync(FULL, has_job)) m = 0;
        bool got = false, inexact = false;
        double ped = 0;
        int l_cur = -1, l_nxt = -1;
        long long it_cur = -1, it_nxt = -1;
        double v_cur[4], v_nxt[4];
        // claim the next idle lane and the next job (warp-uniform), start the loads of its trace
        auto claim = [&](int &l_out, long long &it_out, double (&v)[4]) {
            l_out = -1; it_out = -1;
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) v[c4] = 0;
            if (!m) return;
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
            l_out = __ffs(m) - 1;
            m &= m - 1;
            const int it = (qpos < 32) ? __shfl_sync(FULL, q_item, qpos) : -1;
            if (qpos < 32) qpos++;
            it_out = it;
            if (it >= 0) {
                const double *src = signal + (size_t)it * T;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    const int c = 32 * c4 + lane;
                    if (c < T) v[c4] = src[c];
                }
            }
        };
        claim(l_cur, it_cur, v_cur);
#pragma unroll 1
        while (l_cur >= 0) {   // warp-uniform
            claim(l_nxt, it_nxt, v_nxt);
            // pedestal seed = mean of the first 20 samples (T2:671-677), summed in the reference's order (through shared
            // memory: a shuffle tree inside this loop compiles to five guarded copies of every shuffle)
            __syncwarp();
            if (lane < 20) pedbuf[lane] = v_cur[0];
            __syncwarp();
            double sum = 0;
#pragma unroll
            for (int i = 0; i < 20; i += 2) {
                const double2 pv = *reinterpret_cast<const double2 *>(pedbuf + i);
                sum += pv.x;
                sum += pv.y;
            }
            bool exact = true;
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const int c = 32 * c4 + lane;
                if (c >= MFSTART && c < MFEND && it_cur >= 0) {
                    const float yf = (float)v_cur[c4];
                    exact = exact && ((double)yf == v_cur[c4]) && (fabsf(yf) <= 3.0e38f);   // NaN fails the first test, +-Inf the second
                    ywarp[(c - MFSTART) * FT_LD + l_cur] = yf;
                }
            }
            exact = __all_sync(FULL, exact);
            if (lane == l_cur) {
                if (it_cur >= 0) { item = it_cur; ped = sum / 20; got = true; inexact = !exact; }
                else exhausted = true;
            }
            l_cur = l_nxt; it_cur = it_nxt;
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) v_cur[c4] = v_nxt[c4];
        }
        __syncwarp();
        if (got) {
            bn = (int)(item % B);
            kn = cal.knots + (size_t)bn * KN_LEN + KN_LO;
            const double tref = cal.timeref[bn];
            S.x0[0] = ped;
#pragma unroll
            for (int n = 0; n < N; n++) {
                S.x0[1 + 2 * n] = dsub(wftime[(size_t)item * MAXP + n], tref);   // wftime - timeref   T2:662
                S.x0[2 + 2 * n] = wfampl[(size_t)item * MAXP + n];               // wfampl             T2:663
            }
            has_job = true;
            S.phase = VM_FRESH; S.iters = 0; S.slam = 0;
            S.nev = inexact ? (1 << 20) : 0;   // samples not exact in binary32: the fit goes to the Migrad kernels, which take any doubles
        }
#if NPSWF_VM_LOCKSTEP
        // the warps of a CTA walk the loop body together: the body is larger than the instruction cache next to the SM,
        // and warps that drift apart each stream it from L2 on their own
        if (!__syncthreads_or(has_job)) break;
#else
        if (!__any_sync(FULL, has_job)) break;
#endif

        // ---- one chi2 evaluation (value, gradient, second derivatives along the axes) at the point the lane's
        // minimisation asks for: the seeds, or a point of the current line search
        double trial[P];
#pragma unroll
        for (int i = 0; i < P; i++) trial[i] = S.x0[i] + (S.phase == VM_FRESH ? 0.0 : S.slam * S.dir[i]);
        double fv, g[P], g2[P];
        eval_vm<N, U>(trial, ycol, kn, fv, g, g2);
        bool finished = false, handoff = false;
        if (has_job) {
            c_ev++;
            const int act = vm_advance<P>(S, fv, g, g2);
            finished = act == VM_DONE;
            handoff = act == VM_HANDOFF;
        }
        if (handoff) {   // the Migrad kernels run the fit from its seeds, with the reference's retry / fall-back policy
            cont_list[atomicAdd(cont_count, 1)] = (int)item;
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
                const double oa = S.x0[2 + 2 * n];
                const double ot = dsub(dsub(dadd(dmul(S.x0[1 + 2 * n], kp.dt), corr), cort), accdt);
                wftime[(size_t)item * MAXP + n] = ot;
                wfampl[(size_t)item * MAXP + n] = oa;
                if (n == 0 || fabs(ot) < fabs(bt)) { bt = ot; ba = oa; }
            }
            chi2_out[item] = S.f0 / (double)(NFIT - P);
            if (timewf) timewf[item] = bt;
            if (amplwf) amplwf[item] = ba;
            if (status) status[item] = (uint8_t)(NPSWF_ST_PRESENT | NPSWF_ST_OKTOFIT | NPSWF_ST_FIT_OK1);
            c_ok1++;
            c_it += S.iters;
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
