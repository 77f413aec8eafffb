// Minuit2 Migrad as ROOT::Fit::Fitter drives it from the reference's Fitwf (/root/reference/TEST_2.C:693-773):
// unbounded parameters, Numerical2PGradientCalculator (SetFunction(wfunc, false), T2:746), MnStrategy 1 for the
// first attempt and 2 for the retry (T2:701, 765), tolerance 0.01, Up = 1, call limit 1000 + 100 P + 5 P^2.
//
// This is the product's own scalar implementation of the published algorithm (F. James, MINUIT; Minuit2's
// MnSeedGenerator, InitialGradientCalculator, Numerical2PGradientCalculator, NegativeG2LineSearch,
// VariableMetricBuilder, MnLineSearch, DavidonErrorUpdator, MnHesse, HessianGradientCalculator, MnPosDef), written
// so that ONE executor runs it -- lane 0 of a warp in fit_migrad_kernel (the other lanes only serve chi2
// evaluations), or a host thread in tests/cpp/migrad_core_host.cpp -- on a fixed workspace with no allocation, no
// recursion and no libm beyond sqrt/fabs.  The chi2 comes in as a functor: `double fcn(const double *x)`, counting
// its own calls in `fcn.ncalls`, plus `fcn.pair(x, i, vp, vm, f1, f2)`: f1 = chi2 at x with x[i] = vp, f2 = the same with
// x[i] = vm (x[i] itself is left alone; two calls) -- every central difference of Migrad and MnHesse is such a pair,
// and an implementation may share what the two evaluations have in common as long as each value keeps its bits
// (MG_DEFAULT_PAIR is the plain two-call version).
//
// The translation unit that instantiates this for the device is compiled with -fmad=false: Minuit2 in ROOT is
// built without FMA contraction (x86-64 baseline), and every comparison below (step tolerances, EDM goal, line-search
// brackets) sits on top of these products and sums.
#pragma once

#include <cmath>
#if defined(__CUDACC__)
#define MG_HD __host__ __device__
#else
#define MG_HD
#endif

// the plain implementation of Fcn::pair: two evaluations
#define MG_DEFAULT_PAIR(QUAL)                                                                 \
    QUAL void pair(double *x, int i, double vp, double vm, double &f1, double &f2)            \
    {                                                                                         \
        const double keep = x[i];                                                             \
        x[i] = vp; f1 = (*this)(x);                                                           \
        x[i] = vm; f2 = (*this)(x);                                                           \
        x[i] = keep;                                                                          \
    }

namespace npswf {
namespace mg {

// MnMachinePrecision: eps = 4 * DBL_EPSILON = 2^-50, eps2 = 2 sqrt(eps) = 2^-24 (both exact)
constexpr double EPS = 8.8817841970012523e-16;
constexpr double EPS2 = 5.9604644775390625e-08;

enum CovStatus { COV_POSDEF = 0, COV_MADE_POSDEF = 1, COV_NOT_POSDEF = 2, COV_HESSE_FAILED = 3, COV_INVERT_FAILED = 4, COV_CALL_LIMIT = 5 };
enum MinStatus { MIN_VALID = 0, MIN_ABOVE_MAX_EDM = 1, MIN_CALL_LIMIT = 2 };

struct Strategy {   // MnStrategy::SetLowStrategy / SetMediumStrategy / SetHighStrategy
    int level, grad_ncyc;
    double grad_step_tol, grad_tol;
    int hess_ncyc;
    double hess_step_tol, hess_g2_tol;
    int hess_grad_ncyc;
};
MG_HD inline Strategy make_strategy(int level)
{
    if (level <= 0) return Strategy{0, 2, 0.5, 0.1, 3, 0.5, 0.1, 1};
    if (level == 1) return Strategy{1, 3, 0.3, 0.05, 5, 0.3, 0.05, 2};
    return Strategy{2, 5, 0.1, 0.02, 7, 0.1, 0.02, 6};
}

// std::max / std::min semantics (the second operand wins only on a strict comparison; NaNs behave as in the C++ library)
MG_HD inline double mx(double a, double b) { return (a < b) ? b : a; }
MG_HD inline double mn(double a, double b) { return (b < a) ? b : a; }

template <int PMAX>
struct Work {
    // state of the minimisation (MinimumState): parameters, inverse Hessian estimate, gradient
    double x[PMAX], g[PMAX], g2[PMAX], gs[PMAX];
    double V[PMAX * PMAX];
    // candidate point of an iteration / MnHesse's own derivative arrays
    double xn[PMAX], gn[PMAX], g2n[PMAX], gsn[PMAX];
    double dir[PMAX], xe[PMAX];
    double va[PMAX], vb[PMAX], vc[PMAX], vd[PMAX], hy[PMAX], hd[PMAX];
    double A[PMAX * PMAX], Bm[PMAX * PMAX];
};

struct Scal {   // the scalar part of a MinimumState
    double fval, edm, dcovar;
    int cov;
    MG_HD bool valid() const { return cov == COV_POSDEF || cov == COV_MADE_POSDEF; }
};

struct Result {
    bool valid;          // FunctionMinimum::IsValid(): what Fitter::LeastSquareFit returns as `ok` (T2:755)
    double fval, edm;
    int ncalls, min_status, cov_status;
};

// ---- small dense helpers on n x n matrices stored full, row-major with leading dimension n ----

template <int PMAX>
MG_HD inline void mat_vec(const double *m, const double *v, double *out, int n)
{
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += m[i * n + j] * v[j];
        out[i] = s;
    }
}
MG_HD inline double dotp(const double *a, const double *b, int n)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
// v^T M v, evaluated as v . (M v)
template <int PMAX>
MG_HD inline double quad_form(const double *m, const double *v, double *tmp, int n)
{
    mat_vec<PMAX>(m, v, tmp, n);
    return dotp(v, tmp, n);
}
MG_HD inline double sum_lower_abs(const double *m, int n)   // dasum over the packed triangle of an LASymMatrix
{
    double s = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) s += fabs(m[i * n + j]);
    return s;
}

// MINUIT's mnvert: in-place inverse of a symmetric positive matrix (scaled Gauss-Jordan on the upper triangle).
// s, q, pp: scratch vectors.  Returns 1 on failure.
MG_HD inline int sym_invert(double *m, int n, double *s, double *q, double *pp)
{
    for (int i = 0; i < n; i++) {
        const double si = m[i * n + i];
        if (si <= 0) return 1;
        s[i] = 1.0 / sqrt(si);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) m[i * n + j] *= s[i] * s[j];
    for (int k = 0; k < n; k++) {
        if (m[k * n + k] == 0) return 1;
        q[k] = 1.0 / m[k * n + k];
        pp[k] = 1.0;
        m[k * n + k] = 0.0;
        for (int j = 0; j < k; j++) { pp[j] = m[j * n + k]; q[j] = m[j * n + k] * q[k]; m[j * n + k] = 0.0; }
        for (int j = k + 1; j < n; j++) { pp[j] = m[k * n + j]; q[j] = -m[k * n + j] * q[k]; m[k * n + j] = 0.0; }
        for (int j = 0; j < n; j++)
            for (int l = j; l < n; l++) m[j * n + l] += pp[j] * q[l];
    }
    for (int j = 0; j < n; j++)
        for (int k = 0; k <= j; k++) {
            const double v = m[k * n + j] * s[k] * s[j];
            m[k * n + j] = v;
            m[j * n + k] = v;
        }
    return 0;
}

// smallest and largest eigenvalue of the symmetric matrix m (destroyed): cyclic Jacobi rotations until the
// off-diagonal mass vanishes
MG_HD inline void sym_eigen_minmax(double *m, int n, double *emin, double *emax)
{
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) off += m[i * n + j] * m[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = m[p * n + q];
                if (apq == 0) continue;
                const double theta = (m[q * n + q] - m[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = m[k * n + p], akq = m[k * n + q];
                    m[k * n + p] = c * akp - s * akq;
                    m[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = m[p * n + k], aqk = m[q * n + k];
                    m[p * n + k] = c * apk - s * aqk;
                    m[q * n + k] = s * apk + c * aqk;
                }
            }
    }
    double lo = m[0], hi = m[0];
    for (int i = 1; i < n; i++) {
        const double d = m[i * n + i];
        if (d < lo) lo = d;
        if (d > hi) hi = d;
    }
    *emin = lo;
    *emax = hi;
}

// MnPosDef on the matrix `err` in place; `pm` and `s` are scratch.  Updates *cov to MADE_POSDEF when the
// diagonal had to be inflated.
MG_HD inline void make_posdef(double *err, int n, double *pm, double *s, int *cov)
{
    if (n == 1 && err[0] < EPS) { err[0] = 1.; *cov = COV_MADE_POSDEF; return; }
    if (n == 1 && err[0] > EPS) return;
    const double epspdf = mx(1.e-6, EPS2);
    double dgmin = err[0];
    for (int i = 0; i < n; i++)
        if (err[i * n + i] < dgmin) dgmin = err[i * n + i];
    double dg = 0.;
    if (dgmin <= 0) dg = 0.5 + epspdf - dgmin;
    for (int i = 0; i < n; i++) {
        err[i * n + i] += dg;
        if (err[i * n + i] < 0.) err[i * n + i] = 1.;
        s[i] = 1. / sqrt(err[i * n + i]);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) pm[i * n + j] = err[i * n + j] * s[i] * s[j];
    double pmin, pmax;
    sym_eigen_minmax(pm, n, &pmin, &pmax);
    pmax = mx(fabs(pmin), fabs(pmax));
    if (pmin > epspdf * pmax) return;
    const double padd = 0.001 * pmax - pmin;
    for (int i = 0; i < n; i++) err[i * n + i] *= (1. + padd);
    *cov = COV_MADE_POSDEF;
}

// InitialGradientCalculator for unbounded parameters: gradient guess from the parameter steps `werr`
template <int PMAX>
MG_HD inline void initial_gradient(Work<PMAX> &W, int n, const double *werr, double up)
{
    for (int i = 0; i < n; i++) {
        const double var = W.x[i];
        const double sav = var;
        double sav2 = sav + werr[i];
        const double vplu = sav2 - var;
        sav2 = sav - werr[i];
        const double vmin = sav2 - var;
        const double gsmin = 8. * EPS2 * (fabs(var) + EPS2);
        const double dirin = mx(0.5 * (fabs(vplu) + fabs(vmin)), gsmin);
        const double g2 = 2.0 * up / (dirin * dirin);
        const double gstep = mx(gsmin, 0.1 * dirin);
        W.g[i] = g2 * dirin;
        W.g2[i] = g2;
        W.gs[i] = gstep;
    }
}

// Numerical2PGradientCalculator: central differences at `x` (f(x) = fcnmin), refining the previous derivative
// estimates (g, g2, gs) IN PLACE over up to grad_ncyc cycles per parameter.  xw: scratch copy of x.
template <class Fcn>
MG_HD inline void numerical_gradient(Fcn &fcn, const double *x, double fcnmin, double *g, double *g2, double *gs,
                                     double *xw, int n, const Strategy &st)
{
    const double dfmin = 8. * EPS2 * (fabs(fcnmin) + 1.0);
    const double vrysml = 8. * EPS * EPS;
    for (int i = 0; i < n; i++) xw[i] = x[i];
    for (int i = 0; i < n; i++) {
        const double xtf = xw[i];
        const double epspri = EPS2 + fabs(g[i] * EPS2);
        double stepb4 = 0.;
        double gi = g[i], g2i = g2[i], gsi = gs[i];
        for (int j = 0; j < st.grad_ncyc; j++) {
            const double optstp = sqrt(dfmin / (fabs(g2i) + epspri));
            double step = mx(optstp, fabs(0.1 * gsi));
            const double stpmax = 10. * fabs(gsi);
            if (step > stpmax) step = stpmax;
            const double stpmin = mx(vrysml, 8. * fabs(EPS2 * xw[i]));
            if (step < stpmin) step = stpmin;
            if (fabs((step - stepb4) / step) < st.grad_step_tol) break;
            gsi = step;
            stepb4 = step;
            double fs1, fs2;
            fcn.pair(xw, i, xtf + step, xtf - step, fs1, fs2);
            const double grdb4 = gi;
            gi = 0.5 * (fs1 - fs2) / step;
            g2i = (fs1 + fs2 - 2. * fcnmin) / step / step;
            if (fabs(grdb4 - gi) / (fabs(gi) + dfmin / step) < st.grad_tol) break;
        }
        g[i] = gi; g2[i] = g2i; gs[i] = gsi;
    }
}

struct LinePoint { double x, y; };

// MnLineSearch: minimum of f along x0 + lambda * dir from f0 = f(x0) and the slope gdel = g . dir
// (parabolic interpolation, at most 12 evaluations).  xe: scratch evaluation point.
template <class Fcn>
MG_HD inline LinePoint line_search(Fcn &fcn, const double *x0, double f0, const double *dir, double gdel, double *xe, int n)
{
    double overal = 1000., undral = -100.;
    const double toler = 0.05, slambg = 5., alpha = 2.;
    const int maxiter = 12;
    double slamin = 0.;
    int niter = 1;
    for (int i = 0; i < n; i++) {
        if (dir[i] == 0) continue;
        const double ratio = fabs(x0[i] / dir[i]);
        if (slamin == 0) slamin = ratio;
        if (ratio < slamin) slamin = ratio;
    }
    if (fabs(slamin) < EPS) slamin = EPS;
    slamin *= EPS2;

#define MG_EVAL_AT(lam, out)                                       \
    do {                                                           \
        for (int _i = 0; _i < n; _i++) xe[_i] = x0[_i] + (lam) * dir[_i]; \
        (out) = fcn(xe);                                           \
    } while (0)

    double f1;
    MG_EVAL_AT(1.0, f1);
    niter++;
    double fvmin = f0, xvmin = 0.;
    if (f1 < f0) { fvmin = f1; xvmin = 1.; }
    double toler8 = toler, slamax = slambg, flast = f1, slam = 1.;
    bool iterate = false;
    LinePoint p0{0., f0}, p1{slam, flast};
    double f2 = 0.;
    do {
        iterate = false;
        double denom = 2. * (flast - f0 - gdel * slam) / (slam * slam);
        if (denom != 0) slam = -gdel / denom;
        else { denom = -0.1 * gdel; slam = 1.; }
        if (slam < 0.) slam = slamax;
        if (slam > slamax) slam = slamax;
        if (slam < toler8) slam = toler8;
        if (slam < slamin) return LinePoint{xvmin, fvmin};
        if (fabs(slam - 1.) < toler8 && p1.y < p0.y) return LinePoint{xvmin, fvmin};
        if (fabs(slam - 1.) < toler8) slam = 1. + toler8;
        MG_EVAL_AT(slam, f2);
        niter++;
        if (f2 < fvmin) { fvmin = f2; xvmin = slam; }
        if (fabs(p0.y - fvmin) < fabs(fvmin) * EPS) {
            iterate = true;
            flast = f2;
            toler8 = toler * slam;
            overal = slam - toler8;
            slamax = overal;
            p1 = LinePoint{slam, flast};
        }
    } while (iterate && niter < maxiter);
    if (niter >= maxiter) return LinePoint{xvmin, fvmin};

    LinePoint p2{slam, f2};
    do {
        slamax = mx(slamax, alpha * fabs(xvmin));
        // parabola through p0, p1, p2 (MnParabolaFactory), y = a x^2 + b x + c
        double x1 = p0.x, x2 = p1.x, x3 = p2.x;
        const double dx12 = x1 - x2, dx13 = x1 - x3, dx23 = x2 - x3;
        const double xm = (x1 + x2 + x3) / 3.;
        x1 -= xm; x2 -= xm; x3 -= xm;
        const double y1 = p0.y, y2 = p1.y, y3 = p2.y;
        const double pa = y1 / (dx12 * dx13) - y2 / (dx12 * dx23) + y3 / (dx13 * dx23);
        double pb = -y1 * (x2 + x3) / (dx12 * dx13) + y2 * (x1 + x3) / (dx12 * dx23) - y3 * (x1 + x2) / (dx13 * dx23);
        pb -= 2. * xm * pa;
        if (pa < EPS2) {
            const double slopem = 2. * pa * xvmin + pb;
            if (slopem < 0.) slam = xvmin + slamax;
            else slam = xvmin - slamax;
        } else {
            slam = -pb / (2. * pa);
            if (slam > xvmin + slamax) slam = xvmin + slamax;
            if (slam < xvmin - slamax) slam = xvmin - slamax;
        }
        if (slam > 0.) { if (slam > overal) slam = overal; }
        else { if (slam < undral) slam = undral; }

        double f3 = 0.;
        do {
            iterate = false;
            const double toler9 = mx(toler8, fabs(toler8 * slam));
            if (fabs(p0.x - slam) < toler9 || fabs(p1.x - slam) < toler9 || fabs(p2.x - slam) < toler9)
                return LinePoint{xvmin, fvmin};
            MG_EVAL_AT(slam, f3);
            if (f3 > p0.y && f3 > p1.y && f3 > p2.y) {
                if (slam > xvmin) overal = mn(overal, slam - toler8);
                if (slam < xvmin) undral = mx(undral, slam + toler8);
                slam = 0.5 * (slam + xvmin);
                iterate = true;
                niter++;
            }
        } while (iterate && niter < maxiter);
        if (niter >= maxiter) return LinePoint{xvmin, fvmin};

        const LinePoint p3{slam, f3};
        if (p0.y > p1.y && p0.y > p2.y) p0 = p3;
        else if (p1.y > p0.y && p1.y > p2.y) p1 = p3;
        else p2 = p3;
        if (f3 < fvmin) { fvmin = f3; xvmin = slam; }
        else {
            if (slam > xvmin) overal = mn(overal, slam - toler8);
            if (slam < xvmin) undral = mx(undral, slam + toler8);
        }
        niter++;
    } while (niter < maxiter);
#undef MG_EVAL_AT
    return LinePoint{xvmin, fvmin};
}

MG_HD inline bool any_nonpositive(const double *g2, int n)
{
    for (int i = 0; i < n; i++)
        if (g2[i] <= 0) return true;
    return false;
}

// V = diag(1/g2) (1 where g2 is tiny), the seed's inverse-Hessian guess
template <int PMAX>
MG_HD inline void diag_from_g2(Work<PMAX> &W, int n)
{
    for (int i = 0; i < n * n; i++) W.V[i] = 0.0;
    for (int i = 0; i < n; i++) W.V[i * n + i] = (fabs(W.g2[i]) > EPS2 ? 1. / W.g2[i] : 1.);
}

// NegativeG2LineSearch: while some second derivative is not positive, slide along that axis to where it is
template <int PMAX, class Fcn>
MG_HD inline void negative_g2_line_search(Fcn &fcn, Work<PMAX> &W, Scal &S, int n, const Strategy &st)
{
    if (!any_nonpositive(W.g2, n)) return;
    bool iterate = false;
    unsigned iter = 0;
    do {
        iterate = false;
        for (int i = 0; i < n; i++) {
            if (W.g2[i] <= 0) {
                if (fabs(W.g[i]) < EPS && fabs(W.g2[i]) < EPS) continue;
                for (int k = 0; k < n; k++) W.dir[k] = 0.0;
                if (W.g[i] < 0) W.dir[i] = W.gs[i];
                else W.dir[i] = -W.gs[i];
                const double gdel = W.dir[i] * W.g[i];
                const LinePoint pp = line_search(fcn, W.x, S.fval, W.dir, gdel, W.xe, n);
                for (int k = 0; k < n; k++) W.x[k] += pp.x * W.dir[k];
                S.fval = pp.y;
                numerical_gradient(fcn, W.x, S.fval, W.g, W.g2, W.gs, W.xe, n, st);
                iterate = true;
                break;
            }
        }
    } while (iter++ < 2 * (unsigned)n && iterate);
    diag_from_g2(W, n);
    S.dcovar = 1.;
    S.cov = COV_POSDEF;
    S.edm = 0.5 * quad_form<PMAX>(W.V, W.g, W.va, n);
    if (S.edm < 0) S.cov = COV_NOT_POSDEF;
}

// MnHesse at the current state: second derivatives by finite differences, inverted into V.  On failure only
// (V, S.cov) change -- parameters and gradient stay, as in MnHesse's error returns.
template <int PMAX, class Fcn>
MG_HD inline void hesse(Fcn &fcn, Work<PMAX> &W, Scal &S, int n, const Strategy &st, unsigned maxcalls)
{
    const double amin = fcn(W.x);
    const double aimsag = sqrt(EPS2) * (fabs(amin) + 1.0);
    if (maxcalls == 0) maxcalls = 200 + 100 * n + 5 * n * n;
    double *vh = W.A;
    double *g2 = W.g2n, *gst = W.gsn, *grd = W.gn, *dirin = W.hd, *yy = W.hy, *x = W.xn;
    for (int i = 0; i < n * n; i++) vh[i] = 0.0;
    for (int i = 0; i < n; i++) { g2[i] = W.g2[i]; gst[i] = W.gs[i]; grd[i] = W.g[i]; dirin[i] = W.gs[i]; yy[i] = 0.0; x[i] = W.x[i]; }
    for (int i = 0; i < n; i++) {
        const double xtf = x[i];
        const double dmin = 8. * EPS2 * (fabs(xtf) + EPS2);
        double d = fabs(gst[i]);
        if (d < dmin) d = dmin;
        for (int icyc = 0; icyc < st.hess_ncyc; icyc++) {
            double sag = 0., fs1 = 0., fs2 = 0.;
            bool got = false;
            for (int multpy = 0; multpy < 5; multpy++) {
                fcn.pair(x, i, xtf + d, xtf - d, fs1, fs2);
                sag = 0.5 * (fs1 + fs2 - 2. * amin);
                if (sag != 0) { got = true; break; }
                d *= 10.;
            }
            if (!got) {   // second derivative zero along this axis
                for (int k = 0; k < n * n; k++) W.V[k] = 0.0;
                S.cov = COV_HESSE_FAILED;
                return;
            }
            const double g2bfor = g2[i];
            g2[i] = 2. * sag / (d * d);
            grd[i] = (fs1 - fs2) / (2. * d);
            gst[i] = d;
            dirin[i] = d;
            yy[i] = fs1;
            const double dlast = d;
            d = sqrt(2. * aimsag / fabs(g2[i]));
            if (d < dmin) d = dmin;
            if (fabs((d - dlast) / d) < st.hess_step_tol) break;
            if (fabs((g2[i] - g2bfor) / g2[i]) < st.hess_g2_tol) break;
            d = mn(d, 10. * dlast);
            d = mx(d, 0.1 * dlast);
        }
        vh[i * n + i] = g2[i];
        if ((unsigned)fcn.ncalls > maxcalls) {
            for (int k = 0; k < n * n; k++) W.V[k] = 0.0;
            S.cov = COV_CALL_LIMIT;
            return;
        }
    }
    if (st.level > 0) {
        // HessianGradientCalculator: refine the first derivatives with the steps found above
        const double dfmin = 4. * EPS2 * (fabs(S.fval) + 1.0);
        double *xp = W.xe;
        for (int i = 0; i < n; i++) {
            const double xtf = W.x[i];
            const double dmin = 4. * EPS2 * (xtf + EPS2);
            const double epspri = EPS2 + fabs(grd[i] * EPS2);
            const double optstp = sqrt(dfmin / (fabs(g2[i]) + epspri));
            double d = 0.2 * fabs(gst[i]);
            if (d > optstp) d = optstp;
            if (d < dmin) d = dmin;
            double chgold = 10000.;
            for (int j = 0; j < st.hess_grad_ncyc; j++) {
                for (int k = 0; k < n; k++) xp[k] = W.x[k];
                double fs1, fs2;
                fcn.pair(xp, i, xtf + d, xtf - d, fs1, fs2);
                const double grdold = grd[i];
                const double grdnew = (fs1 - fs2) / (2. * d);
                const double dgmin = EPS * (fabs(fs1) + fabs(fs2)) / d;
                if (grdnew == 0) break;
                const double change = fabs((grdold - grdnew) / grdnew);
                if (change > chgold && j > 1) break;
                chgold = change;
                grd[i] = grdnew;
                gst[i] = d;
                if (change < 0.05) break;
                if (fabs(grdold - grdnew) < dgmin) break;
                if (d < dmin) break;
                d *= 0.2;
            }
        }
    }
    for (int i = 0; i < n; i++) {   // off-diagonal elements
        x[i] += dirin[i];
        for (int j = i + 1; j < n; j++) {
            x[j] += dirin[j];
            const double fs1 = fcn(x);
            const double elem = (fs1 + amin - yy[i] - yy[j]) / (dirin[i] * dirin[j]);
            vh[i * n + j] = elem;
            vh[j * n + i] = elem;
            x[j] -= dirin[j];
        }
        x[i] -= dirin[i];
    }
    int cov = COV_POSDEF;
    make_posdef(vh, n, W.Bm, W.va, &cov);
    if (sym_invert(vh, n, W.va, W.vb, W.vc) != 0) {
        for (int k = 0; k < n * n; k++) W.V[k] = 0.0;
        for (int j = 0; j < n; j++) {
            double t = g2[j];
            if (fabs(t) < EPS2) t = 1.;
            else t = 1. / t;
            W.V[j * n + j] = (t < EPS2 ? 1. : t);
        }
        S.cov = COV_INVERT_FAILED;
        return;
    }
    for (int i = 0; i < n; i++) { W.g[i] = grd[i]; W.g2[i] = g2[i]; W.gs[i] = gst[i]; }
    for (int k = 0; k < n * n; k++) W.V[k] = vh[k];
    if (cov == COV_MADE_POSDEF) { S.cov = COV_MADE_POSDEF; S.dcovar = 1.; }
    else { S.cov = COV_POSDEF; S.dcovar = 0.; }
    S.edm = 0.5 * quad_form<PMAX>(W.V, W.g, W.va, n);
}

// DavidonErrorUpdator: rank-two update of V for the step x -> xn with gradient change g -> gn.  Leaves V and
// dcovar alone when the update is not defined.
template <int PMAX>
MG_HD inline void davidon_update(Work<PMAX> &W, Scal &S, int n)
{
    double *dx = W.va, *dg = W.vb, *vg = W.vc, *w = W.vd, *upd = W.A;
    for (int i = 0; i < n; i++) { dx[i] = W.xn[i] - W.x[i]; dg[i] = W.gn[i] - W.g[i]; }
    const double delgam = dotp(dx, dg, n);
    const double gvg = quad_form<PMAX>(W.V, dg, vg, n);   // vg = V dg
    if (delgam == 0) return;
    if (gvg <= 0) return;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) upd[i * n + j] = dx[i] * dx[j] / delgam - vg[i] * vg[j] / gvg;
    if (delgam > gvg) {   // rank-two term of the BFGS-like complement
        for (int i = 0; i < n; i++) w[i] = dx[i] / delgam - vg[i] / gvg;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) upd[i * n + j] += gvg * w[i] * w[j];
    }
    const double sum_upd = sum_lower_abs(upd, n);
    for (int i = 0; i < n * n; i++) W.V[i] = upd[i] + W.V[i];
    S.dcovar = 0.5 * (S.dcovar + sum_upd / sum_lower_abs(W.V, n));
    S.cov = COV_POSDEF;
}

// VariableMetricBuilder::Minimum, inner loop: iterate from the state in (W, S) until EDM < edmval or the call limit
template <int PMAX, class Fcn>
MG_HD inline int variable_metric(Fcn &fcn, Work<PMAX> &W, Scal &S, int n, unsigned maxfcn, double edmval, const Strategy &st)
{
    double edm = S.edm;
    edm *= (1. + 3. * S.dcovar);
    do {
        mat_vec<PMAX>(W.V, W.g, W.va, n);
        for (int i = 0; i < n; i++) W.dir[i] = -W.va[i];
        double gdel = dotp(W.dir, W.g, n);
        if (gdel > 0.) {
            make_posdef(W.V, n, W.Bm, W.va, &S.cov);
            mat_vec<PMAX>(W.V, W.g, W.va, n);
            for (int i = 0; i < n; i++) W.dir[i] = -W.va[i];
            gdel = dotp(W.dir, W.g, n);
            if (gdel > 0.) return MIN_VALID;
        }
        const LinePoint pp = line_search(fcn, W.x, S.fval, W.dir, gdel, W.xe, n);
        if (fabs(pp.y - S.fval) <= fabs(S.fval) * EPS) break;   // no improvement
        for (int i = 0; i < n; i++) { W.xn[i] = W.x[i] + pp.x * W.dir[i]; W.gn[i] = W.g[i]; W.g2n[i] = W.g2[i]; W.gsn[i] = W.gs[i]; }
        numerical_gradient(fcn, W.xn, pp.y, W.gn, W.g2n, W.gsn, W.xe, n, st);
        edm = 0.5 * quad_form<PMAX>(W.V, W.gn, W.va, n);
        if (edm != edm) return MIN_VALID;
        if (edm < 0.) {
            make_posdef(W.V, n, W.Bm, W.va, &S.cov);
            edm = 0.5 * quad_form<PMAX>(W.V, W.gn, W.va, n);
            if (edm < 0.) return MIN_VALID;
        }
        davidon_update(W, S, n);
        for (int i = 0; i < n; i++) { W.x[i] = W.xn[i]; W.g[i] = W.gn[i]; W.g2[i] = W.g2n[i]; W.gs[i] = W.gsn[i]; }
        S.fval = pp.y;
        S.edm = edm;
        edm *= (1. + 3. * S.dcovar);
    } while (edm > edmval && (unsigned)fcn.ncalls < maxfcn);

    if ((unsigned)fcn.ncalls >= maxfcn) return MIN_CALL_LIMIT;
    if (edm > edmval) {
        if (edm < 10 * edmval) return MIN_VALID;                // "Edm is close to limit"
        if (edm < fabs(EPS2 * S.fval)) return MIN_VALID;          // machine-accuracy limit
        return MIN_ABOVE_MAX_EDM;
    }
    return MIN_VALID;
}

// ModularFunctionMinimizer::Minimize: MnSeedGenerator + VariableMetricBuilder::Minimum.  Parameters start at
// `start` with steps `werr`; on return W.x holds the parameters of the last state.
template <int PMAX, class Fcn>
MG_HD inline Result migrad(Fcn &fcn, Work<PMAX> &W, int n, const double *start, const double *werr, int level,
                           unsigned maxfcn, double tolerance)
{
    const Strategy st = make_strategy(level);
    Scal S;
    if (maxfcn == 0) maxfcn = 200 + 100 * n + 5 * n * n;
    double edmval = tolerance * 1.0;
    if (edmval < EPS2) edmval = EPS2;

    // ---- seed
    for (int i = 0; i < n; i++) W.x[i] = start[i];
    S.fval = fcn(W.x);
    initial_gradient(W, n, werr, 1.0);
    numerical_gradient(fcn, W.x, S.fval, W.g, W.g2, W.gs, W.xe, n, st);
    diag_from_g2(W, n);
    S.dcovar = 1.;
    S.cov = COV_POSDEF;
    S.edm = 0.5 * quad_form<PMAX>(W.V, W.g, W.va, n);
    negative_g2_line_search(fcn, W, S, n, st);
    if (st.level == 2) hesse(fcn, W, S, n, st, 0);

    Result res;
    int ms = MIN_VALID;
#define MG_FINISH()                                          \
    do {                                                     \
        res.fval = S.fval; res.edm = S.edm; res.ncalls = fcn.ncalls; \
        res.min_status = ms; res.cov_status = S.cov;         \
        res.valid = S.valid() && ms == MIN_VALID;            \
        return res;                                          \
    } while (0)

    if ((unsigned)fcn.ncalls >= maxfcn) { ms = MIN_CALL_LIMIT; MG_FINISH(); }
    edmval *= 0.002;
    if (!S.valid()) MG_FINISH();
    if (S.edm < 0.) MG_FINISH();
    double edm = S.edm;
    unsigned maxfcn_eff = maxfcn;
    int ipass = 0;
    bool iterate = false;
    do {
        iterate = false;
        ms = variable_metric(fcn, W, S, n, maxfcn_eff, edmval, st);
        if (ms == MIN_CALL_LIMIT) MG_FINISH();
        if (ipass > 0) {
            if (!(S.valid() && ms == MIN_VALID)) MG_FINISH();
        }
        edm = S.edm;
        if (st.level == 2 || (st.level == 1 && S.dcovar > 0.05)) {
            hesse(fcn, W, S, n, st, maxfcn);
            if (!S.valid()) break;
            edm = S.edm;
            if (edm > edmval) {
                const double machine_limit = fabs(EPS2 * S.fval);
                if (edm >= machine_limit) iterate = true;
            }
        }
        if (ipass == 0) maxfcn_eff = (unsigned)(maxfcn * 1.3);
        ipass++;
    } while (iterate);

    if (edm > 10 * edmval) ms = MIN_ABOVE_MAX_EDM;
    else if (ms == MIN_ABOVE_MAX_EDM) ms = MIN_VALID;   // "Edm has been re-computed after Hesse; now within tolerance"
    MG_FINISH();
#undef MG_FINISH
}

// ---- the chi2 of Fitwf and the attempt / retry policy around Migrad (T2:621-635, 680-688, 755-773) ----

constexpr int FIT_T = 110;        // ntime                      T2:51
constexpr int FIT_X0 = 10;        // first fitted sample        T2:681
constexpr int FIT_NPT = 90;       // fitted samples 10..99      T2:681

// Interval index of GSL's gsl_interp_bsearch over the block's knots: knots[i] <= d < knots[i+1], clamped to [0, T-2]
MG_HD inline int knot_interval(const double *knots, double d)
{
    int lo = 0, hi = FIT_T - 1;
    while (hi > lo + 1) {
        const int i = (hi + lo) / 2;
        if (knots[i] > d) hi = i;
        else lo = i;
    }
    return lo;
}

// One term of ROOT::Fit::Chi2FCN (FitUtil::EvaluateChi2): ((y - f(x; par)) / err)^2 at sample k of the fit window,
// f the model of T2:621-635 on the natural cubic spline `spl` = [109][4] (y, b, c, d per interval).
// knots == nullptr: the knots are the sample indices 0..109 (interval = integer part).
MG_HD inline double chi2_term(int k, const double *par, int N, const double *spl, const double *knots, double y, double w)
{
    const double x = (double)(FIT_X0 + k);
    double val = par[0];
    for (int p = 0; p < N; p++) {
        const double dt0 = x - par[1 + 2 * p];
        if (dt0 > 1 && dt0 < FIT_T - 1) {   // T2:629
            int i;
            double delx;
            if (knots) { i = knot_interval(knots, dt0); delx = dt0 - knots[i]; }
            else { i = (int)dt0; delx = dt0 - (double)i; }
#if defined(__CUDA_ARCH__)
            const double2 c01 = __ldg(reinterpret_cast<const double2 *>(spl + 4 * i));       // 32-byte aligned quads
            const double2 c23 = __ldg(reinterpret_cast<const double2 *>(spl + 4 * i) + 1);
            val += par[2 + 2 * p] * (c01.x + delx * (c01.y + delx * (c23.x + delx * c23.y)));
#else
            const double *c = spl + 4 * i;
            val += par[2 + 2 * p] * (c[0] + delx * (c[1] + delx * (c[2] + delx * c[3])));
#endif
        }
    }
    const double tmp = (y - val) * w;
    return tmp * tmp;
}

// Err of T2:946-956 as ROOT::Fit::BinData stores it (inverse error)
MG_HD inline double inv_err(double y)
{
    double e = sqrt(fabs(y * 4.096 / 2.)) / 4.096;
    if (e < 1.) e = sqrt(fabs(1.0 * 4.096 / 2.)) / 4.096;
    return (e != 0.0) ? 1.0 / e : 0.0;
}

// Seeds of Fitwf (T2:656-677): pedestal = mean of the first 20 samples, pulse times relative to the block's
// reference time, raw amplitudes
MG_HD inline void fit_seeds(const double *trace, double timeref, const double *wftime, const double *wfampl, int N, double *start)
{
    double ped = 0;
    for (int i = 0; i < 20; i++) ped += trace[i];
    ped /= 20;
    start[0] = ped;
    for (int p = 0; p < N; p++) {
        start[1 + 2 * p] = wftime[p] - timeref;
        start[2 + 2 * p] = wfampl[p];
    }
}

struct FitOutcome {
    int status;      // NPSWF_ST_FIT_OK1 (4), NPSWF_ST_FIT_OK2 (8) or NPSWF_ST_FALLBACK (16)
    double fmin;     // chi2 at the minimum (valid fits)
    int ncalls;      // chi2 evaluations of both attempts
};

// Fitwf's minimisation policy: Migrad strategy 1 from the seeds; if the minimum is not valid, strategy 2 from the
// SAME seeds (the fit configuration is only updated after a success, T2:761-768); if that fails too the caller falls
// back to the TSpectrum values.  start[P]: seeds (T2:656-677); werr[P]: scratch for FitConfig's parameter steps,
// 0.3 |value| (0.3 when the value is 0).  The fitted parameters are left in W.x.
template <int PMAX, class Fcn>
MG_HD inline FitOutcome fitwf_minimise(Fcn &fcn, Work<PMAX> &W, int N, const double *start, double *werr)
{
    const int P = 2 * N + 1;
    for (int i = 0; i < P; i++) werr[i] = (start[i] == 0) ? 0.3 : 0.3 * fabs(start[i]);
    const unsigned maxfcn = 1000u + 100u * (unsigned)P + 5u * (unsigned)P * (unsigned)P;   // FitConfig::CreateMinimizer
    FitOutcome out;
    fcn.ncalls = 0;
    Result r = migrad<PMAX>(fcn, W, P, start, werr, 1, maxfcn, 0.01);
    out.ncalls = r.ncalls;
    if (r.valid) { out.status = 4; out.fmin = r.fval; return out; }
    fcn.ncalls = 0;
    r = migrad<PMAX>(fcn, W, P, start, werr, 2, maxfcn, 0.01);
    out.ncalls += r.ncalls;
    out.fmin = r.fval;
    out.status = r.valid ? 8 : 16;
    return out;
}

}  // namespace mg
}  // namespace npswf
