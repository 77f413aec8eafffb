// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see oracle/README.md).
//
// CPU restatement of Minuit2's Migrad as ROOT::Fit::Fitter drives it from the reference
// (/root/reference/TEST_2.C:693-773): unbounded parameters, numerical 2-point gradients
// (Fitter::SetFunction(wfunc,false), T2:746), strategy 1 then strategy 2, tolerance 0.01,
// Up = 1 (least squares).  Minuit2 (ROOT math/minuit2, v6.30) is NOT vendored under
// /root/reference and is not installed here; this file restates the published algorithm
// (F. James, MINUIT; Minuit2 sources: MnSeedGenerator, InitialGradientCalculator,
// Numerical2PGradientCalculator, NegativeG2LineSearch, VariableMetricBuilder, MnLineSearch,
// DavidonErrorUpdator, MnHesse, HessianGradientCalculator, MnPosDef).
// PARITY UNPINNED: it could not be run against Minuit2 itself in this environment.
#include "npswf_oracle.h"
#include "minuit_migrad.hpp"
#include <algorithm>
#include <cfloat>
#include <cmath>

namespace ormn {

namespace {

struct Prec {
    double eps = 4.0 * DBL_EPSILON;               // MnMachinePrecision::fEpsMac
    double eps2 = 2.0 * std::sqrt(4.0 * DBL_EPSILON);  // fEpsMa2
};

struct Strategy {
    int level, gradNCyc;
    double gradStepTol, gradTol;
    int hessNCyc;
    double hessStepTol, hessG2Tol;
    int hessGradNCyc;
};

Strategy make_strategy(int level)
{
    if (level <= 0) return {0, 2, 0.5, 0.1, 3, 0.5, 0.1, 1};
    if (level == 1) return {1, 3, 0.3, 0.05, 5, 0.3, 0.05, 2};
    return {2, 5, 0.1, 0.02, 7, 0.1, 0.02, 6};
}

using Vec = std::vector<double>;

struct Mat {  // dense symmetric n x n, stored full
    int n = 0;
    Vec a;
    Mat() {}
    explicit Mat(int n_) : n(n_), a((size_t)n_ * n_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i * n + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * n + j]; }
    void set_sym(int i, int j, double v) { a[(size_t)i * n + j] = v; a[(size_t)j * n + i] = v; }
};

// LASymMatrix sum_of_elements = dasum over the packed triangle
double sum_packed_abs(const Mat &m)
{
    double s = 0;
    for (int i = 0; i < m.n; i++)
        for (int j = 0; j <= i; j++) s += std::fabs(m(i, j));
    return s;
}

Vec matvec(const Mat &m, const Vec &v)
{
    Vec r(m.n, 0.0);
    for (int i = 0; i < m.n; i++) {
        double s = 0;
        for (int j = 0; j < m.n; j++) s += m(i, j) * v[j];
        r[i] = s;
    }
    return r;
}
double dot(const Vec &a, const Vec &b)
{
    double s = 0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}
double similarity(const Vec &v, const Mat &m) { return dot(v, matvec(m, v)); }

// MINUIT mnvert: in-place inverse of a symmetric matrix; returns 1 on failure
int invert_sym(Mat &m)
{
    const int n = m.n;
    Vec s(n), q(n), pp(n);
    for (int i = 0; i < n; i++) {
        double si = m(i, i);
        if (si <= 0) return 1;
        s[i] = 1.0 / std::sqrt(si);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) m(i, j) *= s[i] * s[j];
    for (int i = 0; i < n; i++) {
        int k = i;
        if (m(k, k) == 0) return 1;
        q[k] = 1.0 / m(k, k);
        pp[k] = 1.0;
        m(k, k) = 0.0;
        for (int j = 0; j < k; j++) { pp[j] = m(j, k); q[j] = m(j, k) * q[k]; m(j, k) = 0.0; }
        for (int j = k + 1; j < n; j++) { pp[j] = m(k, j); q[j] = -m(k, j) * q[k]; m(k, j) = 0.0; }
        for (int j = 0; j < n; j++)
            for (int kk = j; kk < n; kk++) m(j, kk) += pp[j] * q[kk];
    }
    for (int j = 0; j < n; j++)
        for (int k = 0; k <= j; k++) {
            double v = m(k, j) * s[k] * s[j];
            m(k, j) = v;
            m(j, k) = v;
        }
    return 0;
}

// eigenvalues of a symmetric matrix, ascending (cyclic Jacobi; Minuit2 uses mneigen/QL)
Vec eigenvalues(Mat m)
{
    const int n = m.n;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) off += m(i, j) * m(i, j);
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = m(p, q);
                if (apq == 0) continue;
                double theta = (m(q, q) - m(p, p)) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = m(k, p), akq = m(k, q);
                    m(k, p) = c * akp - s * akq;
                    m(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = m(p, k), aqk = m(q, k);
                    m(p, k) = c * apk - s * aqk;
                    m(q, k) = s * apk + c * aqk;
                }
            }
    }
    Vec ev(n);
    for (int i = 0; i < n; i++) ev[i] = m(i, i);
    std::sort(ev.begin(), ev.end());
    return ev;
}

enum ErrStatus { kPosDef, kMadePosDef, kNotPosDef, kHesseFailed, kInvertFailed, kErrCallLimit };

struct MinError {
    Mat invh;
    double dcovar = 1.0;
    ErrStatus st = kPosDef;
    bool valid() const { return st == kPosDef || st == kMadePosDef; }
};

struct Grad {
    Vec g, g2, gstep;
};

struct State {
    Vec x;
    double fval = 0;
    MinError err;
    Grad grad;
    double edm = 0;
    int nfcn = 0;
    bool has_params = true;
    bool valid() const { return has_params && err.valid(); }
};

struct Fcn {
    const FcnBase *f;
    int ncalls = 0;
    double up = 1.0;
    double operator()(const Vec &x)
    {
        ncalls++;
        return (*f)(x.data());
    }
};

double edm_estimate(const Grad &g, const MinError &e) { return 0.5 * similarity(g.g, e.invh); }

// MnPosDef::operator()(MinimumError)
MinError make_posdef(const MinError &e, const Prec &prec)
{
    MinError out = e;
    Mat &err = out.invh;
    const int n = err.n;
    if (n == 1 && err(0, 0) < prec.eps) { err(0, 0) = 1.; out.st = kMadePosDef; return out; }
    if (n == 1 && err(0, 0) > prec.eps) return e;
    double epspdf = std::max(1.e-6, prec.eps2);
    double dgmin = err(0, 0);
    for (int i = 0; i < n; i++)
        if (err(i, i) < dgmin) dgmin = err(i, i);
    double dg = 0.;
    if (dgmin <= 0) dg = 0.5 + epspdf - dgmin;
    Vec s(n);
    Mat p(n);
    for (int i = 0; i < n; i++) {
        err(i, i) += dg;
        if (err(i, i) < 0.) err(i, i) = 1.;
        s[i] = 1. / std::sqrt(err(i, i));
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) p(i, j) = err(i, j) * s[i] * s[j];
    Vec eval = eigenvalues(p);
    double pmin = eval[0], pmax = eval[n - 1];
    pmax = std::max(std::fabs(pmin), std::fabs(pmax));
    if (pmin > epspdf * pmax) { out.dcovar = e.dcovar; out.st = e.st; return out; }
    double padd = 0.001 * pmax - pmin;
    for (int i = 0; i < n; i++) err(i, i) *= (1. + padd);
    out.st = kMadePosDef;
    return out;
}

// InitialGradientCalculator (unbounded parameters: int == ext)
Grad initial_gradient(const Vec &x, const Vec &werr, double up, const Prec &prec)
{
    const int n = (int)x.size();
    Grad g{Vec(n), Vec(n), Vec(n)};
    for (int i = 0; i < n; i++) {
        double var = x[i];
        double sav = var;
        double sav2 = sav + werr[i];
        double vplu = sav2 - var;
        sav2 = sav - werr[i];
        double vmin = sav2 - var;
        double gsmin = 8. * prec.eps2 * (std::fabs(var) + prec.eps2);
        double dirin = std::max(0.5 * (std::fabs(vplu) + std::fabs(vmin)), gsmin);
        double g2 = 2.0 * up / (dirin * dirin);
        double gstep = std::max(gsmin, 0.1 * dirin);
        double grd = g2 * dirin;
        g.g[i] = grd; g.g2[i] = g2; g.gstep[i] = gstep;
    }
    return g;
}

// Numerical2PGradientCalculator::operator()(par, Gradient)
Grad numerical_gradient(Fcn &fcn, const Vec &xin, double fcnmin, const Grad &prev, const Strategy &stra, const Prec &prec)
{
    const double eps2 = prec.eps2, eps = prec.eps;
    const double dfmin = 8. * eps2 * (std::fabs(fcnmin) + fcn.up);
    const double vrysml = 8. * eps * eps;
    const int n = (int)xin.size();
    Vec x = xin;
    Grad out = prev;
    for (int i = 0; i < n; i++) {
        double xtf = x[i];
        double epspri = eps2 + std::fabs(out.g[i] * eps2);
        double stepb4 = 0.;
        for (int j = 0; j < stra.gradNCyc; j++) {
            double optstp = std::sqrt(dfmin / (std::fabs(out.g2[i]) + epspri));
            double step = std::max(optstp, std::fabs(0.1 * out.gstep[i]));
            double stpmax = 10. * std::fabs(out.gstep[i]);
            if (step > stpmax) step = stpmax;
            double stpmin = std::max(vrysml, 8. * std::fabs(eps2 * x[i]));
            if (step < stpmin) step = stpmin;
            if (std::fabs((step - stepb4) / step) < stra.gradStepTol) break;
            out.gstep[i] = step;
            stepb4 = step;
            x[i] = xtf + step;
            double fs1 = fcn(x);
            x[i] = xtf - step;
            double fs2 = fcn(x);
            x[i] = xtf;
            double grdb4 = out.g[i];
            out.g[i] = 0.5 * (fs1 - fs2) / step;
            out.g2[i] = (fs1 + fs2 - 2. * fcnmin) / step / step;
            if (std::fabs(grdb4 - out.g[i]) / (std::fabs(out.g[i]) + dfmin / step) < stra.gradTol) break;
        }
    }
    return out;
}

struct ParabolaPoint { double x, y; };

// MnLineSearch::operator()
ParabolaPoint line_search(Fcn &fcn, const Vec &x0, double f0, const Vec &step, double gdel, const Prec &prec)
{
    double overal = 1000., undral = -100., toler = 0.05, slamin = 0., slambg = 5., alpha = 2.;
    const int maxiter = 12;
    int niter = 1;
    const int n = (int)x0.size();
    for (int i = 0; i < n; i++) {
        if (step[i] == 0) continue;
        double ratio = std::fabs(x0[i] / step[i]);
        if (slamin == 0) slamin = ratio;
        if (ratio < slamin) slamin = ratio;
    }
    if (std::fabs(slamin) < prec.eps) slamin = prec.eps;
    slamin *= prec.eps2;

    auto eval = [&](double lam) {
        Vec xx(n);
        for (int i = 0; i < n; i++) xx[i] = x0[i] + lam * step[i];
        return fcn(xx);
    };
    double f1 = eval(1.0);
    niter++;
    double fvmin = f0, xvmin = 0.;
    if (f1 < f0) { fvmin = f1; xvmin = 1.; }
    double toler8 = toler, slamax = slambg, flast = f1, slam = 1.;
    bool iterate = false;
    ParabolaPoint p0{0., f0}, p1{slam, flast};
    double f2 = 0.;
    do {
        iterate = false;
        double denom = 2. * (flast - f0 - gdel * slam) / (slam * slam);
        if (denom != 0) slam = -gdel / denom;
        else { denom = -0.1 * gdel; slam = 1.; }
        if (slam < 0.) slam = slamax;
        if (slam > slamax) slam = slamax;
        if (slam < toler8) slam = toler8;
        if (slam < slamin) return {xvmin, fvmin};
        if (std::fabs(slam - 1.) < toler8 && p1.y < p0.y) return {xvmin, fvmin};
        if (std::fabs(slam - 1.) < toler8) slam = 1. + toler8;
        f2 = eval(slam);
        niter++;
        if (f2 < fvmin) { fvmin = f2; xvmin = slam; }
        if (std::fabs(p0.y - fvmin) < std::fabs(fvmin) * prec.eps) {
            iterate = true;
            flast = f2;
            toler8 = toler * slam;
            overal = slam - toler8;
            slamax = overal;
            p1 = {slam, flast};
        }
    } while (iterate && niter < maxiter);
    if (niter >= maxiter) return {xvmin, fvmin};

    ParabolaPoint p2{slam, f2};
    do {
        slamax = std::max(slamax, alpha * std::fabs(xvmin));
        // MnParabolaFactory()(p0,p1,p2): y = a x^2 + b x + c
        double x1 = p0.x, x2 = p1.x, x3 = p2.x;
        double dx12 = x1 - x2, dx13 = x1 - x3, dx23 = x2 - x3;
        double xm = (x1 + x2 + x3) / 3.;
        x1 -= xm; x2 -= xm; x3 -= xm;
        double y1 = p0.y, y2 = p1.y, y3 = p2.y;
        double pa = y1 / (dx12 * dx13) - y2 / (dx12 * dx23) + y3 / (dx13 * dx23);
        double pb = -y1 * (x2 + x3) / (dx12 * dx13) + y2 * (x1 + x3) / (dx12 * dx23) - y3 * (x1 + x2) / (dx13 * dx23);
        // double pc = y1 - pa*x1*x1 - pb*x1;  then shifted back by xm:
        pb -= 2. * xm * pa;
        if (pa < prec.eps2) {
            double slopem = 2. * pa * xvmin + pb;
            if (slopem < 0.) slam = xvmin + slamax;
            else slam = xvmin - slamax;
        } else {
            slam = -pb / (2. * pa);  // MnParabola::Min()
            if (slam > xvmin + slamax) slam = xvmin + slamax;
            if (slam < xvmin - slamax) slam = xvmin - slamax;
        }
        if (slam > 0.) { if (slam > overal) slam = overal; }
        else { if (slam < undral) slam = undral; }

        double f3 = 0.;
        do {
            iterate = false;
            double toler9 = std::max(toler8, std::fabs(toler8 * slam));
            if (std::fabs(p0.x - slam) < toler9 || std::fabs(p1.x - slam) < toler9 || std::fabs(p2.x - slam) < toler9)
                return {xvmin, fvmin};
            f3 = eval(slam);
            if (f3 > p0.y && f3 > p1.y && f3 > p2.y) {
                if (slam > xvmin) overal = std::min(overal, slam - toler8);
                if (slam < xvmin) undral = std::max(undral, slam + toler8);
                slam = 0.5 * (slam + xvmin);
                iterate = true;
                niter++;
            }
        } while (iterate && niter < maxiter);
        if (niter >= maxiter) return {xvmin, fvmin};

        ParabolaPoint p3{slam, f3};
        if (p0.y > p1.y && p0.y > p2.y) p0 = p3;
        else if (p1.y > p0.y && p1.y > p2.y) p1 = p3;
        else p2 = p3;
        if (f3 < fvmin) { fvmin = f3; xvmin = slam; }
        else {
            if (slam > xvmin) overal = std::min(overal, slam - toler8);
            if (slam < xvmin) undral = std::max(undral, slam + toler8);
        }
        niter++;
    } while (niter < maxiter);
    return {xvmin, fvmin};
}

bool has_negative_g2(const Grad &g)
{
    for (double v : g.g2)
        if (v <= 0) return true;
    return false;
}

// NegativeG2LineSearch::operator()
State negative_g2_line_search(Fcn &fcn, const State &st, const Strategy &stra, const Prec &prec)
{
    if (!has_negative_g2(st.grad)) return st;
    const int n = (int)st.x.size();
    Grad dgrad = st.grad;
    Vec pa = st.x;
    double pf = st.fval;
    bool iterate = false;
    unsigned iter = 0;
    do {
        iterate = false;
        for (int i = 0; i < n; i++) {
            if (dgrad.g2[i] <= 0) {
                if (std::fabs(dgrad.g[i]) < prec.eps && std::fabs(dgrad.g2[i]) < prec.eps) continue;
                Vec step(n, 0.0);
                if (dgrad.g[i] < 0) step[i] = dgrad.gstep[i];
                else step[i] = -dgrad.gstep[i];
                double gdel = step[i] * dgrad.g[i];
                ParabolaPoint pp = line_search(fcn, pa, pf, step, gdel, prec);
                for (int k = 0; k < n; k++) pa[k] += pp.x * step[k];
                pf = pp.y;
                dgrad = numerical_gradient(fcn, pa, pf, dgrad, stra, prec);
                iterate = true;
                break;
            }
        }
    } while (iter++ < 2 * (unsigned)n && iterate);
    Mat mat(n);
    for (int i = 0; i < n; i++) mat(i, i) = (std::fabs(dgrad.g2[i]) > prec.eps2 ? 1. / dgrad.g2[i] : 1.);
    State out;
    out.x = pa; out.fval = pf; out.grad = dgrad;
    out.err.invh = mat; out.err.dcovar = 1.; out.err.st = kPosDef;
    out.edm = edm_estimate(dgrad, out.err);
    if (out.edm < 0) out.err.st = kNotPosDef;
    out.nfcn = fcn.ncalls;
    return out;
}

// HessianGradientCalculator::DeltaGradient
Grad hessian_gradient(Fcn &fcn, const Vec &x, double fcnmin, const Grad &gin, const Strategy &stra, const Prec &prec)
{
    Grad out = gin;
    const int n = (int)x.size();
    double dfmin = 4. * prec.eps2 * (std::fabs(fcnmin) + fcn.up);
    for (int i = 0; i < n; i++) {
        double xtf = x[i];
        double dmin = 4. * prec.eps2 * (xtf + prec.eps2);
        double epspri = prec.eps2 + std::fabs(out.g[i] * prec.eps2);
        double optstp = std::sqrt(dfmin / (std::fabs(out.g2[i]) + epspri));
        double d = 0.2 * std::fabs(out.gstep[i]);
        if (d > optstp) d = optstp;
        if (d < dmin) d = dmin;
        double chgold = 10000.;
        double dgmin = 0., grdold = 0., grdnew = 0.;
        for (int j = 0; j < stra.hessGradNCyc; j++) {
            Vec xp = x, xm = x;
            xp[i] = xtf + d; xm[i] = xtf - d;
            double fs1 = fcn(xp);
            double fs2 = fcn(xm);
            grdold = out.g[i];
            grdnew = (fs1 - fs2) / (2. * d);
            dgmin = prec.eps * (std::fabs(fs1) + std::fabs(fs2)) / d;
            if (grdnew == 0) break;
            double change = std::fabs((grdold - grdnew) / grdnew);
            if (change > chgold && j > 1) break;
            chgold = change;
            out.g[i] = grdnew;
            out.gstep[i] = d;
            if (change < 0.05) break;
            if (std::fabs(grdold - grdnew) < dgmin) break;
            if (d < dmin) break;
            d *= 0.2;
        }
    }
    return out;
}

// MnHesse::operator()(MnFcn, MinimumState, trafo, maxcalls)
State hesse(Fcn &fcn, const State &st, const Strategy &stra, unsigned maxcalls, const Prec &prec)
{
    const int n = (int)st.x.size();
    double amin = fcn(st.x);
    double aimsag = std::sqrt(prec.eps2) * (std::fabs(amin) + fcn.up);
    if (maxcalls == 0) maxcalls = 200 + 100 * n + 5 * n * n;
    Mat vhmat(n);
    Vec g2 = st.grad.g2, gst = st.grad.gstep, grd = st.grad.g, dirin = st.grad.gstep, yy(n);
    Vec x = st.x;
    State out = st;
    for (int i = 0; i < n; i++) {
        double xtf = x[i];
        double dmin = 8. * prec.eps2 * (std::fabs(xtf) + prec.eps2);
        double d = std::fabs(gst[i]);
        if (d < dmin) d = dmin;
        for (int icyc = 0; icyc < stra.hessNCyc; icyc++) {
            double sag = 0., fs1 = 0., fs2 = 0.;
            bool got = false;
            for (int multpy = 0; multpy < 5; multpy++) {
                x[i] = xtf + d; fs1 = fcn(x);
                x[i] = xtf - d; fs2 = fcn(x);
                x[i] = xtf;
                sag = 0.5 * (fs1 + fs2 - 2. * amin);
                if (sag != 0) { got = true; break; }
                d *= 10.;
            }
            if (!got) {  // L26: second derivative zero
                out.err.invh = Mat(n);
                out.err.st = kHesseFailed;
                out.nfcn = fcn.ncalls;
                return out;
            }
            double g2bfor = g2[i];
            g2[i] = 2. * sag / (d * d);
            grd[i] = (fs1 - fs2) / (2. * d);
            gst[i] = d;
            dirin[i] = d;
            yy[i] = fs1;
            double dlast = d;
            d = std::sqrt(2. * aimsag / std::fabs(g2[i]));
            if (d < dmin) d = dmin;
            if (std::fabs((d - dlast) / d) < stra.hessStepTol) break;
            if (std::fabs((g2[i] - g2bfor) / g2[i]) < stra.hessG2Tol) break;
            d = std::min(d, 10. * dlast);
            d = std::max(d, 0.1 * dlast);
        }
        vhmat(i, i) = g2[i];
        if ((unsigned)fcn.ncalls > maxcalls) {
            out.err.invh = Mat(n);
            out.err.st = kErrCallLimit;
            out.nfcn = fcn.ncalls;
            return out;
        }
    }
    if (stra.level > 0) {
        Grad gin{grd, g2, gst};
        Grad gr = hessian_gradient(fcn, st.x, st.fval, gin, stra, prec);
        grd = gr.g;
        gst = gr.gstep;
    }
    for (int i = 0; i < n; i++) {
        x[i] += dirin[i];
        for (int j = i + 1; j < n; j++) {
            x[j] += dirin[j];
            double fs1 = fcn(x);
            double elem = (fs1 + amin - yy[i] - yy[j]) / (dirin[i] * dirin[j]);
            vhmat.set_sym(i, j, elem);
            x[j] -= dirin[j];
        }
        x[i] -= dirin[i];
    }
    MinError tmp;
    tmp.invh = vhmat; tmp.dcovar = 1.; tmp.st = kPosDef;
    tmp = make_posdef(tmp, prec);
    vhmat = tmp.invh;
    int ifail = invert_sym(vhmat);
    if (ifail != 0) {
        Mat d(n);
        for (int j = 0; j < n; j++) {
            double t = g2[j];
            if (std::fabs(t) < prec.eps2) t = 1.;
            else t = 1. / t;
            d(j, j) = (t < prec.eps2 ? 1. : t);
        }
        out.err.invh = d;
        out.err.st = kInvertFailed;
        out.nfcn = fcn.ncalls;
        return out;
    }
    out.grad = Grad{grd, g2, gst};
    out.err.invh = vhmat;
    if (tmp.st == kMadePosDef) {
        out.err.st = kMadePosDef;
        out.err.dcovar = 1.;  // MinimumError(mat, MnMadePosDef) sets fDCovar = 1
    } else {
        out.err.st = kPosDef;
        out.err.dcovar = 0.;
    }
    out.edm = edm_estimate(out.grad, out.err);
    out.nfcn = fcn.ncalls;
    return out;
}

// DavidonErrorUpdator::Update
MinError davidon_update(const State &s0, const Vec &p1, const Grad &g1)
{
    const Mat &v0 = s0.err.invh;
    const int n = v0.n;
    Vec dx(n), dg(n);
    for (int i = 0; i < n; i++) { dx[i] = p1[i] - s0.x[i]; dg[i] = g1.g[i] - s0.grad.g[i]; }
    double delgam = dot(dx, dg);
    double gvg = similarity(dg, v0);
    if (delgam == 0) return s0.err;
    if (gvg <= 0) return s0.err;
    Vec vg = matvec(v0, dg);
    Mat vupd(n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) vupd(i, j) = dx[i] * dx[j] / delgam - vg[i] * vg[j] / gvg;
    if (delgam > gvg) {
        Vec w(n);
        for (int i = 0; i < n; i++) w[i] = dx[i] / delgam - vg[i] / gvg;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) vupd(i, j) += gvg * w[i] * w[j];
    }
    double sum_upd = sum_packed_abs(vupd);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) vupd(i, j) += v0(i, j);
    MinError e;
    e.invh = vupd;
    e.dcovar = 0.5 * (s0.err.dcovar + sum_upd / sum_packed_abs(vupd));
    e.st = kPosDef;
    return e;
}

enum MinStatus { kMinValid, kMinAboveMaxEdm, kMinCallLimit };

struct InnerResult { MinStatus st; };

// VariableMetricBuilder::Minimum(fcn, gc, seed, result, maxfcn, edmval)  (inner loop)
MinStatus vm_inner(Fcn &fcn, std::vector<State> &result, unsigned maxfcn, double edmval, const Strategy &stra,
                   const Prec &prec)
{
    State s0 = result.back();
    double edm = s0.edm;
    edm *= (1. + 3. * s0.err.dcovar);
    const int n = (int)s0.x.size();
    Vec step(n);
    do {
        Vec hv = matvec(s0.err.invh, s0.grad.g);
        for (int i = 0; i < n; i++) step[i] = -hv[i];
        double gdel = dot(step, s0.grad.g);
        if (gdel > 0.) {
            s0.err = make_posdef(s0.err, prec);
            hv = matvec(s0.err.invh, s0.grad.g);
            for (int i = 0; i < n; i++) step[i] = -hv[i];
            gdel = dot(step, s0.grad.g);
            if (gdel > 0.) {
                result.push_back(s0);
                return kMinValid;
            }
        }
        ParabolaPoint pp = line_search(fcn, s0.x, s0.fval, step, gdel, prec);
        if (std::fabs(pp.y - s0.fval) <= std::fabs(s0.fval) * prec.eps) {
            // no improvement: store the same state (with updated call count)
            State t = s0;
            t.nfcn = fcn.ncalls;
            result.push_back(t);
            break;
        }
        Vec p(n);
        for (int i = 0; i < n; i++) p[i] = s0.x[i] + pp.x * step[i];
        Grad g = numerical_gradient(fcn, p, pp.y, s0.grad, stra, prec);
        MinError tmpe = s0.err;
        edm = edm_estimate(g, tmpe);
        if (std::isnan(edm)) {
            result.push_back(s0);
            return kMinValid;
        }
        if (edm < 0.) {
            s0.err = make_posdef(s0.err, prec);
            edm = edm_estimate(g, s0.err);
            if (edm < 0.) {
                result.push_back(s0);
                return kMinValid;
            }
        }
        MinError e = davidon_update(s0, p, g);
        State s1;
        s1.x = p; s1.fval = pp.y; s1.err = e; s1.grad = g; s1.edm = edm; s1.nfcn = fcn.ncalls;
        s0 = s1;
        result.push_back(s0);
        edm *= (1. + 3. * e.dcovar);
    } while (edm > edmval && (unsigned)fcn.ncalls < maxfcn);

    if (!result.back().valid()) result.back() = s0;
    if ((unsigned)fcn.ncalls >= maxfcn) return kMinCallLimit;
    if (edm > edmval) {
        if (edm < 10 * edmval) return kMinValid;  // "Edm is close to limit - return current minimum"
        if (edm < std::fabs(prec.eps2 * result.back().fval)) return kMinValid;  // machine accuracy limit
        return kMinAboveMaxEdm;
    }
    return kMinValid;
}

}  // namespace

// ModularFunctionMinimizer::Minimize + MnSeedGenerator + VariableMetricBuilder::Minimum (outer)
MigradResult migrad(const FcnBase &f, const std::vector<double> &start, const std::vector<double> &steps,
                    int strategy_level, unsigned maxfcn, double tolerance)
{
    Prec prec;
    Strategy stra = make_strategy(strategy_level);
    Fcn fcn{&f, 0, 1.0};
    const int n = (int)start.size();
    MigradResult res;
    res.par = start;
    if (maxfcn == 0) maxfcn = 200 + 100 * n + 5 * n * n;
    double edmval = tolerance * fcn.up;
    if (edmval < prec.eps2) edmval = prec.eps2;

    // ---- MnSeedGenerator ----
    Vec x = start;
    double fcnmin = fcn(x);
    Grad g0 = initial_gradient(x, steps, fcn.up, prec);
    Grad dgrad = numerical_gradient(fcn, x, fcnmin, g0, stra, prec);
    State seed;
    seed.x = x; seed.fval = fcnmin; seed.grad = dgrad;
    seed.err.invh = Mat(n);
    for (int i = 0; i < n; i++) seed.err.invh(i, i) = (std::fabs(dgrad.g2[i]) > prec.eps2 ? 1. / dgrad.g2[i] : 1.);
    seed.err.dcovar = 1.; seed.err.st = kPosDef;
    seed.edm = edm_estimate(dgrad, seed.err);
    seed.nfcn = fcn.ncalls;
    if (has_negative_g2(dgrad)) seed = negative_g2_line_search(fcn, seed, stra, prec);
    if (stra.level == 2) seed = hesse(fcn, seed, stra, 0, prec);

    auto finish = [&](const State &last, MinStatus ms) {
        res.par = last.x;
        res.fval = last.fval;
        res.edm = last.edm;
        res.ncalls = fcn.ncalls;
        res.above_max_edm = (ms == kMinAboveMaxEdm);
        res.reached_call_limit = (ms == kMinCallLimit);
        res.covar_status = (int)last.err.st;
        res.valid = last.valid() && ms == kMinValid;
        return res;
    };

    if ((unsigned)fcn.ncalls >= maxfcn) return finish(seed, kMinCallLimit);

    // ---- VariableMetricBuilder::Minimum (outer) ----
    edmval *= 0.002;
    std::vector<State> result;
    if (!seed.valid()) return finish(seed, kMinValid);  // FunctionMinimum(seed, up): invalid state
    if (seed.edm < 0.) return finish(seed, kMinValid);  // "Initial matrix not pos.def." (IsValid still from state)
    result.push_back(seed);
    double edm = seed.edm;
    unsigned maxfcn_eff = maxfcn;
    int ipass = 0;
    bool iterate = false;
    MinStatus ms = kMinValid;
    do {
        iterate = false;
        ms = vm_inner(fcn, result, maxfcn_eff, edmval, stra, prec);
        if (ms == kMinCallLimit) return finish(result.back(), ms);
        if (ipass > 0) {
            if (!(result.back().valid() && ms == kMinValid)) return finish(result.back(), ms);
        }
        edm = result.back().edm;
        if (stra.level == 2 || (stra.level == 1 && result.back().err.dcovar > 0.05)) {
            State st = hesse(fcn, result.back(), stra, maxfcn, prec);
            result.push_back(st);
            if (!st.valid()) break;
            edm = st.edm;
            if (edm > edmval) {
                double machine_limit = std::fabs(prec.eps2 * result.back().fval);
                if (edm >= machine_limit) iterate = true;
            }
        }
        if (ipass == 0) maxfcn_eff = (unsigned)(maxfcn * 1.3);
        ipass++;
    } while (iterate);

    if (edm > 10 * edmval) ms = kMinAboveMaxEdm;
    else if (ms == kMinAboveMaxEdm) ms = kMinValid;  // "Edm has been re-computed after Hesse; now within tolerance"
    return finish(result.back(), ms);
}

}  // namespace ormn

namespace {
struct CFcn : ormn::FcnBase {
    oracle_fcn_t f;
    void *user;
    double operator()(const double *p) const override { return f(p, user); }
};
}  // namespace

extern "C" int oracle_migrad(oracle_fcn_t fcn, void *user, int npar, const double *start, const double *step,
                             int strategy, unsigned maxfcn, double tolerance, double *par_out, double *fmin_out,
                             double *edm_out, int32_t *ncalls_out, int32_t *status_out)
{
    CFcn f;
    f.f = fcn; f.user = user;
    std::vector<double> s(start, start + npar), w(step, step + npar);
    ormn::MigradResult r = ormn::migrad(f, s, w, strategy, maxfcn, tolerance);
    for (int i = 0; i < npar; i++) par_out[i] = r.par[i];
    if (fmin_out) *fmin_out = r.fval;
    if (edm_out) *edm_out = r.edm;
    if (ncalls_out) *ncalls_out = r.ncalls;
    if (status_out) *status_out = (r.above_max_edm ? 1 : 0) | (r.reached_call_limit ? 2 : 0) | (r.covar_status << 4);
    return r.valid ? 1 : 0;
}
