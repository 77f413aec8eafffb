// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see oracle/README.md).
//
// CPU restatement of the reference's per-event, per-block waveform path
// (/root/reference/TEST_2.C; "T2:N" below = line N of that file).  PARITY UNPINNED — see
// npswf_oracle.h.  Build with -O2 -ffp-contract=off (the reference is compiled by ACLiC with
// plain g++ -O2 on x86-64: no FMA contraction).
#include "npswf_oracle.h"
#include "det_exp.h"
#include "minuit_migrad.hpp"
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr int T = OR_NTIME, B = OR_NBLOCKS, MAXP = OR_MAXWFPULSES, MFW = OR_MFWIDTH;

struct Spline {  // natural cubic spline, GSL gsl_interp_cspline layout
    std::array<double, T> x, y, c;
    std::array<double, T - 1> b, d;
};

// gsl cspline_init: tridiagonal system for c_i = y''/2 with c_0 = c_{n-1} = 0,
// solved by gsl_linalg_solve_symm_tridiag (LDL^T, Engeln-Muellges + Uhlig p.92)
void spline_build(Spline &s, const double *xa, const double *ya)
{
    const int n = T, max_index = n - 1, sys = max_index - 1;
    std::vector<double> off(sys), diag(sys), g(sys), alpha(sys), gamma(sys), z(sys), cc(sys);
    for (int i = 0; i < n; i++) { s.x[i] = xa[i]; s.y[i] = ya[i]; }
    s.c[0] = 0.0; s.c[max_index] = 0.0;
    for (int i = 0; i < sys; i++) {
        const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
        const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
        const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
        off[i] = h_ip1;
        diag[i] = 2.0 * (h_ip1 + h_i);
        g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
    }
    const int N = sys;
    alpha[0] = diag[0];
    gamma[0] = off[0] / alpha[0];
    for (int i = 1; i < N - 1; i++) {
        alpha[i] = diag[i] - off[i - 1] * gamma[i - 1];
        gamma[i] = off[i] / alpha[i];
    }
    if (N > 1) alpha[N - 1] = diag[N - 1] - off[N - 2] * gamma[N - 2];
    z[0] = g[0];
    for (int i = 1; i < N; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
    for (int i = 0; i < N; i++) cc[i] = z[i] / alpha[i];
    s.c[1 + N - 1] = cc[N - 1];
    for (int i = N - 2; i >= 0; i--) s.c[1 + i] = cc[i] - gamma[i] * s.c[1 + i + 1];
    // coeff_calc, hoisted: identical values to computing them at eval time
    for (int i = 0; i < n - 1; i++) {
        const double dx = xa[i + 1] - xa[i], dy = ya[i + 1] - ya[i];
        s.b[i] = (dy / dx) - dx * (s.c[i + 1] + 2.0 * s.c[i]) / 3.0;
        s.d[i] = (s.c[i + 1] - s.c[i]) / (3.0 * dx);
    }
}

inline int spline_index(const Spline &s, double x)
{
    // gsl_interp_bsearch(x_array, x, 0, size-1): x[i] <= x < x[i+1], clamped to [0, size-2]
    int lo = 0, hi = T - 1;
    while (hi > lo + 1) {
        int i = (hi + lo) / 2;
        if (s.x[i] > x) hi = i;
        else lo = i;
    }
    return lo;
}

inline double spline_eval(const Spline &s, double x)
{
    const int i = spline_index(s, x);
    const double delx = x - s.x[i];
    return s.y[i] + delx * (s.b[i] + delx * (s.c[i] + delx * s.d[i]));
}
inline double spline_deriv(const Spline &s, double x)
{
    const int i = spline_index(s, x);
    const double delx = x - s.x[i];
    return s.b[i] + delx * (2.0 * s.c[i] + 3.0 * delx * s.d[i]);
}

}  // namespace

struct OracleHandle {
    OracleConfig cfg;
    std::vector<double> interpX, interpY, timeref, mfyref, mfint;
    std::vector<float> cortime;
    std::vector<int32_t> preswf;
    std::vector<Spline> spline;
    mutable std::mutex spectrum_mutex;  // T2:49 (only taken with ORACLE_FLAG_FAITHFUL_COST)
};

extern "C" OracleHandle *oracle_create(const OracleConfig *cfg, const OracleCalib *cal)
{
    OracleHandle *h = new OracleHandle;
    h->cfg = *cfg;
    h->interpX.assign(cal->interpX, cal->interpX + (size_t)B * T);
    h->interpY.assign(cal->interpY, cal->interpY + (size_t)B * T);
    h->timeref.assign(cal->timeref, cal->timeref + B);
    h->cortime.assign(cal->cortime, cal->cortime + B);
    h->preswf.assign(cal->preswf, cal->preswf + B);
    h->mfyref.assign((size_t)B * MFW, 0.0);
    h->mfint.assign(B, 0.0);
    h->spline.resize(B);
    for (int i = 0; i < B; i++) {
        if (!h->preswf[i]) continue;
        const double *X = &h->interpX[(size_t)i * T], *Y = &h->interpY[(size_t)i * T];
        // T2:440-451: window of 11 samples centred on the sample whose x equals timeref
        h->mfint[i] = 0;
        for (int it = 0; it < T; it++) {
            if (std::fabs(h->timeref[i] - X[it]) < 0.001) {
                for (int jt = 0; jt < MFW; jt++) {
                    int idx = it + jt - OR_MFLEFT;
                    double v = (idx >= 0 && idx < T) ? Y[idx] : 0.0;  // reference reads out of bounds here; guarded
                    h->mfyref[(size_t)i * MFW + jt] = v;
                    h->mfint[i] += v;
                }
            }
        }
        spline_build(h->spline[i], X, Y);
    }
    return h;
}
extern "C" void oracle_destroy(OracleHandle *h) { delete h; }

extern "C" void oracle_get_mf(const OracleHandle *h, double *mfyref, double *mfint)
{
    std::memcpy(mfyref, h->mfyref.data(), sizeof(double) * B * MFW);
    std::memcpy(mfint, h->mfint.data(), sizeof(double) * B);
}
extern "C" void oracle_get_spline(const OracleHandle *h, int bn, double *y, double *b, double *c, double *d)
{
    const Spline &s = h->spline[bn];
    for (int i = 0; i < T - 1; i++) { y[i] = s.y[i]; b[i] = s.b[i]; c[i] = s.c[i]; d[i] = s.d[i]; }
}
extern "C" double oracle_spline_eval(const OracleHandle *h, int bn, double x) { return spline_eval(h->spline[bn], x); }
extern "C" double oracle_det_exp_c(double x) { return oracle_det_exp(x); }

// FindPulsesMF, matched-filter part (T2:145-179)
extern "C" void oracle_matched_filter(const OracleHandle *h, int bn, const double *sig, double minsignal,
                                      double *mfVals, float *mfhist)
{
    for (int i = 0; i < T; i++) mfVals[i] = 0.0;
    double mfmin = 1.0e6;
    const double *mfy = &h->mfyref[(size_t)bn * MFW];
    const double mfi = h->mfint[bn];
    for (int it = OR_MFLEFT; it < T - OR_MFRIGHT; ++it) {
        double acc = 0.0;
        for (int jt = 0; jt < MFW; ++jt) {
            double raw = sig[bn * T + (it + jt - OR_MFRIGHT)];
            double delta = raw - minsignal;
            double kern = mfy[MFW - 1 - jt];
            acc += (delta * kern) / mfi;
        }
        mfVals[it] = acc;
        if (acc < mfmin) mfmin = acc;
    }
    for (int it = OR_MFLEFT; it < T - OR_MFRIGHT; ++it) mfVals[it] -= mfmin;
    if (mfhist)
        for (int i = 0; i < T; i++) mfhist[i] = (float)mfVals[i];  // TH1F::SetBinContent (T2:174-178)
}

// FindPulsesMF (T2:124-216)
extern "C" int oracle_find_pulses_mf(const OracleHandle *h, int bn, const double *sig, const int32_t *pres,
                                     double minsignal, double *wftime, double *wfampl)
{
    if (pres[bn] == 0) return 0;
    double mfVals[T];
    float hist[T];
    oracle_matched_filter(h, bn, sig, minsignal, mfVals, hist);
    double px[MAXP], py[MAXP];
    int npeaks;
    const int libm = (h->cfg.flags & ORACLE_FLAG_LIBM_EXP) ? 1 : 0;
    if (h->cfg.flags & ORACLE_FLAG_FAITHFUL_COST) {
        std::lock_guard<std::mutex> lock(h->spectrum_mutex);  // T2:186
        npeaks = oracle_tspectrum_search(hist, T, 2, h->cfg.specthres, MAXP, px, py, libm);
    } else {
        npeaks = oracle_tspectrum_search(hist, T, 2, h->cfg.specthres, MAXP, px, py, libm);
    }
    int n = 0;
    for (int ip = 0; ip < npeaks && n < MAXP; ++ip) {
        double xpos = px[ip] - 2.0;
        double ypos = py[ip];
        if (xpos > std::max(OR_MFSTART, 0) && xpos < std::min(OR_MFEND, T - 1) && ypos > h->cfg.mfthres) {
            int ti = static_cast<int>(std::round(xpos));
            double rawAmp = std::abs(sig[bn * T + ti] - minsignal);
            wftime[n] = xpos;
            wfampl[n] = rawAmp;
            n++;
        }
    }
    return n;
}

// PassClusterThreshold (T2:218-278)
extern "C" int oracle_pass_cluster_threshold(const OracleHandle *h, int bn, const double *sig, const int32_t *pres)
{
    const double center = h->timeref[bn] + h->cfg.timerefacc;
    const int row = bn / OR_NCOL, col = bn % OR_NCOL;
    double globalMin = 1e6, maxInWindow = -1e6;
    static const int dR[8] = {0, 0, +1, -1, +1, +1, -1, -1};
    static const int dC[8] = {+1, -1, 0, 0, +1, -1, +1, -1};
    for (int it = 0; it < T; ++it) {
        double sum3x3 = sig[bn * T + it];
        for (int k = 0; k < 8; ++k) {
            int nr = row + dR[k], nc = col + dC[k];
            if (nr < 0 || nr >= OR_NLIN || nc < 0 || nc >= OR_NCOL) continue;
            int nb = nr * OR_NCOL + nc;
            if (pres[nb] == 1) sum3x3 += sig[nb * T + it];
        }
        if (sum3x3 < globalMin) globalMin = sum3x3;
        if (std::abs(double(it) - center) < h->cfg.coinc_width) {
            if (sum3x3 > maxInWindow) maxInWindow = sum3x3;
        }
    }
    return ((maxInWindow - globalMin) > h->cfg.trig_thres) ? 1 : 0;
}

namespace {

// The chi2 objective ROOT::Fit::Chi2FCN evaluates for the BinData built at T2:680-688 with the
// model of T2:621-635 (FitUtil::EvaluateChi2: tmp = (y - f(x)) * invError; chi2 += tmp*tmp).
struct WfChi2 : ormn::FcnBase {
    const Spline *sp;
    int npulse;
    double y[OR_MFEND - OR_MFSTART], inv_err[OR_MFEND - OR_MFSTART];
    double model(double x, const double *par) const
    {
        double val = par[0];
        for (int p = 0; p < npulse; ++p) {
            double dt0 = x - par[1 + 2 * p];
            if (dt0 > 1 && dt0 < T - 1) val += par[2 + 2 * p] * spline_eval(*sp, dt0);
        }
        return val;
    }
    double operator()(const double *par) const override
    {
        double chi2 = 0;
        for (int k = 0; k < OR_MFEND - OR_MFSTART; k++) {
            double fval = model((double)(OR_MFSTART + k), par);
            double tmp = (y[k] - fval) * inv_err[k];
            chi2 += tmp * tmp;
        }
        return chi2;
    }
};

// Independent minimiser of the same chi2 (analytic-Jacobian Levenberg-Marquardt), used
// (a) with ORACLE_FLAG_FIT_LM as a fast CPU mode and (b) by tests to confirm Migrad's minima.
struct LmResult { bool ok; double chi2; int iters; };
std::atomic<long long> g_lm_tries{0}, g_lm_accepts{0}, g_lm_fits{0};
struct LmStatsPrinter { ~LmStatsPrinter() { if (getenv("OR_LM_STATS")) fprintf(stderr, "LM stats: fits %lld tries %lld accepts %lld\n", g_lm_fits.load(), g_lm_tries.load(), g_lm_accepts.load()); } } g_lm_stats_printer;
LmResult lm_minimise(const WfChi2 &f, std::vector<double> &par, int max_iter, double lambda0)
{
    const int P = (int)par.size(), NP = OR_MFEND - OR_MFSTART;
    std::vector<double> JtJ((size_t)P * P), Jtr(P), J(P), A((size_t)P * P), dp(P), trial(P);
    double lambda = lambda0;
    double chi2 = f(par.data());
    g_lm_fits++;
    bool converged = false;
    int it = 0;
    for (; it < max_iter; it++) {
        std::fill(JtJ.begin(), JtJ.end(), 0.0);
        std::fill(Jtr.begin(), Jtr.end(), 0.0);
        for (int k = 0; k < NP; k++) {
            const double x = OR_MFSTART + k, w = f.inv_err[k];
            double val = par[0];
            J[0] = w;
            for (int p = 0; p < f.npulse; p++) {
                double dt0 = x - par[1 + 2 * p];
                if (dt0 > 1 && dt0 < T - 1) {
                    double s = spline_eval(*f.sp, dt0), ds = spline_deriv(*f.sp, dt0);
                    val += par[2 + 2 * p] * s;
                    J[1 + 2 * p] = -par[2 + 2 * p] * ds * w;
                    J[2 + 2 * p] = s * w;
                } else {
                    J[1 + 2 * p] = 0; J[2 + 2 * p] = 0;
                }
            }
            const double r = (f.y[k] - val) * w;
            for (int a = 0; a < P; a++) {
                Jtr[a] += J[a] * r;
                for (int b2 = 0; b2 <= a; b2++) JtJ[(size_t)a * P + b2] += J[a] * J[b2];
            }
        }
        bool accepted = false;
        for (int tries = 0; tries < 30 && !accepted; tries++) {
            for (int a = 0; a < P; a++)
                for (int b2 = 0; b2 <= a; b2++) A[(size_t)a * P + b2] = JtJ[(size_t)a * P + b2];
            for (int a = 0; a < P; a++) A[(size_t)a * P + a] += lambda * (JtJ[(size_t)a * P + a] + 1e-12);
            // Cholesky
            bool pd = true;
            for (int a = 0; a < P && pd; a++) {
                for (int b2 = 0; b2 <= a; b2++) {
                    double s = A[(size_t)a * P + b2];
                    for (int k = 0; k < b2; k++) s -= A[(size_t)a * P + k] * A[(size_t)b2 * P + k];
                    if (a == b2) {
                        if (s <= 0) { pd = false; break; }
                        A[(size_t)a * P + a] = std::sqrt(s);
                    } else A[(size_t)a * P + b2] = s / A[(size_t)b2 * P + b2];
                }
            }
            if (!pd) { lambda = std::max(lambda * 10, 1e-6); continue; }
            for (int a = 0; a < P; a++) {
                double s = Jtr[a];
                for (int k = 0; k < a; k++) s -= A[(size_t)a * P + k] * dp[k];
                dp[a] = s / A[(size_t)a * P + a];
            }
            for (int a = P - 1; a >= 0; a--) {
                double s = dp[a];
                for (int k = a + 1; k < P; k++) s -= A[(size_t)k * P + a] * dp[k];
                dp[a] = s / A[(size_t)a * P + a];
            }
            for (int a = 0; a < P; a++) trial[a] = par[a] + dp[a];
            double c2 = f(trial.data());
            g_lm_tries++;
            if (c2 <= chi2) {
                double rel = (chi2 - c2) / (std::fabs(chi2) + 1e-30);
                par = trial;
                chi2 = c2;
                g_lm_accepts++;
                lambda = std::max(lambda * 0.2, 1e-12);
                accepted = true;
                static const double rel_tol = getenv("OR_LM_RELTOL") ? atof(getenv("OR_LM_RELTOL")) : 1e-9;  // same schedule as the CUDA fit kernels
                if (rel < rel_tol) converged = true;
            } else {
                lambda = std::max(lambda * 10, 1e-6);
            }
        }
        if (!accepted) { converged = true; break; }  // cannot improve: at a (local) minimum to machine precision
        if (converged) break;
    }
    return {converged, chi2, it};
}

// Fitwf (T2:601-828) for one block. wftime/wfampl: the block's 12 slots, in-out.
int fitwf_impl(const OracleHandle *h, int bn, const double *sig, int npulse, double corr_time_HMS, double *wftime,
               double *wfampl, double *chi2_out, int32_t *ncalls_out, double *raw_params)
{
    if (ncalls_out) *ncalls_out = 0;
    if (npulse == 0) {  // T2:605-608
        *chi2_out = -100.0;
        return 0;
    }
    Spline local;
    const Spline *sp = &h->spline[bn];
    if (h->cfg.flags & ORACLE_FLAG_FAITHFUL_COST) {  // T2:612-619: Interpolator rebuilt on every call
        spline_build(local, &h->interpX[(size_t)bn * T], &h->interpY[(size_t)bn * T]);
        sp = &local;
    }
    const int N = std::min(MAXP, npulse), P = 2 * N + 1;
    WfChi2 f;
    f.sp = sp;
    f.npulse = N;
    // Err (T2:946-956) and BinData (T2:680-688)
    for (int ib = OR_MFSTART; ib < OR_MFEND; ++ib) {
        double y = sig[bn * T + ib];
        double e = std::sqrt(std::abs(y * 4.096 / 2.)) / 4.096;
        if (e < 1.) e = std::sqrt(std::abs(1.0 * 4.096 / 2.)) / 4.096;
        f.y[ib - OR_MFSTART] = y;
        f.inv_err[ib - OR_MFSTART] = (e != 0.0) ? 1.0 / e : 0.0;
    }
    // seeds (T2:656-677) and FitConfig::CreateParamsSettings steps (T2:704, 746)
    std::vector<double> start(P), step(P);
    double pedestal = 0;
    for (int i = 0; i < 20; ++i) pedestal += sig[bn * T + i];
    pedestal /= 20;
    start[0] = pedestal;
    for (int p = 0; p < N; p++) {
        start[1 + 2 * p] = wftime[p] - h->timeref[bn];
        start[2 + 2 * p] = wfampl[p];
    }
    for (int i = 0; i < P; i++) step[i] = (start[i] == 0) ? 0.3 : 0.3 * std::fabs(start[i]);

    bool ok = false;
    int status = 0;
    std::vector<double> par = start;
    double fmin = 0;
    int ncalls = 0;
    if (h->cfg.flags & ORACLE_FLAG_FIT_LM) {
        static const double lam0 = getenv("OR_LM_LAMBDA0") ? atof(getenv("OR_LM_LAMBDA0")) : 1e-3;
        LmResult r = lm_minimise(f, par, 60, lam0);
        ok = r.ok; fmin = r.chi2; ncalls = r.iters;
        if (ok) status = OR_ST_FIT_OK1;
        else {
            par = start;
            r = lm_minimise(f, par, 300, 1.0);
            ok = r.ok; fmin = r.chi2; ncalls += r.iters;
            if (ok) status = OR_ST_FIT_OK2;
        }
    } else {
        const unsigned maxfcn = 1000 + 100 * P + 5 * P * P;  // FitConfig::CreateMinimizer default
        ormn::MigradResult r = ormn::migrad(f, start, step, 1, maxfcn, 0.01);  // T2:701, 755
        ok = r.valid; par = r.par; fmin = r.fval; ncalls = r.ncalls;
        if (ok) status = OR_ST_FIT_OK1;
        else {
            r = ormn::migrad(f, start, step, 2, maxfcn, 0.01);  // T2:765-768, same seeds
            ok = r.valid; par = r.par; fmin = r.fval; ncalls += r.ncalls;
            if (ok) status = OR_ST_FIT_OK2;
        }
    }
    if (ncalls_out) *ncalls_out = ncalls;
    if (raw_params)
        for (int i = 0; i < P; i++) raw_params[i] = par[i];
    const double dt = h->cfg.dt, timerefacc = h->cfg.timerefacc;
    if (!ok) {  // T2:774-791
        for (int p = 0; p < npulse; p++)
            wftime[p] = (wftime[p] - h->timeref[bn]) * dt + corr_time_HMS - h->cortime[bn] - timerefacc * dt;
        *chi2_out = -100.;
        return OR_ST_FALLBACK;
    }
    for (int p = 0; p < N; ++p) {  // T2:796-817
        double binOff = par[1 + 2 * p];
        wfampl[p] = par[2 + 2 * p];
        wftime[p] = binOff * dt + corr_time_HMS - h->cortime[bn] - timerefacc * dt;
    }
    *chi2_out = fmin / (double)((OR_MFEND - OR_MFSTART) - P);  // T2:824-827
    return status;
}

void analyze_event(const OracleHandle *h, const double *sig, const int32_t *pres, double corr, int32_t *wfnpulse,
                   double *wftime, double *wfampl, double *chi2, double *timewf, double *amplwf, uint8_t *status,
                   int32_t *ncalls)
{
    for (int i = 0; i < B; i++) {
        wfnpulse[i] = 0; chi2[i] = -100.0; timewf[i] = -100; amplwf[i] = -100; status[i] = 0;  // T2:559-561, 576
        if (ncalls) ncalls[i] = 0;
    }
    for (int i = 0; i < B * MAXP; i++) { wftime[i] = -999; wfampl[i] = -999; }  // T2:583-584
    for (int i = 0; i < B; i++) {  // T2:942-1020
        if (!(pres[i] == 1 && h->preswf[i] == 1)) continue;
        double minsignal = 1e6;  // T2:550, 884
        for (int it = 0; it < T; it++) minsignal = std::min(minsignal, sig[i * T + it]);
        double *wt = wftime + (size_t)i * MAXP, *wa = wfampl + (size_t)i * MAXP;
        int st = OR_ST_PRESENT;
        wfnpulse[i] = oracle_find_pulses_mf(h, i, sig, pres, minsignal, wt, wa);
        bool okToFit = oracle_pass_cluster_threshold(h, i, sig, pres) != 0;
        if (!okToFit) { status[i] = (uint8_t)st; continue; }  // T2:980-986
        st |= OR_ST_OKTOFIT;
        st |= fitwf_impl(h, i, sig, wfnpulse[i], corr, wt, wa, &chi2[i], ncalls ? &ncalls[i] : nullptr, nullptr);
        status[i] = (uint8_t)st;
        for (int p = 0; p < wfnpulse[i]; p++) {  // T2:988-1018
            if (p == 0) { timewf[i] = wt[p]; amplwf[i] = wa[p]; }
            if (p > 0 && std::abs(wt[p]) < std::abs(timewf[i])) { timewf[i] = wt[p]; amplwf[i] = wa[p]; }
        }
    }
}

}  // namespace

extern "C" int oracle_fitwf(const OracleHandle *h, int bn, const double *sig, int npulse, double corr_time_HMS,
                            double *wftime, double *wfampl, double *chi2, int32_t *ncalls, double *raw_params)
{
    return fitwf_impl(h, bn, sig, npulse, corr_time_HMS, wftime, wfampl, chi2, ncalls, raw_params);
}

// Waveform unpack of analyze (/root/reference/TEST_2.C:830-889): packed stream [slot, nsamp, samples...] ->
// signal[1080*110] (zero-filled), pres[1080], minsignal[1080] (init 1e6).  Returns the number of records read.
// Deviations, all marked "do not replicate" in SURVEY.md App. B: pres[] is only written for bloc < 1080 (the
// reference writes past its 1080-entry vector for the 24 non-block slots); a sample index >= 110 is dropped instead
// of spilling into the next block; reads never go past n_words.
extern "C" int oracle_unpack_event(const double *samp, int64_t n_words, double *signal, int32_t *pres, double *minsignal)
{
    const int nslots = 1104;                       // T2:355
    const int64_t Ndata = (int64_t)nslots * (T + 1 + 1);   // T2:356
    for (int i = 0; i < B * T; i++) signal[i] = 0.;          // T2:851
    for (int i = 0; i < B; i++) { pres[i] = 0; if (minsignal) minsignal[i] = 1.0e6; }   // T2:548, 550
    if (n_words > Ndata) return 0;                 // T2:830-836: the event is not processed
    // Int_t bloc, nsamp (T2:553): the doubles are truncated on assignment; NaN / out-of-range give INT_MIN on x86-64
    auto to_int_t = [](double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000; };
    int64_t ns = 0;
    int nrec = 0;
    while (ns + 1 < n_words) {                     // T2:855 (a header is two words)
        int bloc = to_int_t(samp[ns]); ns++;
        const int nsamp = to_int_t(samp[ns]); ns++;
        if (bloc == 2000) bloc = 1080;             // T2:862-865
        if (bloc == 2001) bloc = 1081;
        if (bloc < 0 || bloc > nslots - 0.5) break;   // T2:867-872
        const int b = bloc;
        if (b < B) pres[b] = 1;                    // T2:877 (bounded, see above)
        for (int it = 0; it < nsamp; it++) {       // T2:879-887
            if (b < B && it < T && ns < n_words) {
                signal[b * T + it] = samp[ns];
                if (minsignal) minsignal[b] = std::min(minsignal[b], signal[b * T + it]);
            }
            ns++;
        }
        nrec++;
    }
    return nrec;
}

// Per-event diagnostics that land in the WF tree (T2:1026-1056): ampl[i] = pulse maximum (init -100, T2:591/845),
// integtot = sum of all samples, enertot = sum over the cosmic window 30 < it < 109, both in the reference's
// serial order (block-major).
extern "C" void oracle_event_diagnostics(const double *signal, double *ampl, double *enertot, double *integtot)
{
    double et = 0., it_ = 0.;
    const int binmin = 30, binmax = 109;           // T2:1029-1030
    for (int i = 0; i < B; i++) {
        double sigmax = -100.;
        ampl[i] = -100.;
        for (int it = 0; it < T; it++) {
            const double v = signal[i * T + it];
            it_ += v;                               // T2:1036
            if (it > binmin && it < binmax) et += v;   // T2:1038-1042
            if (v > sigmax) { sigmax = v; ampl[i] = v; }   // T2:1051-1056
        }
    }
    *enertot = et;
    *integtot = it_;
}

// hcana pulse variables and the HMS time correction of analyze (/root/reference/TEST_2.C:893-939), one event.
// adcCounter is renumbered in place as the reference does (T2:895-898).  Deviation (SURVEY.md App. B spirit): the
// reference indexes tdcoffset[nblocks] with counters 1080 / 1081 (first pulse from a scintillator) -- out of bounds;
// here such a first pulse takes offset 0.
extern "C" void oracle_hcana_pulses(int32_t NadcCounter, double *adcCounter, const double *adcSampPulseTime,
                                    const double *adcSampPulseTimeRaw, const double *adcSampPulseAmp, const float *tdcoffset,
                                    const float *timemean2, double *corr_time_HMS, double *Sampampl, double *Samptime)
{
    std::vector<int> Npulse(B, 0);
    for (int i = 0; i < B; i++) { Sampampl[i] = -100; Samptime[i] = -100; }   // T2:569-571
    *corr_time_HMS = 0.;                                                       // T2:557
    for (int iNdata = 0; iNdata < NadcCounter; iNdata++) {
        if (adcCounter[iNdata] == 2000) adcCounter[iNdata] = 1080;
        if (adcCounter[iNdata] == 2001) adcCounter[iNdata] = 1081;
        if (iNdata == 0) {                                                     // T2:901-904
            const int c = (int)(adcCounter[iNdata]);
            const double off = (c >= 0 && c < B) ? (double)tdcoffset[c] : 0.0;
            *corr_time_HMS = adcSampPulseTime[iNdata] - (adcSampPulseTimeRaw[iNdata] / 16.) - off;
        }
        if (adcCounter[iNdata] >= 0 && adcCounter[iNdata] < B) {               // T2:917-938
            const int c = (int)(adcCounter[iNdata]);
            Npulse[c] += 1;
            if (Npulse[c] == 1) {
                Sampampl[c] = adcSampPulseAmp[iNdata];
                Samptime[c] = adcSampPulseTime[iNdata];
            }
            if (Npulse[c] > 1) {
                if (std::abs(Samptime[c] - timemean2[c]) > std::abs(adcSampPulseTime[iNdata] - timemean2[c])) {
                    Sampampl[c] = adcSampPulseAmp[iNdata];
                    Samptime[c] = adcSampPulseTime[iNdata];
                }
            }
        }
    }
}

// h1time / h2time of one event as analyze fills them (T2:988-996): the block loop is run again, keeping what the
// reference reads there -- finter[i]->GetParameter(1 + 2p), i.e. the fitted bin offset after a successful fit
// (T2:818-822) and the seed (wftime - timeref, T2:662) after a failed one.  Returns the number of entries.
extern "C" int oracle_event_times(const OracleHandle *h, const double *sig, const int32_t *pres, double corr, double *h1time,
                                  double *h2time)
{
    int n = 0;
    const double dt = h->cfg.dt, timerefacc = h->cfg.timerefacc;
    for (int i = 0; i < B; i++) {
        if (!(pres[i] == 1 && h->preswf[i] == 1)) continue;
        double minsignal = 1e6;
        for (int it = 0; it < T; it++) minsignal = std::min(minsignal, sig[i * T + it]);
        double wt[MAXP], wa[MAXP], par[2 * MAXP + 1], seed_off[MAXP];
        for (int p = 0; p < MAXP; p++) { wt[p] = -999; wa[p] = -999; }
        const int np = oracle_find_pulses_mf(h, i, sig, pres, minsignal, wt, wa);
        if (!oracle_pass_cluster_threshold(h, i, sig, pres)) continue;         // T2:980-986
        for (int p = 0; p < np; p++) seed_off[p] = wt[p] - h->timeref[i];
        double chi2;
        const int st = fitwf_impl(h, i, sig, np, corr, wt, wa, &chi2, nullptr, par);
        for (int p = 0; p < np; p++) {
            if (wa[p] > 20) {
                const double param = (st & OR_ST_FALLBACK) ? seed_off[p] : par[1 + 2 * p];
                h2time[n] = wt[p];
                h1time[n] = param - timerefacc + corr / dt;
                n++;
            }
        }
    }
    return n;
}

// FindPulsesMF for every block of n_events events, event-parallel (no fits): what the exp / FMA sensitivity study of
// the peak search runs over 10^6 spectra (tools/exp_flip_rate.py).
extern "C" int oracle_find_pulses_batch(const OracleHandle *h, int64_t n_events, const double *signal, const int32_t *pres,
                                        int32_t *wfnpulse, double *wftime, double *wfampl, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const int64_t e = next.fetch_add(1);
            if (e >= n_events) break;
            const double *sig = signal + (size_t)e * B * T;
            const int32_t *pr = pres + (size_t)e * B;
            for (int i = 0; i < B; i++) {
                double *wt = wftime + ((size_t)e * B + i) * MAXP, *wa = wfampl + ((size_t)e * B + i) * MAXP;
                for (int p = 0; p < MAXP; p++) { wt[p] = -999; wa[p] = -999; }
                wfnpulse[(size_t)e * B + i] = 0;
                if (!(pr[i] == 1 && h->preswf[i] == 1)) continue;
                double minsignal = 1e6;
                for (int it = 0; it < T; it++) minsignal = std::min(minsignal, sig[i * T + it]);
                wfnpulse[(size_t)e * B + i] = oracle_find_pulses_mf(h, i, sig, pr, minsignal, wt, wa);
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    return 0;
}

extern "C" int oracle_analyze_batch(const OracleHandle *h, int64_t n_events, const double *signal, const int32_t *pres,
                                    const double *corr_time_HMS, int32_t *wfnpulse, double *wftime, double *wfampl,
                                    double *chi2, double *timewf, double *amplwf, uint8_t *status, int32_t *ncalls,
                                    int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t e = next.fetch_add(1);
            if (e >= n_events) break;
            analyze_event(h, signal + (size_t)e * B * T, pres + (size_t)e * B, corr_time_HMS[e],
                          wfnpulse + (size_t)e * B, wftime + (size_t)e * B * MAXP, wfampl + (size_t)e * B * MAXP,
                          chi2 + (size_t)e * B, timewf + (size_t)e * B, amplwf + (size_t)e * B,
                          status + (size_t)e * B, ncalls ? ncalls + (size_t)e * B : nullptr);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    return 0;
}
