// TEST SHIM (not a product path): compiles the product's scalar Migrad core
// (nps-waveform-analysis_b200/csrc/migrad_core.hpp) for the HOST with a serial chi2 functor, so that the CPU suite can
// check the logic that lane 0 of fit_migrad_kernel executes against the oracle's Migrad restatement bit for bit
// without a GPU.  Built by tests/test_migrad_core_cpu.py with g++ -O2 -ffp-contract=off.
#include <cstdint>
#include "../../nps-waveform-analysis_b200/csrc/migrad_core.hpp"

using namespace npswf::mg;

namespace {
struct SerialFcn {
    const double *spl, *knots;
    int N;
    double y[FIT_NPT], w[FIT_NPT];
    int ncalls = 0;
    double operator()(const double *par)
    {
        ncalls++;
        double chi2 = 0;
        for (int k = 0; k < FIT_NPT; k++) chi2 += chi2_term(k, par, N, spl, knots, y[k], w[k]);
        return chi2;
    }
    MG_DEFAULT_PAIR()
};
}  // namespace

// One Fitwf minimisation (T2:601-773) on the host: trace[110], spline [109][4], N pulses seeded with wftime (bins) /
// wfampl.  Returns the status bits (4 ok, 8 ok on retry, 16 fall-back); par_out[2N+1], *fmin, *ncalls.
extern "C" int mgcore_fitwf(const double *spl, const double *knots, const double *trace, double timeref, int N,
                            const double *wftime, const double *wfampl, double *par_out, double *fmin, int32_t *ncalls)
{
    static thread_local Work<25> W;
    SerialFcn f;
    f.spl = spl; f.knots = knots; f.N = N;
    for (int k = 0; k < FIT_NPT; k++) { f.y[k] = trace[FIT_X0 + k]; f.w[k] = inv_err(f.y[k]); }
    double start[25], werr[25];
    fit_seeds(trace, timeref, wftime, wfampl, N, start);
    const FitOutcome o = fitwf_minimise<25>(f, W, N, start, werr);
    for (int i = 0; i < 2 * N + 1; i++) par_out[i] = W.x[i];
    *fmin = o.fmin;
    *ncalls = o.ncalls;
    return o.status;
}
