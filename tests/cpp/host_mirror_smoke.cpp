// Compiles include/npswf_host.hpp against libnpswf.so and exercises the C++ mirror without a GPU:
// handle creation + derived calibration work on the host, compute calls must fail loudly.
#include <cmath>
#include <cstdio>
#include <vector>
#include "npswf_host.hpp"

int main()
{
    const int B = NPSWF_NBLOCKS, T = NPSWF_NTIME;
    std::vector<double> X(B * T), Y(B * T), tref(B, 35.0);
    std::vector<float> cort(B, 0.5f);
    std::vector<int32_t> preswf(B, 1);
    for (int b = 0; b < B; b++)
        for (int it = 0; it < T; it++) {
            X[b * T + it] = it;
            const double u = it - 30.0;
            Y[b * T + it] = u > 0 ? u * u * std::exp(-u / 3.0) / 4.872 : 0.0;
        }
    for (int b = 0; b < B; b++) {  // timeref = x of the maximum sample (T2:434-438)
        int im = 0;
        for (int it = 0; it < T; it++) if (Y[b * T + it] > Y[b * T + im]) im = it;
        tref[b] = im;
    }
    NpsWfConfig cfg = npswf::Analyzer::defaults();
    NpsWfCalib cal{X.data(), Y.data(), tref.data(), cort.data(), preswf.data()};
    npswf::Analyzer an(cfg, cal);
    std::vector<double> mfy(B * NPSWF_MFWIDTH), mfi(B);
    if (npswf_get_mf_calib(an.raw(), mfy.data(), mfi.data()) != 0 || !(mfi[0] > 0)) return 2;
    if (npswf_device_count() > 0) { std::puts("gpu present: skipping the no-fallback check"); return 0; }
    std::vector<double> sig((size_t)B * T, 0.0), corr(1, 0.0);
    std::vector<int32_t> pres(B, 1);
    int refused = 0;
    try {
        an.analyze(1, sig.data(), pres.data(), corr.data());
    } catch (const std::runtime_error &e) {
        std::printf("expected failure: %s\n", e.what());
        refused++;
    }
    std::vector<double> samp = {3, 110};
    samp.resize(112, 1.0);
    const int64_t offs[2] = {0, 112};
    try {
        an.analyze_packed(1, samp.data(), offs, corr.data());
    } catch (const std::runtime_error &) {
        refused++;
    }
    std::vector<double> ampl(B), et(1), it(1);
    try {
        an.diagnostics(1, sig.data(), ampl.data(), et.data(), it.data());
    } catch (const std::runtime_error &) {
        refused++;
    }
    return refused == 3 ? 0 : 3;  // a CPU fallback would be a bug
}
