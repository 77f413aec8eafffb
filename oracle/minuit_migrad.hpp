// ORACLE — TEST INFRASTRUCTURE ONLY. Interface of the Minuit2-Migrad restatement (minuit_migrad.cpp).
#ifndef ORACLE_MINUIT_MIGRAD_HPP
#define ORACLE_MINUIT_MIGRAD_HPP
#include <vector>
namespace ormn {
struct FcnBase {
    virtual ~FcnBase() {}
    virtual double operator()(const double *par) const = 0;
};
struct MigradResult {
    std::vector<double> par;
    double fval = 0, edm = 0;
    int ncalls = 0;
    bool valid = false;              // FunctionMinimum::IsValid()
    bool above_max_edm = false;      // FunctionMinimum::IsAboveMaxEdm()
    bool reached_call_limit = false; // FunctionMinimum::HasReachedCallLimit()
    int covar_status = 0;            // 0 posdef, 1 made posdef, 2 not posdef, 3 hesse failed, 4 invert failed, 5 call limit
};
// strategy_level 1 or 2 (T2:701, 765); maxfcn as FitConfig::CreateMinimizer computes it; tolerance 0.01
MigradResult migrad(const FcnBase &f, const std::vector<double> &start, const std::vector<double> &steps,
                    int strategy_level, unsigned maxfcn, double tolerance);
}  // namespace ormn
#endif
