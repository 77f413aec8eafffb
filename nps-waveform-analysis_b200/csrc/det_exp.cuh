// Device copy of the deterministic exp (see oracle/det_exp.h for the CPU copy and the rationale):
// pure IEEE-754 mul/add/fma + a 128-entry table of correctly rounded 2^(j/128); < 1 ulp.
// Used by the Markov-smoothing step of the TSpectrum search kernel so that the kernel is
// bit-identical to the CPU oracle by construction.
//
// exp(r) - 1 is evaluated as an even part C(r) plus an odd part S(r).  Negating the argument negates
// kd, r and S exactly and leaves C unchanged, so det_exp_pair() returns exp(q) AND exp(-q) from one
// range reduction and one pair of polynomials (15 FP64 ops instead of 2 x 13) with the very bits the
// single-argument definition gives for q and for -q.
#pragma once
#include "common.cuh"
#include "det_exp_table.h"

namespace npswf {

__device__ const unsigned long long g_det_exp_tab[DET_EXP_N] = DET_EXP_TABLE_BITS;

// 2^(k/128) from the table: the exponent field is adjusted on the high word only
__device__ __forceinline__ double det_exp_scale(int k, const unsigned long long *tab)
{
    const double t = __longlong_as_double((long long)tab[k & (DET_EXP_N - 1)]);
    return __hiloint2double(__double2hiint(t) + ((k >> 7) << 20), __double2loint(t));
}

// `tab` may point to a shared-memory copy of g_det_exp_tab (lane-divergent index).
__device__ __forceinline__ double det_exp(double x, const unsigned long long *tab)
{
    if (!(x == x)) return x;
    if (x > 709.0) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -708.0) return 0.0;
    const double shift = 0x1.8p52;
    const double z = __dmul_rn(DET_EXP_INVLN2N, x);
    double kd = __dadd_rn(z, shift);
    const int k = __double2loint(kd);  // the integer sits in the low mantissa bits (two's complement)
    kd = __dsub_rn(kd, shift);
    double r = __fma_rn(kd, -DET_EXP_LN2HIN, x);
    r = __fma_rn(kd, -DET_EXP_LN2LON, r);
    const double r2 = __dmul_rn(r, r);
    const double c = __fma_rn(r2, 0x1.5555555555555p-5, 0.5);
    const double cm1 = __dmul_rn(r2, c);
    const double s1 = __fma_rn(r2, 0x1.1111111111111p-7, 0x1.5555555555555p-3);
    const double s2 = __dmul_rn(r2, s1);
    const double sn = __fma_rn(r, s2, r);
    const double tmp = __dadd_rn(cm1, sn);
    const double scale = det_exp_scale(k, tab);
    return __fma_rn(scale, tmp, scale);
}

// Constants of the hot FP64 chains live in constant memory: a 64-bit literal operand costs two UMOV issue
// slots every time it is used (the Markov step evaluates ~400 pairs per spectrum), a constant-bank operand
// costs none.  (ptxas folds literals back in however they are written, so they must come from memory.)
struct DetExpConsts {
    double invln2n, nln2hin, nln2lon, c24, c120, c6, c0375;
};
__constant__ DetExpConsts c_det_exp = {DET_EXP_INVLN2N, -DET_EXP_LN2HIN, -DET_EXP_LN2LON, 0x1.5555555555555p-5,
                                       0x1.1111111111111p-7, 0x1.5555555555555p-3, 0.375};

// ep = det_exp(q), em = det_exp(-q), bit for bit, for |q| <= 700 (the caller guards; outside that range the
// result is meaningless but nothing traps).
__device__ __forceinline__ void det_exp_pair(double q, const unsigned long long *tab, double &ep, double &em)
{
    const DetExpConsts &C = c_det_exp;
    const double shift = 0x1.8p52;
    const double z = __dmul_rn(C.invln2n, q);
    double kd = __dadd_rn(z, shift);
    const int k = __double2loint(kd);
    kd = __dsub_rn(kd, shift);
    double r = __fma_rn(kd, C.nln2hin, q);
    r = __fma_rn(kd, C.nln2lon, r);
    const double r2 = __dmul_rn(r, r);
    const double c = __fma_rn(r2, C.c24, 0.5);
    const double cm1 = __dmul_rn(r2, c);
    const double s1 = __fma_rn(r2, C.c120, C.c6);
    const double s2 = __dmul_rn(r2, s1);
    const double sn = __fma_rn(r, s2, r);
    const double tp = __dadd_rn(cm1, sn);
    const double tm = __dsub_rn(cm1, sn);
    const double sp = det_exp_scale(k, tab);
    const double sm = det_exp_scale(-k, tab);
    ep = __fma_rn(sp, tp, sp);
    em = __fma_rn(sm, tm, sm);
}

}  // namespace npswf
