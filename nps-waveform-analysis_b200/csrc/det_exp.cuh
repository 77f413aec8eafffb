// Device copy of the deterministic exp (see oracle/det_exp.h for the CPU copy and the rationale):
// pure IEEE-754 mul/add/fma + a 128-entry table of correctly rounded 2^(j/128); < 1 ulp.
// Used by the Markov-smoothing step of the TSpectrum search kernel so that the kernel is
// bit-identical to the CPU oracle by construction.
#pragma once
#include "common.cuh"
#include "det_exp_table.h"

namespace npswf {

__device__ const unsigned long long g_det_exp_tab[DET_EXP_N] = DET_EXP_TABLE_BITS;

// `tab` may point to a shared-memory copy of g_det_exp_tab (lane-divergent index).
__device__ __forceinline__ double det_exp(double x, const unsigned long long *tab)
{
    if (!(x == x)) return x;
    if (x > 709.0) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -708.0) return 0.0;
    const double shift = 0x1.8p52;
    double z = __dmul_rn(DET_EXP_INVLN2N, x);
    double kd = __dadd_rn(z, shift);
    kd = __dsub_rn(kd, shift);
    int k = __double2int_rn(kd);
    double r = __fma_rn(kd, -DET_EXP_LN2HIN, x);
    r = __fma_rn(kd, -DET_EXP_LN2LON, r);
    double q = __fma_rn(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);
    q = __fma_rn(r, q, 0x1.5555555555555p-3);
    q = __fma_rn(r, q, 0.5);
    double r2 = __dmul_rn(r, r);
    double tmp = __fma_rn(r2, q, r);
    unsigned long long sb = tab[k & (DET_EXP_N - 1)] + ((unsigned long long)(long long)(k >> 7) << 52);
    double scale = __longlong_as_double((long long)sb);
    return __fma_rn(scale, tmp, scale);
}

}  // namespace npswf
