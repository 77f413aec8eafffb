/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md). Not part of the product path.
 *
 * det_exp(): a deterministic double-precision exp() defined purely by IEEE-754
 * operations (mul, add, fma, integer bit ops) and a 128-entry table of correctly
 * rounded 2^(j/128).  Max error < 1 ulp.
 *
 * Why it exists: TSpectrum::SearchHighRes (ROOT hist/spectrum, called at
 * /root/reference/TEST_2.C:188) evaluates TMath::Exp inside its Markov smoothing
 * chain.  glibc's exp and CUDA's exp are both < 1 ulp but differ in the last bit
 * on some inputs, so "bit-exact kernel vs oracle" needs one shared definition.
 * The oracle can be switched to libm exp (ORACLE_FLAG_LIBM_EXP) to show that the
 * discrete outputs (peak count / positions) do not depend on that last bit.
 *
 * The CUDA product carries its own copy of this algorithm
 * (nps-waveform-analysis_b200/csrc/det_exp.cuh); tests check both bitwise.
 */
#ifndef ORACLE_DET_EXP_H
#define ORACLE_DET_EXP_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "det_exp_table.h"

static const uint64_t oracle_det_exp_tab[DET_EXP_N] = DET_EXP_TABLE_BITS;

static inline double oracle_det_exp(double x)
{
    if (!(x == x)) return x;             /* NaN */
    if (x > 709.0) return INFINITY;
    if (x < -708.0) return 0.0;          /* flush: subnormal results are not needed on this path */
    const double shift = 0x1.8p52;
    double z = DET_EXP_INVLN2N * x;
    double kd = z + shift;               /* round-to-nearest-even integer in the low mantissa bits */
    kd = kd - shift;
    int k = (int)kd;                     /* exact: kd is integer-valued */
    double r = fma(kd, -DET_EXP_LN2HIN, x);
    r = fma(kd, -DET_EXP_LN2LON, r);
    /* exp(r) - 1 = C(r) + S(r), |r| <= ln2/256, split into the even part C = r^2 (1/2 + r^2/24) and the odd
       part S = r + r (r^2 (1/6 + r^2/120)).  exp(-x) has kd -> -kd, r -> -r exactly, so C is shared and S
       only changes sign: a GPU lane gets exp(q) and exp(-q) from one reduction and one set of polynomials,
       and this single-argument definition returns the same bits for either of them. */
    double r2 = r * r;
    double c = fma(r2, 0x1.5555555555555p-5, 0.5);
    double cm1 = r2 * c;
    double s1 = fma(r2, 0x1.1111111111111p-7, 0x1.5555555555555p-3);
    double s2 = r2 * s1;
    double sn = fma(r, s2, r);
    double tmp = cm1 + sn;
    uint64_t sb = oracle_det_exp_tab[k & (DET_EXP_N - 1)] + ((uint64_t)(int64_t)(k >> 7) << 52);
    double scale;
    memcpy(&scale, &sb, sizeof scale);
    return fma(scale, tmp, scale);
}
#endif
