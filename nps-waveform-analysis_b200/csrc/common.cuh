// Shared constants and exact-arithmetic helpers for the NPS waveform kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/npswf.h"

namespace npswf {

constexpr int T = NPSWF_NTIME;        // 110 samples          T2:51
constexpr int NCOL = NPSWF_NCOL;      // 30                   T2:54
constexpr int NLIN = NPSWF_NLIN;      // 36                   T2:55
constexpr int B = NPSWF_NBLOCKS;      // 1080                 T2:56
constexpr int MAXP = NPSWF_MAXWFPULSES;  // 12                T2:59
constexpr int MFW = NPSWF_MFWIDTH;    // 11                   T2:67
constexpr int MFLEFT = 5, MFRIGHT = 5;   //                   T2:65-66
constexpr int MFSTART = 10, MFEND = 100; //                   T2:68-69
constexpr int NFIT = MFEND - MFSTART;    // 90 fit points     T2:681
// Convergence tolerance of every fit kernel: stop when a step lowered chi2 by less than FIT_REL_TOL * chi2, or when the
// Gauss-Newton bound of what the next step could gain is below it.
#ifndef NPSWF_FIT_REL_TOL
#define NPSWF_FIT_REL_TOL 1e-9
#endif
constexpr double FIT_REL_TOL = NPSWF_FIT_REL_TOL;
// continuation lists of N >= 4: bit 30 of an entry = the first attempt was cut short in fit_thread_kernel and has to be
// run (again, from the seeds) by the warp-per-fit kernel; without it only the retry is left
constexpr int FIT_CONT_RESTART = 1 << 30;
constexpr int MAXPAR = 2 * MAXP + 1;     // 25

constexpr int ROW_DOUBLES = NCOL * T;              // one detector row of traces
constexpr int ROW_BYTES = ROW_DOUBLES * 8;         // 26 400 B, contiguous in the block-major layout
constexpr int EVENT_DOUBLES = B * T;               // 118 800

// TSpectrum::SearchHighRes fixed shapes for sigma = 2
constexpr int TS_SHIFT = 14;                       // (int)(7*sigma+0.5)
constexpr int TS_S = T + 2 * TS_SHIFT;             // 138 = size_ext
constexpr int TS_LH = 14;                          // lh_gold
constexpr int TS_POSIT = 6;
constexpr int TS_NP = TS_S + 2 * (TS_LH - 1);      // 164 entries of vector p

// The reference is compiled without FMA contraction (g++ -O2, x86-64 baseline).  Every
// operation that must be bit-identical to it goes through these wrappers, which the compiler
// never fuses, so the rest of the code can be built with the default -fmad=true.
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }

// Correctly rounded p / m given r = RN(1/m) (computed once per block on the host):
// two Markstein refinements, 5 FP64 pipe ops instead of the ~12 + branch of a generic division.
// (q1 is faithful, so q2 = RN(q1 + RN(p - q1*m) * r) is the correctly rounded quotient.)
__device__ __forceinline__ double div_by_recip(double p, double m, double r)
{
    double q = __dmul_rn(p, r);
    double e = __fma_rn(-q, m, p);
    q = __fma_rn(e, r, q);
    e = __fma_rn(-q, m, p);
    return __fma_rn(e, r, q);
}

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Knot form of the natural cubic spline for the thread-per-fit kernel: 16 bytes per knot instead of 32 per
// segment, padded with zeros so that a clamped trial time never needs an index check.
constexpr int KN_LO = 128;
constexpr int KN_LEN = 368;

struct DeviceCounters {
    unsigned long long n_present, n_pass_threshold, n_fit_attempted, n_fit_ok_first, n_fit_ok_retry, n_fallback;
    unsigned long long n_pulses, n_peak_buffer_full, n_fit_iterations, n_fit_evals;
};

// Per-device read-only calibration (device pointers)
struct DevCalib {
    const double *mfyref;    // [B][11]
    const double *mfint;     // [B]
    const double *mfrecip;   // [B] RN(1/mfint)
    const double *mfc;       // [B][11] RN(mfyref * RN(1/mfint)): taps of the fast matched-filter evaluation
    const double *mfepsf;    // [B] 2^-47 * sum|mfyref| * |1/mfint| (1 + 1e-9): error-bound factor of that evaluation
    const double *timeref;   // [B]
    const float *cortime;    // [B]
    const int32_t *preswf;   // [B]
    const int32_t *win_lo;   // [B] first time bin with |it - (timeref + timerefacc)| < coinc_width (T2:267); 1 << 20 if none
    const int32_t *win_span; // [B] last - first bin of that window
    const double *spline;    // [B][109][4] = y, b, c, d
    const double2 *knots;    // [B][KN_LEN] = (y_i, c_i = y''_i / 2) at knot i - KN_LO, zero outside 0..109
    const double *knots_x;   // [B][T] the spline's abscissae (interpX, T2:432) when they are not the sample indices; else nullptr
};

struct KParams {
    double specthres, mfthres, trig_thres, dt, timerefacc;
    int coinc_width;
    int fit_max_iter, fit_retry_max_iter;
    int fit_thread_tries;   // LM tries the thread-per-fit kernel runs before handing a fit to the sub-warp kernel
    int search_fused;       // Gold deconvolution of the peak search: 1 fused pass with checked decisions (default), 0 the
                            // reference's arithmetic only, 2 fused pass and then always the exact repeat, 3 as 1 but the debug
                            // taps show the unchecked fused pass (env NPSWF_SEARCH_FUSED; test aids)
};

}  // namespace npswf
