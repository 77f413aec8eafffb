// Front kernel: minsignal + matched filter (FindPulsesMF, T2:145-179) + 3x3 cluster threshold
// (PassClusterThreshold, T2:218-278), one streaming pass over the event's traces.
//
// Layout: signal is [E][1080][110] doubles, block-major, and bn = row*30 + col, so one detector
// row (30 traces) is 26 400 contiguous bytes.  A CTA walks the 36 rows of one event through a
// 4-slot shared-memory ring filled by 1-D bulk TMA copies (cp.async.bulk + mbarrier): while row r
// is processed (it needs rows r-1, r, r+1 for the 3x3 halo) row r+2 is already in flight, and
// every sample is read from HBM exactly once.  Two CTAs fit per SM (105.6 KB ring each).
//
// Results are bit-identical to the reference's arithmetic (per tap (delta*kern)/mfint with a correctly
// rounded quotient, accumulated in tap order; 3x3 sums in the reference's neighbour order); the matched
// filter gets there by a fast evaluation with a rigorous error bound and exact re-evaluation of the few
// outputs whose stored float the bound cannot decide (see mf_block_halfwarp).
#pragma once
#include "common.cuh"

namespace npswf {

constexpr int FRONT_THREADS = 256;
constexpr int FRONT_WARPS = FRONT_THREADS / 32;
constexpr int FRONT_RING = 4;
constexpr size_t FRONT_SMEM = (size_t)FRONT_RING * ROW_BYTES + 64 /*mbarriers*/ + 1088 /*pres bytes*/ + 16;

// flags byte written per (event, block)
constexpr uint8_t FL_PRESENT = 1, FL_OKTOFIT = 2;

// The reference's value of one matched-filter output, operation by operation (T2:153-166): per tap
// (delta * kern) / mfint with a correctly rounded quotient, accumulated in tap order.
__device__ __noinline__ double mf_exact_output(const double *s /* the block's 110 samples (shared memory) */, int it, double mn,
                                               const double *__restrict__ kern, double mfint, double mfrecip)
{
    double a = 0.0;
#pragma unroll 1
    for (int jt = 0; jt < MFW; jt++) {
        const double delta = dsub(s[it + jt - MFLEFT], mn);           // raw - minsignal        T2:159
        const double prod = dmul(delta, kern[MFW - 1 - jt]);          // * reversed kernel      T2:160
        a = dadd(a, div_by_recip(prod, mfint, mfrecip));              // acc += prod / mfint    T2:161
    }
    return a;
}

// Half-warp task: matched filter of one block.  hl = lane & 15 owns outputs it = 5 + 7*hl + o.
//
// What is stored is float(acc[it] - min acc) (T2:170, TH1F storage T2:178), and the float rounding hides almost
// every last-bit difference of the double arithmetic.  So the 1 100 exact divisions per block are not done:
//   * every output is first evaluated as one FMA chain F = sum delta * c, c = RN(kern * RN(1/mfint)) (host);
//     |F - E| <= eps for the reference's value E, eps = 2^-47 * (trace max - min) * sum|kern| / |mfint|
//     (standard rounding-error bounds: 12.1 u S for the reference's order, 13.1 u S for the chain, S = sum |terms|);
//   * every output within 2.5 eps of the smallest F could be the reference's minimum: those (normally one) are
//     recomputed exactly and give the exact minimum Emin;
//   * for every other output w = F - Emin is within eta = 2 eps + 2^-49 |w| of the reference's E - Emin; if
//     float(w - eta) == float(w + eta) the stored float is decided, otherwise (probability ~1e-4 per output) the
//     output is recomputed exactly.
// The result is bit-identical to the exact evaluation by construction, at ~1/4 of its FP64 work.
__device__ __forceinline__ void mf_block_halfwarp(const double *__restrict__ s, int lane_in_warp, bool active,
                                                  const double *__restrict__ kern /*mfyref[b][0..10]*/,
                                                  const double *__restrict__ kc /*mfc[b][0..10]*/, double mfint,
                                                  double mfrecip, double epsf, float *__restrict__ mf_out,
                                                  double *minsig_out)
{
    // the 17 samples this lane's 7 outputs need: indices 7*hl .. 7*hl+16
    const int hl = lane_in_warp & 15;
    double v[17];
    const int base = 7 * hl;
#pragma unroll
    for (int j = 0; j < 17; j++) v[j] = (active && base + j < T) ? s[base + j] : 0.0;
    // minsignal: init 1e6 (T2:550), min over the trace (T2:884); each lane covers its first 7 samples
    double mn = 1.0e6, mx = -1.0e300;
#pragma unroll
    for (int j = 0; j < 7; j++)
        if (base + j < T) { mn = fmin(mn, v[j]); mx = fmax(mx, v[j]); }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    double c[MFW];
#pragma unroll
    for (int j = 0; j < MFW; j++) c[j] = active ? kc[j] : 0.0;
#pragma unroll
    for (int j = 0; j < 17; j++) v[j] = dsub(v[j], mn);   // delta = raw - minsignal (T2:159), shared by both evaluations
    double F[7];
    double fmin_ = 1.0e300;
    bool valid[7];
#pragma unroll
    for (int o = 0; o < 7; o++) {
        double a = 0.0;
#pragma unroll
        for (int jt = 0; jt < MFW; jt++) a = __fma_rn(v[o + jt], c[MFW - 1 - jt], a);
        F[o] = a;
        valid[o] = active && (MFLEFT + base + o < T - MFRIGHT);
        if (valid[o]) fmin_ = fmin(fmin_, a);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) fmin_ = fmin(fmin_, __shfl_xor_sync(0xffffffffu, fmin_, o));
    const double eps = dmul(dsub(mx, mn), epsf);
    const double thr = dadd(fmin_, dmul(2.5, eps));
    // exact values of the candidates for the minimum; the reference's mfmin starts at 1e6 (T2:148).  A candidate
    // is evaluated by its half-warp together: lane jt computes the term of tap jt, then the 11 terms are summed
    // in tap order (the reference's order) -- a 90-cycle chain instead of the 700 of a single lane.
    double E[7];
    double emin = 1.0e6;
    bool cand[7];
    const int half = lane_in_warp & 16;
    const double ktap = (active && hl < MFW) ? kern[MFW - 1 - hl] : 0.0;
#pragma unroll
    for (int o = 0; o < 7; o++) {
        cand[o] = valid[o] && F[o] <= thr;
        E[o] = 0.0;
        unsigned mh = (__ballot_sync(0xffffffffu, cand[o]) >> half) & 0xffffu;
        while (__any_sync(0xffffffffu, mh != 0)) {
            const int L = mh ? (__ffs(mh) - 1) : 0;
            const int it = MFLEFT + 7 * L + o;
            double term = 0.0;
            if (mh && hl < MFW) {
                const double delta = dsub(s[it + hl - MFLEFT], mn);                 // T2:159
                term = div_by_recip(dmul(delta, ktap), mfint, mfrecip);             // T2:160-161
            }
            double a = 0.0;
#pragma unroll
            for (int jt = 0; jt < MFW; jt++) a = dadd(a, __shfl_sync(0xffffffffu, term, half + jt));
            if (mh && hl == L) { E[o] = a; emin = fmin(emin, a); }
            mh &= mh - 1;
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
    if (!active) return;
    if (hl == 0 && minsig_out) *minsig_out = mn;
    if (mf_out) {
#pragma unroll
        for (int o = 0; o < 7; o++) {
            if (!valid[o]) continue;
            const int it = MFLEFT + base + o;
            float out;
            if (cand[o]) {
                out = (float)dsub(E[o], emin);                                   // T2:170, TH1F float storage T2:178
            } else {
                const double w = dsub(F[o], emin);
                const double eta = __fma_rn(fabs(w), 0x1p-49, dmul(2.0, eps));
                const float lo = (float)dsub(w, eta), hi = (float)dadd(w, eta);
                out = lo;
                if (lo != hi) out = (float)dsub(mf_exact_output(s, it, mn, kern, mfint, mfrecip), emin);
            }
            mf_out[it] = out;
        }
        if (hl == 0) {
#pragma unroll
            for (int i = 0; i < MFLEFT; i++) mf_out[i] = 0.f;  // T2:147
        }
        if (hl == 15) {
#pragma unroll
            for (int i = T - MFRIGHT; i < T; i++) mf_out[i] = 0.f;
        }
    }
}

// grid: min(n_events, cap) CTAs, grid-stride over events; block: 256 threads; dyn smem FRONT_SMEM
__global__ void __launch_bounds__(FRONT_THREADS, 2)
front_kernel(const double *__restrict__ signal, const int32_t *__restrict__ pres, long long n_events, DevCalib cal,
             KParams kp, float *__restrict__ mf_out, double *__restrict__ minsig_out, uint8_t *__restrict__ flags_out,
             int do_mf, int do_threshold)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)FRONT_RING * ROW_BYTES);
    uint8_t *pres1 = smem_raw + (size_t)FRONT_RING * ROW_BYTES + 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < FRONT_RING; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase_bits = 0;  // parity per slot (bit i)

    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        const double *ev = signal + (size_t)e * EVENT_DOUBLES;
        // prologue: rows 0,1,2 in flight
        if (tid == 0) {
            for (int r = 0; r < 3; r++) {
                mbar_expect_tx(&bars[r % FRONT_RING], ROW_BYTES);
                tma_load_1d(ring + (size_t)(r % FRONT_RING) * ROW_DOUBLES, ev + (size_t)r * ROW_DOUBLES, ROW_BYTES,
                            &bars[r % FRONT_RING]);
            }
        }
        for (int i = tid; i < B; i += FRONT_THREADS)
            pres1[i] = (pres[(size_t)e * B + i] == 1 && cal.preswf[i] == 1) ? 3 : (pres[(size_t)e * B + i] == 1 ? 1 : 0);
        // pres1 bit0: pres==1 (neighbour gate T2:257); bit1: also preswf==1 (block is analysed, T2:944)
        __syncthreads();

        int waited = 0;  // rows whose barrier this thread has already waited on: rows < waited are visible
        for (int r = 0; r < NLIN; r++) {
            // rows r-1, r, r+1 must have landed
            const int need = (r + 1 < NLIN) ? r + 1 : NLIN - 1;
            while (waited <= need) {
                const int slot = waited % FRONT_RING;
                mbar_wait(&bars[slot], (phase_bits >> slot) & 1u);
                phase_bits ^= (1u << slot);
                waited++;
            }
            const double *rowp[3];
            rowp[0] = (r > 0) ? ring + (size_t)((r - 1) % FRONT_RING) * ROW_DOUBLES : nullptr;
            rowp[1] = ring + (size_t)(r % FRONT_RING) * ROW_DOUBLES;
            rowp[2] = (r + 1 < NLIN) ? ring + (size_t)((r + 1) % FRONT_RING) * ROW_DOUBLES : nullptr;

            if (do_mf) {
                // 30 half-warp tasks: warp w takes block pairs {2w', 2w'+1}
                for (int pair = warp; pair < NCOL / 2; pair += FRONT_WARPS) {
                    const int col = 2 * pair + (lane >> 4);
                    const int bn = r * NCOL + col;
                    const bool active = (pres1[bn] & 2) != 0;
                    const size_t gi = (size_t)e * B + bn;
                    mf_block_halfwarp(rowp[1] + col * T, lane, active, cal.mfyref + (size_t)bn * MFW,
                                      cal.mfc + (size_t)bn * MFW, cal.mfint[bn], cal.mfrecip[bn], cal.mfepsf[bn],
                                      mf_out ? mf_out + gi * T : nullptr, minsig_out ? minsig_out + gi : nullptr);
                }
            }
            for (int col = warp; col < NCOL; col += FRONT_WARPS) {
                const int bn = r * NCOL + col;
                const bool present = (pres1[bn] & 2) != 0;
                bool ok = false;
                if (do_threshold && (present || !do_mf)) {
                    // neighbour order of T2:247-248; the pres gate (T2:257) uses bit0 only
                    const int dR[8] = {0, 0, +1, -1, +1, +1, -1, -1};
                    const int dC[8] = {+1, -1, 0, 0, +1, -1, +1, -1};
                    const double *nb[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int nr = r + dR[k], nc = col + dC[k];
                        const bool in = !(nr < 0 || nr >= NLIN || nc < 0 || nc >= NCOL) && (pres1[nr * NCOL + nc] & 1);
                        nb[k] = in ? rowp[1 + dR[k]] + nc * T : nullptr;
                    }
                    const double *self = rowp[1] + col * T;
                    const double center = dadd(cal.timeref[bn], kp.timerefacc);  // T2:232
                    const double cw = (double)kp.coinc_width;
                    double gmin = 1e6, wmax = -1e6;  // T2:238-239
                    for (int it = lane; it < T; it += 32) {
                        double sum = self[it];
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (nb[k]) sum = dadd(sum, nb[k][it]);
                        gmin = fmin(gmin, sum);
                        if (fabs(dsub((double)it, center)) < cw) wmax = fmax(wmax, sum);  // T2:267
                    }
                    gmin = warp_min(gmin);
                    wmax = warp_max(wmax);
                    ok = dsub(wmax, gmin) > kp.trig_thres;  // T2:277
                }
                if (lane == 0 && flags_out)
                    flags_out[(size_t)e * B + bn] = (present ? FL_PRESENT : 0) | (ok ? FL_OKTOFIT : 0);
            }
            __syncthreads();  // every warp is done with row r-1: its slot (r+3)%4 can be refilled
            if (tid == 0 && r + 3 < NLIN) {
                const int nr = r + 3, slot = nr % FRONT_RING;
                mbar_expect_tx(&bars[slot], ROW_BYTES);
                tma_load_1d(ring + (size_t)slot * ROW_DOUBLES, ev + (size_t)nr * ROW_DOUBLES, ROW_BYTES, &bars[slot]);
            }
        }
        __syncthreads();  // pres1 and the ring are reused by the next event
    }
}

}  // namespace npswf
