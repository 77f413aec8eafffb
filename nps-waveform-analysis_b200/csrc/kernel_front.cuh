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

#ifndef NPSWF_FRONT_WARPS   // 8 (128 registers) or 10 (96 registers: three rounds of ten column tasks instead of four of eight)
#define NPSWF_FRONT_WARPS 8
#endif
constexpr int FRONT_WARPS = NPSWF_FRONT_WARPS;
constexpr int FRONT_THREADS = FRONT_WARPS * 32;
static_assert(FRONT_WARPS == 8 || FRONT_WARPS == 10 || FRONT_WARPS == 12, "the per-row task dealing below is written for 8, 10 or 12 warps");
constexpr int FRONT_RING = 4;
constexpr size_t FRONT_SMEM = (size_t)FRONT_RING * ROW_BYTES + 64 /*mbarriers*/ + 1088 /*pres bytes*/ + 16 + 1024 /*zero trace*/ +
                              2 * 3 * 32 * sizeof(int) /*neighbour offset tables of two rows*/;

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

// min / max of doubles as one compare + select (no NaN canonicalisation: the inputs are finite)
__device__ __forceinline__ double dmin2(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double dmax2(double a, double b) { return b > a ? b : a; }

// Half-warp task: matched filter of one block.  hl = lane & 15 owns outputs it = 5 + 7*hl + o.
//
// What is stored is float(acc[it] - min acc) (T2:170, TH1F storage T2:178), and the float rounding hides almost
// every last-bit difference of the double arithmetic.  So the 1 100 exact divisions per block are not done:
//   * every output is first evaluated as one FMA chain F = sum delta * c, c = RN(kern * RN(1/mfint)) (host);
//     |F - E| <= eps for the reference's value E, eps = 2^-47 * (trace max - min) * sum|kern| / |mfint|
//     (standard rounding-error bounds: 12.1 u S for the reference's order, 13.1 u S for the chain, S = sum |terms|);
//   * every output within 2.5 eps of the smallest F could be the reference's minimum: those (normally one) are
//     recomputed exactly and give the exact minimum Emin;
//   * for every other output w = F - Emin is within eta = 2 eps + 2^-49 |w| <= 2.5 eps of the reference's E - Emin; if
//     float(w - eta) == float(w + eta) the stored float is decided, otherwise (probability ~1e-4 per output) the
//     output is recomputed exactly.
// The result is bit-identical to the exact evaluation by construction, at ~1/4 of its FP64 work.
// `s` may be read up to 12 samples past the end of the trace (the next trace of the ring, or the barrier area):
// such values only reach outputs that are not stored.
__device__ __forceinline__ void mf_block_halfwarp(const double *__restrict__ s, int lane_in_warp, bool active,
                                                  const double *__restrict__ kern /*mfyref[b][0..10]*/,
                                                  const double *__restrict__ kc /*mfc[b][0..10]*/, double mfint,
                                                  double mfrecip, double epsf, float *__restrict__ mf_out,
                                                  double *minsig_out)
{
    const unsigned FULL = 0xffffffffu;
    const int hl = lane_in_warp & 15;
    const int half = lane_in_warp & 16;
    const int base = 7 * hl;
    // the 17 samples this lane's 7 outputs need: indices 7*hl .. 7*hl+16
    double v[17];
#pragma unroll
    for (int j = 0; j < 17; j++) v[j] = s[base + j];
    // minsignal: init 1e6 (T2:550), min over the trace (T2:884); each lane covers its first 7 samples, lane 15 the
    // last 5 (indices 110, 111 do not exist: it re-reads 109)
    double mn = 1.0e6;
    double own[7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        own[j] = (j < 5) ? v[j] : s[min(base + j, T - 1)];
        mn = dmin2(mn, own[j]);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mn = dmin2(mn, __shfl_xor_sync(FULL, mn, o));
    // an upper bound of max(trace) - min(trace) for the error bound: the differences are >= 0, so their order is the
    // order of their high words as integers; the bound is the largest high word + 1 with a zero low word
    int dhi = 0;
#pragma unroll
    for (int j = 0; j < 7; j++) dhi = max(dhi, __double2hiint(dsub(own[j], mn)));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dhi = max(dhi, __shfl_xor_sync(FULL, dhi, o));
    const double spread = __hiloint2double(dhi + 1, 0);
    double c[MFW];
#pragma unroll
    for (int j = 0; j < MFW; j++) c[j] = kc[j];
#pragma unroll
    for (int j = 0; j < 17; j++) v[j] = dsub(v[j], mn);   // delta = raw - minsignal (T2:159), shared by both evaluations
    const int nvalid = active ? min(7, max(0, T - MFRIGHT - MFLEFT - base)) : 0;   // outputs o < nvalid are stored
    double F[7];
    double fmin_ = 1.0e300;
#pragma unroll
    for (int o = 0; o < 7; o++) {
        double a = 0.0;
#pragma unroll
        for (int jt = 0; jt < MFW; jt++) a = __fma_rn(v[o + jt], c[MFW - 1 - jt], a);
        F[o] = a;
        if (o < nvalid) fmin_ = dmin2(fmin_, a);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) fmin_ = dmin2(fmin_, __shfl_xor_sync(FULL, fmin_, o));
    const double eps = dmul(spread, epsf);
    const double thr = dadd(fmin_, dmul(2.5, eps));
    // exact values of the candidates for the minimum; the reference's mfmin starts at 1e6 (T2:148).  A candidate
    // is evaluated by its half-warp together: lane jt computes the term of tap jt, then the 11 terms are summed
    // in tap order (the reference's order) -- a 90-cycle chain instead of the 700 of a single lane.
    double emin = 1.0e6;
    unsigned candmask = 0;   // bit o: output o of this lane was evaluated exactly (value in F[o])
    const double ktap = (hl < MFW) ? kern[MFW - 1 - hl] : 0.0;
#pragma unroll
    for (int o = 0; o < 7; o++) {
        const bool cand = o < nvalid && F[o] <= thr;
        unsigned mh = (__ballot_sync(FULL, cand) >> half) & 0xffffu;
        while (__any_sync(FULL, mh != 0)) {
            const int L = mh ? (__ffs(mh) - 1) : 0;
            const int it = MFLEFT + 7 * L + o;
            double term = 0.0;
            if (hl < MFW) {
                const double delta = dsub(s[it + hl - MFLEFT], mn);                 // T2:159
                term = div_by_recip(dmul(delta, ktap), mfint, mfrecip);             // T2:160-161
            }
            double a = 0.0;
#pragma unroll
            for (int jt = 0; jt < MFW; jt++) a = dadd(a, __shfl_sync(FULL, term, half + jt));
            // every lane of the half-warp holds the same sum a: the minimum is tracked by all of them, no reduction
            if (mh) {
                emin = dmin2(emin, a);
                if (hl == L) { F[o] = a; candmask |= 1u << o; }
            }
            mh &= mh - 1;
        }
    }
    if (!active) return;
    if (hl == 0 && minsig_out) *minsig_out = mn;
    if (mf_out) {
        // one eta for the whole block: |w| <= |F| + |Emin| <= 2 spread sum|c| = 2^48 eps, so 2^-49 |w| <= eps / 2
        const double eta = dmul(2.5, eps);
#pragma unroll
        for (int o = 0; o < 7; o++) {
            if (o < nvalid) {
                const int it = MFLEFT + base + o;
                const double w = dsub(F[o], emin);                  // exact value for a candidate: T2:170
                float out = (float)w;                                // TH1F float storage T2:178
                if (!((candmask >> o) & 1u)) {
                    const float lo = (float)dsub(w, eta), hi = (float)dadd(w, eta);
                    if (lo != hi) out = (float)dsub(mf_exact_output(s, it, mn, kern, mfint, mfrecip), emin);
                }
                mf_out[it] = out;
            }
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
    // a trace of zeros stands in for every neighbour that is outside the grid or absent: sum + 0.0 is the identity,
    // so the 3x3 sums need no predicates
    double *zrow = reinterpret_cast<double *>(smem_raw + (size_t)FRONT_RING * ROW_BYTES + 64 + 1088 + 16);
    const int ZOFF = (int)(zrow - ring);
    // nbr[row parity][dr + 1][col + 1]: ring offset (in doubles) of the trace of block (r + dr, col) if it is inside the grid
    // and present in the data (T2:257), of the zero trace otherwise; built for row r + 1 while row r is processed
    int *nbr = reinterpret_cast<int *>(smem_raw + (size_t)FRONT_RING * ROW_BYTES + 64 + 1088 + 16 + 1024);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto build_nbr = [&](int r) {   // threads 0..95
        const int dr = tid / 32 - 1, cc = (tid & 31) - 1, nr = r + dr;
        const bool in = (unsigned)nr < (unsigned)NLIN && (unsigned)cc < (unsigned)NCOL && (pres1[min(max(nr, 0), NLIN - 1) * NCOL + min(max(cc, 0), NCOL - 1)] & 1);
        nbr[(r & 1) * 96 + tid] = in ? (nr % FRONT_RING) * ROW_DOUBLES + cc * T : ZOFF;
    };

    if (tid == 0) {
        for (int i = 0; i < FRONT_RING; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 128; i += FRONT_THREADS) zrow[i] = 0.0;
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
        if (tid < 96) build_nbr(0);
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
            // ring offsets (in doubles) of the rows r-1, r, r+1
            const int ro[3] = {((r + FRONT_RING - 1) % FRONT_RING) * ROW_DOUBLES, (r % FRONT_RING) * ROW_DOUBLES,
                               ((r + 1) % FRONT_RING) * ROW_DOUBLES};
            if (tid < 96 && r + 1 < NLIN) build_nbr(r + 1);   // published by the barrier at the end of this row
            const int *nb_r = nbr + (r & 1) * 96;

            if (do_mf) {
                // 30 half-warp tasks: warp w takes block pairs {2w', 2w'+1}
                for (int pair = warp; pair < NCOL / 2; pair += FRONT_WARPS) {
                    const int col = 2 * pair + (lane >> 4);
                    const int bn = r * NCOL + col;
                    const bool active = (pres1[bn] & 2) != 0;
                    if (!__any_sync(0xffffffffu, active)) continue;
                    const size_t gi = (size_t)e * B + bn;
                    mf_block_halfwarp(ring + ro[1] + col * T, lane, active, cal.mfyref + (size_t)bn * MFW,
                                      cal.mfc + (size_t)bn * MFW, cal.mfint[bn], cal.mfrecip[bn], cal.mfepsf[bn],
                                      mf_out ? mf_out + gi * T : nullptr, minsig_out ? minsig_out + gi : nullptr);
                }
            }
            // 30 warp tasks.  The matched filter gave warps 0-6 two block pairs and warp 7 one, so the columns are dealt
            // 4-4-4-4-3-3-3-5: every warp ends the row with about the same number of instructions
            // (ten warps: warps 0-4 had two block pairs and take two columns each, warps 5-9 had one and take four)
            const int c_first = (FRONT_WARPS == 12) ? ((warp < 3) ? warp : 3 + 3 * (warp - 3))
                              : (FRONT_WARPS == 10) ? ((warp < 5) ? 2 * warp : 10 + 4 * (warp - 5))
                                                    : ((warp < 4) ? 4 * warp : ((warp < 7) ? 16 + 3 * (warp - 4) : 25));
            const int c_count = (FRONT_WARPS == 12) ? ((warp < 3) ? 1 : 3)
                              : (FRONT_WARPS == 10) ? ((warp < 5) ? 2 : 4) : ((warp < 4) ? 4 : ((warp < 7) ? 3 : 5));
            for (int col = c_first; col < c_first + c_count; col++) {
                const int bn = r * NCOL + col;
                const bool present = (pres1[bn] & 2) != 0;
                bool ok = false;
                if (do_threshold && (present || !do_mf)) {
                    // neighbour order of T2:247-248; the pres gate (T2:257) uses bit0 only
                    const int dR[8] = {0, 0, +1, -1, +1, +1, -1, -1};
                    const int dC[8] = {+1, -1, 0, 0, +1, -1, +1, -1};
                    int off[9];
                    off[0] = ro[1] + col * T + lane;
#pragma unroll
                    for (int k = 0; k < 8; k++) off[1 + k] = nb_r[(dR[k] + 1) * 32 + col + dC[k] + 1] + lane;
                    // window |it - center| < coinc_width (T2:267) as the integer interval [lo, lo + span], evaluated
                    // with the reference's own expression on the host (npswf_create)
                    const int wlo = cal.win_lo[bn];
                    const unsigned wspan = (unsigned)cal.win_span[bn];
                    double gmin = 1e6, wmax = -1e6;  // T2:238-239
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int it = lane + 32 * i;
                        double sum = ring[off[0] + 32 * i];
#pragma unroll
                        for (int k = 1; k < 9; k++) sum = dadd(sum, ring[off[k] + 32 * i]);
                        if (i < 3 || it < T) {
                            gmin = dmin2(gmin, sum);
                            if ((unsigned)(it - wlo) <= wspan) wmax = dmax2(wmax, sum);
                        }
                    }
                    // one butterfly for both: the lower half-warp reduces the minimum, the upper one the negated maximum
                    // (negation is exact and turns max into min)
                    const double nmax = -wmax;
                    const bool up = lane >= 16;
                    double keep = up ? nmax : gmin;
                    keep = dmin2(keep, __shfl_xor_sync(0xffffffffu, up ? gmin : nmax, 16));
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) keep = dmin2(keep, __shfl_xor_sync(0xffffffffu, keep, o));
                    gmin = __shfl_sync(0xffffffffu, keep, 0);
                    wmax = -__shfl_sync(0xffffffffu, keep, 16);
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
