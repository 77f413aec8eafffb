// Front kernel: minsignal + matched filter (FindPulsesMF, T2:145-179) + 3x3 cluster threshold
// (PassClusterThreshold, T2:218-278), one streaming pass over the event's traces.
//
// Layout: signal is [E][1080][110] doubles, block-major, and bn = row*30 + col, so one detector
// row (30 traces) is 26 400 contiguous bytes.  A CTA walks the 36 rows of one event through a
// 4-slot shared-memory ring filled by 1-D bulk TMA copies (cp.async.bulk + mbarrier): while row r
// is processed (it needs rows r-1, r, r+1 for the 3x3 halo) row r+2 is already in flight, and
// every sample is read from HBM exactly once.  Two CTAs fit per SM (105.6 KB ring each).
//
// Arithmetic is bit-faithful to the reference: per tap (delta*kern)/mfint with a correctly
// rounded quotient, accumulated in tap order; 3x3 sums in the reference's neighbour order.
#pragma once
#include "common.cuh"

namespace npswf {

constexpr int FRONT_THREADS = 256;
constexpr int FRONT_WARPS = FRONT_THREADS / 32;
constexpr int FRONT_RING = 4;
constexpr size_t FRONT_SMEM = (size_t)FRONT_RING * ROW_BYTES + 64 /*mbarriers*/ + 1088 /*pres bytes*/ + 16;

// flags byte written per (event, block)
constexpr uint8_t FL_PRESENT = 1, FL_OKTOFIT = 2;

// Half-warp task: matched filter of one block.  hl = lane & 15 owns outputs it = 5 + 7*hl + o.
__device__ __forceinline__ void mf_block_halfwarp(const double *__restrict__ s, int hl, bool active,
                                                  const double *__restrict__ kern /*mfyref[b][0..10]*/, double mfint,
                                                  double mfrecip, float *__restrict__ mf_out, double *minsig_out)
{
    // the 17 samples this lane's 7 outputs need: indices 7*hl .. 7*hl+16
    double v[17];
    const int base = 7 * hl;
#pragma unroll
    for (int j = 0; j < 17; j++) v[j] = (active && base + j < T) ? s[base + j] : 0.0;
    // minsignal: init 1e6 (T2:550), min over the trace (T2:884); each lane covers its first 7 samples
    double mn = 1.0e6;
#pragma unroll
    for (int j = 0; j < 7; j++)
        if (base + j < T) mn = fmin(mn, v[j]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    double k[MFW];
#pragma unroll
    for (int j = 0; j < MFW; j++) k[j] = active ? kern[j] : 0.0;
    double acc[7];
    double mfmin = 1.0e6;  // T2:148
#pragma unroll
    for (int o = 0; o < 7; o++) {
        double a = 0.0;
#pragma unroll
        for (int jt = 0; jt < MFW; jt++) {
            const double delta = dsub(v[o + jt], mn);       // raw - minsignal        T2:159
            const double prod = dmul(delta, k[MFW - 1 - jt]);  // * reversed kernel    T2:160
            a = dadd(a, div_by_recip(prod, mfint, mfrecip));    // acc += prod / mfint  T2:161
        }
        acc[o] = a;
        const int it = MFLEFT + base + o;
        if (it < T - MFRIGHT) mfmin = fmin(mfmin, a);  // T2:164
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mfmin = fmin(mfmin, __shfl_xor_sync(0xffffffffu, mfmin, o));
    if (!active) return;
    if (hl == 0 && minsig_out) *minsig_out = mn;
    if (mf_out) {
#pragma unroll
        for (int o = 0; o < 7; o++) {
            const int it = MFLEFT + base + o;
            if (it < T - MFRIGHT) mf_out[it] = (float)dsub(acc[o], mfmin);  // T2:170, TH1F float storage T2:178
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
                    mf_block_halfwarp(rowp[1] + col * T, lane & 15, active, cal.mfyref + (size_t)bn * MFW, cal.mfint[bn],
                                      cal.mfrecip[bn], mf_out ? mf_out + gi * T : nullptr,
                                      minsig_out ? minsig_out + gi : nullptr);
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
