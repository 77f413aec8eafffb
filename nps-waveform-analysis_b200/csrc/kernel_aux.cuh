// The callers on either side of the hot path (SURVEY.md 8f "next" rows):
//   unpack_kernel  -- the waveform unpack of analyze (T2:851-889): packed hcana stream
//                     [slot, nsamp, samples...] -> signal[1080][110] + pres[1080]
//   diag_kernel    -- the per-event diagnostics that land in the WF tree: ampl[b] = pulse maximum (T2:1051-1056),
//                     enertot (bins 31..108, T2:1038-1042), integtot (T2:1035-1036)
#pragma once
#include "common.cuh"

namespace npswf {

constexpr int UNPACK_THREADS = 256;
constexpr int NSLOTS = 1104;                       // T2:355
constexpr int NDATA_MAX = NSLOTS * (T + 2);        // T2:356

// One CTA per event.  The stream is a list of records [slot, nsamp, nsamp samples] whose positions depend on the
// nsamp fields before them.  Thread 0 walks the headers (the nominal nsamp = 110 makes this a stride-112 walk;
// the loads of a walk that follows the nominal stride are prefetched by all threads first), then the CTA copies
// the samples, one warp per record; records of a slot that occurs more than once are copied in stream order so
// that the slot ends with its later records, as in the reference's sequential loop.
// Reference quirks kept: slots 2000 / 2001 are renumbered 1080 / 1081 (T2:862-865) and therefore carry no block;
// a slot outside [0, 1104) ends the event (T2:867-872); an event with more than 1104*112 words is skipped entirely
// (T2:830-836).  Not kept: pres[bloc] = 1 for bloc in [1080, 1104) writes past the reference's
// 1080-entry vector; samples beyond it = 109 would overwrite the next block.
// Int_t x = <double>: truncation toward zero; NaN and values beyond the int range give INT_MIN, what the x86-64
// conversion of the reference's build produces (`Int_t bloc, nsamp`, T2:553, 857-859)
__device__ __forceinline__ int to_int_t(double v)
{
    return (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;
}

__global__ void __launch_bounds__(UNPACK_THREADS)
unpack_kernel(const double *__restrict__ samp, const long long *__restrict__ offsets, long long base, long long n_events,
              double *__restrict__ signal, int32_t *__restrict__ pres)
{
    __shared__ int s_pos[NSLOTS + 1];     // word index of the record's first sample
    __shared__ short s_bloc[NSLOTS + 1], s_ns[NSLOTS + 1];
    __shared__ int s_nrec, s_more;
    __shared__ long long s_next;          // where the header walk continues in the next round
    __shared__ int s_cnt[NSLOTS];         // records per slot
    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        const double *S = samp + (offsets[e] - base);   // samp holds the words from offset `base` on
        const long long N = offsets[e + 1] - offsets[e];
        double *sig = signal + (size_t)e * EVENT_DOUBLES;
        int32_t *pr = pres + (size_t)e * B;
        for (int i = threadIdx.x; i < EVENT_DOUBLES; i += UNPACK_THREADS) sig[i] = 0.0;   // std::fill(signal, 0)  T2:851
        for (int i = threadIdx.x; i < B; i += UNPACK_THREADS) pr[i] = 0;
        if (N > NDATA_MAX) { __syncthreads(); continue; }                                   // T2:830-836
        // warm the cache lines of the nominal header positions
        for (long long k = threadIdx.x; k * (T + 2) + 1 < N && k < NSLOTS; k += UNPACK_THREADS)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(S + k * (T + 2)));
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        // Rounds of up to NSLOTS records (one round for any well-formed event; a malformed stream of short or empty
        // records can hold up to N / 2 of them): the walk continues until the words run out, as the reference's does.
        for (;;) {
            if (threadIdx.x == 0) {
                long long ns = s_next;
                int nrec = 0, more = 0;
                while (ns + 1 < N) {                  // (a header needs two words)
                    if (nrec == NSLOTS) { more = 1; break; }
                    int bloc = to_int_t(S[ns]);
                    const int nsamp = to_int_t(S[ns + 1]);
                    ns += 2;
                    if (bloc == 2000) bloc = 1080;
                    if (bloc == 2001) bloc = 1081;
                    if (bloc < 0 || bloc > NSLOTS - 1) { ns = N; break; }                   // T2:867-872: ends the event
                    s_pos[nrec] = (int)ns;
                    s_bloc[nrec] = (short)bloc;
                    s_ns[nrec] = (short)max(0, min(nsamp, 32767));
                    nrec++;
                    ns += max(nsamp, 0);
                }
                s_nrec = nrec;
                s_more = more;
                s_next = ns;
            }
            for (int i = threadIdx.x; i < NSLOTS; i += UNPACK_THREADS) s_cnt[i] = 0;
            __syncthreads();
            const int nrec = s_nrec;
            for (int r = threadIdx.x; r < nrec; r += UNPACK_THREADS) atomicAdd(&s_cnt[s_bloc[r]], 1);
            __syncthreads();
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            // slots that occur once in the round (all of them in a well-formed event): one warp per record, any order
            for (int r = warp; r < nrec; r += UNPACK_THREADS / 32) {
                const int b = s_bloc[r];
                if (b < B && s_cnt[b] == 1) {
                    if (lane == 0) pr[b] = 1;                                               // T2:877
                    for (int it = lane; it < s_ns[r] && it < T && s_pos[r] + it < N; it += 32) sig[b * T + it] = S[s_pos[r] + it];
                }
            }
            // a slot that occurs more than once ends with the samples of its later records: stream order, one warp
            if (warp == 0) {
                for (int r = 0; r < nrec; r++) {
                    const int b = s_bloc[r];
                    if (b < B && s_cnt[b] > 1) {
                        if (lane == 0) pr[b] = 1;
                        for (int it = lane; it < s_ns[r] && it < T && s_pos[r] + it < N; it += 32) sig[b * T + it] = S[s_pos[r] + it];
                        __syncwarp();
                    }
                }
            }
            __syncthreads();     // the round's copies are done before the next round (stream order across rounds) / the next event
            if (!s_more) break;
            __syncthreads();
        }
    }
}

constexpr int DIAG_THREADS = 256;

// One CTA per event; one warp per block, lanes over the time bins.  The two per-event sums are accumulated in a
// fixed order (per block: lane partials -> butterfly; per warp: its blocks in ascending order; per event: the 8
// warp partials in warp order), i.e. deterministic, and exact on the ADC lattice (every partial sum is a multiple of
// 1000/4096 mV far below 2^53 lattice units); for arbitrary doubles they agree with the reference's serial sums
// to ~1e-13 relative.
__global__ void __launch_bounds__(DIAG_THREADS)
diag_kernel(const double *__restrict__ signal, long long n_events, double *__restrict__ ampl, double *__restrict__ enertot,
            double *__restrict__ integtot)
{
    __shared__ double s_e[DIAG_THREADS / 32], s_i[DIAG_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        const double *sig = signal + (size_t)e * EVENT_DOUBLES;
        double w_e = 0.0, w_i = 0.0;
        for (int b = warp; b < B; b += DIAG_THREADS / 32) {
            double mx = -100.0, pe = 0.0, pi = 0.0;    // sigmax init T2:845 / ampl init T2:591
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int it = lane + 32 * i;
                if (it < T) {
                    const double v = sig[b * T + it];
                    mx = v > mx ? v : mx;                                                   // T2:1051-1056
                    pi += v;                                                                // T2:1035-1036
                    if (it > 30 && it < 109) pe += v;                                       // T2:1038-1042
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double m2 = __shfl_xor_sync(0xffffffffu, mx, o);
                mx = m2 > mx ? m2 : mx;
                pe += __shfl_xor_sync(0xffffffffu, pe, o);
                pi += __shfl_xor_sync(0xffffffffu, pi, o);
            }
            if (lane == 0 && ampl) ampl[(size_t)e * B + b] = mx;
            w_e += pe;
            w_i += pi;
        }
        if (lane == 0) { s_e[warp] = w_e; s_i[warp] = w_i; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double te = 0.0, ti = 0.0;
            for (int w = 0; w < DIAG_THREADS / 32; w++) { te += s_e[w]; ti += s_i[w]; }
            if (enertot) enertot[e] = te;
            if (integtot) integtot[e] = ti;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Output packing (T2:959-961, 1022, 1289-1296): the reference returns wfampl / wftime as vectors truncated to
// blockOffset[1080] = sum of wfnpulse -- the pulses of the blocks in block order.  Done on the device so that only
// the pulses cross PCIe instead of the padded [1080][12] arrays (of which ~1.5 of 12 slots are used).
//   flat_totals_kernel : pulses per event
//   flat_scan_kernel   : exclusive scan over the events of the chunk (one CTA; n_events <= a few thousand)
//   flat_scatter_kernel: per event, exclusive scan over the blocks (= blockOffset) and the copy
constexpr int FLAT_THREADS = 256;

__device__ __forceinline__ int clamp_npulse(int n) { return n < 0 ? 0 : (n > MAXP ? MAXP : n); }

__global__ void __launch_bounds__(FLAT_THREADS)
flat_totals_kernel(const int32_t *__restrict__ wfnpulse, long long n_events, int *__restrict__ ev_total)
{
    __shared__ int wsum[FLAT_THREADS / 32];
    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        int v = 0;
        for (int b = threadIdx.x; b < B; b += FLAT_THREADS) v += clamp_npulse(wfnpulse[(size_t)e * B + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < FLAT_THREADS / 32; w++) t += wsum[w];
            ev_total[e] = t;
        }
        __syncthreads();
    }
}

// ev_off[0 .. n_events]: exclusive prefix sums of ev_total (int: a chunk holds at most cap * 12 960 pulses)
__global__ void __launch_bounds__(1024) flat_scan_kernel(const int *__restrict__ ev_total, long long n_events, int *__restrict__ ev_off)
{
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const long long per = (n_events + 1023) / 1024;
    const long long lo = (long long)t * per, hi = (lo + per < n_events) ? lo + per : n_events;
    int s = 0;
    for (long long i = lo; i < hi; i++) s += ev_total[i];
    part[t] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partial sums
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = (t == 0) ? 0 : part[t - 1];
    for (long long i = lo; i < hi; i++) {
        ev_off[i] = run;
        run += ev_total[i];
    }
    if (t == 1023) ev_off[n_events] = part[1023];
}

__global__ void __launch_bounds__(FLAT_THREADS)
flat_scatter_kernel(const int32_t *__restrict__ wfnpulse, const double *__restrict__ wftime, const double *__restrict__ wfampl,
                    long long n_events, const int *__restrict__ ev_off, double *__restrict__ flat_t, double *__restrict__ flat_a,
                    int32_t *__restrict__ block_offset /* [E][1081] or null: the reference's blockOffset, T2:959, 1022 */)
{
    __shared__ int wsum[FLAT_THREADS / 32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        const size_t base = (size_t)ev_off[e];
        for (int b0 = 0; b0 < B; b0 += FLAT_THREADS) {
            const int b = b0 + threadIdx.x;
            const int n = (b < B) ? clamp_npulse(wfnpulse[(size_t)e * B + b]) : 0;
            int inc = n;   // inclusive scan within the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            int before = carry;
            for (int w = 0; w < warp; w++) before += wsum[w];
            const int off = before + inc - n;
            if (b < B) {
                if (block_offset) block_offset[(size_t)e * (B + 1) + b] = off;
                for (int p = 0; p < n; p++) {
                    flat_t[base + off + p] = wftime[((size_t)e * B + b) * MAXP + p];
                    flat_a[base + off + p] = wfampl[((size_t)e * B + b) * MAXP + p];
                }
            }
            __syncthreads();
            if (threadIdx.x == FLAT_THREADS - 1) carry = off + n;
            __syncthreads();
        }
        if (threadIdx.x == 0 && block_offset) block_offset[(size_t)e * (B + 1) + B] = carry;
        __syncthreads();
    }
}

}  // namespace npswf
