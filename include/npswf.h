/* npswf.h — C ABI of the B200-native NPS waveform path (libnpswf.so).
 *
 * Drop-in boundary for the per-event, per-block hot path of the reference macro
 * (/root/reference/TEST_2.C, the npsWF.C lineage; "T2:N" = line N of that file).  The
 * reference has no FFI: its stages are free functions / lambdas in one ROOT macro, called once
 * per event by RDataFrame (T2:1305).  This library replaces the `Define("tuple", analyze, ...)`
 * step for a BATCH of events; each entry point cites the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; caller owns every host buffer; return 0 = ok,
 * < 0 = error (message via npswf_last_error); nothing throws across the boundary; a handle is
 * used by one host thread at a time; host-buffer calls are synchronous w.r.t. their outputs.
 * Calls on a handle share its device scratch: the library orders them itself (a device-path call
 * leaves an event behind that the next call -- on any stream, device or host path -- waits for).
 * There is NO CPU fallback: every compute entry point fails with NPSWF_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef NPSWF_H
#define NPSWF_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* shape of the detector / waveform, fixed at compile time as in T2:51-73 */
#define NPSWF_NTIME 110       /* ntime        T2:51 */
#define NPSWF_NCOL 30         /* ncol         T2:54 */
#define NPSWF_NLIN 36         /* nlin         T2:55 */
#define NPSWF_NBLOCKS 1080    /* nblocks      T2:56 */
#define NPSWF_MAXWFPULSES 12  /* maxwfpulses  T2:59 */
#define NPSWF_MFWIDTH 11      /* mfwidth      T2:67 */

enum {
    NPSWF_OK = 0,
    NPSWF_ERR_ARG = -1,
    NPSWF_ERR_CUDA = -2,   /* no usable device / CUDA runtime error */
    NPSWF_ERR_NOMEM = -3,
    NPSWF_ERR_CALIB = -4
};

/* status bits per (event, block) */
enum {
    NPSWF_ST_PRESENT = 1,   /* pres==1 && preswf==1                     T2:944 */
    NPSWF_ST_OKTOFIT = 2,   /* PassClusterThreshold returned true       T2:962 */
    NPSWF_ST_FIT_OK1 = 4,   /* first fit attempt converged              T2:755 */
    NPSWF_ST_FIT_OK2 = 8,   /* retry (tougher configuration) converged  T2:761-768 */
    NPSWF_ST_FALLBACK = 16  /* both failed: TSpectrum values, chi2=-100 T2:774-791 */
};

/* The tunables of T2:64-73, 81, 354 (compile-time constants in the reference). */
typedef struct NpsWfConfig {
    double specthres;    /* 0.02  T2:70  TSpectrum::Search threshold                     */
    double mfthres;      /* 1.5   T2:71  matched-filter amplitude cut (mV)               */
    double trig_thres;   /* 10    T2:72  3x3 sum threshold (mV)                          */
    int32_t coinc_width; /* 20    T2:73  half-width of the coincidence window (bins)     */
    double dt;           /* 4     T2:354 ns per sample                                    */
    double timerefacc;   /* 0     T2:81, 524 (calodist-9.5)/(3e8*1e-9*4), in bins         */
    int32_t n_devices;   /* number of GPUs to shard events over (contiguous ranges); 0 = 1 */
    const int32_t *devices; /* CUDA ordinals [n_devices]; NULL = 0..n_devices-1           */
    int32_t chunk_events;   /* events per internal device chunk; 0 = default (512)        */
    int32_t fit_max_iter;   /* LM iterations of the first attempt; 0 = default            */
    int32_t fit_retry_max_iter; /* LM iterations of the retry; 0 = default                */
    int32_t fit_mode;       /* NPSWF_FIT_FAST (0, default), NPSWF_FIT_MIGRAD or NPSWF_FIT_VM: minimiser of Fitwf, see below */
} NpsWfConfig;

/* Minimiser of Fitwf (T2:693-773).
 * NPSWF_FIT_MIGRAD: the reference's own -- Minuit2 Migrad on numerical gradients, strategy 1, retry with strategy 2
 *   from the same seeds, EDM goal 2e-5, call limit 1000+100P+5P^2 -- re-implemented for the device
 *   (csrc/migrad_core.hpp, fit_migrad_kernel) in the reference's FMA-free arithmetic and summation order: fitted
 *   values, chi2 and the ok / retry / fall-back verdict follow Migrad's path.
 * NPSWF_FIT_VM: Migrad's recursion (seed from the second derivatives, MnLineSearch, Davidon update, EDM stop) with
 *   ANALYTIC gradients instead of Minuit's numerical ones and no MnHesse (fit_vm_thread_kernel, 1-6 pulses; fits that
 *   leave the common path, 7+ pulses and general knots go through the exact kernels): it follows Migrad into the same
 *   minimum and stops where Migrad stops on 99.99 % of ordinary fits (99.5 % with up to 12 pulses near threshold), at
 *   a quarter of MIGRAD's cost.
 * NPSWF_FIT_FAST: Levenberg-Marquardt on analytic spline derivatives (fit_thread_kernel & co.), ~8x cheaper than MIGRAD; it
 *   converges the same chi2 tighter than Migrad's EDM goal, and where the chi2 has several local minima it may end
 *   in another one than Migrad does. */
enum { NPSWF_FIT_FAST = 0, NPSWF_FIT_MIGRAD = 1, NPSWF_FIT_VM = 2 };

/* The calibration globals of T2:74-85 as loaded at T2:360-469.  mfyref / mfint are derived
 * inside npswf_create exactly as T2:440-451 does. */
typedef struct NpsWfCalib {
    const double *interpX;  /* [NBLOCKS][NTIME]  T2:84, 432 */
    const double *interpY;  /* [NBLOCKS][NTIME]             */
    const double *timeref;  /* [NBLOCKS]         T2:77, 437 */
    const float *cortime;   /* [NBLOCKS]         T2:78, 463 (Float_t) */
    const int32_t *preswf;  /* [NBLOCKS]         T2:79, 452 */
} NpsWfCalib;

/* Replaces the atomics nFitFailures / nFitSucceeds (T2:61-62, 1436); summed over all calls. */
typedef struct NpsWfCounters {
    int64_t n_events, n_block_waveforms, n_present, n_pass_threshold;
    int64_t n_fit_attempted;   /* "fitted block-waveforms": present, passed threshold, npulse>0 */
    int64_t n_fit_ok_first, n_fit_ok_retry, n_fallback;
    int64_t n_pulses, n_peak_buffer_full, n_fit_iterations;
    int64_t n_fit_evals;       /* chi2 + normal-equation passes over the 90 points (accepted + rejected + first) */
} NpsWfCounters;

typedef struct npswf_handle npswf_handle;

void npswf_default_config(NpsWfConfig *cfg);
int npswf_create(const NpsWfConfig *cfg, const NpsWfCalib *calib, npswf_handle **out);
void npswf_destroy(npswf_handle *h);
const char *npswf_last_error(const npswf_handle *h); /* h may be NULL: error of the last failed create */
int npswf_get_counters(npswf_handle *h, NpsWfCounters *out);
int npswf_reset_counters(npswf_handle *h);
int npswf_device_count(void);

/* Pinned host allocations for the host-buffer entry points (pageable buffers work, slower). */
void *npswf_host_alloc(size_t bytes);
void npswf_host_free(void *p);

/* ---- analyze(event) for a batch: replaces T2:540-1300 / the Define at T2:1305 for the hot
 * path (block loop T2:942-1022).  Inputs are the unpacked per-event arrays the reference
 * builds at T2:851-889:
 *   signal [E][NBLOCKS][NTIME] mV, block-major (idx bn*ntime+it, T2:549);  pres [E][NBLOCKS];
 *   corr_time_HMS [E] (T2:903).  minsignal is recomputed (min over the trace, init 1e6, T2:884).
 * Outputs (any may be NULL):
 *   wfnpulse [E][NBLOCKS]; wftime/wfampl [E][NBLOCKS][12] padded with -999 (the reference's
 *   scratch init, T2:583-584; flatten with npswf_flatten_event to get T2:1294-1295);
 *   chi2/timewf/amplwf [E][NBLOCKS] (-100 sentinels, T2:559-561); status [E][NBLOCKS].
 * Events are sharded over the handle's devices in contiguous ranges; no collective. */
int npswf_analyze_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                        const double *corr_time_HMS, int32_t *wfnpulse, double *wftime, double *wfampl,
                        double *chi2, double *timewf, double *amplwf, uint8_t *status);

/* Same analysis, with wfampl / wftime in the reference's own output packing (T2:959-961, 1289-1296): the vectors of
 * event e -- the pulses of its blocks in block order, sum of wfnpulse of them -- are
 *   wftime_pool[pulse_offset[e] .. + pulse_count[e]),  wfampl_pool[same range].
 * Packed on the device, so only the pulses cross PCIe (a few % of the padded [1080][12] arrays) and the caller has no
 * flatten pass.  pool_capacity = doubles available in each pool; with n devices, device d's event range uses the
 * d-th share of the pools (offsets are absolute; ranges of different devices need not touch).  A pool that is too
 * small -> NPSWF_ERR_NOMEM (npswf_last_error says how much was missing); E * 12960 always suffices.
 * wfnpulse, chi2, timewf, amplwf, status: as npswf_analyze_batch (may be NULL); n_pulses (may be NULL): total. */
int npswf_analyze_batch_flat(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                             const double *corr_time_HMS, int32_t *wfnpulse, int64_t *pulse_offset, int32_t *pulse_count,
                             double *wftime_pool, double *wfampl_pool, int64_t pool_capacity, double *chi2, double *timewf,
                             double *amplwf, uint8_t *status, int64_t *n_pulses);

/* Same, with inputs as int16 ADC counts (signal = counts * lsb_mV; exact on the 12-bit lattice
 * 1000/4096 mV of T2:357).  Quarter of the PCIe bytes. */
int npswf_analyze_batch_i16(npswf_handle *h, int64_t n_events, const int16_t *counts, double lsb_mV,
                            const int32_t *pres, const double *corr_time_HMS, int32_t *wfnpulse, double *wftime,
                            double *wfampl, double *chi2, double *timewf, double *amplwf, uint8_t *status);

/* int16 ADC counts in, the reference's own output packing out (npswf_analyze_batch_i16 + npswf_analyze_batch_flat):
 * the leanest host transport -- about 250 B per block-waveform cross PCIe instead of 1 100 -- and what a reader that
 * already holds the raw fADC counts should call. */
int npswf_analyze_batch_flat_i16(npswf_handle *h, int64_t n_events, const int16_t *counts, double lsb_mV, const int32_t *pres,
                                 const double *corr_time_HMS, int32_t *wfnpulse, int64_t *pulse_offset, int32_t *pulse_count,
                                 double *wftime_pool, double *wfampl_pool, int64_t pool_capacity, double *chi2, double *timewf,
                                 double *amplwf, uint8_t *status, int64_t *n_pulses);

/* Transport of npswf_analyze_batch's binary64 traces.  The reference's samples are ADC counts * ADCtomV (1000/4096 mV,
 * T2:357), so a chunk is sent to the device as int16 counts whenever that reproduces every double of the chunk bit for
 * bit (checked sample by sample by `n_threads` host threads while the device works on the previous chunk); any other
 * chunk goes over as the caller's doubles.  The kernels see identical traces either way, so results do not depend on
 * the mode.  mode: 0 = always the doubles; 1 = automatic (default: packed while the host threads keep ahead of what
 * the raw upload would do); 2 = packed whenever lossless.  n_threads = 0 and lsb_mV = 0 keep the current values
 * (default: this process's share of the host cores, at most 16; 1000/4096).  Environment: NPSWF_HOST_PACK,
 * NPSWF_HOST_PACK_THREADS set the defaults at npswf_create. */
int npswf_set_host_packing(npswf_handle *h, int mode, int n_threads, double lsb_mV);
/* Host-only tap of the packer (no device needed): counts_out[i] = round(x[i] / lsb_mV) with n_threads host threads;
 * returns 1 if every x[i] == double(counts_out[i]) * lsb_mV with |counts| <= 32767 (the chunk would travel as counts),
 * 0 if not (it would travel as doubles), < 0 on bad arguments. */
int npswf_debug_pack_counts(const double *x, int64_t n, double lsb_mV, int32_t n_threads, int16_t *counts_out);
/* Since npswf_create: chunks with a part sent as int16 counts / chunks whose packing was refused (sent as doubles) /
 * mean packing rate (GB/s of doubles read) / bytes of the caller's doubles that travelled as counts (a quarter of it
 * crossed PCIe).  Any pointer may be NULL. */
int npswf_host_packing_stats(const npswf_handle *h, int64_t *packed_chunks, int64_t *raw_chunks, double *pack_gb_per_s,
                             int64_t *packed_input_bytes);

/* ---- host-side callers of the hot path (plain C++, no device): SURVEY.md 8f-3 / 8f-4 ----
 * npswf_hcana_pulses: T2:893-939 for one event -- corr_time_HMS from the first hcana pulse (T2:903; 0 if there is
 * none, T2:557) and, per block, amplitude and time of the hcana pulse closest to the expected time timemean2[b]
 * (T2:917-938; -100 where a block has none, T2:569-571).  adcCounter is not modified (the reference renumbers the
 * scintillator channels 2000/2001 in place).  tdcoffset, timemean2: [NBLOCKS] Float_t as loaded at T2:368-375 / 526-529.
 * Sampampl / Samptime [NBLOCKS] may be NULL. */
int npswf_hcana_pulses(int32_t n_adc, const double *adcCounter, const double *adcSampPulseTime, const double *adcSampPulseTimeRaw,
                       const double *adcSampPulseAmp, const float *tdcoffset, const float *timemean2, double *corr_time_HMS,
                       double *Sampampl, double *Samptime);
/* npswf_event_times: the h1time / h2time vectors of one event (T2:988-996) from its analysis outputs (padded
 * [NBLOCKS][12] wftime / wfampl, wfnpulse, status): one entry per pulse with wfampl > 20 of every block that passed the
 * cluster threshold, block order.  Returns the number of entries (<= sum of wfnpulse); h1time / h2time hold at least
 * that many doubles (either may be NULL).  cortime: the calibration's [NBLOCKS] Float_t; dt: ns per sample. */
int64_t npswf_event_times(const int32_t *wfnpulse, const double *wftime_padded, const double *wfampl_padded, const uint8_t *status,
                          const float *cortime, double dt, double *h1time, double *h2time);

/* Diagnostics of the NPSWF_FIT_VM mode: how many fits left fit_vm_thread_kernel for the exact Migrad kernels, by reason
 * (0 evaluation limit or trace not exact in binary32, 1 second derivative <= 0 at the seeds, 2 EDM negative / not a
 * number, 3 above the EDM limit, 4 not a descent direction), since the last reset (device 0). */
int npswf_debug_vm_reasons(npswf_handle *h, uint64_t out[8], int reset);

/* Diagnostics of the peak search (TSpectrum::SearchHighRes, T2:187-188): out[0] = spectra whose Gold deconvolution was
 * evaluated with fused multiply-adds and every decision (iteration gates, local maxima, thresholds, integer parts of
 * the centroids) checked against the error margin of that evaluation; out[1] = of those, the spectra with a decision
 * inside the margin, which were repeated with the reference's arithmetic (rounded product, rounded sum).  Summed over
 * the handle's devices since the last reset. */
int npswf_debug_search_fused(npswf_handle *h, uint64_t out[2], int reset);

/* Measured rate of the raw binary64 uploads out of the caller's pinned buffers (CUDA events around the raw part of a
 * chunk, running mean over the devices; 48 until something was measured), and the number of host cores the packer
 * threads / staging buffers are confined to (the cores of each GPU's NUMA node, from sysfs; 0 = topology not visible
 * or NPSWF_NUMA_BIND=0).  Either pointer may be NULL. */
int npswf_host_upload_rate(const npswf_handle *h, double *raw_gb_per_s, int32_t *numa_bound_cpus);

/* Same, all pointers DEVICE memory on the handle's device `dev_slot` (index into cfg.devices),
 * enqueued on `stream` (a cudaStream_t; NULL = the library's own stream) and NOT synchronised. */
int npswf_analyze_batch_device(npswf_handle *h, int32_t dev_slot, int64_t n_events, const double *d_signal,
                               const int32_t *d_pres, const double *d_corr_time_HMS, int32_t *d_wfnpulse,
                               double *d_wftime, double *d_wfampl, double *d_chi2, double *d_timewf,
                               double *d_amplwf, uint8_t *d_status, void *stream);
/* Folds the device-side counters of the last npswf_analyze_batch_device calls into the handle
 * (synchronises the slot's stream). */
int npswf_sync_device(npswf_handle *h, int32_t dev_slot, void *stream);

/* Per-stage device timing (CUDA events on the launching stream around front / search / fit of every
 * chunk; replaces the TStopwatch prints of T2:1121-1124).  Times are folded in at the next
 * npswf_sync_device / end of a host-buffer call.  ms_* are sums over n_chunks chunks. */
int npswf_set_profiling(npswf_handle *h, int on);
int npswf_get_stage_times(npswf_handle *h, double *ms_front, double *ms_search, double *ms_fit, int64_t *n_chunks,
                          int reset);

/* ---- stage-level batch entry points (host buffers), one per reference function ---- */

/* FindPulsesMF (T2:124-216) for every present block of every event: matched filter, TSpectrum
 * search, peak filter.  wftime in TSpectrum bin units (half-integers), wfampl = |s[ti]-min|. */
int npswf_find_pulses_mf_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                               int32_t *wfnpulse, double *wftime, double *wfampl);

/* PassClusterThreshold (T2:218-278) for every block of every event: ok[e][bn] in {0,1}. */
int npswf_pass_cluster_threshold_batch(npswf_handle *h, int64_t n_events, const double *signal,
                                       const int32_t *pres, uint8_t *ok);

/* Fitwf (T2:601-828) for every (event, block) with fit_mask != 0: wfnpulse in; wftime/wfampl
 * in-out ([12] slots per block, TSpectrum seeds in, fitted / fallback values out); chi2, status out. */
int npswf_fitwf_batch(npswf_handle *h, int64_t n_events, const double *signal, const double *corr_time_HMS,
                      const uint8_t *fit_mask, const int32_t *wfnpulse, double *wftime, double *wfampl,
                      double *chi2, uint8_t *status);

/* Matched-filter tap (T2:145-179): mf [E][NBLOCKS][NTIME] float, the TH1F bin contents. */
int npswf_matched_filter_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                               float *mf);

/* TSpectrum tap for tests: runs the search kernel on caller-supplied float histograms
 * hist [n][NTIME]; outputs npeaks [n], pos_x [n][12] (fPositionX before Search()'s bin-centre
 * conversion), smoothed [n][138] (Markov output), decon [n][NTIME] (destVector).  Any output may be NULL. */
int npswf_tspectrum_debug(npswf_handle *h, int64_t n, const float *hist, int32_t *npeaks, double *pos_x,
                          double *smoothed, double *decon);

/* Host helpers (no GPU): derived calibration as the device sees it. */
int npswf_get_mf_calib(const npswf_handle *h, double *mfyref /*[NBLOCKS][11]*/, double *mfint /*[NBLOCKS]*/);
int npswf_get_spline(const npswf_handle *h, double *coef /*[NBLOCKS][109][4] = y,b,c,d*/);
/* Device address of that spline table / timeref on dev_slot (for on-device synthetic generators). */
const double *npswf_device_spline(const npswf_handle *h, int32_t dev_slot);
const double *npswf_device_timeref(const npswf_handle *h, int32_t dev_slot);

/* Output packing of T2:1289-1296: flattened wfampl/wftime of one event (blocks with pulses only,
 * in block order) and blockOffset[NBLOCKS+1] (T2:959-961, 1022).  Returns the total pulse count. */
int64_t npswf_flatten_event(const int32_t *wfnpulse, const double *wftime_padded, const double *wfampl_padded,
                            double *wftime_flat, double *wfampl_flat, int32_t *block_offset);

/* ---- callers on either side of the hot path ------------------------------------------------------- */

/* analyze's waveform unpack (T2:851-889).  samp = the packed hcana stream NPS.cal.fly.adcSampWaveform of n_events
 * events, concatenated: event e is samp[offsets[e] .. offsets[e+1]) (its NSampWaveForm words), a list of records
 * [slot, nsamp, nsamp samples].  Slots 2000/2001 are the scintillator PMs (no block, T2:862-865); a slot outside
 * [0, 1104) ends the event (T2:867-872); an event with more than 1104*112 words is skipped (T2:830-836).
 * Out: signal[E][1080][110] (zero-filled, T2:851), pres[E][1080].  Host buffers. */
int npswf_unpack_batch(npswf_handle *h, int64_t n_events, const double *samp, const int64_t *offsets, double *signal,
                       int32_t *pres);

/* npswf_analyze_batch fed with the packed stream (what the reference's analyze receives, T2:540): the unpack runs
 * on the device, so only the words of present blocks cross PCIe.  Outputs as npswf_analyze_batch. */
int npswf_analyze_batch_packed(npswf_handle *h, int64_t n_events, const double *samp, const int64_t *offsets,
                               const double *corr_time_HMS, int32_t *wfnpulse, double *wftime, double *wfampl,
                               double *chi2, double *timewf, double *amplwf, uint8_t *status);

/* Per-event diagnostics that land in the WF tree: ampl[E][1080] = pulse maximum per block (init -100, T2:591,
 * 1051-1056), enertot[E] = sum of the samples with 30 < it < 109 (T2:1038-1042), integtot[E] = sum of all samples
 * (T2:1035-1036).  The sums are deterministic and exact on the ADC lattice; any output pointer may be NULL.
 * _batch: host buffers; _device: resident buffers on device slot dev_slot, asynchronous on `stream`. */
int npswf_event_diagnostics_batch(npswf_handle *h, int64_t n_events, const double *signal, double *ampl,
                                  double *enertot, double *integtot);
int npswf_event_diagnostics_device(npswf_handle *h, int32_t dev_slot, int64_t n_events, const double *d_signal,
                                   double *d_ampl, double *d_enertot, double *d_integtot, void *stream);

/* Bit-exactness tap: the deterministic exp used by the Markov smoothing kernel, evaluated on
 * the device for n inputs (host buffers). */
int npswf_debug_exp(npswf_handle *h, int64_t n, const double *x, double *y);

/* Bit-exactness tap: the search kernel inlines the refinement chains of IEEE division / square root without
 * the out-of-range slow path.  Runs >= n_trials device-generated operand pairs of the Markov-step domain
 * through both and counts results that differ from __ddiv_rn / __dsqrt_rn:
 * mismatch[0] = b / sqrt(s) chain, mismatch[1] = a / b chain (plus the quotients of the fused deconvolution pass'
 * approximate division that are more than 2 ulp off).  Both must be 0. */
int npswf_debug_exact_ops(npswf_handle *h, int64_t n_trials, uint64_t seed, uint64_t mismatch[2]);

/* Measurement tap: FP64 FMA throughput of device 0 (GFLOP/s, FMA = 2 flop) from a register-resident chain
 * microbenchmark -- the denominator of the FP64-bound stages (TSpectrum search, template fit). */
int npswf_debug_fp64_peak(npswf_handle *h, double *gflops);

#ifdef __cplusplus
}
#endif
#endif
