/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the per-event, per-block waveform path of
 * /root/reference/TEST_2.C (the npsWF.C lineage): analyze() block loop (T2:942-1022),
 * FindPulsesMF (T2:124-216), PassClusterThreshold (T2:218-278), Fitwf (T2:601-828), plus
 * the three un-vendored ROOT calls they make (TSpectrum::Search, Interpolator(kCSPLINE),
 * Fit::Fitter + Minuit2 Migrad), restated from the published algorithms.
 *
 * PARITY UNPINNED: the reference ships no tests/golden vectors and
 * ROOT/GSL/Minuit2 are not installable here, so nothing in this oracle could be checked
 * against an execution of the reference itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may call into this library.  The product (libnpswf.so) never links or loads it.
 */
#ifndef NPSWF_ORACLE_H
#define NPSWF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* compile-time shape, as in the reference's constants block (T2:51-73) */
#define OR_NTIME 110
#define OR_NCOL 30
#define OR_NLIN 36
#define OR_NBLOCKS (OR_NCOL * OR_NLIN)
#define OR_MAXWFPULSES 12
#define OR_MFLEFT 5
#define OR_MFRIGHT 5
#define OR_MFWIDTH 11
#define OR_MFSTART 10
#define OR_MFEND 100

enum {
    ORACLE_FLAG_LIBM_EXP = 1,      /* use libm exp in the Markov smoothing instead of det_exp */
    ORACLE_FLAG_FAITHFUL_COST = 2, /* rebuild the spline per fit + global mutex around the search (T2:186, 612) */
    ORACLE_FLAG_FIT_LM = 4         /* minimise with analytic-gradient LM instead of the Migrad restatement */
};

typedef struct OracleConfig {
    double specthres;   /* 0.02  T2:70 */
    double mfthres;     /* 1.5   T2:71 */
    double trig_thres;  /* 10    T2:72 */
    int32_t coinc_width; /* 20   T2:73 */
    double dt;          /* 4     T2:354 */
    double timerefacc;  /* T2:81, T2:524 */
    int32_t flags;
} OracleConfig;

typedef struct OracleCalib {
    const double *interpX;  /* [B][T] T2:84, 432 */
    const double *interpY;  /* [B][T] */
    const double *timeref;  /* [B]   T2:77, 437 */
    const float *cortime;   /* [B]   T2:78, 463 (Float_t) */
    const int32_t *preswf;  /* [B]   T2:79, 452 */
} OracleCalib;

typedef struct OracleHandle OracleHandle;

/* status bits written per (event, block) */
enum {
    OR_ST_PRESENT = 1,     /* pres && preswf (T2:944) */
    OR_ST_OKTOFIT = 2,     /* PassClusterThreshold true (T2:962) */
    OR_ST_FIT_OK1 = 4,     /* first LeastSquareFit ok (T2:755) */
    OR_ST_FIT_OK2 = 8,     /* retry ok (T2:768) */
    OR_ST_FALLBACK = 16    /* both failed -> TSpectrum values, chi2=-100 (T2:774-791) */
};

OracleHandle *oracle_create(const OracleConfig *cfg, const OracleCalib *cal);
void oracle_destroy(OracleHandle *h);
/* derived calibration, as built at T2:440-451 */
void oracle_get_mf(const OracleHandle *h, double *mfyref /*[B][11]*/, double *mfint /*[B]*/);
/* natural cubic spline coefficients of block bn: y,b,c,d per interval [109] (GSL cspline) */
void oracle_get_spline(const OracleHandle *h, int bn, double *y, double *b, double *c, double *d);
double oracle_spline_eval(const OracleHandle *h, int bn, double x);

/* stage-level entry points (one block of one event) */
void oracle_matched_filter(const OracleHandle *h, int bn, const double *signal_event, double minsignal,
                           double *mfvals /*[T] double*/, float *mfhist /*[T] float, may be NULL*/);
int oracle_find_pulses_mf(const OracleHandle *h, int bn, const double *signal_event, const int32_t *pres,
                          double minsignal, double *wftime /*[12]*/, double *wfampl /*[12]*/);
int oracle_pass_cluster_threshold(const OracleHandle *h, int bn, const double *signal_event, const int32_t *pres);
/* Fitwf for one block: wftime/wfampl are in-out (12 slots). Returns status bits (OK1/OK2/FALLBACK or 0 if npulse==0) */
int oracle_fitwf(const OracleHandle *h, int bn, const double *signal_event, int npulse, double corr_time_HMS,
                 double *wftime, double *wfampl, double *chi2, int32_t *ncalls, double *raw_params /*[25] or NULL*/);

/* analyze(event) block loop for n_events events, event-parallel over n_threads std::threads */
int oracle_analyze_batch(const OracleHandle *h, int64_t n_events, const double *signal /*[E][B][T]*/,
                         const int32_t *pres /*[E][B]*/, const double *corr_time_HMS /*[E]*/,
                         int32_t *wfnpulse /*[E][B]*/, double *wftime /*[E][B][12]*/, double *wfampl /*[E][B][12]*/,
                         double *chi2 /*[E][B]*/, double *timewf /*[E][B]*/, double *amplwf /*[E][B]*/,
                         uint8_t *status /*[E][B]*/, int32_t *ncalls /*[E][B] or NULL*/, int n_threads);

/* FindPulsesMF for all blocks of a batch (no fits), event-parallel */
int oracle_find_pulses_batch(const OracleHandle *h, int64_t n_events, const double *signal, const int32_t *pres,
                             int32_t *wfnpulse, double *wftime /*[E][B][12]*/, double *wfampl, int n_threads);

/* TSpectrum restatement (tspectrum.cpp) */
/* callers on either side of the hot path: waveform unpack T2:830-889, diagnostics T2:1026-1056 */
int oracle_unpack_event(const double *samp, int64_t n_words, double *signal /*[B*T]*/, int32_t *pres /*[B]*/,
                        double *minsignal /*[B] or NULL*/);
void oracle_event_diagnostics(const double *signal /*[B*T]*/, double *ampl /*[B]*/, double *enertot, double *integtot);

/* HMS correction + hcana pulse selection T2:893-939; h1time / h2time of an event T2:988-996 */
void oracle_hcana_pulses(int32_t NadcCounter, double *adcCounter, const double *adcSampPulseTime, const double *adcSampPulseTimeRaw,
                         const double *adcSampPulseAmp, const float *tdcoffset /*[B]*/, const float *timemean2 /*[B]*/,
                         double *corr_time_HMS, double *Sampampl /*[B]*/, double *Samptime /*[B]*/);
int oracle_event_times(const OracleHandle *h, const double *signal_event, const int32_t *pres, double corr_time_HMS,
                       double *h1time /*[B*12]*/, double *h2time /*[B*12]*/);

int oracle_search_highres(const double *source, int ssize, double sigma, double threshold, int decon_iterations,
                          int aver_window, int max_peaks, double *pos_x, double *smoothed_out, double *decon_out,
                          int use_libm_exp);
int oracle_tspectrum_search(const float *hist, int nbins, double sigma, double threshold_frac, int max_peaks,
                            double *position_x, double *position_y, int use_libm_exp);
double oracle_det_exp_c(double x);

/* Minuit2 Migrad restatement on a generic chi2 (minuit_migrad.cpp), for unit tests */
typedef double (*oracle_fcn_t)(const double *par, void *user);
int oracle_migrad(oracle_fcn_t fcn, void *user, int npar, const double *start, const double *step, int strategy,
                  unsigned maxfcn, double tolerance, double *par_out, double *fmin_out, double *edm_out,
                  int32_t *ncalls_out, int32_t *status_out);

#ifdef __cplusplus
}
#endif
#endif
