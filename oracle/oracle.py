"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libnpswf_oracle.so, the CPU restatement of the reference's
per-event, per-block waveform path (/root/reference/TEST_2.C; see npswf_oracle.h).
PARITY UNPINNED: the reference ships no golden vectors and ROOT is not installable here.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libnpswf_oracle.so")

NTIME, NCOL, NLIN, NBLOCKS, MAXP, MFW = 110, 30, 36, 1080, 12, 11
FLAG_LIBM_EXP, FLAG_FAITHFUL_COST, FLAG_FIT_LM = 1, 2, 4
ST_PRESENT, ST_OKTOFIT, ST_FIT_OK1, ST_FIT_OK2, ST_FALLBACK = 1, 2, 4, 8, 16


def build(force=False):
    """Compile the oracle (and the host synthetic generator) with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [("specthres", C.c_double), ("mfthres", C.c_double), ("trig_thres", C.c_double),
                ("coinc_width", C.c_int32), ("dt", C.c_double), ("timerefacc", C.c_double), ("flags", C.c_int32)]


class _Cal(C.Structure):
    _fields_ = [("interpX", C.c_void_p), ("interpY", C.c_void_p), ("timeref", C.c_void_p),
                ("cortime", C.c_void_p), ("preswf", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_Cfg), C.POINTER(_Cal)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_spline_eval.restype = C.c_double
        L.oracle_spline_eval.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.oracle_det_exp_c.restype = C.c_double
        L.oracle_det_exp_c.argtypes = [C.c_double]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def det_exp(x):
    L = lib()
    return np.array([L.oracle_det_exp_c(float(v)) for v in np.atleast_1d(x)])


def search_highres(source, sigma=2.0, threshold=2.0, decon_iterations=3, aver_window=3, max_peaks=12, libm_exp=False):
    """TSpectrum::SearchHighRes restatement. Returns (npeaks, pos_x[npeaks], smoothed[ssize+28], decon[ssize])."""
    L = lib()
    src = _c(source, np.float64)
    n = src.size
    shift = int(7 * sigma + 0.5)
    px = np.zeros(max(max_peaks, 1)); sm = np.zeros(n + 2 * shift); de = np.zeros(n)
    L.oracle_search_highres.restype = C.c_int
    npk = L.oracle_search_highres(_p(src), C.c_int(n), C.c_double(sigma), C.c_double(threshold),
                                  C.c_int(decon_iterations), C.c_int(aver_window), C.c_int(max_peaks),
                                  _p(px), _p(sm), _p(de), C.c_int(1 if libm_exp else 0))
    return npk, px[:npk].copy(), sm, de


def tspectrum_search(hist, sigma=2.0, threshold=0.02, max_peaks=12, libm_exp=False):
    """TSpectrum(max_peaks).Search(h, sigma, "nobackground,nodraw", threshold) on float bin contents."""
    L = lib()
    h = _c(hist, np.float32)
    px = np.zeros(max(max_peaks, 1)); py = np.zeros(max(max_peaks, 1))
    L.oracle_tspectrum_search.restype = C.c_int
    n = L.oracle_tspectrum_search(_p(h), C.c_int(h.size), C.c_double(sigma), C.c_double(threshold),
                                  C.c_int(max_peaks), _p(px), _p(py), C.c_int(1 if libm_exp else 0))
    return n, px[:n].copy(), py[:n].copy()


_FCN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)


def unpack_event(samp):
    """analyze's waveform unpack (T2:830-889) of one event's packed stream -> (signal[1080,110], pres[1080], minsignal[1080])."""
    L = lib()
    sp = _c(samp, np.float64).ravel()
    sig = np.zeros((NBLOCKS, NTIME)); pres = np.zeros(NBLOCKS, np.int32); mn = np.zeros(NBLOCKS)
    L.oracle_unpack_event(_p(sp), C.c_int64(sp.size), _p(sig), _p(pres), _p(mn))
    return sig, pres, mn


def event_diagnostics(signal_event):
    """(ampl[1080], enertot, integtot) of one event (T2:1026-1056)."""
    L = lib()
    sig = _c(signal_event, np.float64).reshape(NBLOCKS, NTIME)
    ampl = np.zeros(NBLOCKS); et = C.c_double(0.0); it = C.c_double(0.0)
    L.oracle_event_diagnostics(_p(sig), _p(ampl), C.byref(et), C.byref(it))
    return ampl, float(et.value), float(it.value)


def hcana_pulses(adcCounter, adcSampPulseTime, adcSampPulseTimeRaw, adcSampPulseAmp, tdcoffset, timemean2):
    """T2:893-939 for one event: (corr_time_HMS, Sampampl[1080], Samptime[1080])."""
    L = lib()
    ac = _c(adcCounter, np.float64).ravel().copy(); pt = _c(adcSampPulseTime, np.float64).ravel()
    pr = _c(adcSampPulseTimeRaw, np.float64).ravel(); pa = _c(adcSampPulseAmp, np.float64).ravel()
    td = _c(tdcoffset, np.float32).ravel(); tm = _c(timemean2, np.float32).ravel()
    corr = C.c_double(0.0); sa = np.zeros(NBLOCKS); st = np.zeros(NBLOCKS)
    L.oracle_hcana_pulses(C.c_int32(ac.size), _p(ac), _p(pt), _p(pr), _p(pa), _p(td), _p(tm), C.byref(corr), _p(sa), _p(st))
    return corr.value, sa, st


def migrad(fcn, start, step, strategy=1, maxfcn=0, tolerance=0.01):
    """Minuit2 Migrad restatement on a Python chi2 callable (unit tests)."""
    L = lib()
    n = len(start)
    cb = _FCN(lambda p, u: float(fcn(np.array([p[i] for i in range(n)]))))
    s = _c(start, np.float64); w = _c(step, np.float64); out = np.zeros(n)
    fmin = C.c_double(); edm = C.c_double(); nc = C.c_int32(); st = C.c_int32()
    L.oracle_migrad.restype = C.c_int
    ok = L.oracle_migrad(cb, None, C.c_int(n), _p(s), _p(w), C.c_int(strategy), C.c_uint(maxfcn),
                         C.c_double(tolerance), _p(out), C.byref(fmin), C.byref(edm), C.byref(nc), C.byref(st))
    return dict(valid=bool(ok), par=out, fval=fmin.value, edm=edm.value, ncalls=nc.value, status=st.value)


class Oracle:
    """Handle over (config, calibration): mirrors the reference's file-scope globals (T2:51-85)."""

    def __init__(self, calib, specthres=0.02, mfthres=1.5, trig_thres=10.0, coinc_width=20, dt=4.0,
                 timerefacc=0.0, flags=0, lib_path=None):
        # lib_path: another build of the same restatement (the FMA-contracted one of `make -C oracle fma`)
        if lib_path is None:
            L = lib()
        else:
            L = C.CDLL(lib_path)
            L.oracle_create.restype = C.c_void_p
            L.oracle_create.argtypes = [C.POINTER(_Cfg), C.POINTER(_Cal)]
            L.oracle_destroy.argtypes = [C.c_void_p]
        self._lib = L
        self._keep = dict(
            interpX=_c(calib["interpX"], np.float64), interpY=_c(calib["interpY"], np.float64),
            timeref=_c(calib["timeref"], np.float64), cortime=_c(calib["cortime"], np.float32),
            preswf=_c(calib["preswf"], np.int32))
        cfg = _Cfg(specthres, mfthres, trig_thres, coinc_width, dt, timerefacc, flags)
        cal = _Cal(*[_p(self._keep[k]) for k in ("interpX", "interpY", "timeref", "cortime", "preswf")])
        self.h = C.c_void_p(L.oracle_create(C.byref(cfg), C.byref(cal)))
        self.timerefacc = timerefacc
        self.dt = dt

    def __del__(self):
        try:
            if self.h:
                self._lib.oracle_destroy(self.h); self.h = None
        except Exception:
            pass

    def mf_calib(self):
        y = np.zeros((NBLOCKS, MFW)); i = np.zeros(NBLOCKS)
        self._lib.oracle_get_mf(self.h, _p(y), _p(i))
        return y, i

    def spline_coeffs(self):
        """[B][109][4] = (y, b, c, d) of the natural cubic spline per unit interval."""
        out = np.zeros((NBLOCKS, NTIME - 1, 4))
        y = np.zeros(NTIME - 1); b = np.zeros(NTIME - 1); c = np.zeros(NTIME - 1); d = np.zeros(NTIME - 1)
        for bn in range(NBLOCKS):
            self._lib.oracle_get_spline(self.h, C.c_int(bn), _p(y), _p(b), _p(c), _p(d))
            out[bn, :, 0] = y; out[bn, :, 1] = b; out[bn, :, 2] = c; out[bn, :, 3] = d
        return out

    def spline_eval(self, bn, x):
        return np.array([self._lib.oracle_spline_eval(self.h, int(bn), float(v)) for v in np.atleast_1d(x)])

    def matched_filter(self, bn, signal_event, minsignal=None):
        sig = _c(signal_event, np.float64).reshape(NBLOCKS, NTIME)
        if minsignal is None:
            minsignal = min(1e6, sig[bn].min())
        mf = np.zeros(NTIME); hist = np.zeros(NTIME, np.float32)
        self._lib.oracle_matched_filter(self.h, C.c_int(bn), _p(sig), C.c_double(minsignal), _p(mf), _p(hist))
        return mf, hist

    def find_pulses_mf(self, bn, signal_event, pres, minsignal=None):
        sig = _c(signal_event, np.float64).reshape(NBLOCKS, NTIME)
        pr = _c(pres, np.int32)
        if minsignal is None:
            minsignal = min(1e6, sig[bn].min())
        t = np.full(MAXP, -999.0); a = np.full(MAXP, -999.0)
        self._lib.oracle_find_pulses_mf.restype = C.c_int
        n = self._lib.oracle_find_pulses_mf(self.h, C.c_int(bn), _p(sig), _p(pr), C.c_double(minsignal), _p(t), _p(a))
        return n, t, a

    def pass_cluster_threshold(self, bn, signal_event, pres):
        sig = _c(signal_event, np.float64).reshape(NBLOCKS, NTIME)
        pr = _c(pres, np.int32)
        self._lib.oracle_pass_cluster_threshold.restype = C.c_int
        return bool(self._lib.oracle_pass_cluster_threshold(self.h, C.c_int(bn), _p(sig), _p(pr)))

    def fitwf(self, bn, signal_event, npulse, wftime, wfampl, corr_time_HMS=0.0):
        sig = _c(signal_event, np.float64).reshape(NBLOCKS, NTIME)
        t = _c(wftime, np.float64).copy(); a = _c(wfampl, np.float64).copy()
        chi2 = C.c_double(); nc = C.c_int32(); raw = np.zeros(25)
        self._lib.oracle_fitwf.restype = C.c_int
        st = self._lib.oracle_fitwf(self.h, C.c_int(bn), _p(sig), C.c_int(npulse), C.c_double(corr_time_HMS),
                                _p(t), _p(a), C.byref(chi2), C.byref(nc), _p(raw))
        return dict(status=st, wftime=t, wfampl=a, chi2=chi2.value, ncalls=nc.value, params=raw[:2 * npulse + 1])

    def event_times(self, signal_event, pres, corr_time_HMS=0.0):
        """(h1time, h2time) of one event as analyze fills them (T2:988-996)."""
        sig = _c(signal_event, np.float64).reshape(NBLOCKS, NTIME)
        pr = _c(pres, np.int32)
        h1 = np.zeros(NBLOCKS * MAXP); h2 = np.zeros(NBLOCKS * MAXP)
        self._lib.oracle_event_times.restype = C.c_int
        n = self._lib.oracle_event_times(self.h, _p(sig), _p(pr), C.c_double(corr_time_HMS), _p(h1), _p(h2))
        return h1[:n].copy(), h2[:n].copy()

    def find_pulses_batch(self, signal, pres, n_threads=1):
        """FindPulsesMF for every block of a batch (no fits): (wfnpulse[E,1080], wftime[E,1080,12], wfampl[E,1080,12])."""
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        n = np.zeros((E, NBLOCKS), np.int32); t = np.zeros((E, NBLOCKS, MAXP)); a = np.zeros((E, NBLOCKS, MAXP))
        self._lib.oracle_find_pulses_batch(self.h, C.c_int64(E), _p(sig), _p(pr), _p(n), _p(t), _p(a), C.c_int(n_threads))
        return n, t, a

    def analyze_batch(self, signal, pres, corr_time_HMS, n_threads=1):
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        co = _c(corr_time_HMS, np.float64).reshape(E)
        out = dict(
            wfnpulse=np.zeros((E, NBLOCKS), np.int32), wftime=np.zeros((E, NBLOCKS, MAXP)),
            wfampl=np.zeros((E, NBLOCKS, MAXP)), chi2=np.zeros((E, NBLOCKS)), timewf=np.zeros((E, NBLOCKS)),
            amplwf=np.zeros((E, NBLOCKS)), status=np.zeros((E, NBLOCKS), np.uint8),
            ncalls=np.zeros((E, NBLOCKS), np.int32))
        self._lib.oracle_analyze_batch(self.h, C.c_int64(E), _p(sig), _p(pr), _p(co), _p(out["wfnpulse"]),
                                   _p(out["wftime"]), _p(out["wfampl"]), _p(out["chi2"]), _p(out["timewf"]),
                                   _p(out["amplwf"]), _p(out["status"]), _p(out["ncalls"]), C.c_int(n_threads))
        return out
