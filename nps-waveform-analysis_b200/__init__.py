"""nps-waveform-analysis_b200 — host-side mirror of the reference's hot-path interface over libnpswf.so.

The reference (/root/reference/TEST_2.C, the npsWF.C lineage) is a compiled ROOT macro; ROOT is not
available in this image, so the host side above the C ABI (include/npswf.h) is this thin ctypes
binding whose method names, argument meaning and sentinel behaviour follow the reference's own
functions: analyze (T2:540), FindPulsesMF (T2:124), PassClusterThreshold (T2:218), Fitwf (T2:601).
A C++ mirror for ROOT macros is in include/npswf_host.hpp and macros/npsWF_gpu.C.

There is no CPU fallback: if lib/libnpswf.so is missing the import fails, and if no B200 is
usable every compute call raises NpsWfError.

Import with:  importlib.import_module("nps-waveform-analysis_b200")
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NPSWF_LIB") or os.path.join(_HERE, "lib", "libnpswf.so")   # NPSWF_LIB: A/B builds of the same ABI

NTIME, NCOL, NLIN, NBLOCKS, MAXWFPULSES, MFWIDTH = 110, 30, 36, 1080, 12, 11
ST_PRESENT, ST_OKTOFIT, ST_FIT_OK1, ST_FIT_OK2, ST_FALLBACK = 1, 2, 4, 8, 16
ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_CALIB = -1, -2, -3, -4
FIT_FAST, FIT_MIGRAD, FIT_VM = 0, 1, 2


class NpsWfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("npswf error %d: %s" % (code, msg))
        self.code = code


class NpsWfConfig(C.Structure):
    _fields_ = [("specthres", C.c_double), ("mfthres", C.c_double), ("trig_thres", C.c_double),
                ("coinc_width", C.c_int32), ("dt", C.c_double), ("timerefacc", C.c_double),
                ("n_devices", C.c_int32), ("devices", C.POINTER(C.c_int32)), ("chunk_events", C.c_int32),
                ("fit_max_iter", C.c_int32), ("fit_retry_max_iter", C.c_int32), ("fit_mode", C.c_int32)]


class NpsWfCalib(C.Structure):
    _fields_ = [("interpX", C.c_void_p), ("interpY", C.c_void_p), ("timeref", C.c_void_p),
                ("cortime", C.c_void_p), ("preswf", C.c_void_p)]


class NpsWfCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_events", "n_block_waveforms", "n_present", "n_pass_threshold", "n_fit_attempted", "n_fit_ok_first",
        "n_fit_ok_retry", "n_fallback", "n_pulses", "n_peak_buffer_full", "n_fit_iterations", "n_fit_evals")]


EXPORTS = [
    "npswf_default_config", "npswf_create", "npswf_destroy", "npswf_last_error", "npswf_get_counters",
    "npswf_reset_counters", "npswf_device_count", "npswf_host_alloc", "npswf_host_free", "npswf_analyze_batch",
    "npswf_analyze_batch_i16", "npswf_analyze_batch_device", "npswf_sync_device", "npswf_find_pulses_mf_batch",
    "npswf_pass_cluster_threshold_batch", "npswf_fitwf_batch", "npswf_matched_filter_batch",
    "npswf_tspectrum_debug", "npswf_get_mf_calib", "npswf_get_spline", "npswf_device_spline",
    "npswf_device_timeref", "npswf_flatten_event", "npswf_debug_exp", "npswf_debug_exact_ops", "npswf_debug_fp64_peak", "npswf_unpack_batch", "npswf_analyze_batch_packed",
    "npswf_event_diagnostics_batch", "npswf_event_diagnostics_device", "npswf_set_profiling",
    "npswf_get_stage_times", "npswf_set_host_packing", "npswf_host_packing_stats", "npswf_debug_pack_counts", "npswf_analyze_batch_flat",
    "npswf_analyze_batch_flat_i16", "npswf_host_upload_rate", "npswf_hcana_pulses", "npswf_event_times", "npswf_debug_vm_reasons",
    "npswf_debug_search_fused",
]

_lib = None


def lib():
    """Load libnpswf.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                              "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.npswf_last_error.restype = C.c_char_p
        L.npswf_last_error.argtypes = [C.c_void_p]
        L.npswf_host_alloc.restype = C.c_void_p
        L.npswf_host_alloc.argtypes = [C.c_size_t]
        L.npswf_host_free.argtypes = [C.c_void_p]
        L.npswf_device_spline.restype = C.c_void_p
        L.npswf_device_spline.argtypes = [C.c_void_p, C.c_int32]
        L.npswf_device_timeref.restype = C.c_void_p
        L.npswf_device_timeref.argtypes = [C.c_void_p, C.c_int32]
        L.npswf_flatten_event.restype = C.c_int64
        L.npswf_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def pack_counts(x, lsb_mV=1000.0 / 4096, n_threads=2):
    """Host-only tap of the lossless int16 transport packer: returns (lossless, counts)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.shape, dtype=np.int16)
    rc = lib().npswf_debug_pack_counts(x.ctypes.data_as(C.c_void_p), C.c_int64(x.size), C.c_double(lsb_mV),
                                       C.c_int32(n_threads), out.ctypes.data_as(C.c_void_p))
    if rc < 0:
        raise NpsWfError(rc, "npswf_debug_pack_counts: bad arguments")
    return bool(rc), out


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'ed memory (npswf_host_alloc); keeps the owner alive."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib().npswf_host_alloc(C.c_size_t(max(n, 1)))
    if not ptr:
        raise NpsWfError(ERR_NOMEM, "npswf_host_alloc failed")
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            try:
                lib().npswf_host_free(C.c_void_p(self.p))
            except Exception:
                pass
    arr_owner = _Owner(ptr)
    return _Pinned(arr, arr_owner, buf)


class _Pinned(np.ndarray):
    def __new__(cls, arr, owner, buf):
        obj = arr.view(cls)
        obj._owner = owner
        obj._buf = buf
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)
            self._buf = getattr(obj, "_buf", None)


def flatten_event(wfnpulse, wftime_padded, wfampl_padded):
    """Output packing of T2:1289-1296: (wftime_flat, wfampl_flat, blockOffset[1081]) of one event."""
    n = _c(wfnpulse, np.int32).reshape(NBLOCKS)
    t = _c(wftime_padded, np.float64).reshape(NBLOCKS, MAXWFPULSES)
    a = _c(wfampl_padded, np.float64).reshape(NBLOCKS, MAXWFPULSES)
    tot = int(np.minimum(n, MAXWFPULSES).sum())
    tf = np.zeros(tot); af = np.zeros(tot); off = np.zeros(NBLOCKS + 1, np.int32)
    got = lib().npswf_flatten_event(_p(n), _p(t), _p(a), _p(tf), _p(af), _p(off))
    assert got == tot
    return tf, af, off


class NpsWf:
    """Handle over (tunables, calibration) — the reference's file-scope globals (T2:51-85)."""

    def __init__(self, calib, specthres=0.02, mfthres=1.5, trig_thres=10.0, coinc_width=20, dt=4.0,
                 timerefacc=0.0, devices=None, chunk_events=0, fit_max_iter=0, fit_retry_max_iter=0, fit_mode=FIT_FAST):
        L = lib()
        cfg = NpsWfConfig()
        L.npswf_default_config(C.byref(cfg))
        cfg.specthres, cfg.mfthres, cfg.trig_thres = specthres, mfthres, trig_thres
        cfg.coinc_width, cfg.dt, cfg.timerefacc = coinc_width, dt, timerefacc
        cfg.chunk_events, cfg.fit_max_iter, cfg.fit_retry_max_iter = chunk_events, fit_max_iter, fit_retry_max_iter
        cfg.fit_mode = fit_mode
        self.chunk_events = chunk_events if chunk_events > 0 else 1184   # library default (events per chunk)
        self._dev = None
        if devices is not None:
            self._dev = (C.c_int32 * len(devices))(*devices)
            cfg.n_devices = len(devices)
            cfg.devices = C.cast(self._dev, C.POINTER(C.c_int32))
        self._keep = dict(
            interpX=_c(calib["interpX"], np.float64), interpY=_c(calib["interpY"], np.float64),
            timeref=_c(calib["timeref"], np.float64), cortime=_c(calib["cortime"], np.float32),
            preswf=_c(calib["preswf"], np.int32))
        cal = NpsWfCalib(*[_p(self._keep[k]) for k in ("interpX", "interpY", "timeref", "cortime", "preswf")])
        h = C.c_void_p()
        rc = L.npswf_create(C.byref(cfg), C.byref(cal), C.byref(h))
        if rc != 0:
            raise NpsWfError(rc, L.npswf_last_error(None).decode())
        self.h = h
        self.n_devices = max(1, cfg.n_devices)

    def close(self):
        if getattr(self, "h", None):
            lib().npswf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise NpsWfError(rc, lib().npswf_last_error(self.h).decode())

    # ---- host-side derived calibration (no GPU needed)
    def mf_calib(self):
        y = np.zeros((NBLOCKS, MFWIDTH)); i = np.zeros(NBLOCKS)
        self._check(lib().npswf_get_mf_calib(self.h, _p(y), _p(i)))
        return y, i

    def spline_coeffs(self):
        out = np.zeros((NBLOCKS, NTIME - 1, 4))
        self._check(lib().npswf_get_spline(self.h, _p(out)))
        return out

    def device_spline_ptr(self, slot=0):
        return lib().npswf_device_spline(self.h, C.c_int32(slot))

    def device_timeref_ptr(self, slot=0):
        return lib().npswf_device_timeref(self.h, C.c_int32(slot))

    def counters(self):
        c = NpsWfCounters()
        self._check(lib().npswf_get_counters(self.h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in NpsWfCounters._fields_}

    def set_profiling(self, on=True):
        self._check(lib().npswf_set_profiling(self.h, C.c_int(1 if on else 0)))

    def stage_times(self, reset=True):
        """Summed CUDA-event times (ms) of the front / search / fit stages and the chunk count."""
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        self._check(lib().npswf_get_stage_times(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(n),
                                                C.c_int(1 if reset else 0)))
        return dict(front_ms=a.value, search_ms=b.value, fit_ms=c.value, chunks=n.value)

    def reset_counters(self):
        self._check(lib().npswf_reset_counters(self.h))

    def set_host_packing(self, mode=1, n_threads=0, lsb_mV=0.0):
        """Transport of analyze()'s binary64 traces: 0 = doubles, 1 = automatic, 2 = int16 counts whenever lossless."""
        self._check(lib().npswf_set_host_packing(self.h, C.c_int(mode), C.c_int(n_threads), C.c_double(lsb_mV)))

    def host_packing_stats(self):
        a, b, r, n = C.c_int64(), C.c_int64(), C.c_double(), C.c_int64()
        self._check(lib().npswf_host_packing_stats(self.h, C.byref(a), C.byref(b), C.byref(r), C.byref(n)))
        return dict(packed_chunks=a.value, raw_chunks=b.value, pack_gb_per_s=r.value, packed_input_bytes=n.value)

    def vm_reasons(self, reset=True):
        """NPSWF_FIT_VM diagnostics: fits handed to the exact Migrad kernels, by reason (see npswf.h)."""
        out = np.zeros(8, np.uint64)
        self._check(lib().npswf_debug_vm_reasons(self.h, _p(out), C.c_int(1 if reset else 0)))
        return [int(v) for v in out]

    def search_fused(self, reset=True):
        """(spectra deconvolved with the fused evaluation, of those repeated with the reference's arithmetic); npswf.h."""
        out = np.zeros(2, np.uint64)
        self._check(lib().npswf_debug_search_fused(self.h, _p(out), C.c_int(1 if reset else 0)))
        return int(out[0]), int(out[1])

    def host_upload_rate(self):
        """(measured GB/s of the raw binary64 uploads, host cores the transport threads are bound to)."""
        r, c = C.c_double(), C.c_int32()
        self._check(lib().npswf_host_upload_rate(self.h, C.byref(r), C.byref(c)))
        return r.value, c.value

    # ---- analyze(event) over a batch (T2:540-1300 hot path)
    @staticmethod
    def alloc_outputs(E, pinned=False):
        mk = pinned_empty if pinned else (lambda s, d: np.empty(s, d))
        return dict(wfnpulse=mk((E, NBLOCKS), np.int32), wftime=mk((E, NBLOCKS, MAXWFPULSES), np.float64),
                    wfampl=mk((E, NBLOCKS, MAXWFPULSES), np.float64), chi2=mk((E, NBLOCKS), np.float64),
                    timewf=mk((E, NBLOCKS), np.float64), amplwf=mk((E, NBLOCKS), np.float64),
                    status=mk((E, NBLOCKS), np.uint8))

    def analyze(self, signal, pres, corr_time_HMS, out=None):
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        co = _c(corr_time_HMS, np.float64).reshape(E)
        o = out if out is not None else self.alloc_outputs(E)
        self._check(lib().npswf_analyze_batch(self.h, C.c_int64(E), _p(sig), _p(pr), _p(co), _p(o["wfnpulse"]),
                                              _p(o["wftime"]), _p(o["wfampl"]), _p(o["chi2"]), _p(o["timewf"]),
                                              _p(o["amplwf"]), _p(o["status"])))
        return o

    @staticmethod
    def alloc_flat_outputs(E, capacity, pinned=False):
        mk = pinned_empty if pinned else (lambda s, d: np.empty(s, d))
        return dict(wfnpulse=mk((E, NBLOCKS), np.int32), pulse_offset=mk((E,), np.int64), pulse_count=mk((E,), np.int32),
                    wftime_pool=mk((capacity,), np.float64), wfampl_pool=mk((capacity,), np.float64),
                    chi2=mk((E, NBLOCKS), np.float64), timewf=mk((E, NBLOCKS), np.float64),
                    amplwf=mk((E, NBLOCKS), np.float64), status=mk((E, NBLOCKS), np.uint8))

    def analyze_flat(self, signal, pres, corr_time_HMS, capacity=None, out=None):
        """analyze() with wfampl / wftime in the reference's own packing (T2:1289-1296): event e's vectors are
        pool[pulse_offset[e] : pulse_offset[e] + pulse_count[e]], pulses in block order.  Packed on the device."""
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        co = _c(corr_time_HMS, np.float64).reshape(E)
        if out is None:
            out = self.alloc_flat_outputs(E, capacity if capacity is not None else E * NBLOCKS * MAXWFPULSES)
        npul = C.c_int64(0)
        self._check(lib().npswf_analyze_batch_flat(
            self.h, C.c_int64(E), _p(sig), _p(pr), _p(co), _p(out["wfnpulse"]), _p(out["pulse_offset"]), _p(out["pulse_count"]),
            _p(out["wftime_pool"]), _p(out["wfampl_pool"]), C.c_int64(out["wftime_pool"].size), _p(out["chi2"]),
            _p(out["timewf"]), _p(out["amplwf"]), _p(out["status"]), C.byref(npul)))
        out["n_pulses"] = npul.value
        return out

    def analyze_flat_i16(self, counts, lsb_mV, pres, corr_time_HMS, capacity=None, out=None):
        """analyze_flat() on int16 ADC counts: the leanest host transport in both directions."""
        cn = _c(counts, np.int16).reshape(-1, NBLOCKS, NTIME)
        E = cn.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        co = _c(corr_time_HMS, np.float64).reshape(E)
        if out is None:
            out = self.alloc_flat_outputs(E, capacity if capacity is not None else E * NBLOCKS * MAXWFPULSES)
        npul = C.c_int64(0)
        self._check(lib().npswf_analyze_batch_flat_i16(
            self.h, C.c_int64(E), _p(cn), C.c_double(lsb_mV), _p(pr), _p(co), _p(out["wfnpulse"]), _p(out["pulse_offset"]),
            _p(out["pulse_count"]), _p(out["wftime_pool"]), _p(out["wfampl_pool"]), C.c_int64(out["wftime_pool"].size),
            _p(out["chi2"]), _p(out["timewf"]), _p(out["amplwf"]), _p(out["status"]), C.byref(npul)))
        out["n_pulses"] = npul.value
        return out

    def analyze_i16(self, counts, lsb_mV, pres, corr_time_HMS, out=None):
        cn = _c(counts, np.int16).reshape(-1, NBLOCKS, NTIME)
        E = cn.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        co = _c(corr_time_HMS, np.float64).reshape(E)
        o = out if out is not None else self.alloc_outputs(E)
        self._check(lib().npswf_analyze_batch_i16(self.h, C.c_int64(E), _p(cn), C.c_double(lsb_mV), _p(pr), _p(co),
                                                  _p(o["wfnpulse"]), _p(o["wftime"]), _p(o["wfampl"]), _p(o["chi2"]),
                                                  _p(o["timewf"]), _p(o["amplwf"]), _p(o["status"])))
        return o

    def unpack(self, samp, offsets):
        """analyze's waveform unpack (T2:851-889) of E packed events -> (signal[E,1080,110], pres[E,1080])."""
        sp = _c(samp, np.float64).ravel()
        of = _c(offsets, np.int64).ravel()
        E = of.size - 1
        sig = np.zeros((E, NBLOCKS, NTIME)); pres = np.zeros((E, NBLOCKS), np.int32)
        self._check(lib().npswf_unpack_batch(self.h, C.c_int64(E), _p(sp), _p(of), _p(sig), _p(pres)))
        return sig, pres

    def analyze_packed(self, samp, offsets, corr_time_HMS, out=None):
        """analyze() on the packed stream NPS.cal.fly.adcSampWaveform (unpacked on the device)."""
        sp = _c(samp, np.float64).ravel()
        of = _c(offsets, np.int64).ravel()
        E = of.size - 1
        co = _c(corr_time_HMS, np.float64).reshape(E)
        o = out if out is not None else self.alloc_outputs(E)
        self._check(lib().npswf_analyze_batch_packed(self.h, C.c_int64(E), _p(sp), _p(of), _p(co), _p(o["wfnpulse"]),
                                                     _p(o["wftime"]), _p(o["wfampl"]), _p(o["chi2"]), _p(o["timewf"]),
                                                     _p(o["amplwf"]), _p(o["status"])))
        return o

    def event_diagnostics(self, signal):
        """(ampl[E,1080], enertot[E], integtot[E]) of the WF tree (T2:1026-1056)."""
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        ampl = np.zeros((E, NBLOCKS)); et = np.zeros(E); it = np.zeros(E)
        self._check(lib().npswf_event_diagnostics_batch(self.h, C.c_int64(E), _p(sig), _p(ampl), _p(et), _p(it)))
        return ampl, et, it

    def analyze_device(self, n_events, d_signal, d_pres, d_corr, d_wfnpulse, d_wftime, d_wfampl, d_chi2, d_timewf,
                       d_amplwf, d_status, stream=0, slot=0):
        """All d_* are raw device pointers (ints); enqueued on `stream`, not synchronised."""
        v = C.c_void_p
        self._check(lib().npswf_analyze_batch_device(
            self.h, C.c_int32(slot), C.c_int64(n_events), v(d_signal), v(d_pres), v(d_corr), v(d_wfnpulse),
            v(d_wftime), v(d_wfampl), v(d_chi2), v(d_timewf), v(d_amplwf), v(d_status), v(stream)))

    def sync_device(self, stream=0, slot=0):
        self._check(lib().npswf_sync_device(self.h, C.c_int32(slot), C.c_void_p(stream)))

    # ---- stage-level, named after the reference's functions
    def FindPulsesMF(self, signal, pres):
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        n = np.zeros((E, NBLOCKS), np.int32); t = np.zeros((E, NBLOCKS, MAXWFPULSES)); a = np.zeros_like(t)
        self._check(lib().npswf_find_pulses_mf_batch(self.h, C.c_int64(E), _p(sig), _p(pr), _p(n), _p(t), _p(a)))
        return n, t, a

    def PassClusterThreshold(self, signal, pres):
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        ok = np.zeros((E, NBLOCKS), np.uint8)
        self._check(lib().npswf_pass_cluster_threshold_batch(self.h, C.c_int64(E), _p(sig), _p(pr), _p(ok)))
        return ok.astype(bool)

    def Fitwf(self, signal, corr_time_HMS, fit_mask, wfnpulse, wftime, wfampl):
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        co = _c(corr_time_HMS, np.float64).reshape(E)
        mk = _c(fit_mask, np.uint8).reshape(E, NBLOCKS)
        n = _c(wfnpulse, np.int32).reshape(E, NBLOCKS)
        t = _c(wftime, np.float64).reshape(E, NBLOCKS, MAXWFPULSES).copy()
        a = _c(wfampl, np.float64).reshape(E, NBLOCKS, MAXWFPULSES).copy()
        chi2 = np.zeros((E, NBLOCKS)); st = np.zeros((E, NBLOCKS), np.uint8)
        self._check(lib().npswf_fitwf_batch(self.h, C.c_int64(E), _p(sig), _p(co), _p(mk), _p(n), _p(t), _p(a),
                                            _p(chi2), _p(st)))
        return dict(wftime=t, wfampl=a, chi2=chi2, status=st)

    def matched_filter(self, signal, pres):
        sig = _c(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
        E = sig.shape[0]
        pr = _c(pres, np.int32).reshape(E, NBLOCKS)
        mf = np.zeros((E, NBLOCKS, NTIME), np.float32)
        self._check(lib().npswf_matched_filter_batch(self.h, C.c_int64(E), _p(sig), _p(pr), _p(mf)))
        return mf

    def tspectrum_debug(self, hist):
        hs = _c(hist, np.float32).reshape(-1, NTIME)
        n = hs.shape[0]
        npk = np.zeros(n, np.int32); px = np.zeros((n, MAXWFPULSES)); sm = np.zeros((n, NTIME + 28)); de = np.zeros((n, NTIME))
        self._check(lib().npswf_tspectrum_debug(self.h, C.c_int64(n), _p(hs), _p(npk), _p(px), _p(sm), _p(de)))
        return npk, px, sm, de

    def debug_exp(self, x):
        xs = _c(x, np.float64).ravel()
        y = np.zeros_like(xs)
        self._check(lib().npswf_debug_exp(self.h, C.c_int64(xs.size), _p(xs), _p(y)))
        return y

    def fp64_peak_gflops(self):
        """Measured FP64 FMA throughput of device 0 (GFLOP/s)."""
        g = C.c_double(0.0)
        self._check(lib().npswf_debug_fp64_peak(self.h, C.byref(g)))
        return float(g.value)

    def debug_exact_ops(self, n_trials, seed=1):
        """(mismatches of the b/sqrt(s) chain, mismatches of the a/b chain) vs IEEE div/sqrt on the device."""
        mm = np.zeros(2, np.uint64)
        self._check(lib().npswf_debug_exact_ops(self.h, C.c_int64(int(n_trials)), C.c_uint64(int(seed)), _p(mm)))
        return int(mm[0]), int(mm[1])


def hcana_pulses(adcCounter, adcSampPulseTime, adcSampPulseTimeRaw, adcSampPulseAmp, tdcoffset, timemean2):
    """T2:893-939 for one event: (corr_time_HMS, Sampampl[1080], Samptime[1080])."""
    ac = _c(adcCounter, np.float64).ravel(); pt = _c(adcSampPulseTime, np.float64).ravel()
    pr = _c(adcSampPulseTimeRaw, np.float64).ravel(); pa = _c(adcSampPulseAmp, np.float64).ravel()
    td = _c(tdcoffset, np.float32).ravel(); tm = _c(timemean2, np.float32).ravel()
    corr = C.c_double(0.0)
    sa = np.zeros(NBLOCKS); stime = np.zeros(NBLOCKS)
    rc = lib().npswf_hcana_pulses(C.c_int32(ac.size), _p(ac), _p(pt), _p(pr), _p(pa), _p(td), _p(tm), C.byref(corr), _p(sa), _p(stime))
    if rc:
        raise NpsWfError(rc, "npswf_hcana_pulses: bad arguments")
    return corr.value, sa, stime


def event_times(wfnpulse, wftime_padded, wfampl_padded, status, cortime, dt=4.0):
    """T2:988-996 for one event: (h1time, h2time) from its analysis outputs."""
    n = _c(wfnpulse, np.int32).reshape(NBLOCKS)
    t = _c(wftime_padded, np.float64).reshape(NBLOCKS, MAXWFPULSES); a = _c(wfampl_padded, np.float64).reshape(NBLOCKS, MAXWFPULSES)
    st = _c(status, np.uint8).reshape(NBLOCKS); ct = _c(cortime, np.float32).reshape(NBLOCKS)
    h1 = np.zeros(NBLOCKS * MAXWFPULSES); h2 = np.zeros(NBLOCKS * MAXWFPULSES)
    L = lib()
    L.npswf_event_times.restype = C.c_int64
    k = L.npswf_event_times(_p(n), _p(t), _p(a), _p(st), _p(ct), C.c_double(dt), _p(h1), _p(h2))
    if k < 0:
        raise NpsWfError(int(k), "npswf_event_times: bad arguments")
    return h1[:k].copy(), h2[:k].copy()


def shard_range(n_events, rank, world_size):
    """Contiguous event range of `rank` (the partition npswf_analyze_batch uses across devices)."""
    return n_events * rank // world_size, n_events * (rank + 1) // world_size
