"""Synthetic calibration and events for the NPS waveform path.

The JLab reference-waveform / timing files the reference loads (/root/reference/TEST_2.C:370-469)
are unavailable, so the derived arrays are replicated synthetically.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NTIME, NBLOCKS, MAXP, MFW, MFLEFT = 110, 1080, 12, 11, 5
LSB = 1000.0 / 4096.0


def make_calibration(seed=20240001):
    """interpX/interpY/timeref/cortime/preswf as the loader at T2:360-469 would produce them."""
    rng = np.random.Generator(np.random.Philox(seed))
    alpha = rng.uniform(1.5, 2.5, NBLOCKS)
    tau_d = rng.uniform(2.5, 4.0, NBLOCKS)
    kpeak = rng.integers(33, 39, NBLOCKS)           # peak sample, production pulses sit near 35.5 (T2:80)
    frac = rng.uniform(-0.4, 0.4, NBLOCKS)
    t = np.arange(NTIME, dtype=np.float64)
    interpX = np.tile(t, (NBLOCKS, 1))
    t0 = (kpeak + frac) - alpha * tau_d
    u = np.clip(t[None, :] - t0[:, None], 0.0, None)
    y = np.where(u > 0, u ** alpha[:, None] * np.exp(-u / tau_d[:, None]), 0.0)
    y /= y.max(axis=1, keepdims=True)               # "normalized to 1" (README.md:60)
    timeref = interpX[np.arange(NBLOCKS), y.argmax(axis=1)].copy()   # T2:434-438
    cortime = rng.uniform(-2.0, 2.0, NBLOCKS).astype(np.float32)
    cortime[cortime == 0] = np.float32(-0.0000001)  # T2:464-467
    preswf = np.ones(NBLOCKS, np.int32)
    # matched-filter gain of a unit pulse sitting on the reference position (for amp_mode 1)
    idx = timeref.astype(int)[:, None] + np.arange(MFW)[None, :] - MFLEFT
    mfy = np.take_along_axis(y, idx, axis=1)
    kappa = (mfy * mfy[:, ::-1]).sum(axis=1) / mfy.sum(axis=1)
    return dict(interpX=interpX, interpY=np.ascontiguousarray(y), timeref=timeref, cortime=cortime,
                preswf=preswf, kappa=kappa)


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("nmin", C.c_int32), ("nmax", C.c_int32), ("amp_mode", C.c_int32),
                ("a_lo", C.c_double), ("a_hi", C.c_double), ("tau_lo", C.c_double), ("tau_hi", C.c_double),
                ("min_sep", C.c_double), ("noise_sigma", C.c_double), ("ped_lo", C.c_double),
                ("ped_hi", C.c_double), ("absent_frac", C.c_double)]


def config_params(config, seed=None, absent_frac=0.0):
    """BASELINE.json configs[0..2]."""
    if config == 1:   # single pulse/block, A ~ logU[5,500], tau ~ U[-10,40]
        return SynthParams(seed or 1, 1, 1, 0, 5.0, 500.0, -10.0, 40.0, 3.0, 0.30, -2.0, 2.0, absent_frac)
    if config == 2:   # 1-3 pulses/block with pile-up, separations >= 3 bins, A ~ logU[3,500]
        return SynthParams(seed or 2, 1, 3, 0, 3.0, 500.0, -10.0, 40.0, 3.0, 0.30, -2.0, 2.0, absent_frac)
    if config == 3:   # up to 12 pulses near the 1.5 mV MF threshold
        return SynthParams(seed or 3, 0, 12, 1, 1.2, 3.0, -20.0, 55.0, 5.0, 0.30, -2.0, 2.0, absent_frac)
    raise ValueError(config)


_host = None


def host_lib():
    global _host
    if _host is None:
        path = os.path.join(_HERE, "libnpswf_synth.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "..", "oracle"), "-s"])
        _host = C.CDLL(path)
    return _host


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def generate_host(params, spline, calib, event0, n_events, n_threads=4, counts=False, truth=False):
    """Events [event0, event0+n) on the host. Returns dict(signal, pres, corr_time_HMS[, counts, truth_*])."""
    spline = np.ascontiguousarray(spline, np.float64)
    timeref = np.ascontiguousarray(calib["timeref"], np.float64)
    kappa = np.ascontiguousarray(calib["kappa"], np.float64)
    E = int(n_events)
    out = dict(signal=np.zeros((E, NBLOCKS, NTIME)), pres=np.zeros((E, NBLOCKS), np.int32),
               corr_time_HMS=np.zeros(E))
    cnt = np.zeros((E, NBLOCKS, NTIME), np.int16) if counts else None
    tn = np.zeros((E, NBLOCKS), np.int32) if truth else None
    tp = np.zeros((E, NBLOCKS, MAXP)) if truth else None
    ta = np.zeros((E, NBLOCKS, MAXP)) if truth else None
    td = np.zeros((E, NBLOCKS)) if truth else None
    host_lib().synth_generate_host(C.byref(params), _p(spline), _p(timeref), _p(kappa), C.c_int64(event0),
                                   C.c_int64(E), _p(out["signal"]), _p(cnt), _p(out["pres"]),
                                   _p(out["corr_time_HMS"]), _p(tn), _p(tp), _p(ta), _p(td), C.c_int(n_threads))
    if counts:
        out["counts"] = cnt
    if truth:
        out.update(truth_n=tn, truth_pos=tp, truth_amp=ta, truth_ped=td)
    return out


_cuda = None


def cuda_lib():
    global _cuda
    if _cuda is None:
        path = os.path.join(_HERE, "libnpswf_synth_cuda.so")
        if not os.path.exists(path):
            raise RuntimeError("synth/libnpswf_synth_cuda.so missing: run __graft_entry__.build()")
        _cuda = C.CDLL(path)
    return _cuda


def generate_device(params, d_spline, d_timeref, d_kappa, event0, n_events, d_signal=0, d_counts=0, d_pres=0,
                    d_corr=0, stream=0):
    """Same generator as a CUDA kernel; all d_* are raw device pointers (ints)."""
    rc = cuda_lib().synth_generate_device(C.byref(params), C.c_void_p(d_spline), C.c_void_p(d_timeref),
                                          C.c_void_p(d_kappa), C.c_int64(event0), C.c_int64(n_events),
                                          C.c_void_p(d_signal), C.c_void_p(d_counts), C.c_void_p(d_pres),
                                          C.c_void_p(d_corr), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError("synth_generate_device: CUDA error %d" % rc)


def pack_events(signal, pres, seed=0, scintillators=True):
    """The packed hcana stream NPS.cal.fly.adcSampWaveform (T2:855-889) of a batch: for every present block a record
    [slot, 110, samples...] in shuffled slot order, plus the two scintillator PM slots 2000 / 2001 (T2:862-865).
    Returns (samp, offsets) with event e = samp[offsets[e]:offsets[e+1]]."""
    rng = np.random.default_rng(seed)
    sig = np.asarray(signal, np.float64).reshape(-1, NBLOCKS, NTIME)
    E = sig.shape[0]
    chunks, offs = [], [0]
    for e in range(E):
        slots = np.nonzero(np.asarray(pres)[e] == 1)[0]
        slots = rng.permutation(slots)
        rec = np.empty((slots.size, NTIME + 2))
        rec[:, 0] = slots; rec[:, 1] = NTIME; rec[:, 2:] = sig[e, slots]
        parts = [rec.ravel()]
        if scintillators:
            sc = np.empty((2, NTIME + 2)); sc[:, 0] = (2000, 2001); sc[:, 1] = NTIME; sc[:, 2:] = rng.normal(50, 5, (2, NTIME))
            parts.insert(rng.integers(0, 2), sc.ravel())
        ev = np.concatenate(parts)
        chunks.append(ev)
        offs.append(offs[-1] + ev.size)
    return np.concatenate(chunks) if chunks else np.zeros(0), np.array(offs, np.int64)
