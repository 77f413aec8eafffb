"""CPU test of the product's scalar Migrad core (nps-waveform-analysis_b200/csrc/migrad_core.hpp), the code lane 0 of
fit_migrad_kernel and every thread of fit_migrad_thread_kernel execute: compiled for the host with a serial chi2
(tests/cpp/migrad_core_host.cpp, g++ -O2 -ffp-contract=off) and compared BIT FOR BIT with the oracle's Migrad
restatement on real fit problems of the three BASELINE configurations -- fitted parameters, chi2, number of chi2
evaluations and the ok / retry / fall-back verdict.  The oracle is the checker; the shim is test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mgcore") / "libmgcore_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-std=c++17", "-shared", "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "migrad_core_host.cpp")])
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _core_fit(core, spl_b, knots, trace, timeref, N, t, a):
    par = np.zeros(25)
    fmin = C.c_double()
    nc = C.c_int32()
    st = core.mgcore_fitwf(_p(np.ascontiguousarray(spl_b)), None if knots is None else _p(np.ascontiguousarray(knots)),
                           _p(np.ascontiguousarray(trace)), C.c_double(timeref), C.c_int(int(N)), _p(t), _p(a), _p(par),
                           C.byref(fmin), C.byref(nc))
    return st, par[:2 * N + 1], fmin.value, nc.value


@pytest.mark.parametrize("cfg,n_events,stride", [(1, 1, 3), (2, 1, 2), (3, 2, 2)])
def test_core_equals_oracle_migrad_bitwise(core, orc, calib, spline, cfg, n_events, stride):
    ev = synth.generate_host(synth.config_params(cfg), spline, calib, 31_000 + cfg, n_events, n_threads=4)
    sig = np.asarray(ev["signal"]).reshape(n_events, 1080, 110)
    n_fit = n_retry = n_fb = 0
    for e in range(n_events):
        for b in range(0, 1080, stride):
            N, t, a = orc.find_pulses_mf(b, sig[e], ev["pres"][e])
            if N == 0 or not orc.pass_cluster_threshold(b, sig[e], ev["pres"][e]):
                continue
            r = orc.fitwf(b, sig[e], N, t, a, 0.0)
            st, par, fmin, nc = _core_fit(core, spline[b], None, sig[e, b], calib["timeref"][b], N, t, a)
            assert st == r["status"] and nc == r["ncalls"], (cfg, e, b, N, st, r["status"], nc, r["ncalls"])
            if st != 16:
                assert np.array_equal(par, r["params"]), (cfg, e, b, N)
                assert fmin / (90 - (2 * N + 1)) == r["chi2"], (cfg, e, b, N)
            n_fit += 1
            n_retry += st == 8
            n_fb += st == 16
    print("cfg%d: %d fits identical to the oracle's Migrad (retries %d, fall-backs %d)" % (cfg, n_fit, n_retry, n_fb))
    assert n_fit > 300


def test_core_on_hard_fits(core, orc, calib, spline):
    """Fits that exercise the rare branches: seeds far from any pulse (negative second derivatives -> NegativeG2LineSearch,
    MnPosDef), a flat trace (zero gradient everywhere), twelve seeds on a single pulse (call limit / fall-back)."""
    rng = np.random.default_rng(5)
    b = 417
    shape = calib["interpY"][b]
    n = 0
    kinds = set()
    for trial in range(60):
        kind = trial % 4
        trace = np.round(rng.normal(0, 0.3, 110) / synth.LSB) * synth.LSB
        if kind == 0:      # pulse, seeds 8-15 bins off
            trace += np.round(80.0 * shape / synth.LSB) * synth.LSB
            N, t, a = 1, np.array([calib["timeref"][b] + rng.uniform(8, 15)] + [-999.0] * 11), np.array([20.0] + [-999.0] * 11)
        elif kind == 1:    # no pulse at all
            N, t, a = 2, np.array([40.5, 60.5] + [-999.0] * 10), np.array([3.0, 2.0] + [-999.0] * 10)
        elif kind == 2:    # twelve seeds, one pulse
            trace += np.round(30.0 * shape / synth.LSB) * synth.LSB
            N, t, a = 12, np.sort(rng.uniform(12, 98, 12)), np.full(12, 2.5)
        else:              # negative amplitude seed
            trace -= np.round(25.0 * shape / synth.LSB) * synth.LSB
            N, t, a = 1, np.array([calib["timeref"][b] + 0.5] + [-999.0] * 11), np.array([10.0] + [-999.0] * 11)
        sig = np.zeros((1080, 110))
        sig[b] = trace
        r = orc.fitwf(b, sig, N, t, a, 0.0)
        st, par, fmin, nc = _core_fit(core, spline[b], None, trace, calib["timeref"][b], N, t.copy(), a.copy())
        assert st == r["status"] and nc == r["ncalls"], (trial, kind, st, r["status"], nc, r["ncalls"])
        if st != 16:
            assert np.array_equal(par, r["params"]), (trial, kind)
        kinds.add((kind, st))
        n += 1
    print("hard fits: %d identical; (kind, verdict) seen: %s" % (n, sorted(kinds)))
    assert len({s for _, s in kinds}) >= 2     # both converged and failed fits occurred
