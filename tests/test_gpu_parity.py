"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libnpswf.so), against the CPU
oracle on the same seeded inputs, and against the committed golden fixtures.

Bar (BASELINE.json north_star): exact equality for peak count / position / order, matched-filter
bin contents and the cluster-threshold decision; for blocks where both fits converge
|dt| <= 0.01 bin, |dA|/A <= 1e-3, relative chi2 <= 1e-3.  The reference minimiser (Migrad) and the
GPU minimiser (analytic LM) are different algorithms, so on multi-pulse blocks a small fraction
lands in a different local minimum of the same chi2; the tests state and bound that fraction."""
import os

import numpy as np
import pytest

import oracle
import synth
from conftest import golden_path

pytestmark = pytest.mark.gpu

TOL_T_BIN, TOL_A_REL, TOL_CHI2_REL = 0.01, 1e-3, 1e-3


def _fit_agreement(ref, got, dt=4.0):
    """Per-block agreement over blocks where both fits converged. Returns (n_both, frac_within_tol)."""
    both = ((ref["status"] & 12) > 0) & ((got["status"] & 12) > 0)
    n = ref["wfnpulse"]
    valid = np.arange(12)[None, None, :] < n[..., None]
    d_t = np.where(valid, np.abs(ref["wftime"] - got["wftime"]) / dt, 0.0).max(axis=-1)
    d_a = np.where(valid, np.abs(ref["wfampl"] - got["wfampl"]) / np.maximum(np.abs(ref["wfampl"]), 1e-300), 0.0).max(axis=-1)
    d_c = np.abs(ref["chi2"] - got["chi2"]) / np.maximum(np.abs(ref["chi2"]), 1e-300)
    good = (d_t <= TOL_T_BIN) & (d_a <= TOL_A_REL) & (d_c <= TOL_CHI2_REL)
    return int(both.sum()), float(good[both].mean()) if both.any() else 1.0, both, good


def test_det_exp_bit_exact(gpu):
    x = np.concatenate([np.random.default_rng(2).uniform(-3, 3, 200000), np.linspace(-1.5, 1.5, 4097), [0.0, -0.0]])
    assert np.array_equal(gpu.debug_exp(x), oracle.det_exp(x))


def test_inlined_div_sqrt_chains_match_ieee(gpu):
    """The search kernel's branch-free division / square-root chains vs __ddiv_rn / __dsqrt_rn, 2e9 operand pairs."""
    assert gpu.debug_exact_ops(2_000_000_000, seed=7) == (0, 0)


@pytest.mark.parametrize("cfg", [1, 2, 3])
def test_matched_filter_bit_exact(gpu, orc, events, cfg):
    ev = events[cfg]
    mf = gpu.matched_filter(ev["signal"], ev["pres"])
    for e in range(ev["signal"].shape[0]):
        for b in range(0, 1080, 3):
            if ev["pres"][e, b]:
                assert np.array_equal(mf[e, b], orc.matched_filter(b, ev["signal"][e])[1]), (cfg, e, b)


def test_matched_filter_bit_exact_adversarial(gpu, orc):
    """The matched filter is evaluated fast + error bound + exact re-evaluation of undecided outputs: arbitrary
    (off-lattice) doubles over 9 orders of magnitude, flat traces (every output ties for the minimum), repeated
    minima, a huge pulse on a quiet baseline, negative pedestals -- every stored float must equal the oracle's."""
    rng = np.random.default_rng(17)
    sig = np.zeros((4, 1080, 110))
    x = np.arange(110.0)
    for b in range(1080):
        kind = b % 9
        if kind == 0:
            sig[:, b] = rng.normal(0.0, 0.3, (4, 110)) * 10.0 ** rng.uniform(-4, 5)
        elif kind == 1:
            sig[:, b] = rng.uniform(-50, 50)                                   # flat
        elif kind == 2:
            sig[:, b] = np.round(rng.normal(0, 1.0, (4, 110)) * 4) / 4        # coarse lattice: many exact ties
        elif kind == 3:
            sig[:, b] = rng.normal(0, 0.3, (4, 110)) + 1.0e4 * np.exp(-0.5 * ((x - rng.uniform(20, 90)) / 2.5) ** 2)
        elif kind == 4:
            sig[:, b] = -200.0 + rng.normal(0, 0.01, (4, 110))
        elif kind == 5:
            sig[:, b] = np.tile(rng.normal(0, 1, 11), 10)                      # periodic with the filter length
        elif kind == 6:
            sig[:, b] = rng.normal(0, 0.3, (4, 110)) + rng.uniform(3, 500) * np.exp(-0.5 * ((x - rng.uniform(15, 95)) / 3.0) ** 2)
        elif kind == 7:
            sig[:, b] = rng.standard_cauchy((4, 110))
        else:
            sig[:, b] = (rng.integers(0, 4096, (4, 110)) - 300) * synth.LSB
    pres = np.ones((4, 1080), np.int32)
    mf = gpu.matched_filter(sig, pres)
    bad = 0
    for e in range(4):
        for b in range(1080):
            ref = orc.matched_filter(b, sig[e])[1]
            if not np.array_equal(mf[e, b], ref):
                bad += 1
                assert bad < 0, (e, b, b % 9, np.nonzero(mf[e, b] != ref)[0][:8], mf[e, b][mf[e, b] != ref][:4], ref[mf[e, b] != ref][:4])


def test_tspectrum_intermediates_bit_exact(gpu, orc, events):
    """Markov smoothing, Gold deconvolution and fPositionX against the oracle, bitwise."""
    hists = []
    for cfg in (1, 2, 3):
        ev = events[cfg]
        for b in range(0, 1080, 11):
            if ev["pres"][0, b]:
                hists.append(orc.matched_filter(b, ev["signal"][0])[1])
    x = np.arange(110)
    edge = np.zeros((6, 110), np.float32)
    edge[1, 50] = 7.0                                                   # single spike
    edge[2, 5:105] = 3.0                                                # flat top
    edge[3] = (5 * np.exp(-0.5 * ((x - 8) / 2.0) ** 2)).astype(np.float32)   # peak at the left edge, non-zero source[0..3]
    edge[4] = (5 * np.exp(-0.5 * ((x - 107) / 2.0) ** 2)).astype(np.float32)  # right edge
    for c in np.arange(8, 104, 7):
        edge[5] += ((10.0 + (c * 37 % 40)) * np.exp(-0.5 * ((x - c) / 2.0) ** 2)).astype(np.float32)  # > 12 peaks
    hists = np.concatenate([np.array(hists, np.float32), edge])
    npk, px, sm, de = gpu.tspectrum_debug(hists)
    for i, h in enumerate(hists):
        n, pos, s, d = oracle.search_highres(h.astype(np.float64))
        assert npk[i] == n, i
        assert np.array_equal(sm[i], s), i
        assert np.array_equal(de[i], d), i
        assert np.array_equal(px[i, :n], pos), i


@pytest.mark.parametrize("cfg", [1, 2, 3])
def test_find_pulses_exact(gpu, orc, events, cfg):
    ev = events[cfg]
    n, t, a = gpu.FindPulsesMF(ev["signal"], ev["pres"])
    cal_preswf = 1
    for e in range(ev["signal"].shape[0]):
        for b in range(1080):
            on, ot, oa = (0, None, None)
            if ev["pres"][e, b] == 1 and cal_preswf:
                on, ot, oa = orc.find_pulses_mf(b, ev["signal"][e], ev["pres"][e])
            assert n[e, b] == on, (cfg, e, b)
            if on:
                assert np.array_equal(t[e, b, :on], ot[:on]) and np.array_equal(a[e, b, :on], oa[:on]), (cfg, e, b)
            assert (t[e, b, on:] == -999).all() and (a[e, b, on:] == -999).all()


def test_search_fused_pass_equals_reference_arithmetic(pkg, calib, orc, monkeypatch):
    """The peak search deconvolves with FMA chains first and repeats a spectrum with the reference's arithmetic
    (rounded product, rounded sum) when a decision lies inside the error margin of that evaluation
    (kernel_search.cuh, gold_block).  Three handles -- exact arithmetic only, the default, and the fused pass
    followed by the exact repeat for EVERY spectrum -- must return the same bits: peak count, order, times
    (integer bins), amplitudes.  Inputs: pile-up near threshold (config 3), 1-3 pulses (config 2), and traces
    built to stress the margins: huge pulses on a quiet baseline (deconvolved tails near the 1e-5 gates of the
    Gold iteration), equal twin pulses (ties between neighbouring channels), plateaus, scaled copies."""
    rng = np.random.default_rng(23)
    sets = []
    for cfg, n_ev in ((3, 6), (2, 6)):
        ev = synth.generate_host(synth.config_params(cfg), orc.spline_coeffs(), calib, 40 + cfg, n_ev, n_threads=4)
        sets.append((ev["signal"], ev["pres"]))
    sig = rng.normal(0.0, 0.3, (4, 1080, 110))
    lsb = 1000.0 / 4096.0
    shape = np.exp(-0.5 * ((np.arange(110)[None, :] - 45.0) / 3.0) ** 2)
    amp = 10.0 ** rng.uniform(0.5, 4.0, (1080, 1))
    sig[0] += amp * shape                                                         # 3 mV .. 10 V single pulses
    sig[1] += amp * (shape + np.roll(shape, 12, axis=1))                          # equal twins 12 bins apart
    sig[2] = np.round((amp * np.clip(shape * 3.0, 0, 1.0)) / lsb) * lsb            # noise-free plateaus on the ADC lattice
    sig[3] = sig[0] * rng.choice([0.25, 0.5, 2.0, 4.0], (1080, 1))                # power-of-two copies
    sets.append((sig, np.ones((4, 1080), np.int32)))
    results = {}
    for mode in (0, 1, 2):
        monkeypatch.setenv("NPSWF_SEARCH_FUSED", str(mode))
        h = pkg.NpsWf(calib)
        h.search_fused(reset=True)
        results[mode] = [h.FindPulsesMF(s_, p_) for s_, p_ in sets]
        fused, redone = h.search_fused(reset=True)
        if mode == 0:
            assert fused == 0
        elif mode == 1:
            assert fused > 10000 and redone <= 0.02 * fused, (fused, redone)
        else:
            assert fused > 10000 and redone == fused
        h.close()
    monkeypatch.delenv("NPSWF_SEARCH_FUSED")
    n_peaks = 0
    for k in range(len(sets)):
        for mode in (1, 2):
            for x, y in zip(results[0][k], results[mode][k]):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), (k, mode)
        n_peaks += int(results[0][k][0].sum())
    assert n_peaks > 20000
    # and the exact-arithmetic handle against the oracle on the stress set
    n, t, a = results[0][2]
    s_, p_ = sets[2]
    for e in range(s_.shape[0]):
        for b in range(0, 1080, 5):
            on, ot, oa = orc.find_pulses_mf(b, s_[e], p_[e])
            assert n[e, b] == on, (e, b)
            if on:
                assert np.array_equal(t[e, b, :on], ot[:on]) and np.array_equal(a[e, b, :on], oa[:on]), (e, b)


def test_search_fused_pass_stays_inside_its_error_budget(pkg, calib, gpu, orc, monkeypatch):
    """The budget behind the fused pass (kernel_search.cuh, gold_block): smoothed spectrum within 24 258 u, deconvolved
    spectrum within 170 200 u = 2^-35.6 of the reference's arithmetic (u = 2^-53), against decision margins of 2^-22.
    NPSWF_SEARCH_FUSED=3 makes the debug taps show the unchecked fused pass; the default handle's taps are the exact
    arithmetic (bitwise equal to the oracle, test_tspectrum_intermediates_bit_exact)."""
    ev = synth.generate_host(synth.config_params(3), orc.spline_coeffs(), calib, 77, 2, n_threads=4)
    ev2 = synth.generate_host(synth.config_params(2), orc.spline_coeffs(), calib, 78, 2, n_threads=4)
    mf = np.concatenate([gpu.matched_filter(e["signal"], e["pres"]).reshape(-1, 110) for e in (ev, ev2)])
    mf = mf[mf.max(axis=1) > 1.5]
    assert len(mf) > 4000
    monkeypatch.setenv("NPSWF_SEARCH_FUSED", "3")
    h3 = pkg.NpsWf(calib)
    monkeypatch.delenv("NPSWF_SEARCH_FUSED")
    n0, p0, s0, d0 = gpu.tspectrum_debug(mf)
    n3, p3, s3, d3 = h3.tspectrum_debug(mf)
    h3.close()
    u = 2.0 ** -53
    worst = {}
    for name, x0, x3, budget in (("smoothed", s0, s3, 24258 * u), ("deconvolved", d0, d3, 170200 * u)):
        assert np.array_equal(x0 == 0, x3 == 0), name   # the zero pattern
        nz = x0 != 0
        rel = np.abs(x3[nz] - x0[nz]) / np.abs(x0[nz])
        worst[name] = float(rel.max())
        assert rel.max() <= budget, (name, rel.max() / u)
    print("fused pass vs exact arithmetic on %d spectra: smoothed spectrum max %.0f u (budget 24 258), deconvolved max %.0f u "
          "(budget 170 200), peak counts equal on %d" % (len(mf), worst["smoothed"] / u, worst["deconvolved"] / u, int((n0 == n3).sum())))
    assert (n0 == n3).mean() > 0.999   # unchecked, so a spectrum with a decision inside the margin may differ


@pytest.mark.parametrize("cfg", [1, 2, 3])
def test_cluster_threshold_exact(gpu, orc, events, cfg):
    ev = events[cfg]
    ok = gpu.PassClusterThreshold(ev["signal"], ev["pres"])
    for e in range(ev["signal"].shape[0]):
        ref = np.array([orc.pass_cluster_threshold(b, ev["signal"][e], ev["pres"][e]) for b in range(1080)])
        assert np.array_equal(ok[e], ref), (cfg, e, np.nonzero(ok[e] != ref)[0][:10])


def test_cluster_threshold_off_lattice_and_near_threshold(gpu, orc):
    """Arbitrary doubles (sum order matters in the last bit) with 3x3 sums hovering around trig_thres."""
    rng = np.random.default_rng(11)
    sig = rng.normal(0.0, 0.4, (2, 1080, 110))
    sig[:, :, 30:45] += rng.uniform(0.55, 0.95, (2, 1080, 1))   # 9 * ~0.75 + noise max ~ 10 mV in the window
    pres = (rng.random((2, 1080)) > 0.1).astype(np.int32)
    ok = gpu.PassClusterThreshold(sig, pres)
    for e in range(2):
        ref = np.array([orc.pass_cluster_threshold(b, sig[e], pres[e]) for b in range(1080)])
        assert 0.05 < ref.mean() < 0.95
        assert np.array_equal(ok[e], ref)


# FAST (LM) mode on the small fixtures (3 events per config: ~3 000 fits, so +-0.5 point of statistical width in
# config 3); the gates at BASELINE size, and the MIGRAD mode's bit-for-bit comparison, are in test_gpu_migrad.py
@pytest.mark.parametrize("cfg,min_frac", [(1, 0.999), (2, 0.993), (3, 0.925)])
def test_analyze_vs_oracle(gpu, orc, events, cfg, min_frac):
    ev = events[cfg]
    ref = orc.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=8)
    got = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    assert np.array_equal(got["wfnpulse"], ref["wfnpulse"])
    assert np.array_equal(got["status"] & 3, ref["status"] & 3)          # present / okToFit
    nofit = (ref["status"] & 28) == 0                                     # not fitted: everything exact
    for k in ("wftime", "wfampl", "chi2", "timewf", "amplwf"):
        assert np.array_equal(got[k][nofit], ref[k][nofit]), k
    n_both, frac, both, good = _fit_agreement(ref, got)
    print("cfg%d: both-converged %d, within tolerance %.5f, GPU fallback %d, oracle fallback %d" % (
        cfg, n_both, frac, int(((got["status"] & 16) > 0).sum()), int(((ref["status"] & 16) > 0).sum())))
    assert n_both > 0.5 * ((ref["status"] & 2) > 0).sum()
    assert frac >= min_frac
    # timewf / amplwf are the pulse with the smallest |wftime| (T2:999-1016)
    fit = (got["status"] & 28) > 0
    tt = np.where(np.arange(12)[None, None, :] < got["wfnpulse"][..., None], np.abs(got["wftime"]), np.inf)
    sel = tt.argmin(axis=-1)
    assert np.array_equal(np.take_along_axis(got["wftime"], sel[..., None], -1)[..., 0][fit], got["timewf"][fit])


def test_fallback_values_exact(gpu, orc, events):
    """Force both attempts to fail (1 LM iteration allowed): outputs must be the TSpectrum values with the
    time converted exactly as T2:779-790 and chi2 = -100."""
    import importlib
    pkg = importlib.import_module("nps-waveform-analysis_b200")
    ev = events[2]
    cal = synth.make_calibration()
    h = pkg.NpsWf(cal, fit_max_iter=1, fit_retry_max_iter=1)
    got = h.analyze(ev["signal"][:1], ev["pres"][:1], ev["corr_time_HMS"][:1])
    n, t, a = gpu.FindPulsesMF(ev["signal"][:1], ev["pres"][:1])
    fb = (got["status"] & 16) > 0
    assert fb.sum() > 500
    tref = cal["timeref"][None, :, None]
    cort = cal["cortime"].astype(np.float64)[None, :, None]
    exp_t = (t - tref) * 4.0 + ev["corr_time_HMS"][:1, None, None] - cort - 0.0 * 4.0
    valid = (np.arange(12)[None, None, :] < n[..., None]) & fb[..., None]
    assert np.array_equal(got["wftime"][valid], exp_t[valid])
    assert np.array_equal(got["wfampl"][valid], a[valid])
    assert (got["chi2"][fb] == -100).all()


def test_i16_entry_point_equals_f64(gpu, events):
    ev = events[2]
    a = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    b = gpu.analyze_i16(ev["counts"], synth.LSB, ev["pres"], ev["corr_time_HMS"])
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_stage_fitwf_equals_pipeline(gpu, events):
    ev = events[2]
    full = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    n, t, a = gpu.FindPulsesMF(ev["signal"], ev["pres"])
    ok = gpu.PassClusterThreshold(ev["signal"], ev["pres"])
    mask = ok & (ev["pres"] == 1)
    r = gpu.Fitwf(ev["signal"], ev["corr_time_HMS"], mask, n, t, a)
    fitted = (full["status"] & 28) > 0
    assert np.array_equal((r["status"] & 28) > 0, fitted)
    assert np.array_equal(r["wftime"][fitted], full["wftime"][fitted])
    assert np.array_equal(r["chi2"][fitted], full["chi2"][fitted])


@pytest.mark.parametrize("name,min_frac", [("cfg1_acc0", 0.999), ("cfg2_acc0", 0.993), ("cfg2_accm5", 0.993),
                                           ("cfg3_acc0", 0.925)])
def test_against_golden_fixture(pkg, name, min_frac):
    g = np.load(golden_path(name + ".npz"))
    c = np.load(golden_path("calib.npz"))
    cal = dict(interpX=np.tile(np.arange(110.0), (1080, 1)), interpY=c["interpY"], timeref=c["timeref"],
               cortime=c["cortime"], preswf=c["preswf"])
    h = pkg.NpsWf(cal, timerefacc=float(g["timerefacc"]))
    got = h.analyze_i16(g["counts"], synth.LSB, g["pres"], g["corr"])
    assert np.array_equal(got["wfnpulse"], g["wfnpulse"])
    assert np.array_equal(got["status"] & 3, g["status"] & 3)
    mf = h.matched_filter(g["counts"].astype(np.float64) * synth.LSB, g["pres"])
    pres7 = g["pres"][:, ::7] == 1
    assert np.array_equal(mf[:, ::7][pres7], g["mfhist_every7"][pres7])
    nofit = (g["status"] & 28) == 0
    for k in ("wftime", "wfampl", "chi2", "timewf", "amplwf"):
        assert np.array_equal(got[k][nofit], g[k][nofit]), k
    ref = {k: g[k] for k in ("status", "wfnpulse", "wftime", "wfampl", "chi2")}
    n_both, frac, _, _ = _fit_agreement(ref, got)
    print("%s: both-converged %d, within tolerance %.5f" % (name, n_both, frac))
    assert frac >= min_frac
    # MIGRAD fit mode: the committed oracle outputs, bit for bit
    hm = pkg.NpsWf(cal, timerefacc=float(g["timerefacc"]), fit_mode=pkg.FIT_MIGRAD)
    gm = hm.analyze_i16(g["counts"], synth.LSB, g["pres"], g["corr"])
    for k in ("wfnpulse", "status", "wftime", "wfampl", "chi2", "timewf", "amplwf"):
        assert np.array_equal(gm[k], g[k]), (name, k)


def test_size_independent_properties(gpu, calib, spline):
    """Larger batch (crosses the internal chunk boundary): determinism, event-permutation equivariance,
    empty batch, and the counters' bookkeeping identities."""
    E = 700
    ev = synth.generate_host(synth.config_params(2, absent_frac=0.03), spline, calib, 50000, E, n_threads=8)
    gpu.reset_counters()
    a = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    c = gpu.counters()
    assert c["n_events"] == E and c["n_block_waveforms"] == E * 1080
    assert c["n_present"] == int((ev["pres"] == 1).sum())
    assert c["n_pass_threshold"] == int(((a["status"] & 2) > 0).sum())
    assert c["n_pulses"] == int(a["wfnpulse"].sum())
    fitted = (a["status"] & 28) > 0
    assert c["n_fit_attempted"] == int(fitted.sum()) == c["n_fit_ok_first"] + c["n_fit_ok_retry"] + c["n_fallback"]
    assert np.array_equal(fitted, ((a["status"] & 2) > 0) & (a["wfnpulse"] > 0))
    b = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    for k in a:
        assert np.array_equal(a[k], b[k]), "non-deterministic " + k
    perm = np.random.default_rng(0).permutation(E)
    p = gpu.analyze(ev["signal"][perm], ev["pres"][perm], ev["corr_time_HMS"][perm])
    for k in a:
        assert np.array_equal(a[k][perm], p[k]), "not event-equivariant " + k
    z = gpu.analyze(np.zeros((0, 1080, 110)), np.zeros((0, 1080), np.int32), np.zeros(0))
    assert z["wfnpulse"].shape == (0, 1080)
    # all blocks absent -> nothing found, sentinels everywhere
    n0 = gpu.analyze(ev["signal"][:2], np.zeros((2, 1080), np.int32), ev["corr_time_HMS"][:2])
    assert (n0["wfnpulse"] == 0).all() and (n0["chi2"] == -100).all() and (n0["status"] == 0).all()


def test_device_entry_point_equals_host_entry_point(gpu, events):
    import torch
    ev = events[2]
    E = ev["signal"].shape[0]
    host = gpu.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    dev = torch.device("cuda:0")
    sig = torch.from_numpy(ev["signal"]).to(dev); pres = torch.from_numpy(ev["pres"]).to(dev)
    corr = torch.from_numpy(ev["corr_time_HMS"]).to(dev)
    o = dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev),
             wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             status=torch.empty((E, 1080), dtype=torch.uint8, device=dev))
    stream = torch.cuda.Stream()
    st = stream.cuda_stream
    gpu.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                       o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(),
                       o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)
    torch.cuda.synchronize()
    for k in host:
        assert np.array_equal(o[k].cpu().numpy(), host[k]), k


def _device_analyze(h, ev, E):
    import torch
    dev = torch.device("cuda:0")
    sig = torch.from_numpy(ev["signal"][:E]).to(dev); pres = torch.from_numpy(ev["pres"][:E]).to(dev)
    corr = torch.from_numpy(ev["corr_time_HMS"][:E]).to(dev)
    o = dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev),
             wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
             chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
             status=torch.empty((E, 1080), dtype=torch.uint8, device=dev))
    stream = torch.cuda.Stream()
    h.analyze_device(E, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                     o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(),
                     o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=stream.cuda_stream)
    h.sync_device(stream=stream.cuda_stream)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in o.items()}


def test_overlapped_chunks_equal_serialised(pkg, calib, spline):
    """Device path with several chunks: chunks on alternating internal streams + fit kernels on side streams
    (profiling off) give bit-identical outputs to the fully serialised order (profiling on)."""
    E = 500
    ev = synth.generate_host(synth.config_params(2, absent_frac=0.02), spline, calib, 70000, E, n_threads=8)
    h = pkg.NpsWf(calib, chunk_events=148)
    h.set_profiling(False)
    a = _device_analyze(h, ev, E)
    h.set_profiling(True)
    b = _device_analyze(h, ev, E)
    t = h.stage_times(reset=True)
    assert t["chunks"] == 4 and t["front_ms"] > 0 and t["search_ms"] > 0 and t["fit_ms"] > 0
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert ((a["status"] & 28) > 0).sum() > 0.8 * E * 1080


def test_thread_fit_handoff_reaches_the_same_minimum(pkg, calib, events, monkeypatch):
    """A fit that the thread-per-fit kernel hands to the sub-warp kernel after 2 tries ends where the same fit
    ends when the thread kernel is allowed to finish it: same status, same minimum to ~1e-6 bin."""
    ev = events[2]
    base = pkg.NpsWf(calib).analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    monkeypatch.setenv("NPSWF_FIT_THREAD_TRIES", "2")
    h2 = pkg.NpsWf(calib)
    monkeypatch.delenv("NPSWF_FIT_THREAD_TRIES")
    got = h2.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    assert np.array_equal(got["wfnpulse"], base["wfnpulse"])
    fit = ((base["status"] & 12) > 0) & (base["wfnpulse"] <= 3)
    assert np.array_equal(got["status"][fit], base["status"][fit])
    valid = (np.arange(12)[None, None, :] < base["wfnpulse"][..., None]) & fit[..., None]
    assert np.abs(got["wftime"][valid] - base["wftime"][valid]).max() < 4e-5          # ns (1e-5 bin)
    # (the thread kernel keeps the weights 1/Err in binary32, the sub-warp kernel in binary64: chi2 differs at 1e-7)
    assert (np.abs(got["chi2"][fit] - base["chi2"][fit]) <= 2e-6 * np.abs(base["chi2"][fit])).all()


def test_off_lattice_traces_are_fitted_unrounded(gpu, events):
    """Samples that are not exact in binary32 must not be rounded: such a trace takes the generic double
    path.  A 1e-13 mV perturbation (far below binary32 resolution) must move the fit by ~1e-13, not by ~1e-8."""
    ev = events[1]
    sig = ev["signal"].copy()
    base = gpu.analyze(sig, ev["pres"], ev["corr_time_HMS"])
    rng = np.random.default_rng(3)
    sig2 = sig + rng.uniform(0.5e-13, 1e-13, sig.shape)
    assert not np.array_equal(sig2.astype(np.float32).astype(np.float64), sig2)
    got = gpu.analyze(sig2, ev["pres"], ev["corr_time_HMS"])
    assert np.array_equal(got["wfnpulse"], base["wfnpulse"])
    fit = ((base["status"] & 12) > 0) & ((got["status"] & 12) > 0)
    assert fit.sum() > 3000
    valid = (np.arange(12)[None, None, :] < base["wfnpulse"][..., None]) & fit[..., None]
    assert np.abs(got["wftime"][valid] - base["wftime"][valid]).max() < 1e-6
    assert np.abs(got["wfampl"][valid] - base["wfampl"][valid]).max() < 1e-6


def test_fp64_peak_tap(gpu):
    g = gpu.fp64_peak_gflops()
    assert 5e3 < g < 1e5, g   # B200: ~3.4e4 GFLOP/s measured


def _malformed_streams(rng):
    """Packed streams exercising every branch of T2:855-889."""
    x = np.arange(110.0)
    def rec(slot, n, vals=None):
        v = rng.normal(0, 1, int(n)) if vals is None else np.asarray(vals, float)
        return np.concatenate([[slot, n], v])
    evs = []
    evs.append(np.concatenate([rec(3, 110), rec(2000, 110), rec(1079, 110), rec(2001, 110), rec(0, 110)]))      # scintillators
    evs.append(np.concatenate([rec(5, 110), rec(1104, 110), rec(6, 110)]))                                      # bad slot ends the event
    evs.append(np.concatenate([rec(5, 110), rec(-1, 110), rec(6, 110)]))                                        # negative slot
    evs.append(np.concatenate([rec(7, 110, x), rec(8, 110), rec(7, 110, -x)]))                                  # repeated slot: last wins
    evs.append(np.concatenate([rec(7, 110, x), rec(7, 40, -x[:40]), rec(9, 110)]))                              # repeated, shorter
    evs.append(np.concatenate([rec(10, 60), rec(11, 110), rec(12, 0), rec(13, 110)]))                           # nsamp != 110
    evs.append(np.concatenate([rec(1090, 110), rec(20, 110)]))                                                   # non-block slot
    evs.append(np.concatenate([rec(30, 110), rec(31, 110)])[:-37])                                               # truncated
    evs.append(np.zeros(0))                                                                                      # empty
    evs.append(np.concatenate([rec(s, 110) for s in range(1104)] + [rec(2000, 110)]))                            # > 1104*112 words: skipped
    evs.append(np.concatenate([rec(s, 110) for s in rng.permutation(1080)]))                                     # full event
    # Int_t truncation of the header words (T2:553, 857-859): -0.5 is slot 0, 2000.7 is the scintillator slot 2000
    evs.append(np.concatenate([rec(-0.5, 110), rec(2000.7, 110), rec(17.9, 110)]))
    evs.append(np.concatenate([rec(40, 110), rec(np.nan, 110), rec(41, 110)]))                                   # NaN slot ends the event
    evs.append(np.concatenate([rec(42, np.nan, []), rec(43, 3e9, []), rec(44, -7, []), rec(45, 110)]))           # nsamp not a count: no samples
    evs.append(np.concatenate([rec(s % 1080, 0, []) for s in range(2500)] + [rec(33, 110), rec(34, 5)]))         # > 1104 records
    evs.append(np.concatenate([rec(7, 20, x[:20] + 1)] + [rec(s % 50, 1, [s]) for s in range(1300)] + [rec(7, 10, -x[:10] - 1)]))  # repeats across rounds
    offs = np.concatenate([[0], np.cumsum([e.size for e in evs])]).astype(np.int64)
    return np.concatenate(evs), offs


def test_unpack_exact(gpu, events):
    """Waveform unpack (T2:851-889) on the device vs the oracle: well-formed shuffled streams and malformed ones."""
    ev = events[2]
    samp, offs = synth.pack_events(ev["signal"], ev["pres"], seed=5)
    sig, pres = gpu.unpack(samp, offs)
    assert np.array_equal(pres, ev["pres"])
    assert np.array_equal(sig, np.where(ev["pres"][..., None] == 1, ev["signal"], 0.0))
    samp, offs = _malformed_streams(np.random.default_rng(9))
    sig, pres = gpu.unpack(samp, offs)
    for e in range(offs.size - 1):
        rs, rp, _ = oracle.unpack_event(samp[offs[e]:offs[e + 1]])
        assert np.array_equal(pres[e], rp), e
        assert np.array_equal(sig[e], rs), e
    assert pres[9].sum() == 0 and pres[10].sum() == 1080 and pres[1].sum() == 1 and pres[6].sum() == 1
    assert pres[11, 0] == 1 and pres[11, 17] == 1 and pres[11].sum() == 2          # -0.5 -> slot 0, 2000.7 -> scintillator, 17.9 -> 17
    assert pres[12].sum() == 1 and pres[13, 45] == 1 and pres[14, 33] == 1 and pres[14, 34] == 1
    assert sig[15, 7, 0] == -1.0 and sig[15, 7, 10] == 11.0                         # the later record of slot 7 wins, 10 samples only


def test_analyze_packed_equals_analyze(gpu, events):
    ev = events[2]
    samp, offs = synth.pack_events(ev["signal"], ev["pres"], seed=6)
    a = gpu.analyze(np.where(ev["pres"][..., None] == 1, ev["signal"], 0.0), ev["pres"], ev["corr_time_HMS"])
    b = gpu.analyze_packed(samp, offs, ev["corr_time_HMS"])
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_event_diagnostics(gpu, events):
    """ampl / enertot / integtot of the WF tree (T2:1026-1056): ampl exact, the sums exact on the ADC lattice and
    within 1e-12 relative for arbitrary doubles."""
    ev = events[2]
    ampl, et, it = gpu.event_diagnostics(ev["signal"])
    for e in range(ev["signal"].shape[0]):
        ra, re, ri = oracle.event_diagnostics(ev["signal"][e])
        assert np.array_equal(ampl[e], ra) and et[e] == re and it[e] == ri
    rng = np.random.default_rng(4)
    sig = rng.normal(0, 50, (2, 1080, 110))
    sig[0, 5] = -500.0                      # below the -100 initial value of sigmax
    ampl, et, it = gpu.event_diagnostics(sig)
    for e in range(2):
        ra, re, ri = oracle.event_diagnostics(sig[e])
        assert np.array_equal(ampl[e], ra)
        scale = np.abs(sig[e]).sum()
        assert abs(et[e] - re) <= 1e-12 * scale and abs(it[e] - ri) <= 1e-12 * scale
    assert ampl[0, 5] == -100.0


def test_two_devices_in_one_process_equal_one_device(pkg, calib, spline):
    """n_devices = 2 (contiguous event ranges, one host thread per device inside the call) gives the same outputs
    as one device.  Skipped on a single-GPU box."""
    if pkg.lib().npswf_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    E = 700
    ev = synth.generate_host(synth.config_params(2, absent_frac=0.02), spline, calib, 90000, E, n_threads=8)
    one = pkg.NpsWf(calib, devices=[0]).analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    h2 = pkg.NpsWf(calib, devices=[0, 1])
    two = h2.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    for k in one:
        assert np.array_equal(one[k], two[k]), k
    c = h2.counters()
    assert c["n_events"] == E and c["n_fit_attempted"] == int(((one["status"] & 28) > 0).sum())
    assert h2.host_packing_stats()["packed_chunks"] >= 2          # each device's host thread packed its own range
    # flat outputs: every device fills its own share of the pools, offsets are absolute
    flat = h2.analyze_flat(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    assert flat["n_pulses"] == int(one["wfnpulse"].sum())
    for e in (0, E // 2 - 1, E // 2, E - 1):
        ft, fa, _ = pkg.flatten_event(one["wfnpulse"][e], one["wftime"][e], one["wfampl"][e])
        o, n = int(flat["pulse_offset"][e]), int(flat["pulse_count"][e])
        assert n == len(ft) and np.array_equal(flat["wftime_pool"][o:o + n], ft) and np.array_equal(flat["wfampl_pool"][o:o + n], fa)
    assert flat["pulse_offset"][E // 2] >= flat["wftime_pool"].size // 2       # the second device's share starts at the middle


def test_host_packing_is_lossless_and_falls_back(pkg, calib, events):
    """npswf_analyze_batch sends lattice chunks as int16 counts (host_pack.hpp) and everything else as the caller's
    doubles; outputs are bit-identical in every mode."""
    ev = events[2]
    sig, pres, corr = ev["signal"], ev["pres"], ev["corr_time_HMS"]
    h = pkg.NpsWf(calib)
    h.set_host_packing(0)
    raw = h.analyze(sig, pres, corr)
    assert h.host_packing_stats()["packed_chunks"] == 0
    h.set_host_packing(2, n_threads=3)
    packed = h.analyze(sig, pres, corr)
    st = h.host_packing_stats()
    assert st["packed_chunks"] >= 1 and st["raw_chunks"] == 0
    for k in raw:
        assert np.array_equal(raw[k], packed[k]), k
    assert (np.signbit(sig) & (sig == 0)).any()   # the sets hold -0.0 samples: they travel as +0.0, no output differs
    # one sample off the lattice / beyond int16 / tiny: the chunk goes over as doubles
    for bad in (0.1, 40000 * 1000.0 / 4096, 2.0 ** -30):
        s2 = sig.copy()
        s2[1, 500, 57] = bad
        h.set_host_packing(0)
        a = h.analyze(s2, pres, corr)
        before = h.host_packing_stats()
        h.set_host_packing(2)
        b = h.analyze(s2, pres, corr)
        after = h.host_packing_stats()
        assert after["raw_chunks"] == before["raw_chunks"] + 1 and after["packed_chunks"] == before["packed_chunks"], bad
        for k in a:
            assert np.array_equal(a[k], b[k]), (bad, k)
    # the one bit pattern the transport does not preserve is the sign of a zero sample.  Adversarial set: rectified
    # traces (their minimum IS a zero sample, so minsignal, the matched-filter differences and the 3x3 sums all see
    # zeros of either sign), signs of the zeros drawn at random -- every output must still be bit-identical
    rng = np.random.default_rng(11)
    s4 = np.abs(sig)
    z = s4 == 0
    s4[z] = np.where(rng.random(int(z.sum())) < 0.5, -0.0, 0.0)
    assert np.signbit(s4[z]).any() and (~np.signbit(s4[z])).any() and (s4.min(axis=-1) == 0).mean() > 0.5
    h.set_host_packing(0)
    a = h.analyze(s4, pres, corr)
    h.set_host_packing(2)
    b = h.analyze(s4, pres, corr)
    assert (a["wfnpulse"] > 0).mean() > 0.3
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(np.signbit(a[k]), np.signbit(b[k])), k
    # another lattice: counts * 0.5 mV
    h.set_host_packing(2, lsb_mV=0.5)
    s3 = np.round(sig * 2.0) / 2.0
    before = h.host_packing_stats()
    c = h.analyze(s3, pres, corr)
    assert h.host_packing_stats()["packed_chunks"] == before["packed_chunks"] + 1
    h.set_host_packing(0)
    d = h.analyze(s3, pres, corr)
    for k in c:
        assert np.array_equal(c[k], d[k]), k


def test_flat_outputs_equal_reference_packing(pkg, calib, events):
    """npswf_analyze_batch_flat: wfampl / wftime packed on the device exactly as the reference packs them
    (T2:959-961, 1289-1296), for one chunk and for many small chunks (deferred pulse copies)."""
    for cfg in (2, 3):
        ev = events[cfg]
        sig, pres, corr = ev["signal"], ev["pres"], ev["corr_time_HMS"]
        E = sig.shape[0]
        for chunk in (0, 1):   # 0: default chunking (one chunk here); 1: the smallest chunks the library makes
            h = pkg.NpsWf(calib, chunk_events=chunk)
            pad = h.analyze(sig, pres, corr)
            flat = h.analyze_flat(sig, pres, corr)
            for k in ("wfnpulse", "chi2", "timewf", "amplwf", "status"):
                assert np.array_equal(pad[k], flat[k]), k
            assert flat["n_pulses"] == int(pad["wfnpulse"].sum())
            assert np.array_equal(flat["pulse_count"], pad["wfnpulse"].sum(axis=1))
            for e in range(E):
                ft, fa, boff = pkg.flatten_event(pad["wfnpulse"][e], pad["wftime"][e], pad["wfampl"][e])
                o, c = int(flat["pulse_offset"][e]), int(flat["pulse_count"][e])
                assert c == len(ft)
                assert np.array_equal(flat["wftime_pool"][o:o + c], ft)
                assert np.array_equal(flat["wfampl_pool"][o:o + c], fa)
            # exact-size pool works, one pulse less is refused with a message
            need = flat["n_pulses"]
            tight = h.analyze_flat(sig, pres, corr, capacity=need)
            assert np.array_equal(tight["wftime_pool"], flat["wftime_pool"][:need])
            with pytest.raises(pkg.NpsWfError) as ei:
                h.analyze_flat(sig, pres, corr, capacity=need - 1)
            assert "pool too small" in str(ei.value)
            again = h.analyze_flat(sig, pres, corr)     # the handle is usable after the refusal
            assert np.array_equal(again["wftime_pool"][:need], flat["wftime_pool"][:need])
    # no block present: empty vectors
    z = h.analyze_flat(sig[:2], np.zeros((2, 1080), np.int32), corr[:2])
    assert z["n_pulses"] == 0 and (z["pulse_count"] == 0).all()


def test_cpp_host_mirror_on_gpu(tmp_path, pkg):
    """The C++ mirror (include/npswf_host.hpp): analyze() through the device-packed flat outputs equals the padded
    C entry point + npswf_flatten_event (tests/cpp/host_mirror_smoke.cpp, GPU branch)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_mirror_smoke")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "host_mirror_smoke.cpp"), "-o", exe,
                           "-L", libdir, "-lnpswf", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "gpu present" in r.stdout, r.stdout + r.stderr


def test_blocks_without_reference_waveform(pkg, calib, events):
    """preswf == 0 blocks (no ref_wf file, T2:452-456): never analysed (T2:944), but their samples still enter the
    3x3 sums of their neighbours, whose gate is pres alone (T2:257)."""
    cal = dict(calib)
    rng = np.random.default_rng(5)
    preswf = np.ones(1080, np.int32)
    preswf[rng.choice(1080, 60, replace=False)] = 0
    cal["preswf"] = preswf
    o = oracle.Oracle(cal)
    h = pkg.NpsWf(cal)
    ev = events[2]
    ref = o.analyze_batch(ev["signal"], ev["pres"], ev["corr_time_HMS"], n_threads=8)
    got = h.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    off = preswf == 0
    assert (got["wfnpulse"][:, off] == 0).all() and (got["chi2"][:, off] == -100).all() and (got["status"][:, off] == 0).all()
    assert np.array_equal(got["wfnpulse"], ref["wfnpulse"])
    assert np.array_equal(got["status"] & 3, ref["status"] & 3)
    nofit = (ref["status"] & 28) == 0
    for k in ("wftime", "wfampl", "chi2", "timewf", "amplwf"):
        assert np.array_equal(got[k][nofit], ref[k][nofit]), k
    n_both, frac, _, _ = _fit_agreement(ref, got)
    assert n_both > 1000 and frac >= 0.99
    # a 6 mV pulse next to an 8 mV pulse in a switched-off block: passes the 10 mV threshold only through the
    # neighbour's samples, which count as long as the neighbour is present in the data (pres), preswf or not
    bx = int(np.flatnonzero(off)[np.flatnonzero(off) % 30 > 0][0])    # a switched-off block with a left neighbour
    lsb = synth.LSB
    sig = np.zeros((1, 1080, 110))
    sig[0, bx - 1] = np.round(6.0 * calib["interpY"][bx - 1] / lsb) * lsb
    sig[0, bx] = np.round(8.0 * calib["interpY"][bx] / lsb) * lsb
    pres1 = np.zeros((1, 1080), np.int32)
    pres1[0, bx - 1] = pres1[0, bx] = 1
    for pres_nb, expect in ((1, 2), (0, 0)):
        pres1[0, bx] = pres_nb
        r = o.analyze_batch(sig, pres1, np.zeros(1), n_threads=1)
        g = h.analyze(sig, pres1, np.zeros(1))
        assert (r["status"][0, bx - 1] & 2) == expect and (g["status"][0, bx - 1] & 2) == expect
        assert g["status"][0, bx] == 0 and g["wfnpulse"][0, bx] == 0
        assert np.array_equal(g["wfnpulse"], r["wfnpulse"])


def test_full_size_configs_by_properties(pkg, calib):
    """BASELINE configs[1] (100 000 events, 1-3 pulses) and configs[2] (10 000 events, up to 12 pulses) at full size,
    generated on the device in slices: the bookkeeping identities hold, the results do not depend on how the events
    are cut into chunks (bitwise, whole output arrays compared on the device), and the rates that the configurations
    are designed to produce are there (every block fitted in config 2; peak-buffer-full and fall-backs in config 3)."""
    import torch
    dev = torch.device("cuda:0")
    ha = pkg.NpsWf(calib)                      # default chunks (1 184 events)
    hb = pkg.NpsWf(calib, chunk_events=444)    # other chunking
    stream = torch.cuda.Stream()
    st = stream.cuda_stream
    d_spl = torch.from_numpy(ha.spline_coeffs()).to(dev)
    d_tref = torch.from_numpy(calib["timeref"]).to(dev)
    d_kap = torch.from_numpy(calib["kappa"]).to(dev)
    names = (("wfnpulse", torch.int32, ()), ("wftime", torch.float64, (12,)), ("wfampl", torch.float64, (12,)),
             ("chi2", torch.float64, ()), ("timewf", torch.float64, ()), ("amplwf", torch.float64, ()),
             ("status", torch.uint8, ()))
    for cfg, total, slice_events in ((2, 100_000, 12_500), (3, 10_000, 5_000)):
        sig = torch.empty((slice_events, 1080, 110), dtype=torch.float64, device=dev)
        pres = torch.empty((slice_events, 1080), dtype=torch.int32, device=dev)
        corr = torch.empty((slice_events,), dtype=torch.float64, device=dev)
        oa = {k: torch.empty((slice_events, 1080) + sh, dtype=dt, device=dev) for k, dt, sh in names}
        ob = {k: torch.empty((slice_events, 1080) + sh, dtype=dt, device=dev) for k, dt, sh in names}
        ha.reset_counters()
        n_fitted = n_pulses = n_pass = 0
        for first in range(0, total, slice_events):
            synth.generate_device(synth.config_params(cfg), d_spl.data_ptr(), d_tref.data_ptr(), d_kap.data_ptr(), first,
                                  slice_events, sig.data_ptr(), 0, pres.data_ptr(), corr.data_ptr(), st)
            for h, o in ((ha, oa), (hb, ob)):
                h.analyze_device(slice_events, sig.data_ptr(), pres.data_ptr(), corr.data_ptr(), o["wfnpulse"].data_ptr(),
                                 o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(),
                                 o["timewf"].data_ptr(), o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=st)
                h.sync_device(stream=st)
            torch.cuda.synchronize()
            for k, _, _ in names:
                assert torch.equal(oa[k], ob[k]), "cfg %d slice %d: %s depends on the chunking" % (cfg, first, k)
            fitted = (oa["status"] & 28) > 0
            assert torch.equal(fitted, ((oa["status"] & 2) > 0) & (oa["wfnpulse"] > 0))
            assert bool(((oa["chi2"] == -100.0) == ((oa["status"] & 12) == 0)).all())     # chi2 sentinel <=> no converged fit
            assert int(oa["wfnpulse"].max()) <= 12 and int(oa["wfnpulse"].min()) >= 0
            n_fitted += int(fitted.sum()); n_pulses += int(oa["wfnpulse"].sum()); n_pass += int(((oa["status"] & 2) > 0).sum())
        c = ha.counters()
        assert c["n_block_waveforms"] == total * 1080 and c["n_present"] == total * 1080
        assert c["n_fit_attempted"] == n_fitted == c["n_fit_ok_first"] + c["n_fit_ok_retry"] + c["n_fallback"]
        assert c["n_pulses"] == n_pulses and c["n_pass_threshold"] == n_pass
        print("cfg%d full size: %d events, fitted %.4f of the blocks, %.3f pulses per fitted block, fall-backs %d, "
              "peak buffer full %d" % (cfg, total, n_fitted / (total * 1080.0), n_pulses / max(1, n_fitted),
                                       c["n_fallback"], c["n_peak_buffer_full"]))
        if cfg == 2:
            assert n_fitted > 0.999 * total * 1080 and c["n_fallback"] < 1e-3 * n_fitted
        else:
            assert c["n_peak_buffer_full"] > 0 and c["n_fallback"] + c["n_fit_ok_retry"] > 0
        del sig, pres, corr, oa, ob
        torch.cuda.empty_cache()


def test_flat_i16_entry_point(gpu, events):
    """npswf_analyze_batch_flat_i16 (int16 counts in, reference packing out) equals the binary64 flat call."""
    ev = events[2]
    a = gpu.analyze_flat(ev["signal"], ev["pres"], ev["corr_time_HMS"])
    b = gpu.analyze_flat_i16(ev["counts"], synth.LSB, ev["pres"], ev["corr_time_HMS"])
    assert a["n_pulses"] == b["n_pulses"] > 0
    n = a["n_pulses"]
    for k in ("wfnpulse", "pulse_offset", "pulse_count", "chi2", "timewf", "amplwf", "status"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["wftime_pool"][:n], b["wftime_pool"][:n]) and np.array_equal(a["wfampl_pool"][:n], b["wfampl_pool"][:n])


def test_device_calls_on_different_streams_are_ordered(pkg, calib, spline):
    """Two device-path calls enqueued back to back on DIFFERENT streams share the handle's scratch, job lists and fit
    streams: the second one orders itself behind the first (last-use event), so both give what they give alone; a
    host-buffer call right after them waits for them too."""
    import torch
    E = 400
    ev = [synth.generate_host(synth.config_params(2, absent_frac=0.02), spline, calib, 95_000 + 1000 * i, E, n_threads=8) for i in range(2)]
    h = pkg.NpsWf(calib)
    alone = [_device_analyze(h, ev[i], E) for i in range(2)]
    dev = torch.device("cuda:0")
    ins, outs, streams = [], [], [torch.cuda.Stream(), torch.cuda.Stream()]
    for i in range(2):
        ins.append((torch.from_numpy(ev[i]["signal"]).to(dev), torch.from_numpy(ev[i]["pres"]).to(dev),
                    torch.from_numpy(ev[i]["corr_time_HMS"]).to(dev)))
        outs.append(dict(wfnpulse=torch.empty((E, 1080), dtype=torch.int32, device=dev),
                         wftime=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
                         wfampl=torch.empty((E, 1080, 12), dtype=torch.float64, device=dev),
                         chi2=torch.empty((E, 1080), dtype=torch.float64, device=dev),
                         timewf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
                         amplwf=torch.empty((E, 1080), dtype=torch.float64, device=dev),
                         status=torch.empty((E, 1080), dtype=torch.uint8, device=dev)))
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(2):
            o = outs[i]
            h.analyze_device(E, ins[i][0].data_ptr(), ins[i][1].data_ptr(), ins[i][2].data_ptr(), o["wfnpulse"].data_ptr(),
                             o["wftime"].data_ptr(), o["wfampl"].data_ptr(), o["chi2"].data_ptr(), o["timewf"].data_ptr(),
                             o["amplwf"].data_ptr(), o["status"].data_ptr(), stream=streams[i].cuda_stream)
        host = h.analyze(ev[0]["signal"][:50], ev[0]["pres"][:50], ev[0]["corr_time_HMS"][:50])   # no explicit sync before it
        torch.cuda.synchronize()
        for i in range(2):
            for k in alone[i]:
                assert np.array_equal(outs[i][k].cpu().numpy(), alone[i][k]), (rep, i, k)
        for k in host:
            assert np.array_equal(host[k], alone[0][k][:50]), (rep, k)


def test_non_finite_samples_end_as_fallback(pkg, gpu, orc, calib, events):
    """A NaN / Inf sample inside the fit window makes chi2 not a number: no descent step exists, and the fit must end
    as a failed fit (TSpectrum values, chi2 = -100, T2:774-791) -- not as a 'converged' fit with a NaN chi2.  Driven
    through the stage-level Fitwf entry point with the seeds of the clean traces."""
    ev = events[1]
    sig0, pres, corr = ev["signal"][:1], ev["pres"][:1], ev["corr_time_HMS"][:1]
    n, t, a = gpu.FindPulsesMF(sig0, pres)
    ok = gpu.PassClusterThreshold(sig0, pres)
    blocks = np.flatnonzero(ok[0] & (n[0] > 0))[:40]
    assert blocks.size >= 20
    sig = sig0.copy()
    for j, b in enumerate(blocks):
        sig[0, b, 60 + j % 30] = np.nan if j % 2 == 0 else np.inf
    mask = np.zeros((1, 1080), np.uint8)
    mask[0, blocks] = 1
    r = gpu.Fitwf(sig, corr, mask, n, t, a)        # FAST (LM) mode: a failed fit
    for b in blocks:
        assert (r["status"][0, b] & 28) == 16 and r["chi2"][0, b] == -100.0, (b, r["status"][0, b], r["chi2"][0, b])
        assert np.array_equal(r["wfampl"][0, b, :n[0, b]], a[0, b, :n[0, b]])
    # MIGRAD mode follows the Migrad restatement wherever it goes (a NaN EDM ends VariableMetricBuilder with the
    # current state): verdict and values equal the oracle's, NaN for NaN
    hm = pkg.NpsWf(calib, fit_mode=pkg.FIT_MIGRAD)
    rm = hm.Fitwf(sig, corr, mask, n, t, a)
    for b in blocks[:12]:
        o = orc.fitwf(b, sig[0], n[0, b], t[0, b], a[0, b], corr[0])
        assert (rm["status"][0, b] & 28) == o["status"], (b, rm["status"][0, b], o["status"])
        assert np.array_equal(rm["chi2"][0, b], o["chi2"], equal_nan=True)
        assert np.array_equal(rm["wftime"][0, b, :n[0, b]], o["wftime"][:n[0, b]], equal_nan=True)


def test_event_times_h1time_h2time(pkg, gpu, orc, calib, events):
    """h1time / h2time (T2:988-996) from the analysis outputs (npswf_event_times) against the oracle, which reads the
    fit parameters as the reference does.  MIGRAD mode: h2time exact, h1time = (wftime + cortime) / dt within 1e-12
    relative of parameter - timerefacc + corr/dt (one rounding of the round trip through the corrected time)."""
    hm = pkg.NpsWf(calib, fit_mode=pkg.FIT_MIGRAD)
    n_seen = 0
    for cfg in (2, 3):
        ev = events[cfg]
        got = hm.analyze(ev["signal"], ev["pres"], ev["corr_time_HMS"])
        for e in range(ev["signal"].shape[0]):
            h1, h2 = pkg.event_times(got["wfnpulse"][e], got["wftime"][e], got["wfampl"][e], got["status"][e], calib["cortime"])
            r1, r2 = orc.event_times(ev["signal"][e], ev["pres"][e], ev["corr_time_HMS"][e])
            assert np.array_equal(h2, r2), (cfg, e)
            assert h1.shape == r1.shape and np.allclose(h1, r1, rtol=1e-12, atol=1e-12), (cfg, e, np.abs(h1 - r1).max())
            n_seen += h1.size
    assert n_seen > 1000
