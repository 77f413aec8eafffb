"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/npswf.h declares,
derives the same calibration as the oracle, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "npswf.h")).read()
    declared = sorted(set(re.findall(r"\b(npswf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), "libnpswf.so does not export %s" % name
    assert sorted(pkg.EXPORTS) == declared


def test_struct_layout_matches_header(pkg):
    assert C.sizeof(pkg.NpsWfConfig) == 80
    assert C.sizeof(pkg.NpsWfCalib) == 40
    assert C.sizeof(pkg.NpsWfCounters) == 96


def test_derived_calibration_matches_oracle(pkg, calib, orc):
    h = pkg.NpsWf(calib)
    y, i = h.mf_calib()
    oy, oi = orc.mf_calib()
    assert np.array_equal(y, oy) and np.array_equal(i, oi)      # T2:440-451, bitwise
    assert np.abs(h.spline_coeffs() - orc.spline_coeffs()).max() < 1e-14


def test_bad_calibration_is_rejected(pkg, calib):
    bad = dict(calib)
    bad["interpX"] = calib["interpX"] * 0.5
    with pytest.raises(pkg.NpsWfError) as ei:
        pkg.NpsWf(bad)
    assert ei.value.code == pkg.ERR_CALIB


def test_no_cpu_fallback(pkg, calib):
    """Without a usable GPU every compute entry point must fail loudly with NPSWF_ERR_CUDA."""
    if pkg.lib().npswf_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    h = pkg.NpsWf(calib)
    sig = np.zeros((1, 1080, 110)); pres = np.ones((1, 1080), np.int32)
    for call in (lambda: h.analyze(sig, pres, np.zeros(1)), lambda: h.FindPulsesMF(sig, pres),
                 lambda: h.PassClusterThreshold(sig, pres), lambda: h.matched_filter(sig, pres),
                 lambda: h.tspectrum_debug(np.zeros((1, 110), np.float32)), lambda: h.search_fused(),
                 lambda: h.debug_exact_ops(1000)):
        with pytest.raises(pkg.NpsWfError) as ei:
            call()
        assert ei.value.code == pkg.ERR_CUDA


def test_flatten_event_matches_reference_packing(pkg):
    """blockOffset / flattened wfampl, wftime as at T2:959-961, 1022, 1294-1295."""
    rng = np.random.default_rng(5)
    n = rng.integers(0, 5, 1080).astype(np.int32)
    n[rng.random(1080) < 0.5] = 0
    t = rng.normal(size=(1080, 12)); a = rng.normal(size=(1080, 12))
    tf, af, off = pkg.flatten_event(n, t, a)
    assert off[0] == 0 and off[-1] == n.sum() and np.array_equal(np.diff(off), n)
    exp_t = np.concatenate([t[b, :n[b]] for b in range(1080)])
    exp_a = np.concatenate([a[b, :n[b]] for b in range(1080)])
    assert np.array_equal(tf, exp_t) and np.array_equal(af, exp_a)
    tf0, af0, off0 = pkg.flatten_event(np.zeros(1080, np.int32), t, a)   # empty event
    assert tf0.size == 0 and off0[-1] == 0


def test_shard_range_partitions_events(pkg):
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            r = [pkg.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_synth_lattice_and_determinism(calib, spline):
    import synth
    p = synth.config_params(2, absent_frac=0.1)
    a = synth.generate_host(p, spline, calib, 5, 2, n_threads=2, counts=True)
    b = synth.generate_host(p, spline, calib, 6, 1, n_threads=1, counts=True)
    assert np.array_equal(a["signal"][1], b["signal"][0])               # keyed by (seed, event, block)
    assert np.array_equal(a["counts"] * synth.LSB, a["signal"])         # exact on the ADC lattice
    assert (a["signal"][a["pres"] == 0] == 0).all() and 0.05 < (a["pres"] == 0).mean() < 0.15


def test_cpp_host_mirror_compiles_and_refuses_cpu(tmp_path, pkg):
    """include/npswf_host.hpp (the C++ mirror a ROOT macro would include) builds against libnpswf.so."""
    import subprocess
    exe = str(tmp_path / "host_mirror_smoke")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_smoke.cpp"), "-o", exe,
                           "-L", libdir, "-lnpswf", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_host_packer_accepts_exactly_the_lattice(pkg, calib, spline):
    """The lossless int16 transport (csrc/host_pack.*): a chunk is packed iff every double is count * lsb."""
    import synth
    ev = synth.generate_host(synth.config_params(2), spline, calib, 77, 2, n_threads=2, counts=True)
    sig = ev["signal"]
    for nt in (1, 3, 8):
        ok, k = pkg.pack_counts(sig, synth.LSB, n_threads=nt)
        assert ok and np.array_equal(k, ev["counts"])
        assert np.array_equal(k.astype(np.float64) * synth.LSB, sig)      # what the device computes from the counts
    lsb = synth.LSB
    edge = np.array([32767 * lsb, -32767 * lsb, 0.0, -0.0, lsb, -lsb] * 3)   # odd length: vector body + scalar tail
    ok, k = pkg.pack_counts(edge, lsb)
    assert ok and k[0] == 32767 and k[1] == -32767 and k[2] == 0 and k[3] == 0
    for bad in (0.1, 32768 * lsb, -32768 * lsb, np.nan, np.inf, -np.inf, 2.0 ** -40, lsb * (1 + 2.0 ** -52), 1e300):
        for pos in (0, 5, len(edge) - 1):
            x = edge.copy()
            x[pos] = bad
            assert not pkg.pack_counts(x, lsb)[0], (bad, pos)
    big = np.tile(sig.reshape(-1), 3)
    big[big.size // 2 + 12345] += 2.0 ** -20
    assert not pkg.pack_counts(big, lsb, n_threads=4)[0]
    # other lattices
    assert pkg.pack_counts(np.arange(-50, 50) * 0.5, 0.5)[0]
    assert not pkg.pack_counts(np.arange(-50, 50) * 0.25, 0.5)[0]
    assert pkg.pack_counts(np.zeros(0), lsb)[0]


def test_root_macro_typechecks_against_host_mirror():
    """macros/npsWF_gpu.C cannot run here (no ROOT), but it must at least compile against include/npswf_host.hpp:
    type-checked with stub ROOT headers (tests/cpp/mock_root), and it must fill the 17 WF columns of T2:1387."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(root, "tests", "cpp", "mock_root"),
                           "-I", os.path.join(root, "include"), "-x", "c++", os.path.join(root, "macros", "npsWF_gpu.C")])
    src = open(os.path.join(root, "macros", "npsWF_gpu.C")).read()
    for col in ("chi2", "ampl", "amplwf", "wfnpulse", "Sampampl", "Samptime", "timewf", "enertot", "integtot", "pres",
                "corr_time_HMS", "h1time", "h2time", "runnum", "evt", "wfampl", "wftime"):
        assert 'Branch("%s"' % col in src, col
