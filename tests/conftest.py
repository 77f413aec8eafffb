import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_pkg():
    return importlib.import_module("nps-waveform-analysis_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def calib():
    import synth
    return synth.make_calibration()


@pytest.fixture(scope="session")
def orc(calib):
    import oracle
    return oracle.Oracle(calib)


@pytest.fixture(scope="session")
def spline(orc):
    return orc.spline_coeffs()


@pytest.fixture(scope="session")
def events(calib, spline):
    """Small seeded event sets per BASELINE config (oracle finishes each in seconds)."""
    import synth
    out = {}
    for cfg, n, absent in ((1, 3, 0.0), (2, 3, 0.05), (3, 3, 0.02)):
        out[cfg] = synth.generate_host(synth.config_params(cfg, absent_frac=absent), spline, calib, 100 * cfg, n,
                                       n_threads=4, counts=True, truth=True)
    return out


@pytest.fixture(scope="session")
def gpu(pkg, calib):
    """Product handle on cuda:0 through the C ABI."""
    h = pkg.NpsWf(calib)
    if pkg.lib().npswf_device_count() < 1:
        pytest.fail("GPU test selected but libnpswf.so sees no CUDA device (no CPU fallback exists)")
    return h


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", name)
