"""Host-side callers of the hot path: npswf_hcana_pulses (HMS time correction + hcana pulse
selection, T2:893-939) against the oracle's restatement.  No GPU needed: plain host code of libnpswf.so."""
import numpy as np

import oracle


def _event(rng, n, with_scint=True):
    counter = rng.integers(0, 1080, n).astype(np.float64)
    if with_scint and n > 3:
        counter[rng.integers(1, n)] = 2000.0
        counter[rng.integers(1, n)] = 2001.0
    # repeated blocks: several hcana pulses in one block (pile-up), the one closest to timemean2 must win
    if n > 6:
        counter[n // 2] = counter[0]
        counter[n // 2 + 1] = counter[0]
        counter[n - 1] = counter[1]
    t = rng.uniform(100.0, 240.0, n)
    traw = rng.uniform(0.0, 4096.0, n)
    amp = rng.uniform(1.0, 500.0, n)
    return counter, t, traw, amp


def test_hcana_pulses_equal_oracle(pkg):
    rng = np.random.default_rng(12)
    tdc = rng.uniform(-30, 30, 1080).astype(np.float32)
    for acc in (0.0, -5.0):
        timemean2 = np.full(1080, 170 + acc * 4.0, np.float32)          # T2:526-529
        for n in (0, 1, 2, 7, 40, 300, 2000):
            c, t, traw, amp = _event(rng, n)
            got = pkg.hcana_pulses(c, t, traw, amp, tdc, timemean2)
            ref = oracle.hcana_pulses(c, t, traw, amp, tdc, timemean2)
            assert got[0] == ref[0], (acc, n)
            assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]), (acc, n)
            if n == 0:
                assert got[0] == 0.0 and (got[1] == -100).all() and (got[2] == -100).all()
            else:
                assert got[0] == t[0] - traw[0] / 16. - np.float64(tdc[int(c[0])])


def test_hcana_pulses_closest_to_expected_time_wins(pkg):
    tdc = np.zeros(1080, np.float32)
    tm = np.full(1080, 170.0, np.float32)
    c = np.array([5, 5, 5, 9, 9.0])
    t = np.array([150.0, 171.0, 160.0, 169.0, 100.0])
    amp = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    corr, sa, st = pkg.hcana_pulses(c, t, np.zeros(5), amp, tdc, tm)
    assert sa[5] == 2.0 and st[5] == 171.0 and sa[9] == 4.0 and st[9] == 169.0
    assert (np.delete(sa, [5, 9]) == -100).all()
    # a first pulse from a scintillator channel: the reference reads tdcoffset out of bounds; offset 0 here
    corr, _, _ = pkg.hcana_pulses(np.array([2000.0]), np.array([200.0]), np.array([160.0]), np.array([3.0]), tdc + 7, tm)
    assert corr == 200.0 - 10.0
