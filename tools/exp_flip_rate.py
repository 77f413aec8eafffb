#!/usr/bin/env python3
"""How much does "exact vs the oracle" say about "exact vs ROOT"?  Two things in the oracle's TSpectrum restatement
are known NOT to be bit-identical to a ROOT build: (a) the Markov step uses a deterministic exp shared with the CUDA
kernel instead of glibc's exp (they differ in the last bit for some arguments); (b) the oracle is compiled without
FMA contraction, while a ROOT built with -march=native may contract.  This tool runs the peak search (FindPulsesMF:
matched filter -> float histogram -> SearchHighRes -> peak filter) over >= 10^6 spectra of the three BASELINE
configurations under all three variants and counts the block-waveforms whose peak list (count, positions, order,
amplitudes) differs from the checker's.  CPU only; test infrastructure.

Usage: python tools/exp_flip_rate.py [spectra_per_config=340000] > profiles/r2_exp_fma_flip_rate.txt"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402


def main():
    per_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 340_000
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "fma"])
    fma_lib = os.path.join(ROOT, "oracle", "_build", "libnpswf_oracle_fma.so")
    cal = synth.make_calibration()
    base = oracle.Oracle(cal)                                   # the checker: det_exp, no contraction
    libm = oracle.Oracle(cal, flags=oracle.FLAG_LIBM_EXP)       # glibc exp in the Markov step
    fma = oracle.Oracle(cal, flags=oracle.FLAG_LIBM_EXP, lib_path=fma_lib)   # glibc exp + FMA contraction everywhere
    spl = base.spline_coeffs()
    threads = os.cpu_count() or 1
    print("variant A: glibc exp instead of the deterministic exp; variant B: glibc exp and -ffp-contract=fast -mfma (whole oracle)")
    tot = dict(n=0, a=0, b=0, a_count=0, b_count=0)
    for cfg in (1, 2, 3):
        n_ev = (per_cfg + 1079) // 1080
        n_spec = n_a = n_b = n_ac = n_bc = n_pulses = 0
        for first in range(0, n_ev, 64):
            n = min(64, n_ev - first)
            ev = synth.generate_host(synth.config_params(cfg), spl, cal, 60_000_000 + 1_000_000 * cfg + first, n, n_threads=threads)
            r0 = base.find_pulses_batch(ev["signal"], ev["pres"], n_threads=threads)
            ra = libm.find_pulses_batch(ev["signal"], ev["pres"], n_threads=threads)
            rb = fma.find_pulses_batch(ev["signal"], ev["pres"], n_threads=threads)
            da = (r0[0] != ra[0]) | (r0[1] != ra[1]).any(-1) | (r0[2] != ra[2]).any(-1)
            db = (r0[0] != rb[0]) | (r0[1] != rb[1]).any(-1) | (r0[2] != rb[2]).any(-1)
            n_spec += da.size; n_a += int(da.sum()); n_b += int(db.sum())
            n_ac += int((r0[0] != ra[0]).sum()); n_bc += int((r0[0] != rb[0]).sum()); n_pulses += int(r0[0].sum())
        print("config %d: %8d spectra, %9d pulses | peak lists differing: A %d (count differs %d)  B %d (count differs %d)" % (
            cfg, n_spec, n_pulses, n_a, n_ac, n_b, n_bc))
        tot["n"] += n_spec; tot["a"] += n_a; tot["b"] += n_b; tot["a_count"] += n_ac; tot["b_count"] += n_bc
    print("total: %d spectra | A: %d differ (%.3g per spectrum) | B: %d differ (%.3g per spectrum)" % (
        tot["n"], tot["a"], tot["a"] / tot["n"], tot["b"], tot["b"] / tot["n"]))


if __name__ == "__main__":
    main()
