"""A/B of host packer builds on one box: packing rate of pack_counts() with 16 threads over a 1.2 GB lattice array."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("nps-waveform-analysis_b200")
import ctypes as C
n = 1184 * 1080 * 110
lsb = 1000.0 / 4096
x = pkg.pinned_empty((n,), np.float64)
x[:] = (np.arange(n, dtype=np.int64) * 2654435761 % 4000 - 500).astype(np.float64) * lsb
out = pkg.pinned_empty((n,), np.int16)
L = pkg.lib()
for nt in (8, 16):
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        rc = L.npswf_debug_pack_counts(x.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_double(lsb), C.c_int32(nt), out.ctypes.data_as(C.c_void_p))
        ts.append(time.perf_counter() - t0)
    print(os.environ.get("NPSWF_LIB", "default").split("/")[-1], "threads", nt, "rc", rc, "best %.1f GB/s median %.1f GB/s" % (n * 8 / min(ts) / 1e9, n * 8 / sorted(ts)[3] / 1e9), flush=True)
