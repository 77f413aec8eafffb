"""Builds libnpswf.so (the C-ABI product library) and the device-side synthetic generator for sm_100a.

nvcc cross-compiles without a GPU; the .so files are git-ignored but travel to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))] + [os.path.join(ROOT, "include", "npswf.h")]
    out = os.path.join(HERE, "lib", "libnpswf.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if force or _stale(out, srcs):
        # two device translation units: npswf_api.cu (default FMA contraction; its bit-exact parts use explicit
        # __dmul_rn / __dadd_rn) and npswf_migrad.cu, whose whole arithmetic follows an FMA-free x86-64 build of
        # Minuit2 (-fmad=false).  Compiled side by side, then linked with the host packer.
        obj = os.path.join(HERE, "lib", "obj")
        os.makedirs(obj, exist_ok=True)
        vflag = ["-Xptxas", "-v"] if verbose else []
        units = [("npswf_api.cu", []), ("npswf_migrad.cu", ["-fmad=false"]), ("host_pack.cpp", []), ("host_event.cpp", [])]
        hdr = os.path.join(ROOT, "include", "npswf.h")
        migrad_only = {"npswf_migrad.cu", "kernel_fit_migrad.cuh", "migrad_core.hpp", "host_event.cpp"}
        deps = {
            "npswf_migrad.cu": [os.path.join(csrc, f) for f in ("npswf_migrad.cu", "kernel_fit_migrad.cuh", "migrad_core.hpp",
                                                                "migrad_launch.hpp", "common.cuh")] + [hdr],
            "host_pack.cpp": [os.path.join(csrc, f) for f in ("host_pack.cpp", "host_pack.hpp")],
            "host_event.cpp": [os.path.join(csrc, "host_event.cpp"), hdr],
            "npswf_api.cu": [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f not in migrad_only] + [hdr],
        }
        procs = []
        for name, extra in units:
            o = os.path.join(obj, os.path.splitext(name)[0] + ".o")
            if force or _stale(o, deps[name]):
                cmd = [NVCC] + ARCH + [f for f in COMMON if f != "-shared"] + vflag + extra + ["-c", "-o", o, os.path.join(csrc, name)]
                procs.append((cmd, subprocess.Popen(cmd)))
        for cmd, pr in procs:
            if pr.wait() != 0:
                raise subprocess.CalledProcessError(pr.returncode, cmd)
        objs = [os.path.join(obj, os.path.splitext(name)[0] + ".o") for name, _ in units]
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", out] + objs)
    syn = os.path.join(ROOT, "synth")
    sout = os.path.join(syn, "libnpswf_synth_cuda.so")
    ssrcs = [os.path.join(syn, "synth_cuda.cu"), os.path.join(syn, "npswf_synth.h")]
    if force or _stale(sout, ssrcs):
        subprocess.check_call([NVCC] + ARCH + COMMON + ["-o", sout, ssrcs[0]])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
