"""Builds the C++ host side above the C ABI: cbench_mini (the CBench scalar loop driving the plugin
classes) linked against ../libzfb.so.  Called from __graft_entry__.build()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
EXE = os.path.join(HERE, "cbench_mini")
SHIM = os.path.join(PKG, "libzfb_cbench.so")


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("cbench_mini.cpp", "cbench_shim.cpp", "gpuMetrics.hpp", "zfpB200Compressor.hpp",
                                            "cbench_api.hpp", "mpi_single.hpp")]
    lib = os.path.join(PKG, "libzfb.so")
    stale = (not os.path.exists(EXE)) or (not os.path.exists(SHIM)) or any(os.path.getmtime(s) > os.path.getmtime(EXE) for s in srcs + [lib])
    if not (force or stale):
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", HERE, srcs[0], "-o", EXE,
           "-L", PKG, "-lzfb", "-Wl,-rpath," + PKG, "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cbench_mini build failed:\n" + r.stderr[-3000:])
    # the plugin classes behind a C interface (bench.py's end-to-end leg drives them through ctypes)
    cmd = [cxx, "-std=c++14", "-O2", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", HERE,
           os.path.join(HERE, "cbench_shim.cpp"), "-o", SHIM, "-L", PKG, "-lzfb", "-Wl,-rpath," + PKG, "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libzfb_cbench.so build failed:\n" + r.stderr[-3000:])
    return EXE


if __name__ == "__main__":
    print(build(force=True))
