"""The C++ host side above the C ABI: ZFPCompressorB200 (CompressorInterface) and the GPU metric classes
(MetricInterface), driven by cbench_mini the way CBench/main.cpp drives its plugins."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CB = os.path.join(ROOT, "vizaly-foresight_b200", "cbench")
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference/CBench"


def build_mini():
    import importlib.util

    spec = importlib.util.spec_from_file_location("zfb_cbench_build", os.path.join(CB, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()


def test_mini_driver_builds_and_fails_loudly_without_gpu(pkg):
    import torch

    exe = build_mini()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "synth", "8", "8", "8", "8"], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no CPU fallback" in (r.stdout + r.stderr)


def test_plugins_compile_against_reference_headers(pkg):
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    out = os.path.join(ROOT, "build", "dropin_check")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    pk = os.path.join(ROOT, "vizaly-foresight_b200")
    cmd = [cxx, "-std=c++14", "-w", "-I", os.path.join(ROOT, "oracle", "mpi_stub"), "-I", os.path.join(REF, "compressors"),
           "-I", os.path.join(REF, "metrics"), "-I", os.path.join(REF, "utils"), "-I", os.path.join(ROOT, "include"),
           "-I", CB, os.path.join(ROOT, "tests", "dropin_check.cpp"), "-o", out, "-L", pk, "-lzfb", "-Wl,-rpath," + pk,
           "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([out], capture_output=True, text=True)
    assert "zfp_gpu bits8 psnr" in r.stdout
    assert "histogram_script identical" in r.stdout
    import torch

    if not torch.cuda.is_available():
        assert "compress rc 0" in r.stdout and "compression failed" in r.stdout


@pytest.mark.gpu
def test_mini_driver_matches_oracle(oracle):
    exe = build_mini()
    f = np.fromfile(os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), dtype=np.float32).reshape(32, 32, 32)
    for bits, extra in (("8", []), ("14", ["--hist"]), ("4.9", [])):   # "bits" is int-truncated: 4.9 -> 4
        r = subprocess.run([exe, os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), "32", "32", "32", bits,
                            "--field", "baryon_density"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        lines = r.stdout.splitlines()
        assert lines[0].startswith("Compressor_field__params, name, absolute_error, relative_error, mse, psnr, minmax, "
                                   "Compression Throughput(MB/s)")
        row = [t.strip() for t in lines[1].split(",")]
        assert row[0] == "zfp_gpu_baryon_density__bits" + bits
        mb = oracle.maxbits(np.float32, 3, int(float(bits)))
        g = oracle.decompress(oracle.compress(f, mb), f.shape, np.float32, mb)
        m = oracle.metrics(f, g)
        got = [float(t) for t in row[2:7]]
        want = [m["absolute_error"], m["relative_error"], m["mse"], m["psnr"], m["minmax"]]
        for a, b in zip(got, want):
            assert abs(a - b) <= 1.01e-5 * abs(b), (bits, got, want)   # ostream default: 6 significant digits
        assert abs(float(row[9]) - f.nbytes / oracle.stream_bytes(f.shape, mb)) < 1e-3
        text = "\n".join(lines[2:])
        assert "-Max Abs Error:" in text and "-psnr:" in text and "-minmax:" in text and "-mse:" in text
        if extra:
            assert text.count("[histogram script:") == 3


@pytest.mark.gpu
def test_dropin_check_runs_on_gpu_if_built():
    exe = os.path.join(ROOT, "build", "dropin_check")
    if not os.path.exists(exe):
        pytest.skip("dropin_check is built only where the reference tree exists")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "compress rc 1" in r.stdout


@pytest.mark.gpu
def test_mini_driver_abs_and_rel_modes_match_oracle(oracle):
    """The CPU zfp plugin's parameters ("abs", "rel": zfp/zfpCompressor.hpp:83-93) through the same plugin class."""
    exe = build_mini()
    f = np.fromfile(os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), dtype=np.float32).reshape(32, 32, 32)
    for arg, p in (("abs=0.01", oracle.params_accuracy(0.01)), ("rel=18", oracle.params_precision(18)),
                   ("abs=1.5", oracle.params_accuracy(1.5))):
        r = subprocess.run([exe, os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), "32", "32", "32", arg,
                            "--field", "baryon_density", "--name", "zfp"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        row = [t.strip() for t in r.stdout.splitlines()[1].split(",")]
        assert row[0] == "zfp_baryon_density__" + arg.replace("=", "")
        s = oracle.compress_ex(f, p)
        g, _ = oracle.decompress_ex(s, f.shape, np.float32, p)
        m = oracle.metrics(f, g)
        got = [float(t) for t in row[2:7]]
        want = [m["absolute_error"], m["relative_error"], m["mse"], m["psnr"], m["minmax"]]
        for a, b in zip(got, want):
            assert abs(a - b) <= 1.01e-5 * abs(b), (arg, got, want)
        assert abs(float(row[9]) - f.nbytes / len(s)) <= 1e-5 * f.nbytes / len(s) + 1e-3, (arg, row[9], f.nbytes / len(s))
    # no parameter at all: abs = 1E-3 under both names (zfpCompressor.hpp:80, zfpCompressorGpu.hpp:80-100)
    p = oracle.params_accuracy(1e-3)
    s = oracle.compress_ex(f, p)
    for name in ("zfp", "zfp_gpu"):
        r = subprocess.run([exe, os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), "32", "32", "32", "default",
                            "--field", "baryon_density", "--name", name], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        row = [t.strip() for t in r.stdout.splitlines()[1].split(",")]
        assert abs(float(row[9]) - f.nbytes / len(s)) <= 1e-5 * f.nbytes / len(s) + 1e-3, (name, row[9], f.nbytes / len(s))


@pytest.mark.gpu
@pytest.mark.parametrize("name,arg", [("quant", "0.05"), ("zfp_gpu", "8"), ("zfp", "abs=0.01")])
def test_mini_driver_several_scalars_at_the_same_addresses(oracle, name, arg):
    """main.cpp re-allocates `data` and `decompdata` for every scalar (main.cpp:395-398) and the allocator returns
    the same addresses: every CSV row must carry ITS scalar's metrics, also when the compressor is not ours and only
    the GPU metric classes see the arrays (the device copies are keyed by host address)."""
    exe = build_mini()
    path = os.path.join(GOLD, "nyx_z255_32_baryon_density.f32")
    f = np.fromfile(path, dtype=np.float32).reshape(32, 32, 32)
    r = subprocess.run([exe, path, "32", "32", "32", arg, "--field", "rho", "--name", name, "--scalars", "3", "--hist"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    for sc in range(3):
        x = (f * np.float32(sc + 1)).astype(np.float32)
        if name == "quant":
            step = np.float32(float(arg))
            g = (np.rint(x / step).astype(np.int32).astype(np.float32) * step).astype(np.float32)
        elif name == "zfp_gpu":
            mb = oracle.maxbits(np.float32, 3, int(arg))
            g = oracle.decompress(oracle.compress(x, mb), x.shape, np.float32, mb)
        else:
            p = oracle.params_accuracy(0.01)
            g, _ = oracle.decompress_ex(oracle.compress_ex(x, p), x.shape, np.float32, p)
        m = oracle.metrics(x, g)
        row = [t.strip() for t in lines[1 + sc].split(",")]
        assert row[0].startswith(name + "_rho%d__" % sc)
        got = [float(t) for t in row[2:7]]
        want = [m["absolute_error"], m["relative_error"], m["mse"], m["psnr"], m["minmax"]]
        for a, b in zip(got, want):
            assert abs(a - b) <= 1.01e-5 * abs(b), (name, sc, got, want)
