"""The device block codec (zfb_codec.cuh) compiled for the CPU (tests/host_emul.cpp) must produce the
oracle's streams and decoded values bit for bit.  This pins the kernel arithmetic without a GPU."""
import ctypes as C
import os
import glob
import subprocess

import numpy as np
import pytest

from conftest import ROOT, smooth_field

_lib = None


def emul():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, "build", "libhost_emul.so")
        src = os.path.join(ROOT, "tests", "host_emul.cpp")
        hdrs = glob.glob(os.path.join(ROOT, "vizaly-foresight_b200", "csrc", "*.cuh"))
        os.makedirs(os.path.dirname(so), exist_ok=True)
        if (not os.path.exists(so)) or max(os.path.getmtime(p) for p in [src] + hdrs) > os.path.getmtime(so):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src],
                           check=True)
        L = C.CDLL(so)
        for f in (L.emul_compress, L.emul_decompress):
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p,
                          C.c_size_t]
        _lib = L
    return _lib


def emul_roundtrip(O, f, mb):
    d, nx, ny, nz = O.dims_of(f.shape)
    t = 1 if f.dtype == np.float32 else 2
    nbytes = O.stream_bytes(f.shape, mb)
    s = np.zeros(nbytes // 8 + 1, dtype=np.uint64)
    assert emul().emul_compress(t, d, nx, ny, nz, mb, f.ctypes.data, s.ctypes.data, nbytes // 8) == 0
    g = np.empty_like(f)
    assert emul().emul_decompress(t, d, nx, ny, nz, mb, s.ctypes.data, g.ctypes.data, nbytes // 8) == 0
    return s.view(np.uint8)[:nbytes], g


def test_transposes_are_lsb_transposes():
    L = emul()
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2 ** 32, size=32, dtype=np.uint64).astype(np.uint32)
    b = a.copy()
    L.emul_transpose32(b.ctypes.data_as(C.c_void_p))
    for k in range(32):
        for i in range(32):
            assert (int(b[k]) >> i) & 1 == (int(a[i]) >> k) & 1
    a = rng.integers(0, 2 ** 63, size=64, dtype=np.uint64) * 2 + rng.integers(0, 2, size=64, dtype=np.uint64)
    b = a.copy()
    L.emul_transpose64(b.ctypes.data_as(C.c_void_p))
    for k in range(64):
        for i in range(64):
            assert (int(b[k]) >> i) & 1 == (int(a[i]) >> k) & 1


CASES = [
    (np.float32, (10,)), (np.float32, (4099,)), (np.float32, (5, 7)), (np.float32, (64, 64)),
    (np.float32, (5, 6, 7)), (np.float32, (33, 33, 33)), (np.float32, (16, 32, 64)),
    (np.float64, (1001,)), (np.float64, (17, 19)), (np.float64, (5, 6, 7)), (np.float64, (16, 16, 32)),
]


@pytest.mark.parametrize("dtype,shape", CASES)
def test_device_codec_equals_oracle(oracle, dtype, shape):
    O = oracle
    kinds = {
        "smooth": smooth_field(shape, seed=11, dtype=dtype, amp=100.0),
        "noise": (900 * np.random.default_rng(5).standard_normal(shape)).astype(dtype),
        "lognormal": np.exp(2.0 * smooth_field(shape, seed=12, dtype=np.float64)).astype(dtype),
        "sparse": (smooth_field(shape, seed=13, dtype=dtype) * (np.random.default_rng(6).random(shape) > 0.9)).astype(dtype),
    }
    rates = (1, 3, 4, 8, 16, 32) if dtype == np.float32 else (1, 5, 8, 16, 32, 64)
    for label, f in kinds.items():
        for rate in rates:
            for wra in (1, 0):
                mb = O.maxbits(dtype, len(shape), rate, wra)
                s_ref = O.compress(f, mb)
                g_ref = O.decompress(s_ref, shape, dtype, mb)
                s, g = emul_roundtrip(O, f, mb)
                assert np.array_equal(s, s_ref), (label, rate, wra, mb)
                assert np.array_equal(g.view(np.uint8), g_ref.view(np.uint8)), (label, rate, wra, mb)


def test_device_codec_extremes(oracle):
    O = oracle
    rng = np.random.default_rng(3)
    cases = {
        "zeros": np.zeros((8, 8, 8), np.float32),
        "tiny": (rng.standard_normal((8, 8, 8)) * 1e-35).astype(np.float32),  # scale overflow edge (A.5)
        "denorm": (rng.standard_normal((8, 8, 8)) * 1e-42).astype(np.float32),
        "huge": (rng.standard_normal((8, 8, 8)) * 1e37).astype(np.float32),
        "mixed": (rng.standard_normal((8, 8, 8)) * 10.0 ** rng.integers(-30, 30, (8, 8, 8))).astype(np.float32),
        "one_nonzero": np.pad(np.ones((1, 1, 1), np.float32), ((3, 4), (2, 5), (0, 7))),
        "neg": -np.abs(rng.standard_normal((8, 8, 8))).astype(np.float32),
    }
    for label, f in cases.items():
        for mb in (64, 256, 512, 1024, 2048, 100, 9, 37):
            s_ref = O.compress(f, mb)
            g_ref = O.decompress(s_ref, f.shape, np.float32, mb)
            s, g = emul_roundtrip(O, f, mb)
            assert np.array_equal(s, s_ref), (label, mb)
            assert np.array_equal(g.view(np.uint8), g_ref.view(np.uint8)), (label, mb)
    fd = (rng.standard_normal((4, 8, 8)) * 1e-300).astype(np.float64)
    for mb in (64, 512, 4096, 12, 77):
        s_ref = O.compress(fd, mb)
        s, g = emul_roundtrip(O, fd, mb)
        assert np.array_equal(s, s_ref), mb
        assert np.array_equal(g.view(np.uint8), O.decompress(s_ref, fd.shape, np.float64, mb).view(np.uint8)), mb


def test_fast_coder_budget_boundaries(oracle):
    """float/3D fast coder (zfb_codec_f3.cuh): every budget from 9 to 2200 bits on fields that drive the
    coder through its table steps, its exact single-bit steps (budget end, last coefficients) and the
    lazily finished lower planes."""
    O = oracle
    rng = np.random.default_rng(42)
    base = smooth_field((4, 8, 8), seed=3, amp=7.0)
    fields = [
        base,
        (rng.standard_normal((4, 8, 8)) * 900).astype(np.float32),
        (base * (rng.random((4, 8, 8)) > 0.8)).astype(np.float32),
        np.where(rng.random((4, 8, 8)) > 0.97, 1.0e3, 1.0e-3).astype(np.float32),
        np.full((4, 8, 8), -3.25, np.float32),
        np.cumsum(np.ones((4, 8, 8), np.float32)).reshape(4, 8, 8),
    ]
    for f in fields:
        for mb in list(range(9, 140)) + list(range(140, 2200, 37)) + [256, 512, 1024, 2048]:
            s_ref = O.compress(f, mb)
            s, g = emul_roundtrip(O, f, mb)
            assert np.array_equal(s, s_ref), mb
            assert np.array_equal(g.view(np.uint8), O.decompress(s_ref, f.shape, np.float32, mb).view(np.uint8)), mb


def test_fast_1d_coder(oracle):
    """float/1D fast coder (zfb_codec_f1.cuh): budgets 32/64/96/128, HACC-like and adversarial arrays."""
    O = oracle
    rng = np.random.default_rng(9)
    n = 4096
    arrays = {
        "uniform256": (256 * rng.random(n)).astype(np.float32),
        "normal900": (900 * rng.standard_normal(n)).astype(np.float32),
        "smooth": smooth_field((n,), seed=2, amp=50.0),
        "ramp": np.arange(n, dtype=np.float32) * 0.37,
        "const": np.full(n, 7.5, np.float32),
        "zeros": np.zeros(n, np.float32),
        "sparse": (rng.standard_normal(n) * (rng.random(n) > 0.9)).astype(np.float32),
        "tiny": (rng.standard_normal(n) * 1e-36).astype(np.float32),
        "mixed": (rng.standard_normal(n) * 10.0 ** rng.integers(-20, 20, n)).astype(np.float32),
        "one_of_four": np.tile(np.array([3.0, 0, 0, 0], np.float32), n // 4),
        "last_of_four": np.tile(np.array([0, 0, 0, -5.0], np.float32), n // 4),
        "golden_hacc": np.fromfile(os.path.join(ROOT, "tests", "golden", "hacc_m000_x_4096.f32"), dtype=np.float32),
        # long runs of planes with 1 or 2 significant coefficients (run-based head coder)
        "dc_plus_tiny": (1000.0 + 1e-4 * rng.standard_normal(n)).astype(np.float32),
        "dc_plus_small": (1000.0 + 1e-1 * rng.standard_normal(n)).astype(np.float32),
        "steps": np.repeat((rng.standard_normal(n // 4) * 50).astype(np.float32), 4),
        "slopes": (np.arange(n) * np.repeat(10.0 ** rng.integers(-6, 3, n // 4), 4)).astype(np.float32),
        "last_differs": (np.repeat((rng.random(n // 4) * 9).astype(np.float32), 4)
                         + np.tile(np.array([0, 0, 0, 1e-3], np.float32), n // 4)),
        "pairwise": np.repeat((rng.standard_normal(n // 2) * 3).astype(np.float32), 2),
    }
    # cubic blocks whose four polynomial terms differ by up to 2^36 in size: runs of every length at n = 1 and n = 2 (longer
    # than one 32-bit window, i.e. the loop behind the three straight-line head steps), budgets that end inside a run, an
    # event or the interleave tail
    xs = np.arange(4, dtype=np.float64) - 1.5
    for tag, spread in (("poly_wide", 36), ("poly_narrow", 6)):
        mags = 2.0 ** -rng.integers(0, spread, (4 * n // 4, 4)).astype(np.float64)
        sign = rng.choice([-1.0, 1.0], mags.shape)
        c = mags * sign * rng.random(mags.shape) * 1e3
        blocks = c[:, 0:1] + c[:, 1:2] * xs + c[:, 2:3] * xs ** 2 + c[:, 3:4] * xs ** 3
        arrays[tag] = blocks.astype(np.float32).ravel()
    for label, f in arrays.items():
        for mb in (32, 64, 96, 128):
            s_ref = O.compress(f, mb)
            s, g = emul_roundtrip(O, f, mb)
            assert np.array_equal(s, s_ref), (label, mb)
            assert np.array_equal(g.view(np.uint8), O.decompress(s_ref, f.shape, np.float32, mb).view(np.uint8)), (label, mb)


def test_fast_f64_coder_budget_boundaries(oracle):
    """double/3D fast coder (zfb_codec_d3.cuh): budgets 12..4200 bits, lazily finished plane groups,
    reduced precision for tiny exponents (kmin > 0)."""
    O = oracle
    rng = np.random.default_rng(43)
    base = smooth_field((4, 8, 8), seed=5, dtype=np.float64, amp=7.0)
    fields = [
        base,
        rng.standard_normal((4, 8, 8)) * 900,
        base * (rng.random((4, 8, 8)) > 0.8),
        np.where(rng.random((4, 8, 8)) > 0.97, 1.0e3, 1.0e-3),
        np.full((4, 8, 8), -3.25),
        rng.standard_normal((4, 8, 8)) * 1e-305,     # emax near the bottom: precision < 64 planes
        rng.standard_normal((4, 8, 8)) * 4e-308,
        rng.standard_normal((4, 8, 8)) * 1e300,
    ]
    for f in fields:
        f = np.ascontiguousarray(f, dtype=np.float64)
        for mb in list(range(12, 100, 3)) + list(range(100, 4200, 97)) + [64, 512, 1024, 2048, 4096]:
            s_ref = O.compress(f, mb)
            s, g = emul_roundtrip(O, f, mb)
            assert np.array_equal(s, s_ref), mb
            assert np.array_equal(g.view(np.uint8), O.decompress(s_ref, f.shape, np.float64, mb).view(np.uint8)), mb


# --------------------------------------------------------------------------- variable-rate modes
def _var_protos():
    L = emul()
    L.emul_compress_var.restype = C.c_size_t
    L.emul_compress_var.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]
    L.emul_decompress_var.restype = C.c_size_t
    L.emul_decompress_var.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t]
    return L


def emul_var_roundtrip(O, f, params):
    L = _var_protos()
    d, nx, ny, nz = O.dims_of(f.shape)
    t = 1 if f.dtype == np.float32 else 2
    cap = O.maximum_size(f.shape, f.dtype, params)
    prm = np.array([params.minbits, params.maxbits, params.maxprec, params.minexp & 0xffffffff], dtype=np.uint32)
    s = np.zeros(cap // 8 + 2, dtype=np.uint64)
    nb = ((nx + 3) // 4) * ((ny + 3) // 4 if d >= 2 else 1) * ((nz + 3) // 4 if d >= 3 else 1)
    offs = np.zeros(nb + 1, dtype=np.uint64)
    got = L.emul_compress_var(t, d, nx, ny, nz, prm.ctypes.data, f.ctypes.data, s.ctypes.data, cap // 8, offs.ctypes.data)
    assert got
    g = np.empty_like(f)
    used = L.emul_decompress_var(t, d, nx, ny, nz, prm.ctypes.data, s.ctypes.data, g.ctypes.data, cap // 8)
    assert used == got
    return s.view(np.uint8)[:got], g, offs


VAR_MODES = [("accuracy", 1e-3), ("accuracy", 0.05), ("accuracy", 3.0), ("accuracy", 1e4), ("accuracy", 0.0),
             ("precision", 1), ("precision", 7), ("precision", 16), ("precision", 23), ("precision", 32),
             ("precision", 40), ("precision", 64)]


def _var_params(O, kind, v):
    return O.params_accuracy(v) if kind == "accuracy" else O.params_precision(v)


@pytest.mark.parametrize("dtype,shape", CASES)
def test_device_codec_variable_rate_equals_oracle(oracle, dtype, shape):
    """The device codec's general (minbits, maxbits, maxprec, minexp) form, block lengths included, against the
    oracle in zfp's fixed-accuracy and fixed-precision modes ("abs" / "rel" of the CPU zfp plugin)."""
    O = oracle
    rng = np.random.default_rng(len(shape) * 13 + np.dtype(dtype).itemsize)
    f = (smooth_field(shape, seed=5, amp=40.0).astype(np.float64) + 0.01 * rng.standard_normal(shape)).astype(dtype)
    f.ravel()[:: max(1, f.size // 9)] *= 1e-4  # mixed magnitudes: some blocks fall under the accuracy floor
    f[tuple(slice(0, min(4, n)) for n in shape)] = 0  # an all-zero block
    for kind, v in VAR_MODES:
        p = _var_params(O, kind, v)
        s_ref, offs_ref = O.compress_ex(f, p, with_offsets=True)
        g_ref, used = O.decompress_ex(s_ref, f.shape, dtype, p)
        assert used == len(s_ref)
        s, g, offs = emul_var_roundtrip(O, f, p)
        assert len(s) == len(s_ref), (kind, v)
        assert np.array_equal(offs, offs_ref), (kind, v)
        assert np.array_equal(s, s_ref), (kind, v)
        assert np.array_equal(g.view(np.uint8), g_ref.view(np.uint8)), (kind, v)
        if kind == "accuracy" and v > 0:
            assert np.abs(g.astype(np.float64) - f.astype(np.float64)).max() <= v


def test_device_codec_expert_parameters(oracle):
    """zfp's expert mode (zfp_stream_set_params): budget, padding floor, plane limit and accuracy floor all active
    at once -- every exit of the coder loops in one stream."""
    O = oracle
    rng = np.random.default_rng(77)
    for dtype, shape in ((np.float32, (9, 10, 11)), (np.float32, (16, 16, 16)), (np.float64, (6, 9, 10)),
                         (np.float32, (37,)), (np.float64, (11, 13))):
        f = (smooth_field(shape, seed=8, amp=300.0).astype(np.float64) * (rng.random(shape) > 0.3)).astype(dtype)
        for minbits, maxbits, maxprec, minexp in ((64, 300, 20, -20), (1, 130, 64, -1074), (200, 200, 9, -5),
                                                  (17, 2000, 31, -30), (1, 16657, 13, -2), (96, 97, 64, -1074)):
            p = O.Params(minbits, maxbits, maxprec, minexp)
            s_ref, offs_ref = O.compress_ex(f, p, with_offsets=True)
            g_ref, used = O.decompress_ex(s_ref, f.shape, dtype, p)
            assert used == len(s_ref)
            lens = np.diff(offs_ref.astype(np.int64))
            assert lens.min() >= minbits and lens.max() <= max(maxbits, minbits)
            s, g, offs = emul_var_roundtrip(O, f, p)
            assert np.array_equal(offs, offs_ref), (dtype, shape, minbits, maxbits, maxprec, minexp)
            assert np.array_equal(s, s_ref), (dtype, shape, minbits, maxbits, maxprec, minexp)
            assert np.array_equal(g.view(np.uint8), g_ref.view(np.uint8)), (dtype, shape, minbits, maxbits, maxprec, minexp)


def test_device_codec_random_sweep(oracle):
    """Seeded random shapes / types / parameter sets through the CPU build of the device codec (fixed rate, accuracy,
    precision, expert) against the oracle: the GPU-less counterpart of tests/test_gpu_fuzz.py."""
    O = oracle
    rng = np.random.default_rng(424242)
    for case in range(120):
        d = int(rng.integers(1, 4))
        shape = tuple(int(rng.integers(1, 14 if d == 3 else (40 if d == 2 else 300))) for _ in range(d))
        dtype = np.float32 if rng.random() < 0.6 else np.float64
        kind = int(rng.integers(0, 4))
        if kind == 0:
            f = rng.standard_normal(shape) * 10.0 ** int(rng.integers(-8, 9))
        elif kind == 1:
            f = smooth_field(shape, seed=case, amp=float(10.0 ** rng.uniform(-2, 3))).astype(np.float64)
        elif kind == 2:
            f = rng.standard_normal(shape) * (rng.random(shape) > 0.75)
        else:
            f = np.exp(rng.standard_normal(shape) * 5.0)
        f = np.ascontiguousarray(f.astype(dtype))
        sel = rng.random()
        if sel < 0.35:
            rate = float(rng.choice([1, 2.5, 4, 8, 13, 16, 32]))
            mb = O.maxbits(dtype, d, rate, int(rng.random() < 0.7))
            s_ref = O.compress(f, mb)
            s, g = emul_roundtrip(O, f, mb)
            assert np.array_equal(s, s_ref), (case, shape, dtype, mb)
            assert np.array_equal(g.view(np.uint8), O.decompress(s_ref, shape, dtype, mb).view(np.uint8)), (case, shape, mb)
            continue
        if sel < 0.6:
            p = O.params_accuracy(float(10.0 ** rng.uniform(-8, 4)))
        elif sel < 0.8:
            p = O.params_precision(int(rng.integers(1, 66)))
        else:
            hdr = 9 if dtype == np.float32 else 12
            maxbits = int(rng.integers(hdr, 700))
            cap = 140 if (d == 1 and dtype == np.float32) else 4000
            p = O.Params(int(rng.integers(0, min(maxbits, cap) + 1)), maxbits, int(rng.integers(1, 65)), int(rng.integers(-60, 10)))
        s_ref, offs_ref = O.compress_ex(f, p, with_offsets=True)
        g_ref, _ = O.decompress_ex(s_ref, shape, dtype, p)
        s, g, offs = emul_var_roundtrip(O, f, p)
        assert np.array_equal(offs, offs_ref), (case, shape, dtype, p.as_tuple())
        assert np.array_equal(s, s_ref), (case, shape, dtype, p.as_tuple())
        assert np.array_equal(g.view(np.uint8), g_ref.view(np.uint8)), (case, shape, dtype, p.as_tuple())
