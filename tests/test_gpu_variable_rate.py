"""GPU parity of the variable-rate path (zfp's fixed-accuracy / fixed-precision modes, the CPU zfp plugin's
"abs" / "rel": CBench/compressors/zfp/zfpCompressor.hpp:83-93,117-120) against the oracle, through the C ABI.

Bars: stream bytes, stream length, block index and decoded values BIT-EXACT; metrics within 1e-6 relative.
"""
import os

import numpy as np
import pytest

from conftest import smooth_field

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL = 1e-6


def _torch():
    import torch

    return torch


def _field(shape, dtype, seed=5):
    rng = np.random.default_rng(seed + len(shape))
    f = (smooth_field(shape, seed=seed, amp=40.0).astype(np.float64) + 0.01 * rng.standard_normal(shape)).astype(dtype)
    f.ravel()[:: max(1, f.size // 9)] *= 1e-4
    f[tuple(slice(0, min(4, n)) for n in shape)] = 0
    return f


def _modes(pkg, oracle):
    out = []
    for tol in (1e-3, 0.05, 3.0, 1e4, 0.0):
        out.append(("abs=%g" % tol, pkg.mode_accuracy(tol), oracle.params_accuracy(tol)))
    for prec in (1, 7, 16, 23, 32, 64):
        out.append(("rel=%d" % prec, pkg.mode_precision(prec), oracle.params_precision(prec)))
    return out


SHAPES = [(np.float32, (10,)), (np.float32, (4099,)), (np.float32, (70001,)), (np.float64, (33,)), (np.float32, (5, 7)), (np.float32, (64, 64)),
          (np.float32, (5, 6, 7)), (np.float32, (33, 33, 33)), (np.float32, (16, 32, 64)), (np.float32, (64, 64, 64)),
          (np.float64, (1001,)), (np.float64, (17, 19)), (np.float64, (5, 6, 7)), (np.float64, (16, 16, 32))]


@pytest.mark.parametrize("dtype,shape", SHAPES)
def test_variable_rate_bit_exact(codec, pkg, oracle, dtype, shape):
    torch = _torch()
    f = _field(shape, dtype)
    d_f = torch.from_numpy(f).cuda()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    ni, unit = pkg.index_count(shape), pkg.index_unit(len(shape))
    assert ni == -(-pkg.block_count(shape) // unit) and unit == (8 if len(shape) == 1 else 1)
    for name, m, p in _modes(pkg, oracle):
        s_ref, offs_ref = oracle.compress_ex(f, p, with_offsets=True)
        g_ref, _ = oracle.decompress_ex(s_ref, shape, dtype, p)
        index = torch.zeros(ni + 1, dtype=torch.int64, device="cuda")
        s, nbytes = codec.encode_var(d_f, m, index=index)
        assert nbytes == len(s_ref), name
        # one entry per `unit` blocks (1D: groups of 8), the total number of bits last
        want = np.concatenate([offs_ref[:-1][::unit], offs_ref[-1:]])
        assert np.array_equal(index.cpu().numpy().view(np.uint64), want), name
        assert np.array_equal(s.cpu().numpy(), s_ref), name
        # decode with the caller's index, fused metrics
        out = codec.decode_var(s, shape, tdt, m, index=index, orig=d_f)
        part = codec.partials()
        assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), name
        if dtype == np.float32:
            ref = oracle.metrics(f, g_ref)
            v = pkg.metric_values(part)
            assert v["absolute_error"] == ref["absolute_error"], name
            assert abs(v["mse"] - ref["mse"]) <= REL * max(ref["mse"], 1e-300), name
        # decode a bare copy of the stream: the index is rebuilt by walking the blocks
        bare = s.clone()
        out2 = codec.decode_var(bare, shape, tdt, m)
        assert torch.equal(out2.view(torch.uint8), out.view(torch.uint8)), name


def test_context_keeps_the_index_between_encode_and_decode(codec, pkg, oracle):
    torch = _torch()
    f = _field((32, 48, 64), np.float32, seed=9)
    d_f = torch.from_numpy(f).cuda()
    m, p = pkg.mode_accuracy(0.01), oracle.params_accuracy(0.01)
    before = codec.launches
    s, nbytes = codec.encode_var(d_f, m)           # no index tensor: the context keeps it
    enc_launches = codec.launches - before
    before = codec.launches
    out = codec.decode_var(s, f.shape, torch.float32, m)
    assert codec.launches - before == 1, "resident index must be used (no index walk)"
    assert enc_launches == 6
    g_ref, _ = oracle.decompress_ex(oracle.compress_ex(f, p), f.shape, np.float32, p)
    assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8))
    assert float((out - d_f).abs().max()) <= 0.01


def test_variable_rate_host_entry_points(codec, pkg, oracle):
    f = _field((40, 36, 52), np.float32, seed=11)
    for m, p in ((pkg.mode_accuracy(0.05), oracle.params_accuracy(0.05)), (pkg.mode_precision(14), oracle.params_precision(14))):
        s_ref = oracle.compress_ex(f, p)
        g_ref, _ = oracle.decompress_ex(s_ref, f.shape, np.float32, p)
        s = codec.compress_var_host(f, m)
        assert np.array_equal(s, s_ref)
        out = codec.decompress_var_host(s, f.shape, np.float32, m, orig=f)
        assert np.array_equal(out.view(np.uint8), g_ref.view(np.uint8))
        part = codec.metrics_host(f, out)  # fused partials are reused
        ref = oracle.metrics(f, g_ref)
        v = pkg.metric_values(part)
        assert v["absolute_error"] == ref["absolute_error"]
        assert abs(v["psnr"] - ref["psnr"]) <= REL * abs(ref["psnr"])
        # a stream that did not come from this context (fresh host buffer): index rebuilt on the device
        out2 = codec.decompress_var_host(s.copy(), f.shape, np.float32, m)
        assert np.array_equal(out2.view(np.uint8), g_ref.view(np.uint8))
    codec.release()


def test_shipped_input_parameters_on_reference_toy_data(codec, pkg, oracle):
    """The reference's own toy inputs with the parameters its shipped JSON files use
    (inputs/hacc/hacc_cbench_test.json: zfp abs 0.003 on HACC positions; inputs/nyx: abs / rel sweeps)."""
    torch = _torch()
    hacc = np.fromfile(os.path.join(GOLD, "hacc_m000_x_4096.f32"), dtype=np.float32)
    nyx = np.fromfile(os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), dtype=np.float32).reshape(32, 32, 32)
    for f, modes in ((hacc, [("abs", 0.003), ("abs", 0.1), ("rel", 16), ("rel", 24)]),
                     (nyx, [("abs", 0.01), ("abs", 1.0), ("rel", 12), ("rel", 20)])):
        d_f = torch.from_numpy(f).cuda()
        for kind, v in modes:
            m = pkg.mode_accuracy(v) if kind == "abs" else pkg.mode_precision(v)
            p = oracle.params_accuracy(v) if kind == "abs" else oracle.params_precision(v)
            s_ref = oracle.compress_ex(f, p)
            g_ref, _ = oracle.decompress_ex(s_ref, f.shape, np.float32, p)
            s, _ = codec.encode_var(d_f, m)
            assert np.array_equal(s.cpu().numpy(), s_ref), (kind, v)
            out = codec.decode_var(s, f.shape, torch.float32, m, orig=d_f)
            assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), (kind, v)
            if kind == "abs":
                assert pkg.metric_values(codec.partials())["absolute_error"] <= v


def test_variable_rate_outputs_stay_in_bounds_and_bad_arguments(codec, pkg):
    torch = _torch()
    G = 4096
    f = torch.randn((20, 24, 28), dtype=torch.float32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    m = pkg.mode_accuracy(0.02)
    cap = pkg.maximum_size(f.shape, f.dtype, m)
    buf = torch.full((cap + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
    s, nbytes = codec.encode_var(f, m, out=buf[G:G + cap])
    assert bool((buf[:G] == 0xA5).all()) and bool((buf[G + cap:] == 0xA5).all())
    assert bool((buf[G + nbytes:G + cap] == 0xA5).all()), "bytes behind the stream must not be touched"
    obuf = torch.full((f.numel() * 4 + 2 * G,), 0x5A, dtype=torch.uint8, device="cuda")
    out = obuf[G:G + f.numel() * 4].view(torch.float32).view(f.shape)
    codec.decode_var(s, f.shape, torch.float32, m, out=out)
    codec.synchronize()
    assert bool((obuf[:G] == 0x5A).all()) and bool((obuf[G + f.numel() * 4:] == 0x5A).all())
    assert float((out - f).abs().max()) <= 0.02
    bad = pkg.Mode(0, 0, 0, 0)
    with pytest.raises(pkg.ZfbError):
        codec.encode_var(f, bad)
    with pytest.raises(pkg.ZfbError):  # stream buffer below zfb_maximum_size()
        pkg.lib().zfb_encode_var_dev  # noqa: B018  (symbol exists)
        small = torch.empty(cap // 2, dtype=torch.uint8, device="cuda")
        import ctypes as C
        nb = C.c_size_t()
        rc = pkg.lib().zfb_encode_var_dev(codec._h, 1, 3, 28, 24, 20, C.byref(m), f.data_ptr(), small.data_ptr(),
                                          small.numel(), None, C.byref(nb))
        assert rc == -1
        raise pkg.ZfbError("EINVAL as expected")


def test_expert_parameters_and_fixed_rate_through_the_general_path(codec, pkg, oracle):
    """Budget, padding floor, plane limit and accuracy floor together (zfp_stream_set_params), and fixed rate run
    through the general path: the stream must be the one the fixed-rate kernels write."""
    torch = _torch()
    rng = np.random.default_rng(77)
    for dtype, shape in ((np.float32, (9, 10, 11)), (np.float32, (16, 16, 32)), (np.float64, (6, 9, 10)),
                         (np.float32, (37,)), (np.float64, (11, 13))):
        f = (smooth_field(shape, seed=8, amp=300.0).astype(np.float64) * (rng.random(shape) > 0.3)).astype(dtype)
        d_f = torch.from_numpy(f).cuda()
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        for prm in ((64, 300, 20, -20), (1, 130, 64, -1074), (200, 200, 9, -5), (17, 2000, 31, -30), (96, 97, 64, -1074)):
            if len(shape) == 1 and prm[0] > 140:
                continue  # 1D: a padding floor above the longest codable block is rejected (see below)
            m, p = pkg.Mode(*prm), oracle.Params(*prm)
            s_ref = oracle.compress_ex(f, p)
            g_ref, _ = oracle.decompress_ex(s_ref, shape, dtype, p)
            s, _ = codec.encode_var(d_f, m)
            assert np.array_equal(s.cpu().numpy(), s_ref), (dtype, shape, prm)
            out = codec.decode_var(s, shape, tdt, m)
            assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), (dtype, shape, prm)
        m = pkg.mode_rate(dtype, len(shape), 8)
        s, nbytes = codec.encode_var(d_f, m)
        s_fixed = codec.encode(d_f, m.maxbits)
        assert nbytes == s_fixed.numel() and torch.equal(s, s_fixed), (dtype, shape)
    with pytest.raises(pkg.ZfbError):  # minbits > maxbits
        codec.encode_var(d_f, pkg.Mode(9000, 300, 64, -1074))
    with pytest.raises(pkg.ZfbError):  # 1D float: padding floor above the 140-bit block bound
        codec.encode_var(torch.zeros(64, device="cuda"), pkg.Mode(200, 200, 9, -5))


def test_variable_rate_extreme_values(codec, pkg, oracle):
    """Denormals, values around the float range limits, mixed magnitudes, negative-only and all-zero fields in the
    accuracy / precision modes, float and double, 1D and 3D."""
    torch = _torch()
    rng = np.random.default_rng(3)
    cases = {
        "zeros": np.zeros((8, 8, 8), np.float32),
        "tiny": (rng.standard_normal((8, 8, 8)) * 1e-35).astype(np.float32),
        "denorm": (rng.standard_normal((8, 8, 8)) * 1e-42).astype(np.float32),
        "huge": (rng.standard_normal((8, 8, 8)) * 1e37).astype(np.float32),
        "mixed": (rng.standard_normal((8, 8, 8)) * 10.0 ** rng.integers(-30, 30, (8, 8, 8))).astype(np.float32),
        "neg": -np.abs(rng.standard_normal((8, 8, 8))).astype(np.float32),
        "line": (rng.standard_normal(4099) * 10.0 ** rng.integers(-20, 20, 4099)).astype(np.float32),
        "dmixed": rng.standard_normal((6, 8, 12)) * 10.0 ** rng.integers(-200, 200, (6, 8, 12)),
        "ddenorm": rng.standard_normal((4, 8, 8)) * 1e-310,
    }
    for label, f in cases.items():
        d_f = torch.from_numpy(f).cuda()
        tdt = torch.float32 if f.dtype == np.float32 else torch.float64
        for m, p in ((pkg.mode_accuracy(1e-3), oracle.params_accuracy(1e-3)), (pkg.mode_accuracy(1e-40), oracle.params_accuracy(1e-40)),
                     (pkg.mode_accuracy(1e30), oracle.params_accuracy(1e30)), (pkg.mode_precision(11), oracle.params_precision(11)),
                     (pkg.mode_precision(64), oracle.params_precision(64))):
            s_ref = oracle.compress_ex(f, p)
            g_ref, _ = oracle.decompress_ex(s_ref, f.shape, f.dtype, p)
            s, _ = codec.encode_var(d_f, m)
            assert np.array_equal(s.cpu().numpy(), s_ref), (label, m.as_tuple())
            out = codec.decode_var(s, f.shape, tdt, m)
            assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), (label, m.as_tuple())


def test_variable_rate_known_answer_hashes_on_gpu(codec, pkg):
    """The committed known-answer vectors (tests/golden/codec_var_kat.json) without the oracle in the loop."""
    import hashlib
    import json

    import make_golden

    torch = _torch()
    kat = json.load(open(os.path.join(GOLD, "codec_var_kat.json")))
    for case in kat["cases"]:
        f = make_golden.case_input(case)
        assert hashlib.sha256(f.tobytes()).hexdigest() == case["input_sha256"], case["name"]
        m = pkg.mode_accuracy(case["value"]) if case["mode"] == "abs" else pkg.mode_precision(case["value"])
        assert list(m.as_tuple()) == case["params"], case["name"]
        d_f = torch.from_numpy(f).cuda()
        tdt = torch.float32 if f.dtype == np.float32 else torch.float64
        s, nbytes = codec.encode_var(d_f, m)
        assert nbytes == case["stream_bytes"], case["name"]
        assert hashlib.sha256(s.cpu().numpy().tobytes()).hexdigest() == case["stream_sha256"], case["name"]
        out = codec.decode_var(s, f.shape, tdt, m)
        assert hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest() == case["decoded_sha256"], case["name"]


_SLAB_SCRIPT = r"""
import sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import __graft_entry__ as g
import oracle as O
from conftest import smooth_field
pkg = g.load_package()
codec = pkg.Codec(0)
rng = np.random.default_rng(5)
cases = [(np.float32, (70, 96, 128)), (np.float32, (3000001,)), (np.float64, (37, 64, 96)), (np.float32, (301, 1500))]
for dtype, shape in cases:
    f = (smooth_field(shape, seed=3, amp=20.0).astype(np.float64) + 0.01 * rng.standard_normal(shape)).astype(dtype)
    for m, p in ((pkg.mode_accuracy(0.01), O.params_accuracy(0.01)), (pkg.mode_precision(13), O.params_precision(13))):
        s_ref = O.compress_ex(f, p)
        g_ref, _ = O.decompress_ex(s_ref, shape, dtype, p)
        s = codec.compress_var_host(f, m)
        assert np.array_equal(s, s_ref), ("stream", dtype, shape)
        out = codec.decompress_var_host(s, shape, dtype, m, orig=f)
        assert np.array_equal(out.view(np.uint8), g_ref.view(np.uint8)), ("decoded", dtype, shape)
        if dtype == np.float32:
            v = pkg.metric_values(codec.metrics_host(f, out)); r = O.metrics(f, g_ref)
            assert v["absolute_error"] == r["absolute_error"] and abs(v["mse"] - r["mse"]) <= 1e-6 * r["mse"], (v, r)
        out2 = codec.decompress_var_host(s.copy(), shape, dtype, m)   # foreign buffer: whole-stream path
        assert np.array_equal(out2.view(np.uint8), g_ref.view(np.uint8)), ("decoded bare", dtype, shape)
print("SLABS OK", len(cases))
"""


def test_variable_rate_host_pipeline_with_many_slabs():
    """The host entry points cut a field into slabs whose streams join at arbitrary bits; with 1 MiB slabs
    (ZFB_SLAB_MB, read once per process, hence the subprocess) a 3-12 MiB field becomes 3-12 chained slabs."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ZFB_SLAB_MB="1")
    r = subprocess.run([sys.executable, "-c", _SLAB_SCRIPT], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SLABS OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_coarse_index_expansion(codec, pkg, oracle):
    """A stream kept for later needs only every stride-th index entry: the full index comes back by walking the
    segments in parallel, and a coarse index that does not fit the stream is refused."""
    torch = _torch()
    for dtype, shape in ((np.float32, (40, 48, 64)), (np.float32, (70001,)), (np.float64, (12, 20, 28)), (np.float32, (90, 70))):
        f = _field(shape, dtype, seed=13)
        d_f = torch.from_numpy(f).cuda()
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        m, p = pkg.mode_accuracy(0.02), oracle.params_accuracy(0.02)
        ni = pkg.index_count(shape)
        full = torch.zeros(ni + 1, dtype=torch.int64, device="cuda")
        s, _ = codec.encode_var(d_f, m, index=full)
        for stride in (1, 7, 128, ni + 5):
            coarse = torch.cat([full[:-1][::stride], full[-1:]]).contiguous()
            got = codec.index_expand(s, shape, tdt, m, coarse, stride)
            assert torch.equal(got, full), (dtype, shape, stride)
        out = codec.decode_var(s.clone(), shape, tdt, m, index=got)
        g_ref, _ = oracle.decompress_ex(oracle.compress_ex(f, p), shape, dtype, p)
        assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), (dtype, shape)
        if ni > 20:
            wrong = torch.cat([full[:-1][::7], full[-1:]]).contiguous()
            wrong[1] += 3  # a kept entry that is not a block boundary
            with pytest.raises(pkg.ZfbError):
                codec.index_expand(s, shape, tdt, m, wrong, 7)
