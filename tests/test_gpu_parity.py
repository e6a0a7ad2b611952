"""GPU parity tests proper (run with -m gpu on the B200 box).  Everything goes through the C ABI
(include/zfb.h) via the ctypes wrapper; the oracle is only the checker.

Bars: compressed streams and decompressed values BIT-EXACT with the oracle; metrics within 1e-6 relative
(reduction order differs), max/min quantities exact; histogram counts exact.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT, smooth_field

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL = 1e-6  # tolerance north_star states for the metrics


def _torch():
    import torch

    return torch


def gpu_roundtrip(codec, f, mb, with_metrics=True):
    torch = _torch()
    d_f = torch.from_numpy(f).cuda()
    stream = codec.encode(d_f, mb)
    tdt = torch.float32 if f.dtype == np.float32 else torch.float64
    out = codec.decode(stream, f.shape, tdt, mb, orig=d_f if with_metrics else None)
    part = codec.partials() if with_metrics else None
    return stream.cpu().numpy(), out.cpu().numpy(), part


def assert_bits_equal(a, b, what):
    assert a.shape == b.shape, what
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), what


def check_metrics(pkg, oracle, part, f, g_ref, what):
    m_ref = oracle.metrics(f, g_ref)
    v = pkg.metric_values(part)
    assert v["n"] == f.size
    for k in ("absolute_error", "minmax", "min", "max"):      # max/min of identical floats: exact
        assert v[k] == m_ref[k], (what, k, v[k], m_ref[k])
    for k in ("relative_error", "mse", "psnr"):
        if np.isinf(m_ref[k]) or np.isnan(m_ref[k]):   # lossless round trip: mse 0, psnr inf
            assert v[k] == m_ref[k] or (np.isnan(v[k]) and np.isnan(m_ref[k])), (what, k, v[k], m_ref[k])
            continue
        assert abs(v[k] - m_ref[k]) <= REL * abs(m_ref[k]) + 1e-300, (what, k, v[k], m_ref[k])
    assert abs(v["sum_abs"] - m_ref["abs_sum"]) <= REL * abs(m_ref["abs_sum"]) + 1e-300, what
    assert abs(v["sum_rel"] - m_ref["rel_sum"]) <= REL * abs(m_ref["rel_sum"]) + 1e-300, what


SMALL = [
    (np.float32, (10,)), (np.float32, (4099,)), (np.float32, (5, 7)), (np.float32, (64, 64)), (np.float32, (5, 6, 7)),
    (np.float32, (33, 33, 33)), (np.float32, (16, 32, 64)), (np.float32, (1, 1, 1)), (np.float32, (4, 4, 4)),
    (np.float64, (1001,)), (np.float64, (17, 19)), (np.float64, (5, 6, 7)), (np.float64, (16, 16, 32)),
]


@pytest.mark.parametrize("dtype,shape", SMALL)
def test_small_fields_bit_exact(codec, pkg, oracle, dtype, shape):
    kinds = {
        "smooth": smooth_field(shape, seed=11, dtype=dtype, amp=100.0),
        "noise": (900 * np.random.default_rng(5).standard_normal(shape)).astype(dtype),
        "lognormal": np.exp(2.0 * smooth_field(shape, seed=12, dtype=np.float64)).astype(dtype),
    }
    rates = (1, 3, 4, 8, 16, 32) if dtype == np.float32 else (1, 5, 8, 16, 32, 64)
    for label, f in kinds.items():
        for rate in rates:
            for wra in (1, 0):
                mb = oracle.maxbits(dtype, len(shape), rate, wra)
                assert mb == pkg.maxbits(dtype, len(shape), rate, wra)
                s_ref = oracle.compress(f, mb)
                g_ref = oracle.decompress(s_ref, shape, dtype, mb)
                s, g, part = gpu_roundtrip(codec, f, mb, with_metrics=(dtype == np.float32))
                assert np.array_equal(s, s_ref), (label, rate, wra)
                assert_bits_equal(g, g_ref, (label, rate, wra))
                if dtype == np.float32:
                    check_metrics(pkg, oracle, part, f, g_ref, (label, rate, wra))


TILE_SHAPES = [(8, 8, 256), (8, 8, 1028), (36, 40, 72), (64, 64, 64), (4, 4, 4), (12, 8, 516), (32, 256, 512),
               # ny / nz not multiples of 4: tile kernels on the full block rows + edge kernel on the partial ones
               (7, 9, 12), (10, 6, 256), (33, 34, 64), (5, 4, 1028), (4, 5, 260), (9, 130, 520)]


@pytest.mark.parametrize("shape", TILE_SHAPES)
@pytest.mark.parametrize("rate", [4, 8, 16, 3, 32])
def test_tile_kernels_bit_exact(codec, pkg, oracle, shape, rate):
    """3D float fields with dims % 4 == 0 take the TMA tile kernels (boxes straddling row ends included)."""
    f = np.exp(1.2 * smooth_field(shape, seed=rate, dtype=np.float64)).astype(np.float32)
    mb = pkg.maxbits(np.float32, 3, rate)
    s_ref = oracle.compress(f, mb)
    g_ref = oracle.decompress(s_ref, shape, np.float32, mb)
    s, g, part = gpu_roundtrip(codec, f, mb)
    assert np.array_equal(s, s_ref)
    assert_bits_equal(g, g_ref, "decoded")
    check_metrics(pkg, oracle, part, f, g_ref, (shape, rate))
    # decode without metrics takes the other kernel instantiation
    s2, g2, _ = gpu_roundtrip(codec, f, mb, with_metrics=False)
    assert_bits_equal(g2, g_ref, "decoded (no metrics)")


def test_extreme_values(codec, pkg, oracle):
    rng = np.random.default_rng(3)
    cases = {
        "zeros": np.zeros((8, 8, 8), np.float32),
        "tiny": (rng.standard_normal((8, 8, 8)) * 1e-35).astype(np.float32),
        "denorm": (rng.standard_normal((8, 8, 8)) * 1e-42).astype(np.float32),
        "huge": (rng.standard_normal((8, 8, 8)) * 1e37).astype(np.float32),
        "mixed": (rng.standard_normal((8, 8, 8)) * 10.0 ** rng.integers(-30, 30, (8, 8, 8))).astype(np.float32),
        "neg": -np.abs(rng.standard_normal((8, 8, 8))).astype(np.float32),
        "ragged": (rng.standard_normal((7, 9, 11)) * 5).astype(np.float32),
    }
    for label, f in cases.items():
        for mb in (64, 256, 512, 1024, 2048, 100, 9, 37):
            s_ref = oracle.compress(f, mb)
            g_ref = oracle.decompress(s_ref, f.shape, np.float32, mb)
            s, g, _ = gpu_roundtrip(codec, f, mb, with_metrics=False)
            assert np.array_equal(s, s_ref), (label, mb)
            assert_bits_equal(g, g_ref, (label, mb))


def test_known_answer_hashes_on_gpu(codec, oracle):
    import make_golden

    kat = json.load(open(os.path.join(GOLD, "codec_kat.json")))
    for case in kat["cases"]:
        f = make_golden.case_input(case)
        assert hashlib.sha256(f.tobytes()).hexdigest() == case["input_sha256"], case["name"]
        s, g, _ = gpu_roundtrip(codec, f, case["maxbits"], with_metrics=False)
        assert len(s) == case["stream_bytes"]
        assert hashlib.sha256(s.tobytes()).hexdigest() == case["stream_sha256"], case["name"]
        assert hashlib.sha256(g.tobytes()).hexdigest() == case["decoded_sha256"], case["name"]


def test_reference_toy_data(codec, pkg, oracle):
    """Real inputs lifted from the reference's toy files (testing/data), all metrics included."""
    nyx = np.fromfile(os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), dtype=np.float32).reshape(32, 32, 32)
    hacc = np.fromfile(os.path.join(GOLD, "hacc_m000_x_4096.f32"), dtype=np.float32)
    for f, rates in ((nyx, (4, 8, 14, 16)), (hacc, (8, 16, 24))):
        for rate in rates:
            mb = pkg.maxbits(np.float32, f.ndim, int(rate))   # "bits" is int-truncated, zfpCompressorGpu.hpp:83,99
            s_ref = oracle.compress(f, mb)
            g_ref = oracle.decompress(s_ref, f.shape, np.float32, mb)
            s, g, part = gpu_roundtrip(codec, f, mb)
            assert np.array_equal(s, s_ref)
            assert_bits_equal(g, g_ref, rate)
            check_metrics(pkg, oracle, part, f, g_ref, rate)


def test_metrics_match_reference_vectors(codec, pkg, oracle):
    """Stand-alone metric kernel against values produced by the reference's own classes (metrics_kat.json)."""
    torch = _torch()
    kat = json.load(open(os.path.join(GOLD, "metrics_kat.json")))
    for case in kat["cases"]:
        rng = np.random.default_rng(case["seed"])
        n, kind = case["n"], case["kind"]
        if kind == "noise":
            o = rng.standard_normal(n).astype(np.float32) * 3
            a = (o + 0.01 * rng.standard_normal(n)).astype(np.float32)
        elif kind == "big":
            o = (rng.standard_normal(n) * 1e4).astype(np.float32)
            a = (o * (1 + 1e-3 * rng.standard_normal(n))).astype(np.float32)
        else:
            side = int(round(n ** (1 / 3)))
            o = smooth_field((side, side, side), seed=case["seed"], amp=50.0)
            mb = 512 if kind == "codec8" else 256
            a = oracle.decompress(oracle.compress(o, mb), o.shape, np.float32, mb)
        part = codec.metrics(torch.from_numpy(o.ravel()).cuda(), torch.from_numpy(a.ravel()).cuda())
        v = pkg.metric_values(part)
        assert v["absolute_error"] == case["absolute_error"]
        assert v["minmax"] == case["minmax"]
        for k in ("relative_error", "mse", "psnr"):
            assert abs(v[k] - case[k]) <= REL * abs(case[k]), (kind, k, v[k], case[k])


def test_histograms_match_oracle(codec, pkg, oracle):
    torch = _torch()
    f = smooth_field((32, 32, 64), seed=4, amp=30.0)
    mb = 256
    g = oracle.decompress(oracle.compress(f, mb), f.shape, np.float32, mb)
    m = oracle.metrics(f, g, hist=True)
    d_f, d_g = torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda()
    n = f.size
    cnt, oob = codec.histogram(0, 0.0, m["absolute_error"], d_f, d_g)
    assert int(oob.item()) == 0 and int(cnt.sum().item()) == n
    assert np.array_equal((cnt.cpu().numpy().astype(np.float32) / np.float32(n)), m["abs_hist"])
    cnt, oob = codec.histogram(1, 0.0, m["relative_error"], d_f, d_g)
    assert int(cnt.sum().item()) == n
    assert np.array_equal((cnt.cpu().numpy().astype(np.float32) / np.float32(n)), m["rel_hist"])
    cnt, oob = codec.histogram(2, m["min"], m["max"], None, d_g)
    assert int(cnt.sum().item()) == n
    assert np.array_equal(cnt.cpu().numpy().astype(np.float32), m["minmax_hist"])
    assert int(oob.item()) == m["minmax_oob"]


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("shape", [(32, 64, 256), (16, 24, 520), (9, 10, 36), (4099,)])
def test_value_range_histogram_fused_into_the_codec_passes(pkg, oracle, shape, fused):
    """minmaxMetric.hpp:62-133 without a second pass: the encode pass harvests the field's min / max, the decode pass
    bins what it decodes.  Whole-block float 3D shapes take the fused tile kernels, the others the stand-alone
    passes behind the same two calls; min / max exact, counts equal to the reference loop bin for bin."""
    torch = _torch()
    os.environ["ZFB_FUSED_HIST"] = fused
    codec = pkg.Codec(0)
    f = np.exp(1.3 * smooth_field(shape, seed=31, amp=1.0)).astype(np.float32)
    mb = pkg.maxbits(np.float32, len(shape), 8)
    d_f = torch.from_numpy(f).cuda()
    s_ref = oracle.compress(f, mb)
    g_ref = oracle.decompress(s_ref, f.shape, np.float32, mb)
    m = oracle.metrics(f, g_ref, hist=True)
    launches0 = codec.launches
    stream = codec.encode_range(d_f, mb)
    lo, hi = codec.range()
    assert (lo, hi) == (float(f.min()), float(f.max())) == (m["min"], m["max"])
    assert np.array_equal(stream.cpu().numpy(), s_ref)
    out, cnt, oob = codec.decode_hist(stream, f.shape, torch.float32, mb, lo, hi, orig=d_f)
    if len(shape) == 3 and all(e % 4 == 0 for e in shape):
        # encode + finalize (min / max harvested in the encode kernel), decode + finalize, and the stand-alone value pass
        # unless the binning is fused into the decode kernel as well
        assert codec.launches - launches0 == (4 if fused == "1" else 5)
    assert_bits_equal(out.cpu().numpy(), g_ref, "decoded")
    check_metrics(pkg, oracle, codec.partials(), f, g_ref, "fused with histogram")
    assert int(cnt.sum().item()) == f.size
    assert np.array_equal(cnt.cpu().numpy().astype(np.float32), m["minmax_hist"])
    assert int(oob.item()) == m["minmax_oob"]
    codec.close()
    del os.environ["ZFB_FUSED_HIST"]


def test_histogram_bin_edges_are_exact(codec):
    """Values on and next to bin edges: the reciprocal estimate in hist_kernel must fall back to the reference's
    exact double division wherever a few ulp could move a value across an edge (absoluteError.hpp:119-121,
    minmaxMetric.hpp:112)."""
    torch = _torch()
    rng = np.random.default_rng(3)
    for hi in (1.0, 0.3, 7.123456789, 1e-3, 123456.0):
        binw = hi / 1024.0
        j = rng.integers(0, 1025, size=60000)
        base = (j * binw).astype(np.float32)
        err = np.concatenate([base, np.nextafter(base, np.float32(0)), np.nextafter(base, np.float32(np.inf)),
                              (rng.random(60000) * hi).astype(np.float32), np.zeros(1000, np.float32),
                              np.full(1000, hi, np.float32)]).astype(np.float32)
        err = err[(err >= 0)]
        # which = 0: |o - a| with a = 0 is exactly err
        q = err.astype(np.float64) / binw
        pos = q.astype(np.int64)
        bad = (q >= 1025.0) | ((pos - (pos >= 1024)) >= 1024)
        pos = np.where(pos >= 1024, pos - 1, pos)
        pos = np.where(pos >= 1024, 1023, pos)
        want = np.bincount(pos, minlength=1024)
        d_o = torch.from_numpy(err).cuda()
        d_a = torch.zeros_like(d_o)
        cnt, oob = codec.histogram(0, 0.0, hi, d_o, d_a)
        assert np.array_equal(cnt.cpu().numpy().astype(np.int64), want), hi
        assert int(oob.item()) == int(bad.sum()), hi
        # which = 2: values over [lo, hi]
        lo = -0.25 * hi
        v = (err.astype(np.float64) * 1.25 + lo).astype(np.float32)
        rng_ = hi - lo
        q = (rng_ * ((v.astype(np.float64) - lo) / rng_)) / (rng_ / 1024.0)
        ok = q >= 0
        pos = np.where(ok, q, 0).astype(np.int64)
        over = ok & (q >= 1025.0)
        pos = np.where(over, 1023, pos)
        pos = np.where(~over & (pos >= 1024), pos - 1, pos)
        pos = np.where(pos >= 1024, 1023, pos)
        want = np.bincount(pos, minlength=1024)
        cnt, oob = codec.histogram(2, lo, hi, None, torch.from_numpy(v).cuda())
        assert np.array_equal(cnt.cpu().numpy().astype(np.int64), want), ("minmax", hi)


def test_generic_and_tile_paths_agree(pkg, oracle):
    """ZFB_FORCE_GENERIC=1 routes a tile-eligible field through the generic kernels: same bits."""
    torch = _torch()
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    f = smooth_field((16, 64, 512), seed=8, amp=10.0)
    mb = 512
    os.environ["ZFB_FORCE_GENERIC"] = "1"
    try:
        c2 = pkg.Codec(0)
    finally:
        del os.environ["ZFB_FORCE_GENERIC"]
    s_ref = oracle.compress(f, mb)
    s, g, part = gpu_roundtrip(c2, f, mb)
    assert np.array_equal(s, s_ref)
    check_metrics(pkg, oracle, part, f, oracle.decompress(s_ref, f.shape, np.float32, mb), "generic")
    c2.close()


@pytest.mark.parametrize("dtype,shape,mb", [(np.float32, (16, 64, 516), 512), (np.float32, (12, 20, 132), 256),
                                            (np.float32, (8, 8, 8), 1280), (np.float64, (16, 32, 260), 512),
                                            (np.float64, (8, 12, 36), 2048)])
def test_warp_specialised_and_tile_decode_agree(pkg, oracle, dtype, shape, mb):
    """The warp-specialised decode kernels (default) and the tile kernels (ZFB_DEC_WS=0) decode the same stream to the
    same bits and the same metrics, both equal to the oracle's: partial warp-tiles (nx / 4 not a multiple of 32), wide
    slots, tiny fields with fewer warp-tiles than parser warps."""
    torch = _torch()
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    f = smooth_field(shape, seed=21, amp=50.0, dtype=dtype)
    s_ref = oracle.compress(f, mb)
    g_ref = oracle.decompress(s_ref, f.shape, dtype, mb)
    outs = []
    for ws in ("1", "0"):
        os.environ["ZFB_DEC_WS"] = ws
        try:
            c2 = pkg.Codec(0)
        finally:
            del os.environ["ZFB_DEC_WS"]
        s, g, part = gpu_roundtrip(c2, f, mb)
        assert np.array_equal(s, s_ref)
        assert_bits_equal(g, g_ref, "decode, ZFB_DEC_WS=" + ws)
        if dtype == np.float32:
            check_metrics(pkg, oracle, part, f, g_ref, "ZFB_DEC_WS=" + ws)
        else:  # metrics of doubles are computed in double (test_f64_tile_kernels_bit_exact)
            d = f - g_ref
            v = pkg.metric_values(part)
            assert v["n"] == f.size and v["absolute_error"] == float(np.abs(d).max())
            assert v["max"] == float(f.max()) and v["min"] == float(f.min())
            assert abs(v["mse"] - float(np.mean(d * d))) <= 1e-9 * float(np.mean(d * d)) + 1e-300
        outs.append(g)
        c2.close()
    assert_bits_equal(outs[0], outs[1], "warp-specialised against tile kernel")


def test_host_entry_points_and_fused_metrics(codec, pkg, oracle):
    """zfb_compress_host / zfb_decompress_host / zfb_metrics_host, multi-slab pipeline included."""
    shape = (96, 512, 512)   # 96 MiB -> two slabs in the host pipeline
    f = smooth_field(shape, seed=2, amp=1000.0)
    mb = 512
    s = codec.compress_host(f, mb)
    g = codec.decompress_host(s, shape, np.float32, mb, orig=f)
    part = codec.metrics_host(f, g)       # the fused partials, no re-upload
    s_ref = oracle.compress(f, mb)
    g_ref = oracle.decompress(s_ref, shape, np.float32, mb)
    assert np.array_equal(s, s_ref)
    assert_bits_equal(g, g_ref, "host decode")
    check_metrics(pkg, oracle, part, f, g_ref, "fused host")
    # stand-alone: different buffers -> uploads both and runs the metric kernel
    g2 = g.copy()
    part2 = codec.metrics_host(f, g2)
    check_metrics(pkg, oracle, part2, f, g_ref, "standalone host")
    cnt, oob = codec.histogram_host(0, 0.0, part.max_abs, f, g2)
    assert int(cnt.sum()) == f.size
    codec.release()
    # 1D, ragged length, unaligned slots
    h = (256 * np.random.default_rng(1).random(100003)).astype(np.float32)
    for mb1 in (64, 32, 45):
        s1 = codec.compress_host(h, mb1)
        g1 = codec.decompress_host(s1, h.shape, np.float32, mb1, orig=h)
        assert np.array_equal(s1, oracle.compress(h, mb1)), mb1
        assert_bits_equal(g1, oracle.decompress(s1, h.shape, np.float32, mb1), mb1)
    codec.release()


def test_host_entry_points_pinned_and_pageable_buffers(codec, pkg, oracle):
    """The same result whether the caller's buffers are pageable (CBench: new T[] / malloc, main.cpp:262-280,395; staged
    through the pinned double buffers by the copy threads) or pinned (direct DMA); the stream zfb_compress_host left on
    the device is reused by zfb_decompress_host only while the host buffer still holds it."""
    torch = _torch()
    shape = (72, 512, 512)   # 72 MiB -> two slabs, the second one shorter
    f = smooth_field(shape, seed=21, amp=50.0)
    mb = 512
    s_ref = oracle.compress(f, mb)
    g_ref = oracle.decompress(s_ref, shape, np.float32, mb)
    # pageable
    s = codec.compress_host(f, mb)
    g = codec.decompress_host(s, shape, np.float32, mb, orig=f)
    assert np.array_equal(s, s_ref)
    assert_bits_equal(g, g_ref, "pageable")
    check_metrics(pkg, oracle, codec.metrics_host(f, g), f, g_ref, "pageable fused")
    # pinned
    tf = torch.from_numpy(f).pin_memory()
    ts = torch.empty(len(s_ref), dtype=torch.uint8).pin_memory()
    tg = torch.empty(shape, dtype=torch.float32).pin_memory()
    s2 = codec.compress_host(tf.numpy(), mb, out=ts.numpy())
    g2 = codec.decompress_host(s2, shape, np.float32, mb, orig=tf.numpy(), out=tg.numpy())
    assert np.array_equal(s2, s_ref)
    assert_bits_equal(g2, g_ref, "pinned")
    check_metrics(pkg, oracle, codec.metrics_host(tf.numpy(), g2), f, g_ref, "pinned fused")
    # the caller overwrites the stream buffer between the two calls: the resident copy must not be trusted
    f3 = smooth_field(shape, seed=22, amp=3.0)
    s3_ref = oracle.compress(f3, mb)
    s = codec.compress_host(f, mb)
    s[:] = s3_ref
    g3 = codec.decompress_host(s, shape, np.float32, mb)
    assert_bits_equal(g3, oracle.decompress(s3_ref, shape, np.float32, mb), "rewritten stream buffer")
    codec.release()


def test_host_cache_invalidation_for_reused_addresses(codec, pkg, oracle):
    """CBench frees and re-allocates `data` / `decompdata` per scalar at what are usually the same addresses
    (main.cpp:395-398): with a compressor that is not ours the GPU metrics must not return the previous scalar's
    numbers.  Codec.metrics_host drops the device copies unless decompress_host registered the pair."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal(1 << 18).astype(np.float32)
    b = (a + 0.01 * rng.standard_normal(a.size)).astype(np.float32)
    p1 = pkg.metric_values(codec.metrics_host(a, b))
    r1 = oracle.metrics(a, b)
    assert abs(p1["mse"] - r1["mse"]) <= 1e-6 * r1["mse"]
    # "next scalar": new contents at the same addresses
    a[:] = 100.0 * rng.standard_normal(a.size).astype(np.float32)
    b[:] = (a + 0.5 * rng.standard_normal(a.size)).astype(np.float32)
    p2 = pkg.metric_values(codec.metrics_host(a, b))
    r2 = oracle.metrics(a, b)
    for k in ("absolute_error", "relative_error", "mse", "psnr", "minmax"):
        assert abs(p2[k] - r2[k]) <= 1e-6 * abs(r2[k]) + 1e-300, (k, p2[k], r2[k])
    # explicit reuse keeps the device copies (the second metric of the same scalar)
    p3 = pkg.metric_values(codec.metrics_host(a, b, reuse=True))
    assert p3 == p2
    codec.release()


def test_ragged_1d_unaligned_slots_into_dirty_buffer(codec, pkg, oracle):
    """1D line path with a trailing partial block whose 96-bit slot spans two 64-bit words: everything behind the full
    blocks' slots is cleared before the partial block is OR-ed in (the stream buffer is reused and never zeroed)."""
    torch = _torch()
    for n, mb in ((11, 96), (4099, 96), (4098, 45), (10, 64)):
        h = (256 * np.random.default_rng(n).random(n)).astype(np.float32)
        d_h = torch.from_numpy(h).cuda()
        nbytes = pkg.stream_bytes(h.shape, mb)
        out = torch.full((nbytes + 16,), 0xFF, dtype=torch.uint8, device="cuda")
        s = codec.encode(d_h, mb, out=out)
        assert np.array_equal(s.cpu().numpy()[:nbytes], oracle.compress(h, mb)), (n, mb)


def test_1d_line_kernels_long_runs_and_tight_budgets(codec, pkg, oracle):
    """1D line kernels (three straight-line head steps + loop-free tails for 32/64-bit slots, zfb_codec_f1.cuh): cubic blocks
    whose polynomial terms differ by up to 2^36 give runs of every length at n = 1 and n = 2 (longer than one window: the
    loop behind the steps), and budgets of 32 / 64 / 96 / 128 bits end inside runs, events and the interleave tail."""
    torch = _torch()
    rng = np.random.default_rng(77)
    nb = 1 << 15
    xs = np.arange(4, dtype=np.float64) - 1.5
    for spread in (36, 6):
        mags = 2.0 ** -rng.integers(0, spread, (nb, 4)).astype(np.float64)
        c = mags * rng.choice([-1.0, 1.0], mags.shape) * rng.random(mags.shape) * 1e3
        h = (c[:, 0:1] + c[:, 1:2] * xs + c[:, 2:3] * xs ** 2 + c[:, 3:4] * xs ** 3).astype(np.float32).ravel()
        d_h = torch.from_numpy(h).cuda()
        for mb in (32, 64, 96, 128):
            s_ref = oracle.compress(h, mb)
            s = codec.encode(d_h, mb)
            assert np.array_equal(s.cpu().numpy(), s_ref), (spread, mb)
            o = codec.decode(s, h.shape, torch.float32, mb, orig=d_h)
            g_ref = oracle.decompress(s_ref, h.shape, np.float32, mb)
            assert_bits_equal(o.cpu().numpy(), g_ref, (spread, mb))
            check_metrics(pkg, oracle, codec.partials(), h, g_ref, "1D long runs %d/%d" % (spread, mb))


def test_misaligned_1d_arrays_take_the_scalar_path(codec, pkg, oracle):
    """A contiguous 1D slice such as field[1:] is only 4-byte aligned: no 16-byte accesses (a misaligned vector load is
    a sticky fault that kills the context)."""
    torch = _torch()
    base = torch.from_numpy((256 * np.random.default_rng(8).random(4101)).astype(np.float32)).cuda()
    for off in (1, 2, 3):
        d_h = base[off:off + 4096]
        assert d_h.data_ptr() % 16 != 0
        h = d_h.cpu().numpy()
        mb = 64
        s = codec.encode(d_h, mb)
        assert np.array_equal(s.cpu().numpy(), oracle.compress(h, mb))
        obuf = torch.empty(4101, dtype=torch.float32, device="cuda")
        o = codec.decode(s, h.shape, torch.float32, mb, orig=d_h, out=obuf[off:off + 4096])
        g_ref = oracle.decompress(oracle.compress(h, mb), h.shape, np.float32, mb)
        assert_bits_equal(o.cpu().numpy(), g_ref, off)
        check_metrics(pkg, oracle, codec.partials(), h, g_ref, "misaligned 1D")


def test_slab_streams_concatenate(codec, pkg):
    """slabs along the slowest axis (thickness % 4 == 0) concatenate to the whole-field stream."""
    torch = _torch()
    f = torch.from_numpy(smooth_field((64, 64, 128), seed=6)).cuda()
    mb = 512
    whole = codec.encode(f, mb).clone()
    parts = [codec.encode(f[z:z + 16].contiguous(), mb).clone() for z in range(0, 64, 16)]
    assert torch.equal(whole, torch.cat(parts))


def test_nyx_512_full_size_against_oracle(codec, pkg, oracle):
    """BASELINE config 3 at full size: 512^3 float32, rate 8; oracle runs slab-parallel on the host cores."""
    import ctypes as C
    torch = _torch()
    shape = (512, 512, 512)
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.from_numpy(smooth_field((64, 64, 64), seed=1, amp=1.0)).cuda()
    f = torch.nn.functional.interpolate(base[None, None], size=shape, mode="trilinear", align_corners=False)[0, 0]
    f = torch.exp(1.5 * f + 0.05 * torch.randn(shape, device="cuda", generator=g)).contiguous()
    mb = pkg.maxbits(np.float32, 3, 8)
    stream = codec.encode(f, mb)
    out = codec.decode(stream, shape, torch.float32, mb, orig=f)
    part = codec.partials()
    h = f.cpu().numpy()
    nslab = 32
    st = np.zeros(stream.numel(), dtype=np.uint8)
    dec = np.empty_like(h)
    oracle.lib().zfo_roundtrip_slabs(1, 3, 512, 512, 512 // nslab, mb, h.ctypes.data, st.ctypes.data, dec.ctypes.data,
                                     nslab, 0, None, None)
    assert np.array_equal(stream.cpu().numpy(), st)
    assert_bits_equal(out.cpu().numpy(), dec, "512^3 decode")
    v = pkg.metric_values(part)
    d = (h.astype(np.float32) - dec).astype(np.float64)
    assert abs(v["mse"] - float(np.mean(d * d))) <= REL * float(np.mean(d * d))
    assert v["absolute_error"] == float(np.abs(h - dec).max())
    assert v["max"] == float(h.max()) and v["min"] == float(h.min())


def _oracle_slabs(oracle, h, mb, nslab):
    """Oracle round trip of a field cut into nslab pieces along the slowest axis (thickness % 4 == 0, so the piece
    streams concatenate to the whole-field stream), one host thread per piece."""
    shape = h.shape
    code = 1 if h.dtype == np.float32 else 2
    d = len(shape)
    thick = shape[0] // nslab
    assert thick * nslab == shape[0] and thick % 4 == 0
    nx = shape[-1] if d > 1 else thick
    ny = shape[1] if d == 3 else (thick if d == 2 else 1)
    nz = thick if d == 3 else 1
    sbytes = oracle.stream_bytes((thick,) + tuple(shape[1:]), mb)
    st = np.zeros(nslab * sbytes, dtype=np.uint8)
    dec = np.empty_like(h)
    oracle.lib().zfo_roundtrip_slabs(code, d, nx, ny, nz, mb, h.ctypes.data, st.ctypes.data, dec.ctypes.data, nslab, 0,
                                     None, None)
    return st, dec


def _bench_field(shape, dtype, kind):
    """The very field bench.py times (rank 0)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("zfb_bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    torch = _torch()
    return b.make_field(torch, shape, dtype, kind, 0, torch.device("cuda", 0))


def test_bench_workload_nyx_slab_against_oracle(codec, pkg, oracle):
    """The headline workload itself (bench.py nyx2048-slab: 256 x 2048 x 2048 float32, rate 8, rank 0's field), stream
    and decoded values bit for bit against the oracle on the host cores, fused metrics against the decoded pair."""
    torch = _torch()
    shape = (256, 2048, 2048)
    f = _bench_field(shape, "float32", "lognormal")
    mb = pkg.maxbits(np.float32, 3, 8)
    stream = codec.encode(f, mb)
    out = codec.decode(stream, shape, torch.float32, mb, orig=f)
    v = pkg.metric_values(codec.partials())
    h = f.cpu().numpy()
    st, dec = _oracle_slabs(oracle, h, mb, 64)
    assert np.array_equal(stream.cpu().numpy(), st)
    del st
    assert_bits_equal(out.cpu().numpy(), dec, "slab decode")
    assert v["absolute_error"] == float(np.abs(h - dec).max())
    assert v["max"] == float(h.max()) and v["min"] == float(h.min())
    m = oracle.metrics(h[:16].ravel(), dec[:16].ravel())  # formulas on a piece; sums over the whole field below
    assert m["absolute_error"] <= v["absolute_error"]
    d = (h - dec).astype(np.float64)
    mse = float(np.mean(d * d))
    assert abs(v["mse"] - mse) <= REL * mse


def test_bench_workload_hacc_1d_against_oracle(codec, pkg, oracle):
    """bench.py hacc1d (BASELINE configs[1]): 268 435 456 float32, "bits" 8 and 16 (both 64-bit slots under CBench's
    wra, zfpCompressorGpu.hpp:129) and the unaligned 32-bit slots of rate 8 without wra."""
    torch = _torch()
    n = 268435456
    f = _bench_field((n,), "float32", "uniform")
    h = f.cpu().numpy()
    for rate, wra in ((8, 1), (16, 1), (8, 0)):
        mb = pkg.maxbits(np.float32, 1, rate, wra)
        stream = codec.encode(f, mb)
        out = codec.decode(stream, (n,), torch.float32, mb, orig=f)
        v = pkg.metric_values(codec.partials())
        st, dec = _oracle_slabs(oracle, h, mb, 64)
        assert np.array_equal(stream.cpu().numpy(), st), (rate, wra)
        del st, stream
        assert_bits_equal(out.cpu().numpy(), dec, ("hacc decode", rate, wra))
        assert v["absolute_error"] == float(np.abs(h - dec).max())
        del out, dec


def test_bench_workload_f64_slab_against_oracle(codec, pkg, oracle):
    """bench.py vpic1024-f64 (BASELINE configs[4]): 256 x 1024 x 1024 float64 at three rates of the sweep."""
    torch = _torch()
    shape = (256, 1024, 1024)
    f = _bench_field(shape, "float64", "smooth")
    h = f.cpu().numpy()
    for rate in (1, 8, 32):
        mb = pkg.maxbits(np.float64, 3, rate)
        stream = codec.encode(f, mb)
        out = codec.decode(stream, shape, torch.float64, mb, orig=f)
        v = pkg.metric_values(codec.partials())
        st, dec = _oracle_slabs(oracle, h, mb, 64)
        assert np.array_equal(stream.cpu().numpy(), st), rate
        del st, stream
        assert_bits_equal(out.cpu().numpy(), dec, ("f64 decode", rate))
        assert v["absolute_error"] == float(np.abs(h - dec).max()), rate
        del out, dec


def test_bad_arguments_return_einval(codec, pkg):
    import ctypes as C
    torch = _torch()
    t = torch.zeros(64, device="cuda")
    L = pkg.lib()
    assert L.zfb_encode_dev(codec._h, 1, 4, 4, 4, 4, 512, t.data_ptr(), t.data_ptr()) == -1   # dims 4
    assert L.zfb_encode_dev(codec._h, 3, 1, 64, 1, 1, 64, t.data_ptr(), t.data_ptr()) == -1   # dtype
    assert L.zfb_encode_dev(codec._h, 1, 1, 64, 1, 1, 8, t.data_ptr(), t.data_ptr()) == -1    # maxbits < 1+EBITS
    assert L.zfb_encode_dev(codec._h, 1, 1, 64, 1, 1, 64, None, t.data_ptr()) == -1


F64_TILE_SHAPES = [(8, 8, 256), (8, 8, 1028), (36, 40, 72), (4, 4, 4), (7, 9, 12), (9, 130, 520), (16, 64, 512)]


@pytest.mark.parametrize("shape", F64_TILE_SHAPES)
@pytest.mark.parametrize("rate", [1, 4, 8, 16, 32])
def test_f64_tile_kernels_bit_exact(codec, pkg, oracle, shape, rate):
    """3D double fields (nx % 4 == 0) take the f64 TMA tile kernels; metrics for doubles are computed in
    double (documented deviation: the reference reinterprets the bytes as float)."""
    f = np.exp(1.2 * smooth_field(shape, seed=rate + 3, dtype=np.float64))
    mb = pkg.maxbits(np.float64, 3, rate)
    s_ref = oracle.compress(f, mb)
    g_ref = oracle.decompress(s_ref, shape, np.float64, mb)
    s, g, part = gpu_roundtrip(codec, f, mb)
    assert np.array_equal(s, s_ref)
    assert_bits_equal(g, g_ref, "decoded")
    d = f - g_ref
    v = pkg.metric_values(part)
    assert v["n"] == f.size
    assert v["absolute_error"] == float(np.abs(d).max())
    assert v["max"] == float(f.max()) and v["min"] == float(f.min())
    mse = float(np.mean(d * d))
    assert abs(v["mse"] - mse) <= 1e-9 * mse + 1e-300
    s2, g2, _ = gpu_roundtrip(codec, f, mb, with_metrics=False)
    assert_bits_equal(g2, g_ref, "decoded (no metrics)")


def test_nyx_2048_full_size_properties(codec, pkg):
    """BASELINE's headline field at FULL size on one GPU (2048^3 float32 = 32 GiB, ~70 GiB resident): the
    size-independent properties -- the whole-field stream is the concatenation of the 8 slab streams, slab
    decodes equal the whole-field decode, slab metric partials combine to the whole-field partials, and the
    error obeys the fixed-rate bound seen on small fields."""
    torch = _torch()
    free, total = torch.cuda.mem_get_info()
    if free < 76 * 2 ** 30:
        pytest.skip("needs ~76 GiB of free device memory")
    n = 2048
    g = torch.Generator(device="cuda").manual_seed(7)
    f = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
    xs = torch.arange(n, device="cuda", dtype=torch.float32) / n
    for z0 in range(0, n, 64):
        zz = (torch.arange(z0, z0 + 64, device="cuda", dtype=torch.float32) / n)[:, None, None]
        a = torch.sin(6.2831853 * (3 * zz + 2 * xs[None, :, None] + 5 * xs[None, None, :]))
        a = a + 0.5 * torch.cos(6.2831853 * (7 * zz - 4 * xs[None, :, None]))
        a = a + 0.05 * torch.randn(a.shape, generator=g, device="cuda")
        f[z0:z0 + 64] = torch.exp(1.5 * a)
        del a
    mb = pkg.maxbits(np.float32, 3, 8)
    stream = codec.encode(f, mb)
    assert stream.numel() == (n // 4) ** 3 * 64
    out = codec.decode(stream, (n, n, n), torch.float32, mb, orig=f)
    whole = codec.partials()
    sl = n // 8
    parts = []
    sbytes = stream.numel() // 8
    for r in range(8):
        fs = f[r * sl:(r + 1) * sl]
        ss = codec.encode(fs, mb)
        assert torch.equal(ss, stream[r * sbytes:(r + 1) * sbytes]), r          # slab-concatenation identity
        os_ = codec.decode(ss, (sl, n, n), torch.float32, mb, orig=fs)
        assert torch.equal(os_, out[r * sl:(r + 1) * sl]), r
        parts.append(codec.partials())
        del ss, os_
    comb = pkg.combine_partials(parts)
    assert comb.n == whole.n == n ** 3
    for k in ("max_abs", "max_rel", "max_orig", "neg_min_orig"):
        assert getattr(comb, k) == getattr(whole, k), k
    for k in ("sum_sq", "sum_abs", "sum_rel"):
        assert abs(getattr(comb, k) - getattr(whole, k)) <= 1e-9 * abs(getattr(whole, k)), k
    v = pkg.metric_values(whole)
    assert v["absolute_error"] == float((f - out).abs().max().item())
    assert v["psnr"] > 60.0


GUARD_CASES = [("float32", (64, 64, 64), 8), ("float32", (36, 40, 72), 8), ("float32", (13, 9, 7), 8),
               ("float32", (12, 8, 516), 4), ("float32", (4099,), 16), ("float32", (1 << 16,), 8),
               ("float32", (33, 70), 8), ("float64", (16, 16, 64), 8), ("float64", (9, 10, 11), 16),
               ("float64", (1001,), 16), ("float32", (20, 20, 20), 5.5)]


@pytest.mark.parametrize("dtype,shape,rate", GUARD_CASES)
def test_outputs_stay_inside_exact_size_buffers(codec, pkg, dtype, shape, rate):
    """Every kernel family writes exactly stream_bytes / field bytes: guard bands on both sides stay intact
    (the pool has no compute-sanitizer, so this is the out-of-bounds check)."""
    torch = _torch()
    G = 4096
    tdt = getattr(torch, dtype)
    mb = pkg.maxbits(dtype, len(shape), rate)
    nbytes = pkg.stream_bytes(shape, mb)
    g = torch.Generator(device="cuda").manual_seed(5)
    f = torch.randn(shape, dtype=tdt, device="cuda", generator=g)
    sbuf = torch.full((nbytes + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
    s = codec.encode(f, mb, out=sbuf[G:G + nbytes])
    assert s.data_ptr() == sbuf.data_ptr() + G
    fbytes = f.numel() * f.element_size()
    obuf = torch.full((fbytes + 2 * G,), 0x5A, dtype=torch.uint8, device="cuda")
    out = obuf[G:G + fbytes].view(tdt).view(shape)
    codec.decode(s, shape, tdt, mb, orig=f, out=out)
    codec.synchronize()
    assert bool((sbuf[:G] == 0xA5).all()) and bool((sbuf[G + nbytes:] == 0xA5).all()), "stream guard overwritten"
    assert bool((obuf[:G] == 0x5A).all()) and bool((obuf[G + fbytes:] == 0x5A).all()), "field guard overwritten"
    # and the guarded result equals the unguarded one
    s2 = codec.encode(f, mb)
    assert torch.equal(s, s2)
    ref = codec.decode(s2, shape, tdt, mb)
    assert torch.equal(out.contiguous().view(torch.uint8), ref.view(torch.uint8)), "guarded decode"


def test_cbench_brick_partition_against_oracle(codec, pkg, oracle):
    """CBench-parity sharding (getPartition, dataLoaderInterface.hpp:123-215): the 8 ranks of a 2x2x2 split each
    compress their own brick as a contiguous local array; every brick matches the oracle bit for bit (odd brick
    extents: partial blocks), and the ranks' partials combine to the metrics of the whole field."""
    torch = _torch()
    ext = (40, 52, 60)  # 2x2x2 bricks of 20 x 26 x 30
    f = smooth_field(ext, seed=21, amp=25.0)
    mb = pkg.maxbits(np.float32, 3, 8)
    parts, g_all = [], np.empty_like(f)
    for r in range(8):
        cnt, off = pkg.partition(r, 8, ext)
        sl = tuple(slice(o, o + c) for o, c in zip(off, cnt))
        brick = np.ascontiguousarray(f[sl])
        d_b = torch.from_numpy(brick).cuda()
        s = codec.encode(d_b, mb)
        s_ref = oracle.compress(brick, mb)
        assert np.array_equal(s.cpu().numpy(), s_ref), r
        out = codec.decode(s, brick.shape, torch.float32, mb, orig=d_b)
        g_ref = oracle.decompress(s_ref, brick.shape, np.float32, mb)
        assert_bits_equal(out.cpu().numpy(), g_ref, "brick %d" % r)
        parts.append(codec.partials())
        g_all[sl] = g_ref
    total = pkg.combine_partials(parts)
    check_metrics(pkg, oracle, total, f, g_all, "8 bricks combined")
