"""CPU-side checks of the C ABI: the library loads, exports every symbol include/zfb.h declares, its
pure-host arithmetic matches the oracle, and compute entry points refuse to run without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.LIB_PATH), "run __graft_entry__.build() first"
    L = pkg.lib()
    syms = pkg.declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), s


def test_no_torch_types_in_header():
    text = open(os.path.join(ROOT, "include", "zfb.h")).read()
    assert "torch" not in text.lower().replace("no c++/torch types", "")
    assert 'extern "C"' in text


def test_maxbits_and_sizes_match_oracle(pkg, oracle):
    for dt in (np.float32, np.float64):
        for dims in (1, 2, 3):
            for rate in (0.5, 1, 3, 4, 7.5, 8, 14, 16, 32):
                for wra in (0, 1, 8):
                    assert pkg.maxbits(dt, dims, rate, wra) == oracle.maxbits(dt, dims, rate, wra)
    for shape in [(10,), (4099,), (5, 7), (64, 64), (5, 6, 7), (33, 33, 33), (256, 2048, 2048)]:
        for mb in (9, 64, 100, 512, 1024):
            assert pkg.stream_bytes(shape, mb) == oracle.stream_bytes(shape, mb)
            assert pkg.stream_capacity(shape, mb) == oracle.stream_capacity(shape, mb)


def _ref_partition(rank, nranks, ext):
    """Transcription of the splitting rule of getPartition (dataLoaderInterface.hpp:123-215) in Python."""
    parts = [((0, 0, 0), tuple(ext))]
    axis = 0
    while len(parts) < nranks:
        cur = len(parts)
        for _ in range(cur):
            lo, hi = parts.pop(0)
            if hi[axis] - lo[axis] <= 1:
                parts.append((lo, hi))
                continue
            mid = (lo[axis] + hi[axis]) // 2
            h1 = list(hi); h1[axis] = mid
            l2 = list(lo); l2[axis] = mid
            parts.append((lo, tuple(h1)))
            parts.append((tuple(l2), hi))
            if len(parts) >= nranks:
                break
        axis = (axis + 1) % 3
    lo, hi = parts[rank]
    return tuple(h - l for l, h in zip(lo, hi)), lo


def test_partition_matches_cbench_rule(pkg):
    for ext in [(2048, 2048, 2048), (32, 32, 32), (512, 256, 100), (7, 5, 3)]:
        for nranks in (1, 2, 4, 8, 16):
            for r in range(nranks):
                assert pkg.partition(r, nranks, ext) == _ref_partition(r, nranks, ext)
    # 2 ranks: slabs along the slowest axis; 8 ranks: 2x2x2 bricks
    assert pkg.partition(1, 2, (2048, 2048, 2048)) == ((1024, 2048, 2048), (1024, 0, 0))
    assert pkg.partition(7, 8, (2048, 2048, 2048)) == ((1024, 1024, 1024), (1024, 1024, 1024))


def test_metric_value_formulas(pkg):
    p = pkg.Partials(sum_sq=4.0, sum_abs=6.0, sum_rel=3.0, max_abs=1.5, max_rel=0.75, max_orig=10.0,
                     neg_min_orig=2.0, n=16)
    v = pkg.metric_values(p)
    assert v["absolute_error"] == 1.5 and v["relative_error"] == 0.75
    assert v["mse"] == 0.25
    assert abs(v["psnr"] - 10 * np.log10(100.0 / 0.25)) < 1e-12   # uses max of original, psnrError.hpp:85
    assert v["minmax"] == 10.0 and v["min"] == -2.0
    q = pkg.Partials(sum_sq=1.0, sum_abs=1.0, sum_rel=1.0, max_abs=2.5, max_rel=0.1, max_orig=3.0,
                     neg_min_orig=7.0, n=4)
    c = pkg.combine_partials([p, q])
    assert (c.sum_sq, c.sum_abs, c.sum_rel, c.max_abs, c.max_rel, c.max_orig, c.neg_min_orig, c.n) == \
        (5.0, 7.0, 4.0, 2.5, 0.75, 10.0, 7.0, 20)


def test_compute_refuses_without_gpu(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ZfbError):
        pkg.Codec(0)          # ZFB_ENODEV: there is no CPU fallback
    h = C.c_void_p()
    assert pkg.lib().zfb_create(C.byref(h), 0) == -2


def test_product_path_never_touches_the_oracle():
    """Nothing under the package directory or include/ may import, link or call oracle/ code."""
    bad = []
    for base in ("vizaly-foresight_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if "liboracle" in t or "zfo_" in t or "import oracle" in t or "oracle/" in t:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_mode_parameters_and_maximum_size_match_oracle(pkg, oracle):
    """zfp_stream_set_accuracy / set_precision / set_rate and zfp_stream_maximum_size (pure host arithmetic)."""
    for tol in (0.0, 1e-6, 0.003, 0.05, 0.5, 1.0, 3.0, 1000.0, 2.0 ** -20, 2.0 ** 10):
        assert pkg.mode_accuracy(tol).as_tuple() == oracle.params_accuracy(tol).as_tuple(), tol
    for prec in (0, 1, 8, 16, 23, 32, 40, 64, 100):
        assert pkg.mode_precision(prec).as_tuple() == oracle.params_precision(prec).as_tuple(), prec
    for dt in (np.float32, np.float64):
        for dims, rate in ((1, 8), (2, 4.5), (3, 8), (3, 16)):
            assert pkg.mode_rate(dt, dims, rate).as_tuple() == oracle.params_rate(dt, dims, rate).as_tuple()
        for shape in [(10,), (4099,), (5, 7), (64, 64), (5, 6, 7), (33, 33, 33)]:
            for m, p in ((pkg.mode_accuracy(0.01), oracle.params_accuracy(0.01)),
                         (pkg.mode_precision(12), oracle.params_precision(12)),
                         (pkg.mode_rate(dt, len(shape), 8), oracle.params_rate(dt, len(shape), 8))):
                assert pkg.maximum_size(shape, dt, m) == oracle.maximum_size(shape, dt, p)
    assert pkg.block_count((5, 6, 7)) == 2 * 2 * 2 and pkg.block_count((4099,)) == 1025
    with pytest.raises(pkg.ZfbError):
        pkg.mode_accuracy(-1.0)


def test_extra_metric_values_from_partials(pkg):
    """PSNR over the value range and NRMSE come out of the same partials as the five CBench metrics."""
    import math

    p = pkg.Partials(4.0e-3, 1.0, 1.0, 0.5, 0.5, 10.0, 2.0, 1000)      # max 10, min -2, sum of squares 4e-3 over 1000
    v = pkg.metric_values(p)
    mse = 4.0e-6
    assert abs(v["mse"] - mse) < 1e-18
    assert abs(v["psnr"] - 10 * math.log10(100.0 / mse)) < 1e-9          # CBench: max(original)^2 / mse
    assert abs(v["psnr_range"] - 10 * math.log10(144.0 / mse)) < 1e-9    # (max - min)^2 / mse
    assert abs(v["nrmse"] - math.sqrt(mse) / 12.0) < 1e-15
    # the reference's running maximum starts at -999999999 (psnrError.hpp:58): an all-negative field cannot go below it
    q = pkg.Partials(1.0, 1.0, 1.0, 0.5, 0.5, -1.7976931348623157e308, 5.0, 10)
    assert abs(pkg.metric_values(q)["psnr"] - 10 * math.log10(999999999.0 ** 2 / 0.1)) < 1e-9
