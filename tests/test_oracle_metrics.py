"""oracle/metrics_ref.c against the reference's OWN metric classes (oracle/_ref, built from
CBench/metrics/*.hpp) and against tests/golden/metrics_kat.json generated from that build."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import smooth_field

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("absolute_error", "relative_error", "mse", "psnr", "minmax")


def _pairs():
    rng = np.random.default_rng(7)
    o = rng.standard_normal(5000).astype(np.float32) * 3
    yield "noise", o, (o + 0.01 * rng.standard_normal(o.size)).astype(np.float32)
    o = (rng.standard_normal(3000) * 1e4).astype(np.float32)
    yield "big", o, (o * (1 + 1e-3 * rng.standard_normal(o.size))).astype(np.float32)
    o = np.abs(rng.standard_normal(2048)).astype(np.float32) * 0.5  # mostly |o| < 1: abs-error branch of relError
    yield "small", o, (o + 1e-3).astype(np.float32)
    o = rng.standard_normal(777).astype(np.float32)
    yield "identical", o, o.copy()


@pytest.mark.parametrize("hist", [False, True])
def test_restatement_equals_reference_classes(oracle, hist):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for label, o, a in _pairs():
        mine = oracle.metrics(o, a, hist=hist)
        for name in NAMES:
            r = oracle.ref_metric(name, o, a, hist=hist)
            if np.isnan(r["value"]) or np.isinf(r["value"]):
                assert np.isnan(mine[name]) or np.isinf(mine[name]), (label, name)
            else:
                assert mine[name] == r["value"], (label, name)  # same arithmetic, same order: bit-equal
        if hist:
            for name, key in (("absolute_error", "abs_hist"), ("relative_error", "rel_hist"), ("minmax", "minmax_hist")):
                r = oracle.ref_metric(name, o, a, hist=True)
                if not r["extra"]:
                    assert not mine[key].any()
                    continue
                y, lo, hi = oracle.parse_hist_script(r["extra"])
                if name == "minmax" and mine["minmax_oob"]:
                    continue  # reference indexes out of bounds here (UB); restatement clamps
                # the script prints std::to_string(float): 6 decimals
                assert np.allclose(y, mine[key].astype(np.float64), rtol=0, atol=5.1e-7), (label, name)


def test_restatement_matches_committed_reference_vectors(oracle):
    path = os.path.join(GOLD, "metrics_kat.json")
    kat = json.load(open(path))
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
        m = oracle.metrics(o, a, hist=True)
        for name in NAMES:
            assert m[name] == case[name], (kind, name)
        for name, key in (("absolute_error", "abs_hist"), ("relative_error", "rel_hist")):
            # committed hash is of the values as printed by the reference (6 decimals)
            printed = np.array([float("%f" % v) for v in m[key]], dtype=np.float64)
            assert hashlib.sha256(printed.tobytes()).hexdigest() == case[name + "_hist_sha256"], (kind, name)


def test_rel_error_is_float_rounded(oracle):
    """relError<float> returns T=float (relativeError.hpp:66-75): the quotient is rounded to float."""
    o = np.array([3.0], dtype=np.float32)
    a = np.array([2.0], dtype=np.float32)
    m = oracle.metrics(o, a)
    assert m["relative_error"] == float(np.float32(1.0 / 3.0))
    o = np.array([0.5], dtype=np.float32)  # |o| < 1 -> absolute error
    m = oracle.metrics(o, np.array([0.25], dtype=np.float32))
    assert m["relative_error"] == 0.25


def test_psnr_uses_max_of_original(oracle):
    o = np.array([1.0, -8.0, 2.0, 4.0], dtype=np.float32)
    a = o + np.float32(0.5)
    m = oracle.metrics(o, a)
    assert abs(m["psnr"] - 10 * np.log10(16.0 / 0.25)) < 1e-12
    assert m["min"] == -8.0 and m["max"] == 4.0 and m["minmax"] == 4.0
