"""Seeded random sweep over shapes, types, rates and modes: GPU streams and decoded values against the oracle.
Complements the hand-picked cases of test_gpu_parity.py / test_gpu_variable_rate.py (dispatch corners: tile vs
edge vs generic kernels, unaligned slots, partial blocks in every direction, tiny and ragged extents)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_shape(rng):
    d = int(rng.integers(1, 4))
    if d == 1:
        return (int(rng.integers(1, 5000)),)
    if d == 2:
        return (int(rng.integers(1, 70)), int(rng.integers(1, 90)))
    # 3D: bias towards extents that take the tile path (x a multiple of 4, long rows) but keep ragged ones
    nx = int(rng.choice([int(rng.integers(1, 40)), 4 * int(rng.integers(1, 80)), 256, 260, 512]))
    return (int(rng.integers(1, 20)), int(rng.integers(1, 24)), nx)


def _rand_field(rng, shape, dtype):
    kind = int(rng.integers(0, 5))
    if kind == 0:
        f = rng.standard_normal(shape) * 10.0 ** int(rng.integers(-6, 7))
    elif kind == 1:
        idx = np.indices(shape).sum(0)
        f = np.sin(0.05 * idx) * 100.0 + 0.001 * rng.standard_normal(shape)
    elif kind == 2:
        f = rng.standard_normal(shape) * (rng.random(shape) > 0.7)       # sparse
    elif kind == 3:
        f = np.exp(rng.standard_normal(shape) * 4.0)                      # wide dynamic range
    else:
        f = np.full(shape, rng.standard_normal())                         # constant
    return np.ascontiguousarray(f.astype(dtype))


@pytest.mark.parametrize("chunk", range(6))
def test_random_cases_against_oracle(codec, pkg, oracle, chunk):
    import torch

    rng = np.random.default_rng(20261019 + chunk)
    for case in range(25):
        dtype = np.float32 if rng.random() < 0.65 else np.float64
        shape = _rand_shape(rng)
        f = _rand_field(rng, shape, dtype)
        d_f = torch.from_numpy(f).cuda()
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        what = (chunk, case, dtype.__name__, shape)
        if rng.random() < 0.6:  # fixed rate, CBench's aligned slots or the unaligned extension
            rate = float(rng.choice([1, 2, 3, 4, 5.5, 8, 11, 16, 24, 32]))
            mb = pkg.maxbits(dtype, len(shape), rate, int(rng.random() < 0.8))
            s_ref = oracle.compress(f, mb)
            s = codec.encode(d_f, mb)
            assert np.array_equal(s.cpu().numpy(), s_ref), what + (mb,)
            out = codec.decode(s, shape, tdt, mb, orig=d_f)
            g_ref = oracle.decompress(s_ref, shape, dtype, mb)
            assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), what + (mb,)
        else:
            if rng.random() < 0.5:
                tol = float(10.0 ** rng.uniform(-6, 3))
                m, p = pkg.mode_accuracy(tol), oracle.params_accuracy(tol)
            else:
                prec = int(rng.integers(1, 40))
                m, p = pkg.mode_precision(prec), oracle.params_precision(prec)
            s_ref = oracle.compress_ex(f, p)
            s, nbytes = codec.encode_var(d_f, m)
            assert nbytes == len(s_ref), what + (m.as_tuple(),)
            assert np.array_equal(s.cpu().numpy(), s_ref), what + (m.as_tuple(),)
            out = codec.decode_var(s, shape, tdt, m, orig=d_f)
            g_ref, _ = oracle.decompress_ex(s_ref, shape, dtype, p)
            assert np.array_equal(out.cpu().numpy().view(np.uint8), g_ref.view(np.uint8)), what + (m.as_tuple(),)
