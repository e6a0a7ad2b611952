"""pytest wiring: the `gpu` marker, import paths, and shared field generators."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def smooth_field(shape, seed=0, dtype=np.float32, amp=1.0):
    """Deterministic smooth-ish field: a few low-k Fourier modes + 5% white noise."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(s, dtype=np.float64) / max(s, 1) for s in shape], indexing="ij")
    g = np.zeros(shape, dtype=np.float64)
    for _ in range(8):
        k = rng.integers(1, 5, size=len(shape))
        ph = rng.uniform(0, 2 * np.pi)
        arg = sum(2 * np.pi * kk * gg for kk, gg in zip(k, grids)) + ph
        g += rng.standard_normal() * np.cos(arg)
    g += 0.05 * rng.standard_normal(shape)
    return (amp * g).astype(dtype)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g

    return g.load_package()


@pytest.fixture(scope="session")
def codec(pkg):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Codec(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O
