"""Shell-binned power spectrum (SURVEY 8 f3 extra: the quantity PAT's NYX analysis plots as a ratio between the original
and every decompressed field, Analysis/pat/nyx/cinema.py:84-130).  CPU: the oracle statement against known answers and
the file layout PAT reads.  GPU: zfb_power_spectrum_bin_dev through the C ABI against the oracle."""
import csv
import os

import numpy as np
import pytest

from conftest import smooth_field


def pat_extract_csv_col(filename, delimiter_char, colpos):
    # Analysis/pat/utils/file_utilities.py:95-107, restated (the GPU box has no reference tree)
    with open(filename) as csv_file:
        return [float(row[colpos]) for row in csv.reader(csv_file, delimiter=delimiter_char)]


def test_oracle_plane_wave_and_noise(oracle):
    n = 32
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    f = np.cos(2 * np.pi * (3 * x + 4 * y) / n)  # |k| = 5: two modes (k and -k), amplitude N/2 each
    k, pk, cnt = oracle.power_spectrum(f)
    assert cnt[0] == 1 and cnt.sum() == np.sum(np.floor(np.sqrt(sum(np.meshgrid(*(np.fft.fftfreq(n, 1.0 / n) ** 2,) * 3, indexing="ij"))) + 0.5) < n // 2 + 1)
    assert abs(pk[5] * cnt[5] - 0.5) < 1e-12          # 2 modes x (1/2)^2
    assert np.all(pk[np.arange(len(pk)) != 5] < 1e-20)
    rng = np.random.default_rng(1)
    g = rng.standard_normal((n, n, n))
    k, pk, cnt = oracle.power_spectrum(g)
    assert abs(np.sum(pk * cnt) / np.sum(cnt) * n ** 3 - 1.0) < 0.05  # white noise: flat spectrum, variance / N per mode
    assert np.all(np.abs(k[1:] - np.arange(1, len(k))) < 0.51)


def test_ps3d_file_is_what_pat_reads(oracle, pkg, tmp_path):
    f = smooth_field((16, 16, 16), seed=3, dtype=np.float32)
    k, pk, cnt = oracle.power_spectrum(f)
    path = os.path.join(tmp_path, "orig_rhob_ps3d.txt")
    pkg.write_ps3d(path, k, pk, cnt)
    kk = pat_extract_csv_col(path, " ", 2)
    pp = pat_extract_csv_col(path, " ", 3)
    assert np.allclose(kk, k[cnt > 0]) and np.allclose(pp, pk[cnt > 0])
    ratio = [i / j for i, j in zip(pp, pp)]  # cinema.py:120
    assert all(r == 1.0 for r in ratio if r == r)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(32, 32, 32), (24, 40, 64), (33, 20, 18)])
def test_gpu_power_spectrum_matches_oracle(oracle, pkg, codec, shape):
    import torch

    f = np.exp(smooth_field(shape, seed=5, dtype=np.float64)).astype(np.float32)
    k_ref, pk_ref, cnt_ref = oracle.power_spectrum(f)
    k, pk, cnt = codec.power_spectrum(torch.from_numpy(f).cuda())
    k, pk, cnt = k.cpu().numpy(), pk.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, cnt_ref)                      # every mode in the same shell, mirror modes counted
    assert np.allclose(k, k_ref, rtol=1e-6, atol=1e-6)
    assert np.allclose(pk, pk_ref, rtol=2e-4, atol=1e-12 * pk_ref.max())  # float32 transform against float64


@pytest.mark.gpu
def test_gpu_spectrum_ratio_of_a_decompressed_field(oracle, pkg, codec):
    """The analysis itself: compress at two rates, the spectrum ratio decompressed / original stays near 1 at low k and
    is closer to 1 at the higher rate."""
    import torch

    shape = (64, 64, 64)
    f = np.exp(2.0 * smooth_field(shape, seed=9, dtype=np.float64)).astype(np.float32)
    d_f = torch.from_numpy(f).cuda()
    _, pk0, _ = codec.power_spectrum(d_f)
    err = {}
    for rate in (4, 16):
        mb = pkg.maxbits(np.float32, 3, rate)
        out = codec.decode(codec.encode(d_f, mb), shape, torch.float32, mb)
        _, pk1, _ = codec.power_spectrum(out)
        ratio = (pk1 / pk0).cpu().numpy()
        err[rate] = np.abs(ratio[1:9] - 1.0).max()
        g_ref = oracle.decompress(oracle.compress(f, mb), shape, np.float32, mb)
        _, pk1_ref, _ = oracle.power_spectrum(g_ref)
        _, pk0_ref, _ = oracle.power_spectrum(f)
        assert np.allclose(ratio[1:9], (pk1_ref / pk0_ref)[1:9], rtol=1e-3)
    assert err[16] < err[4] and err[16] < 1e-3
