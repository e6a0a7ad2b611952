"""The loaders / writers either side of the hot path (csrc/zfb_io.cuh behind the C ABI): GenericIO particle files
(HACC), contiguous HDF5 datasets (NYX) and raw arrays (VPIC .gda), on the reference's own toy inputs
(testing/data/m000.full.mpicosmo.50, z255_32.h5; copied to tests/golden by make_golden.py).  The CPU part covers the
parsers, the CRC-64 and the rank split; the GPU part (-m gpu) the file -> device paths and CBench's config #1."""
import os
import struct

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(os.path.dirname(__file__), "golden")
GIO = os.path.join(GOLD, "m000.full.mpicosmo.50")
H5 = os.path.join(GOLD, "z255_32.h5")
REF = "/root/reference/testing/data"


def test_genericio_header_and_columns(pkg):
    g = pkg.GenericIOFile(GIO)
    assert g.header_crc_ok                      # GenericIO.cxx:968-969: CRC over header + stored CRC == -1
    names = [v["name"] for v in g.vars]
    assert names[:6] == ["x", "y", "z", "vx", "vy", "vz"]      # inputs/hacc/hacc_cbench_test.json "scalars"
    assert g.nelems == sum(g.rank_elems) == 32768
    assert g.nranks == 32
    assert g.phys_scale[0] == 256.0
    # -np 2 split of HACCDataLoader.hpp:189-194: equal shares of the file ranks, 16 360 / 16 408 particles
    (a0, a1), (b0, b1) = g.rank_range(0, 2), g.rank_range(1, 2)
    assert (a0, a1, b0, b1) == (0, 16, 16, 32)
    assert (sum(g.rank_elems[a0:a1]), sum(g.rank_elems[b0:b1])) == (16360, 16408)
    # every column, every block CRC verified on the way
    raw = open(GIO, "rb").read()
    ranks_start = struct.unpack_from("<Q", raw, 88)[0]
    for k, v in enumerate(g.vars):
        col = g.read(v["name"])
        assert col.size == 32768 and col.dtype.itemsize == v["size"]
        # independent check against the bytes in the file: rank r's block of variable k
        pos = 0
        for r in range(g.nranks):
            n_r = struct.unpack_from("<Q", raw, ranks_start + 48 * r + 24)[0]
            start = struct.unpack_from("<Q", raw, ranks_start + 48 * r + 32)[0]
            off = start + sum(n_r * w["size"] + 8 for w in g.vars[:k])
            assert col[pos:pos + n_r].tobytes() == raw[off:off + n_r * v["size"]], (v["name"], r)
            pos += n_r
    x = g.read("x")
    assert 0.0 <= float(x.min()) and float(x.max()) < 256.0
    # the fixture round 1 tests were built from: the first file rank's x, tiled
    old = np.fromfile(os.path.join(GOLD, "hacc_m000_x_4096.f32"), dtype=np.float32)
    n0 = g.rank_elems[0]
    assert np.array_equal(old[:min(n0, 4096)], x[:min(n0, 4096)])
    # split read = slices of the whole column
    assert np.array_equal(g.read("vx", 0, 16), g.read("vx")[:16360])
    assert np.array_equal(g.read("vx", 16, 32), g.read("vx")[16360:])
    g.close()


def test_genericio_fixture_is_the_reference_file():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    assert open(GIO, "rb").read() == open(os.path.join(REF, "m000.full.mpicosmo.50"), "rb").read()
    assert open(H5, "rb").read() == open(os.path.join(REF, "z255_32.h5"), "rb").read()


def test_genericio_crc_detects_corruption(pkg, tmp_path):
    raw = bytearray(open(GIO, "rb").read())
    g = pkg.GenericIOFile(GIO)
    ranks_start = struct.unpack_from("<Q", raw, 88)[0]
    start0 = struct.unpack_from("<Q", raw, ranks_start + 32)[0]
    g.close()
    raw[start0 + 5] ^= 0x40          # one bit of rank 0's x block
    p = tmp_path / "bad.gio"
    p.write_bytes(bytes(raw))
    g = pkg.GenericIOFile(str(p))
    with pytest.raises(pkg.ZfbError):
        g.read("x")
    assert g.read("x", check_crc=False).size == 32768      # "CRC64 skip": the bytes are still readable
    assert g.read("y").size == 32768                        # other blocks are intact
    g.close()
    raw[start0 + 5] ^= 0x40
    raw[100] ^= 1                    # header corruption: reported, not fatal
    p.write_bytes(bytes(raw))
    assert not pkg.GenericIOFile(str(p)).header_crc_ok


def test_genericio_write_back_round_trip(pkg, tmp_path):
    """HACCDataLoader::writeData for one rank: decompressed columns beside untouched ones, readable again with CRCs."""
    g = pkg.GenericIOFile(GIO)
    cols = {v["name"]: g.read(v["name"]) for v in g.vars}
    cols["x"] = (cols["x"] + np.float32(0.001)).astype(np.float32)      # stands for a decompressed scalar
    out = str(tmp_path / "zfp__m000.full.mpicosmo.50")
    pkg.gio_write(out, cols, origin=g.phys_origin, scale=g.phys_scale)
    h = pkg.GenericIOFile(out)
    assert h.header_crc_ok and h.nranks == 1 and h.nelems == 32768
    assert [v["name"] for v in h.vars] == [v["name"] for v in g.vars]
    assert [v["size"] for v in h.vars] == [v["size"] for v in g.vars]
    assert h.phys_scale == g.phys_scale and h.phys_origin == g.phys_origin
    for k in cols:
        assert np.array_equal(h.read(k), cols[k]), k           # read() verifies each block's CRC
    xv = [v for v in h.vars if v["name"] == "x"][0]
    assert xv["flags"] & 4 and xv["is_float"]
    h.close()
    g.close()


def test_hdf5_contiguous_dataset_lookup(pkg):
    """/native_fields/<name> of the toy NYX file: six 32^3 float32 datasets; offsets as in SURVEY.md 8c."""
    want = {"baryon_density": 4712, "dark_matter_density": 137832, "temperature": 268904, "velocity_x": 399976,
            "velocity_y": 531048, "velocity_z": 662120}
    raw = open(H5, "rb").read()
    for name, off in want.items():
        d = pkg.h5_dataset(H5, "/native_fields/" + name)
        assert d["shape"] == (32, 32, 32) and d["elem_size"] == 4 and d["is_float"] and d["bytes"] == 131072
        assert d["offset"] == off
    rho = np.frombuffer(raw, dtype="<f4", count=32768, offset=want["baryon_density"])
    assert np.array_equal(rho, np.fromfile(os.path.join(GOLD, "nyx_z255_32_baryon_density.f32"), dtype=np.float32))
    with pytest.raises(pkg.ZfbError):
        pkg.h5_dataset(H5, "/native_fields/no_such_field")
    with pytest.raises(pkg.ZfbError):
        pkg.h5_dataset(GIO, "/native_fields/baryon_density")


# ------------------------------------------------------------------------------------------------ GPU
def _partition_slices(pkg, rank, world, ext):
    cnt, off = pkg.partition(rank, world, ext)
    return tuple(off), tuple(cnt)


@pytest.mark.gpu
def test_genericio_columns_land_on_the_device(pkg):
    codec = pkg.Codec(0)
    g = pkg.GenericIOFile(GIO)
    for name in ("x", "vz", "id"):
        if g.find(name) < 0:
            continue
        for r0, r1 in ((0, 32), (0, 16), (16, 32), (3, 4)):
            d = codec.gio_read_dev(g, name, r0, r1)
            assert np.array_equal(d.cpu().numpy().view(np.uint8), g.read(name, r0, r1).view(np.uint8)), (name, r0, r1)
    g.close()
    codec.close()


@pytest.mark.gpu
def test_hdf5_and_raw_hyperslabs_land_on_the_device(pkg, tmp_path):
    """NYXDataLoader's hyperslab per MPI rank (getPartition: slabs for 2 ranks, bricks for 4 and 8) and a raw .gda
    array, read straight into device memory; then the decompressed write-back into a copy of the file."""
    import shutil
    import torch

    codec = pkg.Codec(0)
    d = pkg.h5_dataset(H5, "/native_fields/temperature")
    full = np.frombuffer(open(H5, "rb").read(), dtype="<f4", count=32768, offset=d["offset"]).reshape(32, 32, 32)
    for world in (1, 2, 4, 8):
        for rank in range(world):
            off, cnt = _partition_slices(pkg, rank, world, d["shape"])
            t = codec.read_block_dev(H5, d["offset"], torch.float32, d["shape"], off, cnt)
            want = full[off[0]:off[0] + cnt[0], off[1]:off[1] + cnt[1], off[2]:off[2] + cnt[2]]
            assert np.array_equal(t.cpu().numpy(), want), (world, rank)
    # raw array (VPIC .gda): the same call at offset 0, float64
    rng = np.random.default_rng(5)
    a = rng.standard_normal((12, 20, 36))
    gda = str(tmp_path / "field.gda")
    a.tofile(gda)
    t = codec.read_block_dev(gda, 0, torch.float64, a.shape, (4, 8, 12), (8, 12, 24))
    assert np.array_equal(t.cpu().numpy(), a[4:12, 8:20, 12:36])
    # write-back: every rank's brick of the "decompressed" field into a copy of the HDF5 file (NYX) -> the dataset is
    # replaced, everything else in the file is untouched
    out = str(tmp_path / "zfp__z255_32.h5")
    shutil.copyfile(H5, out)
    new = (full * np.float32(1.5)).astype(np.float32)
    for rank in range(8):
        off, cnt = _partition_slices(pkg, rank, 8, d["shape"])
        brick = torch.from_numpy(np.ascontiguousarray(new[off[0]:off[0] + cnt[0], off[1]:off[1] + cnt[1], off[2]:off[2] + cnt[2]])).cuda()
        codec.write_block_dev(out, d["offset"], brick, d["shape"], off)
    raw0, raw1 = open(H5, "rb").read(), open(out, "rb").read()
    assert len(raw0) == len(raw1)
    assert raw1[d["offset"]:d["offset"] + d["bytes"]] == new.tobytes()
    assert raw1[:d["offset"]] == raw0[:d["offset"]] and raw1[d["offset"] + d["bytes"]:] == raw0[d["offset"] + d["bytes"]:]
    assert pkg.h5_dataset(out, "/native_fields/temperature")["offset"] == d["offset"]
    codec.close()


@pytest.mark.gpu
def test_config_1_toy_hacc_all_scalars_two_ranks(pkg, oracle):
    """BASELINE configs[0] as written (inputs/hacc/hacc_cbench_test.json): all six scalars of the toy file, split over
    two MPI ranks as HACCDataLoader does (16 360 / 16 408 particles), zfp "abs": 1E-2 (the file's setting) and the
    fixed-rate "bits" the GPU plugin adds, absolute_error + minmax histograms for x and vx, partials combined over the
    ranks.  Columns go file -> device; every stream and decoded value against the oracle."""
    import torch

    codec = pkg.Codec(0)
    g = pkg.GenericIOFile(GIO)
    for name in ("x", "y", "z", "vx", "vy", "vz"):
        whole = g.read(name)
        for setting in ("abs", "bits"):
            parts, decs = [], []
            for rank in range(2):
                r0, r1 = g.rank_range(rank, 2)
                d_col = codec.gio_read_dev(g, name, r0, r1)
                h_col = g.read(name, r0, r1)
                assert d_col.numel() == (16360, 16408)[rank]
                if setting == "abs":
                    mode, prm = pkg.mode_accuracy(1e-2), oracle.params_accuracy(1e-2)
                    s, _ = codec.encode_var(d_col, mode)
                    out = codec.decode_var(s, h_col.shape, torch.float32, mode, orig=d_col)
                    s_ref = oracle.compress_ex(h_col, prm)
                    g_ref, _ = oracle.decompress_ex(s_ref, h_col.shape, np.float32, prm)
                else:
                    mb = pkg.maxbits(np.float32, 1, 8)          # "bits": 8 -> 64-bit slots under CBench's wra
                    s = codec.encode(d_col, mb)
                    out = codec.decode(s, h_col.shape, torch.float32, mb, orig=d_col)
                    s_ref = oracle.compress(h_col, mb)
                    g_ref = oracle.decompress(s_ref, h_col.shape, np.float32, mb)
                assert np.array_equal(s.cpu().numpy(), s_ref), (name, setting, rank)
                assert np.array_equal(out.cpu().numpy().view(np.uint32), g_ref.view(np.uint32)), (name, setting, rank)
                parts.append(codec.partials())
                decs.append((d_col, out, g_ref))
            comb = pkg.combine_partials(parts)              # what MPI_Allreduce / zfb_allreduce_partials produce
            v = pkg.metric_values(comb)
            g_all = np.concatenate([t[2] for t in decs])
            m = oracle.metrics(whole, g_all, hist=True)
            for k in ("absolute_error", "relative_error", "mse", "psnr", "minmax"):
                assert abs(v[k] - m[k]) <= 1e-6 * abs(m[k]) + 1e-300, (name, setting, k, v[k], m[k])
            if setting == "abs":
                assert v["absolute_error"] <= 1e-2
            if name in ("x", "vx"):                          # "histogram": ["x", "vx"] for absolute_error and minmax
                cnt_abs = np.zeros(1024, dtype=np.int64)
                cnt_mm = np.zeros(1024, dtype=np.int64)
                for d_col, out, _ in decs:
                    c, _oob = codec.histogram(0, 0.0, v["absolute_error"], d_col, out)
                    cnt_abs += c.cpu().numpy().astype(np.int64)
                    c, _oob = codec.histogram(2, v["min"], v["max"], None, out)
                    cnt_mm += c.cpu().numpy().astype(np.int64)
                assert np.array_equal(cnt_abs.astype(np.float32) / np.float32(whole.size), m["abs_hist"]), (name, setting)
                assert np.array_equal(cnt_mm.astype(np.float32), m["minmax_hist"]), (name, setting)
    g.close()
    codec.close()
