"""Generates tests/golden/codec_kat.json, tests/golden/codec_var_kat.json and tests/golden/metrics_kat.json.

* codec_kat.json   -- SHA-256 of stream and decoded bytes from oracle/zfp_ref.c on seeded inputs
                      (regression pins of OUR restatement: the reference ships no codec vectors).
* metrics_kat.json -- metric values produced by the REFERENCE's own classes (oracle/_ref, built from
                      /root/reference/CBench/metrics/*.hpp) on seeded inputs; these do pin parity.
* nyx_z255_32_baryon_density.f32 / hacc_m000_x_4096.f32 -- small real inputs lifted out of the
  reference's toy data files (testing/data/z255_32.h5 offset 4712; m000.full.mpicosmo.50 first
  file-rank x column), so GPU-box tests have real-data inputs without /root/reference.

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = [
    # name, dtype, shape, rate, wra, kind, seed
    ("f32_1d_n10_r16", "float32", [10], 16, 1, "smooth", 1),
    ("f32_1d_n4099_r8", "float32", [4099], 8, 1, "normal900", 2),
    ("f32_1d_n1000_r8_wra0", "float32", [1000], 8, 0, "uniform256", 3),
    ("f32_2d_5x7_r14", "float32", [5, 7], 14, 1, "smooth", 4),
    ("f32_2d_64x64_r8", "float32", [64, 64], 8, 1, "smooth", 5),
    ("f32_3d_5x6x7_r4", "float32", [5, 6, 7], 4, 1, "smooth", 6),
    ("f32_3d_5x6x7_r8", "float32", [5, 6, 7], 8, 1, "smooth", 6),
    ("f32_3d_5x6x7_r16", "float32", [5, 6, 7], 16, 1, "smooth", 6),
    ("f32_3d_33_r8", "float32", [33, 33, 33], 8, 1, "lognormal", 7),
    ("f32_3d_64_r4", "float32", [64, 64, 64], 4, 1, "smooth", 8),
    ("f32_3d_64_r8", "float32", [64, 64, 64], 8, 1, "smooth", 8),
    ("f32_3d_64_r16", "float32", [64, 64, 64], 16, 1, "smooth", 8),
    ("f32_3d_32_noise_r8", "float32", [32, 32, 32], 8, 1, "normal900", 9),
    ("f32_3d_16_zero_r8", "float32", [16, 16, 16], 8, 1, "zero", 0),
    ("f64_3d_5x6x7_r8", "float64", [5, 6, 7], 8, 1, "smooth", 10),
    ("f64_3d_32_r1", "float64", [32, 32, 32], 1, 1, "smooth", 11),
    ("f64_3d_32_r13", "float64", [32, 32, 32], 13, 1, "smooth", 11),
    ("f64_3d_32_r32", "float64", [32, 32, 32], 32, 1, "smooth", 11),
    ("f64_1d_n1001_r16", "float64", [1001], 16, 1, "normal900", 12),
    ("f64_2d_17x19_r9", "float64", [17, 19], 9, 1, "smooth", 13),
]


VAR_CASES = [
    # name, dtype, shape, mode ("abs" tolerance | "rel" precision), value, kind, seed
    ("f32_1d_n4099_abs1e-2", "float32", [4099], "abs", 1e-2, "uniform256", 31),
    ("f32_1d_n1000_rel12", "float32", [1000], "rel", 12, "normal900", 32),
    ("f32_2d_33x70_abs0.5", "float32", [33, 70], "abs", 0.5, "smooth", 33),
    ("f32_3d_5x6x7_abs1e-3", "float32", [5, 6, 7], "abs", 1e-3, "smooth", 34),
    ("f32_3d_33_abs0.05", "float32", [33, 33, 33], "abs", 0.05, "lognormal", 35),
    ("f32_3d_64_rel16", "float32", [64, 64, 64], "rel", 16, "smooth", 36),
    ("f32_3d_32_noise_abs10", "float32", [32, 32, 32], "abs", 10.0, "normal900", 37),
    ("f32_3d_16_zero_abs1", "float32", [16, 16, 16], "abs", 1.0, "zero", 0),
    ("f64_3d_9x10x11_abs1e-6", "float64", [9, 10, 11], "abs", 1e-6, "smooth", 38),
    ("f64_3d_32_rel30", "float64", [32, 32, 32], "rel", 30, "smooth", 39),
    ("f64_1d_n1001_abs1e-3", "float64", [1001], "abs", 1e-3, "normal900", 40),
]


def case_input(case):
    from conftest import smooth_field

    shape = tuple(case["shape"])
    dt = np.dtype(case["dtype"])
    rng = np.random.default_rng(case["seed"])
    k = case["kind"]
    if k == "smooth":
        return smooth_field(shape, seed=case["seed"], dtype=dt, amp=100.0)
    if k == "normal900":
        return (900.0 * rng.standard_normal(shape)).astype(dt)
    if k == "uniform256":
        return (256.0 * rng.random(shape)).astype(dt)
    if k == "lognormal":
        return np.exp(1.5 * smooth_field(shape, seed=case["seed"], dtype=np.float64)).astype(dt)
    if k == "zero":
        return np.zeros(shape, dtype=dt)
    raise ValueError(k)


def main():
    import oracle as O

    O.build(force=True)
    out = {"generator": "tests/golden/make_golden.py", "oracle": "oracle/zfp_ref.c (restatement; parity unpinned)",
           "cases": []}
    for name, dt, shape, rate, wra, kind, seed in CASES:
        case = {"name": name, "dtype": dt, "shape": shape, "rate": rate, "wra": wra, "kind": kind, "seed": seed}
        f = case_input(case)
        mb = O.maxbits(f.dtype, len(shape), rate, wra)
        s = O.compress(f, mb)
        g = O.decompress(s, f.shape, f.dtype, mb)
        case["maxbits"] = mb
        case["stream_bytes"] = int(len(s))
        case["input_sha256"] = hashlib.sha256(f.tobytes()).hexdigest()
        case["stream_sha256"] = hashlib.sha256(s.tobytes()).hexdigest()
        case["decoded_sha256"] = hashlib.sha256(g.tobytes()).hexdigest()
        out["cases"].append(case)
    json.dump(out, open(os.path.join(HERE, "codec_kat.json"), "w"), indent=1)

    # variable-rate modes (zfp fixed accuracy / precision: the CPU zfp plugin's "abs" / "rel")
    vout = {"generator": "tests/golden/make_golden.py", "oracle": "oracle/zfp_ref.c zfo_compress_ex (restatement; parity unpinned)",
            "cases": []}
    for name, dt, shape, mode, value, kind, seed in VAR_CASES:
        case = {"name": name, "dtype": dt, "shape": shape, "mode": mode, "value": value, "kind": kind, "seed": seed}
        f = case_input(case)
        p = O.params_accuracy(value) if mode == "abs" else O.params_precision(value)
        s, offs = O.compress_ex(f, p, with_offsets=True)
        g, _ = O.decompress_ex(s, f.shape, f.dtype, p)
        case["params"] = list(p.as_tuple())
        case["stream_bytes"] = int(len(s))
        case["total_bits"] = int(offs[-1])
        case["input_sha256"] = hashlib.sha256(f.tobytes()).hexdigest()
        case["stream_sha256"] = hashlib.sha256(s.tobytes()).hexdigest()
        case["index_sha256"] = hashlib.sha256(offs.tobytes()).hexdigest()
        case["decoded_sha256"] = hashlib.sha256(g.tobytes()).hexdigest()
        vout["cases"].append(case)
    json.dump(vout, open(os.path.join(HERE, "codec_var_kat.json"), "w"), indent=1)

    # real-data inputs out of the reference's toy files
    ref = "/root/reference/testing/data"
    if os.path.isdir(ref):
        h5 = np.fromfile(os.path.join(ref, "z255_32.h5"), dtype=np.uint8)
        h5[4712:4712 + 131072].tofile(os.path.join(HERE, "nyx_z255_32_baryon_density.f32"))
        gio = open(os.path.join(ref, "m000.full.mpicosmo.50"), "rb").read()
        # GenericIO header (GenericIO.cxx:380-441): locate the first rank's data start from the header
        import struct
        # GlobalHeader: Magic[8], HeaderSize, NElems, Dims[3], NVars, VarsSize, VarsStart, NRanks,
        # RanksSize, RanksStart ...; RankHeader: Coords[3], NElems, Start, GlobalRank
        ranks_start = struct.unpack_from("<Q", gio, 88)[0]
        n0 = struct.unpack_from("<Q", gio, ranks_start + 24)[0]
        start0 = struct.unpack_from("<Q", gio, ranks_start + 32)[0]
        x = np.frombuffer(gio, dtype="<f4", count=int(n0), offset=int(start0))
        reps = int(np.ceil(4096 / max(1, n0)))
        np.tile(x, reps)[:4096].astype(np.float32).tofile(os.path.join(HERE, "hacc_m000_x_4096.f32"))
        # the two toy INPUT files themselves (data, not code): the loaders of csrc/zfb_io.cuh are tested on the real
        # GenericIO / HDF5 byte layouts, also on the GPU box where /root/reference does not exist
        import shutil
        shutil.copyfile(os.path.join(ref, "m000.full.mpicosmo.50"), os.path.join(HERE, "m000.full.mpicosmo.50"))
        shutil.copyfile(os.path.join(ref, "z255_32.h5"), os.path.join(HERE, "z255_32.h5"))

    # metric vectors from the reference's own classes
    if O.ref_lib() is not None:
        mk = {"generator": "tests/golden/make_golden.py", "source": "CBench/metrics/*.hpp via oracle/_ref",
              "cases": []}
        for seed, n, kind in [(21, 4096, "noise"), (22, 20000, "codec8"), (23, 32768, "codec4"), (24, 1000, "big")]:
            rng = np.random.default_rng(seed)
            if kind == "noise":
                o = rng.standard_normal(n).astype(np.float32) * 3
                a = (o + 0.01 * rng.standard_normal(n)).astype(np.float32)
            elif kind == "big":
                o = (rng.standard_normal(n) * 1e4).astype(np.float32)
                a = (o * (1 + 1e-3 * rng.standard_normal(n))).astype(np.float32)
            else:
                from conftest import smooth_field
                side = int(round(n ** (1 / 3)))
                side -= side % 4
                o = smooth_field((side, side, side), seed=seed, amp=50.0)
                mb = 512 if kind == "codec8" else 256
                a = O.decompress(O.compress(o, mb), o.shape, np.float32, mb)
                o, a = o.ravel(), a.ravel()
            entry = {"seed": seed, "n": int(o.size), "kind": kind}
            for name in ("absolute_error", "relative_error", "mse", "psnr", "minmax"):
                r = O.ref_metric(name, o, a, hist=True)
                entry[name] = r["value"]
                entry[name + "_log"] = r["log"]
                if r["extra"]:
                    y, lo, hi = O.parse_hist_script(r["extra"])
                    entry[name + "_hist_sha256"] = hashlib.sha256(
                        np.asarray(y, dtype=np.float64).tobytes()).hexdigest()
                    entry[name + "_hist_range"] = [lo, hi]
            mk["cases"].append(entry)
        json.dump(mk, open(os.path.join(HERE, "metrics_kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
