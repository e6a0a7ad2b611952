#!/usr/bin/env python
"""CPU model of the embedded coder's SIMT cost on the benchmark field (no GPU needed).

For a sub-slab of the synthetic NYX field of bench.py the script computes, per 4x4x4 block, the bit planes of the
zfp transform, then per plane the quantities the coder loops depend on (n before the plane,
position of the last one, code length).  32 x-consecutive blocks form a warp, exactly as in the tile kernels, and the
script compares loop structures: "per plane" (the warp pays the slowest lane in every plane) against "flat" (the warp
pays the slowest lane over the whole block) for different step sizes.

  python tools/coder_model.py [--rate 8] [--nz 8] [--ny 256]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PERM3 = [0, 1, 4, 16, 20, 17, 5, 2, 8, 32, 21, 6, 18, 24, 9, 33, 36, 3, 12, 48, 22, 25, 37, 40, 34, 10, 7, 19, 28, 13, 49, 52,
         41, 38, 26, 23, 29, 53, 11, 35, 44, 14, 50, 56, 42, 27, 39, 45, 30, 54, 57, 60, 51, 15, 43, 46, 58, 61, 55, 31, 62,
         59, 47, 63]


def fwd_lift(v, idx):
    x, y, z, w = (v[:, i].copy() for i in idx)
    x += w; x >>= 1; w -= x
    z += y; z >>= 1; y -= z
    x += z; x >>= 1; z -= x
    w += y; w >>= 1; y -= w
    w += y >> 1; y -= w >> 1
    for i, a in zip(idx, (x, y, z, w)):
        v[:, i] = a


def planes_of(field):
    """field (nz, ny, nx) float32, dims % 4 == 0 -> (nblocks, 32) uint64 bit planes (plane k of block b), block order z, y, x"""
    nz, ny, nx = field.shape
    b = field.reshape(nz // 4, 4, ny // 4, 4, nx // 4, 4).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 64)
    mx = np.abs(b).max(axis=1)
    e = np.where(mx > 0, np.frexp(mx)[1], -127).astype(np.int32)
    e = np.maximum(e, -126)
    s = np.ldexp(np.float32(1), (30 - e)).astype(np.float32)
    with np.errstate(all="ignore"):
        iv = (b * s[:, None]).astype(np.float32).astype(np.int32)
    for z in range(4):
        for y in range(4):
            fwd_lift(iv, [16 * z + 4 * y + i for i in range(4)])
    for x in range(4):
        for z in range(4):
            fwd_lift(iv, [16 * z + 4 * i + x for i in range(4)])
    for y in range(4):
        for x in range(4):
            fwd_lift(iv, [16 * i + 4 * y + x for i in range(4)])
    u = ((iv[:, PERM3].astype(np.int64) + 0xaaaaaaaa) & 0xffffffff) ^ 0xaaaaaaaa
    u = u.astype(np.uint64)
    pl = np.zeros((u.shape[0], 32), dtype=np.uint64)
    for k in range(32):
        bits = (u >> np.uint64(k)) & np.uint64(1)
        pl[:, k] = (bits << np.arange(64, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)
    return pl, mx > 0


def msb64(v):
    """index of the highest set bit, -1 for 0"""
    r = np.full(v.shape, -1, dtype=np.int64)
    vv = v.copy()
    for s in (32, 16, 8, 4, 2, 1):
        m = (vv >> np.uint64(s)) != 0
        r[m] += s
        vv[m] >>= np.uint64(s)
    r[v != 0] += 1
    return r


def popc64(v):
    c = np.zeros(v.shape, dtype=np.int64)
    for i in range(64):
        c += ((v >> np.uint64(i)) & np.uint64(1)).astype(np.int64)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rate", type=int, default=8)
    ap.add_argument("--nz", type=int, default=8)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--field", default="lognormal")
    args = ap.parse_args()
    import torch
    import bench

    full = (256, 2048, 2048)
    # the generator is separable in z-chunks; build only the first nz planes and the first ny rows
    f = bench.make_field(torch, (args.nz, full[1], full[2]), "float32", args.field, 0, "cpu").numpy()[:, :args.ny, :]
    pl, nonzero = planes_of(np.ascontiguousarray(f))
    nb = pl.shape[0]
    maxbits = 64 * args.rate
    budget = maxbits - 9
    n = np.zeros(nb, dtype=np.int64)
    used = np.zeros(nb, dtype=np.int64)
    alive = nonzero.copy()
    # per plane records
    rec = []
    for k in range(31, -1, -1):
        x = pl[:, k]
        act = alive & (used < budget)
        y = np.where(n < 64, x >> np.minimum(n, 63).astype(np.uint64), 0).astype(np.uint64)
        y[n >= 64] = 0
        t = (y != 0) & (n < 64)
        plast = msb64(y)  # relative
        pc = popc64(y)
        plast_abs = n + plast
        code = np.where(n < 64, 1 + np.where(t, plast + 1 + pc - np.where(plast_abs == 63, 2, 0), 0), 0)
        length = n + code
        room = budget - used
        code_eff = np.clip(room - n, 0, None)
        code_eff = np.minimum(code, code_eff)
        rec.append(dict(k=k, act=act.copy(), n=n.copy(), t=t & act, bytes=np.where(t & act, plast // 8 + 1, 0),
                        code=np.where(act, code_eff, 0), verb=np.where(act, np.minimum(n, room), 0)))
        used = np.where(act, used + np.minimum(length, room), used)
        n = np.where(act & t, plast_abs + 1, n)
    W = nb // 32
    def warp(a):
        return a[: W * 32].reshape(W, 32)

    planes_per_lane = sum(warp(r["act"]).astype(int) for r in rec)
    print("blocks %d warps %d rate %d" % (nb, W, args.rate))
    print("planes coded per lane: mean %.2f  warp-max mean %.2f" % (planes_per_lane.mean(), planes_per_lane.max(1).mean()))
    nn = np.stack([r["n"] for r in rec])
    # encoder: 2-byte expansion steps
    for name, stepfn in (("enc 2-byte steps", lambda r: (r["bytes"] + 1) // 2),
                         ("enc 4-byte steps", lambda r: (r["bytes"] + 3) // 4),
                         ("dec 8-bit steps", lambda r: (r["code"] + 7) // 8),
                         ("dec 16-bit steps", lambda r: (r["code"] + 15) // 16),
                         ("dec 32-bit steps", lambda r: (r["code"] + 31) // 32)):
        per_plane_max = sum(warp(stepfn(r)).max(1) for r in rec)
        lane_tot = sum(warp(stepfn(r)) for r in rec)
        flat = sum(np.maximum(warp(stepfn(r)), warp(r["act"]).astype(int)) for r in rec)  # >= 1 iteration per plane
        print("%-18s lane mean %.1f | per-plane loop: warp iterations %.1f | flat(step only) warp max %.1f | flat(max(1,steps)) warp max %.1f"
              % (name, lane_tot.mean(), per_plane_max.mean(), lane_tot.max(1).mean(), flat.max(1).mean()))
    # distribution of n at the plane where the budget ends, and of planes with n == 64
    v64 = sum(warp(r["act"] & (r["n"] >= 64)).astype(int) for r in rec)
    print("verbatim (n == 64) planes per lane: mean %.2f" % v64.mean())
    for r in rec[:16]:
        a = r["act"]
        if a.sum() == 0:
            continue
        print("plane %2d active %.2f  n mean %5.1f  t %.2f  bytes mean %.2f  code mean %5.1f  verb mean %5.1f" % (
            r["k"], a.mean(), r["n"][a].mean(), r["t"][a].mean(), r["bytes"][a].mean(), r["code"][a].mean(), r["verb"][a].mean()))


if __name__ == "__main__":
    main()
