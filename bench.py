#!/usr/bin/env python
"""bench.py -- the hot path of CBench on B200: fixed-rate zfp compress + decompress fused with the five
error metrics (main.cpp:262-352), on synthetic fields of the shapes BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--rate R]

A "step" is one pass of compress -> decompress+metrics (-> NCCL all-reduce of the metric partials when
N > 1) over one field that is already resident in HBM.  Workloads (weak scaling: every GPU owns one
independent sub-field, exactly as every CBench MPI rank does):

  nyx2048-slab  (default) 256 x 2048 x 2048 float32 slab per GPU -- at N = 8 the slabs are the NYX 2048^3
                field of BASELINE.json's headline target; zfp 3D fixed-rate 8 bpv
  nyx512        512^3 float32 per GPU (BASELINE configs[2])
  hacc1d        268 435 456 float32 particles per GPU, 1D (BASELINE configs[1]); "bits" 8 -> 64-bit slots
  vpic1024-f64  256 x 1024 x 1024 float64 slab per GPU (BASELINE configs[4])

Rank 0 prints ONE JSON line (see the task contract).  `--impl reference` times the CPU restatement of the
reference path (oracle/, serial zfp per rank + the five metric passes, one OpenMP thread per "rank") on the
host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "nyx2048-slab": dict(shape=(256, 2048, 2048), full=(2048, 2048, 2048), dtype="float32", rate=8, kind="lognormal",
                         desc="NYX 2048^3 float32 slab-sharded along the slowest axis: one 256x2048x2048 slab per GPU "
                              "(8 slabs = the full field at N=8), zfp 3D fixed-rate + fused abs/rel/mse/psnr/minmax"),
    "nyx512": dict(shape=(512, 512, 512), dtype="float32", rate=8, kind="lognormal",
                   desc="NYX 512^3 float32 field per GPU, zfp 3D fixed-rate + fused metrics"),
    "hacc1d": dict(shape=(268435456,), dtype="float32", rate=8, kind="uniform",
                   desc="HACC 1D particle array, 256Mi float32 per GPU, zfp 1D fixed-rate (wra: 64-bit slots) + fused metrics"),
    "vpic1024-f64": dict(shape=(256, 1024, 1024), dtype="float64", rate=8, kind="smooth",
                         desc="VPIC-sized 1024^3 float64 slab per GPU (256x1024x1024), zfp 3D fixed-rate"),
}


def make_field(torch, shape, dtype, kind, seed, device, zoff=None):
    """Deterministic synthetic field generated on the device.  zoff = position of the slab's
    first plane along z in units of 2048 planes (default: slab `seed` of 256-plane slabs)."""
    g = torch.Generator(device=device).manual_seed(20261019 + seed)
    tdt = torch.float32 if dtype == "float32" else torch.float64
    if kind == "uniform":  # HACC positions: U[0, 256)
        return (256.0 * torch.rand(shape, generator=g, device=device, dtype=torch.float32)).to(tdt)
    if kind == "velocity":  # HACC velocities: N(0, 900^2)
        return (900.0 * torch.randn(shape, generator=g, device=device, dtype=torch.float32)).to(tdt)
    if kind == "noise":  # coder worst case: white noise
        return torch.randn(shape, generator=g, device=device, dtype=torch.float32).to(tdt)
    if kind == "zero":  # coder best case: every block empty
        return torch.zeros(shape, device=device, dtype=tdt)
    nz, ny, nx = shape
    cg = torch.Generator(device="cpu").manual_seed(1000 + seed)
    out = torch.empty(shape, device=device, dtype=tdt)
    zs = torch.arange(nz, device=device, dtype=torch.float32) / max(nz, 1)
    ys = torch.arange(ny, device=device, dtype=torch.float32) / 2048.0
    xs = torch.arange(nx, device=device, dtype=torch.float32) / 2048.0
    modes = []
    for _ in range(12):
        k = torch.randint(1, 9, (3,), generator=cg).tolist()
        ph = float(torch.rand(1, generator=cg)) * 6.2831853
        amp = float(torch.randn(1, generator=cg)) / (1 + 0.3 * sum(k))
        modes.append((k, ph, amp))
    chunk = max(4, min(nz, (64 << 20) // (ny * nx * 4) // 4 * 4))
    for z0 in range(0, nz, chunk):
        z1 = min(nz, z0 + chunk)
        acc = torch.zeros((z1 - z0, ny, nx), device=device, dtype=torch.float32)
        for (kz, ky, kx), ph, amp in modes:
            arg = (6.2831853 * kz * (zs[z0:z1] * (nz / 2048.0) + (seed * 0.125 if zoff is None else zoff)))[:, None, None] \
                + (6.2831853 * ky * ys)[None, :, None] + (6.2831853 * kx * xs)[None, None, :] + ph
            acc += amp * torch.cos(arg)
        acc += 0.05 * torch.randn(acc.shape, generator=g, device=device, dtype=torch.float32)
        if kind == "lognormal":
            acc = torch.exp(1.5 * acc)
        out[z0:z1] = acc.to(tdt)
        del acc
    return out


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([t.strip() for t in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for j, nme in enumerate(names):
                if len(r) > 3 + j and r[3 + j].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, wl):
    """--impl reference: the CPU restatement of the reference path on the host cores (rank 0 only): serial zfp per
    "rank" + the five metric passes, one OpenMP thread per rank, on a bounded sample of the SAME workload -- the same
    generator (make_field) with the same parameters, cut to slabs of about 32 MiB so that a step takes seconds."""
    import ctypes as C
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    world = int(os.environ.get("WORLD_SIZE", "1"))
    shape, dtype, rate = wl["shape"], wl["dtype"], args.rate or wl["rate"]
    if args.scaling == "strong":
        shape = (wl["full"][0] // world,) + tuple(wl["full"][1:])
    d = len(shape)
    npdt = np.float32 if dtype == "float32" else np.float64
    esz = 4 if dtype == "float32" else 8
    mb = O.maxbits(npdt, d, int(rate) if float(rate).is_integer() else rate, args.wra)
    nranks = min(cores, 64)
    # bounded sample: each "rank" gets a slab of ~32 MiB of the benchmark field (same generator, same parameters)
    if d == 3:
        thick = max(4, min(shape[0], ((32 << 20) // (shape[1] * shape[2] * esz)) // 4 * 4))
        sshape = (thick, shape[1], shape[2])
    else:
        sshape = (min(shape[0] // 4 * 4, (32 << 20) // esz),)
    kind = args.field or wl["kind"]
    slabs = []
    for r in range(nranks):
        zoff = (r * sshape[0] / 2048.0) if d == 3 else None
        slabs.append(make_field(torch, sshape, dtype, kind, r if d == 1 else 0, "cpu", zoff).numpy())
    field = np.ascontiguousarray(np.stack(slabs))
    del slabs
    sample_bytes = field.nbytes

    def one_step():
        if dtype == "float32":
            r = O.cbench_step(field, mb, nthreads=nranks)
            return r["wall"], (r["t_compress"], r["t_decompress"], r["t_metrics"])
        # float64: the reference's metric classes read the bytes as float (meaningless): codec timing only
        dd, nx, ny, nz = O.dims_of(sshape)
        sbytes = int(O.lib().zfo_stream_bytes(dd, nx, ny, nz, mb))
        st = np.zeros(nranks * sbytes + 16, dtype=np.uint8)
        out = np.empty_like(field)
        te, td = C.c_double(), C.c_double()
        w = O.lib().zfo_roundtrip_slabs(2, dd, nx, ny, nz, mb, field.ctypes.data, st.ctypes.data, out.ctypes.data, nranks,
                                        nranks, C.byref(te), C.byref(td))
        return w, (te.value, td.value, 0.0)

    for _ in range(max(1, min(args.warmup, 1))):
        one_step()
    t, parts = [], []
    for _ in range(args.steps):
        w, p3 = one_step()
        t.append(w)
        parts.append(p3)
    wall = sum(t)
    value = sample_bytes * args.steps / wall / 1e9
    n = int(np.prod(shape))
    cbytes = int(O.stream_bytes(shape, mb))
    sample = "%d ranks (OpenMP threads) x %s %s slab each of the benchmark field = %.1f MiB per step%s" % (
        nranks, "x".join(map(str, sshape)), dtype, sample_bytes / 2 ** 20,
        "" if dtype == "float32" else "; codec only (the reference's metrics are float-only)")
    line = {
        "impl": "reference", "metric": "uncompressed_GBps_compress_decompress_metrics", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
        # the same keys and values as the GPU arm's config (the workload is the same; the CPU times a bounded sample of it)
        "config": {"workload": args.workload if args.scaling == "weak" else args.workload.replace("-slab", "") + "-strong",
                   "description": wl["desc"], "shape_per_gpu": list(shape), "rate_bpv": rate, "maxbits": mb,
                   "mode": args.mode or "rate", "wra": args.wra, "field": kind, "histogram": "none",
                   "compressed_bytes_per_gpu": cbytes,
                   "l2": "inputs (%.1f GiB per step) are far larger than the 126 MB L2; no flush needed" % (n * esz / 2 ** 30),
                   "exchange": "zfb_allreduce_partials: one ncclAllGather of the 64-byte partials + rank-ordered combine" if world > 1 else "none (1 GPU)"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": nranks, "kind": "port", "sample": sample,
                         "t_compress_s": parts[-1][0], "t_decompress_s": parts[-1][1], "t_metrics_s": parts[-1][2],
                         "note": "CPU restatement of serial zfp 0.5.5 fixed-rate + the five CBench metric passes (oracle/, "
                                 "-O3 -march=native); one OpenMP thread per CBench rank; libzfp / MPI are not in the image"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """Multi-GPU runs: pin this rank to the CPUs local to its GPU so that pinned host buffers (first touch) and the
    staging copies sit on the GPU's NUMA node; without it 8 ranks share one node's memory and PCIe root."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        dom, rest = out.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:], rest)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"cpus": len(allowed), "node": open(path.replace("local_cpulist", "numa_node")).read().strip()}
    except Exception:
        pass
    return None


def run_e2e(torch, dist, codec, pkg, field, shape, dtype, tdt, mb, cbytes, field_bytes, world, dev, args, rate):
    """End to end through the plugin CLASSES with the buffers CBench really hands them (main.cpp:262-352, 395): the
    field sits in pageable host memory like the loader's `new T[]`, ZFPCompressorB200::compress / decompress malloc
    their outputs inside the timed region, the five GPU metric objects are created and executed per step, and the
    decompressed array is std::free'd -- libzfb_cbench.so (cbench/cbench_shim.cpp) is that sequence behind a C call.
    Per step: H2D = the field, D2H = the stream + the decompressed field (the stream stays resident on the device
    between compress and decompress; the fused metric partials come back as 64 bytes)."""
    import ctypes as C
    import numpy as np

    shim = C.CDLL(os.path.join(ROOT, "vizaly-foresight_b200", "libzfb_cbench.so"))
    shim.zfbc_open.restype = C.c_void_p
    shim.zfbc_open.argtypes = [C.c_char_p]
    shim.zfbc_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    shim.zfbc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t,
                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_size_t)]
    shim.zfbc_close.argtypes = [C.c_void_p]
    os.environ.setdefault("ZFB_DEVICE", str(dev.index))
    h = shim.zfbc_open(b"zfp_gpu")
    shim.zfbc_set(h, b"bits", str(int(rate)).encode())
    if args.wra != 1:
        shim.zfbc_set(h, b"wra", str(args.wra).encode())
    h_in = np.empty(shape, dtype=np.float32 if dtype == "float32" else np.float64)  # malloc'd, pageable
    h_in[...] = field.cpu().numpy()
    n3 = list(shape) + [0] * (3 - len(shape))
    met = (C.c_double * 5)()
    sec = (C.c_double * 3)()
    cb = C.c_size_t()
    dt = b"float" if dtype == "float32" else b"double"

    def e2e_step():
        if shim.zfbc_step(h, h_in.ctypes.data, dt, n3[0], n3[1], n3[2], met, sec, C.byref(cb)) != 0:
            raise RuntimeError("plugin step failed")

    e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    parts = [0.0, 0.0, 0.0]
    for _ in range(args.e2e_steps):
        e2e_step()
        for j in range(3):
            parts[j] += sec[j]
    w = time.perf_counter() - w0
    wt = torch.tensor([w], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    w = float(wt.item())
    shim.zfbc_close(h)
    del h_in
    out = {"value": world * field_bytes * args.e2e_steps / w / 1e9, "unit": "GB/s",
           "h2d_bytes_per_step": world * field_bytes, "d2h_bytes_per_step": world * (field_bytes + int(cb.value)),
           "steps": args.e2e_steps, "ms_per_step": 1e3 * w / args.e2e_steps,
           "ms_compress": 1e3 * parts[0] / args.e2e_steps, "ms_decompress": 1e3 * parts[1] / args.e2e_steps,
           "ms_metrics": 1e3 * parts[2] / args.e2e_steps,
           "metrics_check": {"absolute_error": met[0], "relative_error": met[1], "mse": met[2], "psnr": met[3], "minmax": met[4]},
           "api": "ZFPCompressorB200::compress + ::decompress + five gpu metric objects (cbench/), pageable `new T[]`-style "
                  "input, malloc'd outputs allocated inside the timed region; PCIe + host-copy bound"}
    return out


def run_e2e_pinned(torch, dist, codec, pkg, field, shape, dtype, tdt, mb, cbytes, field_bytes, world, dev, args):
    """The same two C-ABI calls on PINNED caller buffers allocated outside the timed region (direct DMA, no staging
    copies, no page faults): the ceiling the pageable path is measured against."""
    h_in = torch.empty(shape, dtype=tdt, pin_memory=True)
    h_in.copy_(field)
    h_stream = torch.empty(cbytes + 16, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(shape, dtype=tdt, pin_memory=True)
    code = 1 if dtype == "float32" else 2

    def e2e_step():
        codec.compress_host_ptr(code, shape, mb, h_in.data_ptr(), h_stream.data_ptr())
        codec.decompress_host_ptr(code, shape, mb, h_stream.data_ptr(), h_out.data_ptr(), h_in.data_ptr())
        return codec.partials()

    e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    w = time.perf_counter() - w0
    wt = torch.tensor([w], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    w = float(wt.item())
    codec.release()
    del h_in, h_out, h_stream
    return {"value": world * field_bytes * args.e2e_steps / w / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": world * field_bytes, "d2h_bytes_per_step": world * (field_bytes + cbytes),
            "steps": args.e2e_steps, "ms_per_step": 1e3 * w / args.e2e_steps,
            "api": "zfb_compress_host + zfb_decompress_host (fused metrics), pinned host buffers; PCIe bound"}


def run_e2e_var(torch, codec, pkg, field, shape, dtype, tdt, mode, field_bytes, args):
    """The same end-to-end measurement for a variable-rate mode (N = 1 sweeps): zfb_compress_var_host +
    zfb_decompress_var_host on pinned host buffers, stream length as the state between the two calls."""
    import ctypes as C

    h_in = torch.empty(shape, dtype=tdt, pin_memory=True)
    h_in.copy_(field)
    cap = pkg.maximum_size(shape, "float32" if dtype == "float32" else "float64", mode)
    h_stream = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(shape, dtype=tdt, pin_memory=True)
    code = 1 if dtype == "float32" else 2
    d, nx, ny, nz = pkg.dims_of(shape)
    L = pkg.lib()
    nbytes = C.c_size_t()

    def e2e_step():
        codec._check(L.zfb_compress_var_host(codec._h, code, d, nx, ny, nz, C.byref(mode), h_in.data_ptr(), h_stream.data_ptr(),
                                             cap, C.byref(nbytes)), "zfb_compress_var_host")
        codec._check(L.zfb_decompress_var_host(codec._h, code, d, nx, ny, nz, C.byref(mode), h_stream.data_ptr(), nbytes.value,
                                               h_out.data_ptr(), h_in.data_ptr()), "zfb_decompress_var_host")
        return codec.partials()

    e2e_step()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    w = time.perf_counter() - w0
    codec.release()
    cb = int(nbytes.value)
    return {"value": field_bytes * args.e2e_steps / w / 1e9, "unit": "GB/s", "h2d_bytes_per_step": field_bytes + cb,
            "d2h_bytes_per_step": field_bytes + cb, "steps": args.e2e_steps, "ms_per_step": 1e3 * w / args.e2e_steps,
            "api": "zfb_compress_var_host + zfb_decompress_var_host (fused metrics), pinned host buffers, slab pipeline"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nyx2048-slab", choices=sorted(WORKLOADS))
    ap.add_argument("--rate", type=float, default=0)
    ap.add_argument("--wra", type=int, default=1, help="1 = CBench semantics (64-bit aligned slots), 0 = unaligned slots")
    ap.add_argument("--field", default="", choices=["", "lognormal", "smooth", "uniform", "velocity", "noise", "zero"],
                    help="override the workload's synthetic field kind (sweeps only)")
    ap.add_argument("--mode", default="", help="variable-rate sweep only: abs:<tolerance> (zfp fixed accuracy) or "
                                               "rel:<precision> (fixed precision); default = fixed rate")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU owns one slab of the workload's per-GPU shape; strong: the workload's FULL "
                         "field (nyx2048-slab: 2048^3, 32 GiB) is cut into N slabs, so N = 1 processes all of it")
    ap.add_argument("--hist", action="store_true",
                    help="also produce the minmax metric's value-range histogram (minmaxMetric.hpp:92-133) inside the step: "
                         "min / max harvested by the encode pass, bins filled by the decode pass, counters and range "
                         "combined across ranks over NCCL")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    pkg = ge.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    shape, dtype, rate = wl["shape"], wl["dtype"], args.rate or wl["rate"]
    zoff = None
    if args.scaling == "strong":
        full = wl.get("full")
        if not full or full[0] % (4 * world):
            raise SystemExit("--scaling strong needs a workload with a full field whose slowest extent divides by 4 N")
        shape = (full[0] // world,) + tuple(full[1:])
        zoff = rank * shape[0] / 2048.0
    d = len(shape)
    npdt = np.float32 if dtype == "float32" else np.float64
    tdt = torch.float32 if dtype == "float32" else torch.float64
    esz = 4 if dtype == "float32" else 8
    mb = pkg.maxbits(npdt, d, int(rate) if float(rate).is_integer() else rate, args.wra)
    field = make_field(torch, shape, dtype, args.field or wl["kind"], rank, dev, zoff)
    n = field.numel()
    field_bytes = n * esz
    cbytes = pkg.stream_bytes(shape, mb)
    var_mode = None
    if args.mode:
        kind, _, val = args.mode.partition(":")
        var_mode = pkg.mode_accuracy(float(val)) if kind == "abs" else pkg.mode_precision(int(val))
        cbytes = pkg.maximum_size(shape, npdt, var_mode)  # capacity; the real size is data dependent (set below)
        args.no_cpu = True
    stream = torch.empty(cbytes + 16, dtype=torch.uint8, device=dev)
    out = torch.empty(shape, dtype=tdt, device=dev)

    codec = pkg.Codec(local)
    codec.use_torch_stream()
    want_metrics = dtype == "float32" or True
    if world > 1:
        codec.comm_init(dist, dev)  # the context's own NCCL communicator (id broadcast through the job's process group)

    def allreduce_partials():
        # the only exchange step: ONE ncclAllGather of the 64-byte partials + a rank-ordered combine on the device,
        # issued by the C ABI on the stream the kernels run on (zfb_allreduce_partials, csrc/zfb_comm.cuh)
        codec.allreduce_partials()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]

    var_bytes = [0]
    hist_range = [0.0, 1.0]
    hist_counts = torch.zeros(1025, dtype=torch.int64, device=dev)

    def step(i=None):
        if i is not None:
            ev[i][0].record()
        if var_mode is not None:
            s_var, var_bytes[0] = codec.encode_var(field, var_mode, out=stream)
        elif args.hist:
            codec.encode_range(field, mb, out=stream)
            if world > 1:
                codec.allreduce_range()
            hist_range[:] = codec.range()   # the one host synchronisation of the step (CBench is synchronous there too)
        else:
            codec.encode(field, mb, out=stream)
        if i is not None:
            ev[i][1].record()
        if var_mode is not None:
            codec.decode_var(s_var, shape, tdt, var_mode, orig=field, out=out)
        elif args.hist:
            codec.decode_hist(stream, shape, tdt, mb, hist_range[0], hist_range[1], orig=field, out=out, counts=hist_counts)
        else:
            codec.decode(stream, shape, tdt, mb, orig=field if want_metrics else None, out=out)
        if i is not None:
            ev[i][2].record()
        if world > 1:
            allreduce_partials()
            if args.hist:
                codec.allreduce_counts(hist_counts)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = codec.launches
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for i in range(args.steps):
        step(i)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = codec.launches - launches0
    ms_total = t0.elapsed_time(t1)
    enc_ms = sum(ev[i][0].elapsed_time(ev[i][1]) for i in range(args.steps)) / args.steps
    dec_ms = sum(ev[i][1].elapsed_time(ev[i][2]) for i in range(args.steps)) / args.steps
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    value = world * field_bytes * args.steps / (ms_total * 1e-3) / 1e9
    vals = pkg.metric_values(codec.partials())

    # ---- end to end through the host-pointer C ABI (what the CBench plugin calls) ------------------
    e2e = None
    if not args.no_e2e and field_bytes > (8 << 30):
        try:  # the plugin leg needs ~3 host copies of the field (input, malloc'd output, pinned arm): skip if the host is small
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 0
        if avail < 4 * field_bytes * (world if world > 1 else 1):
            args.no_e2e = True
    if not args.no_e2e:
        try:
            if var_mode is not None:
                e2e = run_e2e_var(torch, codec, pkg, field, shape, dtype, tdt, var_mode, field_bytes, args) if world == 1 else None
            else:
                e2e = run_e2e(torch, dist, codec, pkg, field, shape, dtype, tdt, mb, cbytes, field_bytes, world, dev, args, rate)
                e2e["pinned"] = run_e2e_pinned(torch, dist, codec, pkg, field, shape, dtype, tdt, mb, cbytes, field_bytes, world, dev, args)
        except Exception as ex:  # e.g. pinned host memory exhausted with many ranks: report, do not lose the bench line
            e2e = {"value": None, "unit": "GB/s", "error": str(ex)[:200]}
        ok = torch.tensor([0 if e2e.get("value") is None else 1], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and e2e.get("value") is not None:
            e2e = {"value": None, "unit": "GB/s", "error": "e2e failed on another rank"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (decode + fused metrics) ---------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    if var_mode is not None:
        cbytes = var_bytes[0]
    dec_bytes = cbytes + 2 * field_bytes
    enc_bytes = cbytes + field_bytes
    roof = {"bound": "hbm", "kernel": "decode+metrics", "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9, "peak": peak,
            "unit": "GB/s", "traffic": None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_launch": dec_bytes, "ms_per_launch": dec_ms,
            "encode": {"achieved": enc_bytes / (enc_ms * 1e-3) / 1e9, "algorithmic_bytes_per_launch": enc_bytes,
                       "ms_per_launch": enc_ms},
            "round_trip": {"achieved": (enc_bytes + dec_bytes) / ((enc_ms + dec_ms) * 1e-3) / 1e9,
                           "bytes_per_value": (enc_bytes + dec_bytes) / n}}
    # DRAM traffic of one launch, from the committed ncu --set full capture of this workload (null otherwise)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_nyx2048slab.json")))
        if args.workload == "nyx2048-slab" and int(rate) == 8:
            roof["traffic"] = tr["kernels"]["decode+metrics"]["traffic"]
            roof["traffic_source"] = ("profiles/r02_traffic_nyx2048slab.json: ncu dram__bytes_read.sum + dram__bytes_write.sum of ONE --set full "
                                      "capture of this command with these kernels (tools/gpu_profile_final.sh); not re-measured by this run")
            roof["encode"]["traffic"] = tr["kernels"]["encode"]["traffic"]
    except Exception:
        pass
    roof["frac"] = roof["achieved"] / peak
    roof["encode"]["frac"] = roof["encode"]["achieved"] / peak
    roof["round_trip"]["frac"] = roof["round_trip"]["achieved"] / peak

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu and dtype == "float64" and d == 3:
        # float64: codec timing only (the reference's metric classes read the bytes as float)
        import ctypes as C
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O

        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 1
        nranks = min(cores, 64)
        thick = max(4, min(shape[0], ((32 << 20) // (shape[1] * shape[2] * 8)) // 4 * 4))
        sample_arr = torch.stack([field[(r * thick) % (shape[0] - thick + 1):][:thick] for r in range(nranks)]).cpu().numpy()
        sbytes = int(O.lib().zfo_stream_bytes(3, shape[2], shape[1], thick, mb))
        st = np.zeros(nranks * sbytes + 16, dtype=np.uint8)
        dec = np.empty_like(sample_arr)
        te, td = C.c_double(), C.c_double()
        args_rt = (2, 3, shape[2], shape[1], thick, mb, sample_arr.ctypes.data, st.ctypes.data, dec.ctypes.data, nranks, nranks)
        O.lib().zfo_roundtrip_slabs(*args_rt, C.byref(te), C.byref(td))
        w = O.lib().zfo_roundtrip_slabs(*args_rt, C.byref(te), C.byref(td))
        cpu = {"value": sample_arr.nbytes / w / 1e9, "unit": "GB/s", "cores": nranks, "kind": "port",
               "sample": "%d ranks x %dx%dx%d float64 slabs of the benchmark field (%.0f MiB); serial-zfp restatement, codec "
                         "only, one OpenMP thread per rank" % (nranks, thick, shape[1], shape[2], sample_arr.nbytes / 2 ** 20),
               "t_compress_s": te.value, "t_decompress_s": td.value, "t_metrics_s": 0.0}
    if world == 1 and not args.no_cpu and dtype == "float32":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O

        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 1
        nranks = min(cores, 64)
        if d == 3:
            thick = max(4, min(shape[0], ((32 << 20) // (shape[1] * shape[2] * 4)) // 4 * 4))
            slabs = [field[(r * thick) % (shape[0] - thick + 1):][:thick] for r in range(nranks)]
            sample_arr = torch.stack(slabs).cpu().numpy()
            sdesc = "%d ranks x %dx%dx%d float32 slabs of the benchmark field" % (nranks, thick, shape[1], shape[2])
        else:
            m = min(shape[0] // nranks // 4 * 4, 8 << 20)
            sample_arr = field[: nranks * m].reshape(nranks, m).cpu().numpy()
            sdesc = "%d ranks x %d float32 values of the benchmark array" % (nranks, m)
        O.cbench_step(sample_arr[: max(1, nranks // 8)], mb, nthreads=nranks)  # touch pages / warm up
        r = O.cbench_step(sample_arr, mb, nthreads=nranks)
        cpu = {"value": sample_arr.nbytes / r["wall"] / 1e9, "unit": "GB/s", "cores": nranks, "kind": "port",
               "sample": sdesc + " (%.0f MiB); serial-zfp restatement + five metric passes per rank, one OpenMP thread "
                                 "per rank" % (sample_arr.nbytes / 2 ** 20),
               "t_compress_s": r["t_compress"], "t_decompress_s": r["t_decompress"], "t_metrics_s": r["t_metrics"]}

    line = {
        "metric": "uncompressed_GBps_compress_decompress_metrics", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
        "config": {"workload": args.workload if args.scaling == "weak" else args.workload.replace("-slab", "") + "-strong", "description": wl["desc"], "shape_per_gpu": list(shape),
                   "rate_bpv": rate if var_mode is None else 8.0 * cbytes / n, "maxbits": mb if var_mode is None else None,
                   "mode": args.mode or "rate", "wra": args.wra, "field": args.field or wl["kind"],
                   "histogram": ("minmax value-range histogram fused into the step: %d values binned, %d out of range"
                                 % (int(hist_counts[:1024].sum().item()), int(hist_counts[1024].item()))) if args.hist else "none",
                   "compressed_bytes_per_gpu": cbytes,
                   "l2": "inputs (%.1f GiB per step) are far larger than the 126 MB L2; no flush needed" % (field_bytes / 2 ** 30),
                   "exchange": "zfb_allreduce_partials: one ncclAllGather of the 64-byte partials + rank-ordered combine" if world > 1 else "none (1 GPU)"},
        "numa_binding": numa, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        "enc_ms": enc_ms, "dec_ms": dec_ms,
        "metrics_check": {k: vals[k] for k in ("absolute_error", "relative_error", "mse", "psnr", "minmax")},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
