"""oracle/oracle.py -- TEST INFRASTRUCTURE: ctypes door to the CPU checker.

* ``liboracle.so``          zfp_ref.c (codec restatement, PARITY UNPINNED -- see its header)
                            + metrics_ref.c (metric restatement, pinned against the reference build)
* ``_ref/libcbench_ref.so`` the reference's own metric headers (CBench/metrics/*.hpp) built by
                            oracle/Makefile against mpi_stub/mpi.h (present only if built where
                            /root/reference exists; the prebuilt file travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FLOAT, DOUBLE = 1, 2
NBINS = 1024
_lib = None
_ref = None


def build(force=False):
    """make -C oracle (also builds _ref/ when /root/reference is present)."""
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("zfp_ref.c", "metrics_ref.c", "cbench_step_ref.c", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    # built with -march=native: a library that travelled from another machine is rebuilt for this CPU
    stamp = os.path.join(HERE, "liboracle.cpu")
    cpu = _cpu_id()
    if os.path.exists(stamp) and open(stamp).read() != cpu:
        stale = True
    if force or stale or not os.path.exists(stamp):
        r = subprocess.run(["make", "-B", "-C", HERE, "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:  # e.g. a compiler that does not know this CPU: the portable baseline
            subprocess.run(["make", "-B", "-C", HERE, "liboracle.so", "MARCH=x86-64-v2"], check=True, capture_output=True)
        open(stamp, "w").write(cpu)
    ref_so = os.path.join(HERE, "_ref", "libcbench_ref.so")
    if os.path.isdir("/root/reference/CBench/metrics") and (force or not os.path.exists(ref_so)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def _cpu_id():
    """Model name + instruction-set flags of this host (what -march=native depends on)."""
    model, flags = "", ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            elif line.startswith("flags") and not flags:
                flags = " ".join(sorted(line.split(":", 1)[1].split()))
            if model and flags:
                break
    except OSError:
        pass
    import hashlib
    return model + " " + hashlib.sha1(flags.encode()).hexdigest()[:12]


class Params(C.Structure):
    """zfp_stream's four parameters (zfo_params in zfp_ref.c)."""
    _fields_ = [("minbits", C.c_uint), ("maxbits", C.c_uint), ("maxprec", C.c_uint), ("minexp", C.c_int)]

    def as_tuple(self):
        return (self.minbits, self.maxbits, self.maxprec, self.minexp)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        sz, u, i, vp, dp = C.c_size_t, C.c_uint, C.c_int, C.c_void_p, C.POINTER(C.c_double)
        L.zfo_maxbits.restype = u
        L.zfo_maxbits.argtypes = [i, i, C.c_double, i]
        L.zfo_stream_bytes.restype = sz
        L.zfo_stream_bytes.argtypes = [i, sz, sz, sz, u]
        L.zfo_stream_capacity.restype = sz
        L.zfo_stream_capacity.argtypes = [i, sz, sz, sz, u]
        L.zfo_compress.restype = sz
        L.zfo_compress.argtypes = [i, i, sz, sz, sz, u, vp, vp]
        L.zfo_decompress.restype = sz
        L.zfo_decompress.argtypes = [i, i, sz, sz, sz, u, vp, vp]
        pp = C.POINTER(Params)
        L.zfo_set_rate.restype = None
        L.zfo_set_rate.argtypes = [pp, i, i, C.c_double, i]
        L.zfo_set_accuracy.restype = None
        L.zfo_set_accuracy.argtypes = [pp, C.c_double]
        L.zfo_set_precision.restype = None
        L.zfo_set_precision.argtypes = [pp, u]
        L.zfo_maximum_size.restype = sz
        L.zfo_maximum_size.argtypes = [i, i, sz, sz, sz, pp]
        L.zfo_compress_ex.restype = sz
        L.zfo_compress_ex.argtypes = [i, i, sz, sz, sz, pp, vp, vp, sz, vp]
        L.zfo_decompress_ex.restype = sz
        L.zfo_decompress_ex.argtypes = [i, i, sz, sz, sz, pp, vp, vp]
        L.zfo_perm.restype = C.POINTER(C.c_ubyte)
        L.zfo_perm.argtypes = [i]
        for name in ("zfo_fwd_xform_i32", "zfo_inv_xform_i32", "zfo_fwd_xform_i64", "zfo_inv_xform_i64"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, i]
        L.zfo_encode_ints_u32.restype = u
        L.zfo_encode_ints_u32.argtypes = [vp, u, u, u, vp]
        L.zfo_decode_ints_u32.restype = u
        L.zfo_decode_ints_u32.argtypes = [vp, u, u, u, vp]
        L.zfo_roundtrip_slabs.restype = C.c_double
        L.zfo_roundtrip_slabs.argtypes = [i, i, sz, sz, sz, u, vp, vp, vp, i, i, dp, dp]
        L.zfo_cbench_step.restype = C.c_double
        L.zfo_cbench_step.argtypes = [i, sz, sz, sz, u, vp, vp, vp, i, i, dp, dp]
        L.zfo_abs_error.restype = None
        L.zfo_abs_error.argtypes = [vp, vp, sz, dp, vp]
        L.zfo_rel_error.restype = None
        L.zfo_rel_error.argtypes = [vp, vp, sz, dp, vp]
        L.zfo_mse.restype = C.c_double
        L.zfo_mse.argtypes = [vp, vp, sz, dp]
        L.zfo_psnr.restype = C.c_double
        L.zfo_psnr.argtypes = [vp, vp, sz]
        L.zfo_minmax.restype = None
        L.zfo_minmax.argtypes = [vp, vp, sz, dp, vp, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def ref_lib():
    """The reference's own metric classes, or None when oracle/_ref was never built."""
    global _ref
    if _ref is None:
        build()
        p = os.path.join(HERE, "_ref", "libcbench_ref.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.cbref_metric.restype = C.c_double
        R.cbref_metric.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_char_p,
                                   C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_double)]
        R.cbref_python_histogram.restype = C.c_size_t
        R.cbref_python_histogram.argtypes = [C.c_size_t, C.c_float, C.c_float, C.c_void_p, C.c_char_p,
                                             C.c_size_t]
        _ref = R
    return _ref


# --------------------------------------------------------------------------- codec
def _type_of(a):
    if a.dtype == np.float32:
        return FLOAT
    if a.dtype == np.float64:
        return DOUBLE
    raise TypeError(a.dtype)


def dims_of(shape):
    """CBench order: shape = (n[0] slowest ... n[d-1] fastest) -> (dims, nx, ny, nz)
    (zfpCompressorGpu.hpp:110-118: zfp_field_3d(input, type, n[2], n[1], n[0]))."""
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    if d == 1:
        return 1, shape[0], 1, 1
    if d == 2:
        return 2, shape[1], shape[0], 1
    if d == 3:
        return 3, shape[2], shape[1], shape[0]
    raise ValueError("dims 1..3 only")


def maxbits(dtype, dims, rate, wra=1):
    t = FLOAT if np.dtype(dtype) == np.float32 else DOUBLE
    return int(lib().zfo_maxbits(t, dims, float(rate), int(wra)))


def stream_bytes(shape, mb):
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfo_stream_bytes(d, nx, ny, nz, mb))


def stream_capacity(shape, mb):
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfo_stream_capacity(d, nx, ny, nz, mb))


def compress(a, mb):
    """a: C-contiguous float32/float64 array, CBench axis order.  Returns uint8 stream."""
    a = np.ascontiguousarray(a)
    d, nx, ny, nz = dims_of(a.shape)
    n = int(lib().zfo_stream_bytes(d, nx, ny, nz, mb))
    out = np.zeros(n + 16, dtype=np.uint8)
    got = lib().zfo_compress(_type_of(a), d, nx, ny, nz, mb, a.ctypes.data, out.ctypes.data)
    if got != n:
        raise RuntimeError("oracle compress failed")
    return out[:n]


def decompress(stream, shape, dtype, mb):
    d, nx, ny, nz = dims_of(shape)
    out = np.empty(shape, dtype=dtype)
    buf = np.zeros(len(stream) + 16, dtype=np.uint8)
    buf[: len(stream)] = stream
    t = FLOAT if np.dtype(dtype) == np.float32 else DOUBLE
    if not lib().zfo_decompress(t, d, nx, ny, nz, mb, buf.ctypes.data, out.ctypes.data):
        raise RuntimeError("oracle decompress failed")
    return out


# ---- general stream parameters: fixed accuracy ("abs") / fixed precision ("rel") of the CPU zfp plugin
def params_rate(dtype, dims, rate, wra=1):
    p = Params()
    lib().zfo_set_rate(C.byref(p), FLOAT if np.dtype(dtype) == np.float32 else DOUBLE, dims, float(rate), int(wra))
    return p


def params_accuracy(tolerance):
    p = Params()
    lib().zfo_set_accuracy(C.byref(p), float(tolerance))
    return p


def params_precision(precision):
    p = Params()
    lib().zfo_set_precision(C.byref(p), int(precision))
    return p


def maximum_size(shape, dtype, params):
    d, nx, ny, nz = dims_of(shape)
    t = FLOAT if np.dtype(dtype) == np.float32 else DOUBLE
    return int(lib().zfo_maximum_size(t, d, nx, ny, nz, C.byref(params)))


def compress_ex(a, params, with_offsets=False):
    """Any mode.  Returns the uint8 stream (length = zfp_compress's return) [, block bit offsets (nblocks+1)]."""
    a = np.ascontiguousarray(a)
    d, nx, ny, nz = dims_of(a.shape)
    cap = maximum_size(a.shape, a.dtype, params)
    out = np.zeros(cap + 16, dtype=np.uint8)
    nb = ((nx + 3) // 4) * ((ny + 3) // 4 if d >= 2 else 1) * ((nz + 3) // 4 if d >= 3 else 1)
    offs = np.zeros(nb + 1, dtype=np.uint64) if with_offsets else None
    got = int(lib().zfo_compress_ex(_type_of(a), d, nx, ny, nz, C.byref(params), a.ctypes.data, out.ctypes.data, cap,
                                    offs.ctypes.data if with_offsets else None))
    if not got or got > cap:
        raise RuntimeError("oracle compress_ex failed")
    return (out[:got], offs) if with_offsets else out[:got]


def decompress_ex(stream, shape, dtype, params):
    d, nx, ny, nz = dims_of(shape)
    out = np.empty(shape, dtype=dtype)
    buf = np.zeros(len(stream) + 64, dtype=np.uint8)
    buf[: len(stream)] = stream
    t = FLOAT if np.dtype(dtype) == np.float32 else DOUBLE
    got = int(lib().zfo_decompress_ex(t, d, nx, ny, nz, C.byref(params), buf.ctypes.data, out.ctypes.data))
    if not got:
        raise RuntimeError("oracle decompress_ex failed")
    return out, got


def perm(dims):
    p = lib().zfo_perm(dims)
    return np.array([p[i] for i in range(4 ** dims)], dtype=np.int64)


# --------------------------------------------------------------------------- metrics
def metrics(orig, approx, hist=False):
    """All five CBench metrics on float32 arrays via the restatement (metrics_ref.c)."""
    o = np.ascontiguousarray(orig, dtype=np.float32).ravel()
    a = np.ascontiguousarray(approx, dtype=np.float32).ravel()
    n = o.size
    L = lib()
    out = {}
    v = (C.c_double * 3)()
    h = np.zeros(NBINS, dtype=np.float32)
    L.zfo_abs_error(o.ctypes.data, a.ctypes.data, n, v, h.ctypes.data if hist else None)
    out["absolute_error"] = v[0]
    out["abs_sum"], out["abs_mean"] = v[1], v[2]
    if hist:
        out["abs_hist"] = h.copy()
    L.zfo_rel_error(o.ctypes.data, a.ctypes.data, n, v, h.ctypes.data if hist else None)
    out["relative_error"] = v[0]
    out["rel_sum"], out["rel_mean"] = v[1], v[2]
    if hist:
        out["rel_hist"] = h.copy()
    ss = C.c_double()
    out["mse"] = L.zfo_mse(o.ctypes.data, a.ctypes.data, n, C.byref(ss))
    out["sum_sq"] = ss.value
    out["psnr"] = L.zfo_psnr(o.ctypes.data, a.ctypes.data, n)
    mm = (C.c_double * 2)()
    oob = C.c_uint64()
    L.zfo_minmax(o.ctypes.data, a.ctypes.data, n, mm, h.ctypes.data if hist else None, C.byref(oob))
    out["min"], out["max"] = mm[0], mm[1]
    out["minmax"] = mm[1]
    if hist:
        out["minmax_hist"] = h.copy()
        out["minmax_oob"] = oob.value
    return out


def ref_metric(name, orig, approx, hist=False):
    """One metric through the reference's own class (None if oracle/_ref is absent)."""
    R = ref_lib()
    if R is None:
        return None
    o = np.ascontiguousarray(orig, dtype=np.float32).ravel()
    a = np.ascontiguousarray(approx, dtype=np.float32).ravel()
    log = C.create_string_buffer(4096)
    extra = C.create_string_buffer(1 << 16)
    loc = C.c_double()
    v = R.cbref_metric(name.encode(), o.ctypes.data, a.ctypes.data, o.size, int(hist), log, 4096, extra,
                       1 << 16, C.byref(loc))
    return {"value": v, "local": loc.value, "log": log.value.decode(), "extra": extra.value.decode()}


def cbench_step(field_slabs, mb, nthreads=0):
    """CBench scalar iteration on CPU: field_slabs[r] is rank r's float32 sub-field (all the same shape).
    Returns dict(wall, t_compress, t_decompress, t_metrics, values)."""
    a = np.ascontiguousarray(field_slabs, dtype=np.float32)
    nranks = a.shape[0]
    d, nx, ny, nz = dims_of(a.shape[1:])
    sbytes = int(lib().zfo_stream_bytes(d, nx, ny, nz, mb))
    stream = np.zeros(nranks * sbytes + 16, dtype=np.uint8)
    out = np.empty_like(a)
    times = (C.c_double * 3)()
    vals = (C.c_double * 5)()
    w = lib().zfo_cbench_step(d, nx, ny, nz, mb, a.ctypes.data, stream.ctypes.data, out.ctypes.data, nranks,
                              int(nthreads), times, vals)
    return {"wall": w, "t_compress": times[0], "t_decompress": times[1], "t_metrics": times[2],
            "values": list(vals), "out": out, "stream": stream[: nranks * sbytes]}


def parse_hist_script(text):
    """Pull y=[...], minVal, maxVal back out of a python_histogram() script (utils.hpp:321-352)."""
    y = lo = hi = None
    for line in text.splitlines():
        if line.startswith("y=["):
            y = np.array([float(t) for t in line[3:-1].split(",")], dtype=np.float64)
        elif line.startswith("minVal="):
            lo = float(line[7:])
        elif line.startswith("maxVal="):
            hi = float(line[7:])
    return y, lo, hi


def power_spectrum(field, nbins=None):
    """CPU statement of zfb_power_spectrum_bin_dev + Codec.power_spectrum (TEST INFRASTRUCTURE): the shell-binned power
    spectrum PAT's NYX analysis compares between original and decompressed fields (Analysis/pat/nyx/cinema.py:84-130 --
    columns 2, 3 of gimlet's *_ps3d.txt; gimlet is not in the reference tree, so the normalisation is ours and only the
    ratio of two spectra is meaningful to PAT).  Full complex transform in float64, shells b = round(|k|), every mode
    counted once.  Returns (k, pk, count)."""
    f = np.asarray(field, dtype=np.float64)
    nz, ny, nx = f.shape
    if nbins is None:
        nbins = min(nx, ny, nz) // 2 + 1
    F = np.fft.fftn(f)
    p = (F.real ** 2 + F.imag ** 2).ravel()
    kz = np.fft.fftfreq(nz, 1.0 / nz)[:, None, None]
    ky = np.fft.fftfreq(ny, 1.0 / ny)[None, :, None]
    kx = np.fft.fftfreq(nx, 1.0 / nx)[None, None, :]
    # the device folds kx onto the half-spectrum (0 .. nx/2) and kz, ky of n/2 onto +n/2: |k| is the same
    kk = np.sqrt(kx * kx + ky * ky + kz * kz).ravel()
    b = np.floor(kk + 0.5).astype(np.int64)
    m = b < nbins
    cnt = np.bincount(b[m], minlength=nbins).astype(np.float64)
    sk = np.bincount(b[m], weights=kk[m], minlength=nbins)
    sp = np.bincount(b[m], weights=p[m], minlength=nbins)
    safe = np.maximum(cnt, 1.0)
    n = float(nx) * ny * nz
    return sk / safe, sp / safe / (n * n), cnt
