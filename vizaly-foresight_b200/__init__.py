"""vizaly-foresight_b200 -- B200-native fixed-rate zfp codec + fused CBench error metrics.

Python here is plumbing only: it loads the C-ABI shared library (include/zfb.h, built from
csrc/*.cu for sm_100a) with ctypes and passes raw device / host pointers to it.  torch is used for
device memory and streams, numpy for host arrays.  There is NO CPU fallback: if the CUDA library is
missing or no B200-class device is present, construction fails loudly.

The directory name has a hyphen, so import it through ``__graft_entry__.load_package()`` (tests/conftest.py
does this) under the module name ``vizaly_foresight_b200``.
"""
import ctypes as C
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libzfb.so")
HEADER = os.path.join(ROOT, "include", "zfb.h")
FLOAT, DOUBLE = 1, 2
NBINS = 1024
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


class ZfbError(RuntimeError):
    pass


class Partials(C.Structure):
    _fields_ = [("sum_sq", C.c_double), ("sum_abs", C.c_double), ("sum_rel", C.c_double), ("max_abs", C.c_double),
                ("max_rel", C.c_double), ("max_orig", C.c_double), ("neg_min_orig", C.c_double), ("n", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Mode(C.Structure):
    """zfb_mode of include/zfb.h: zfp_stream's (minbits, maxbits, maxprec, minexp)."""
    _fields_ = [("minbits", C.c_uint), ("maxbits", C.c_uint), ("maxprec", C.c_uint), ("minexp", C.c_int)]

    def as_tuple(self):
        return (self.minbits, self.maxbits, self.maxprec, self.minexp)


def build_library(force=False, verbose=False):
    """nvcc -> vizaly-foresight_b200/libzfb.so (in-tree, so it travels to the GPU box)."""
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, "zfb_api.cu")] + sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")) + [HEADER]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if not (force or stale):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, srcs[0]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise ZfbError("nvcc failed:\n" + r.stderr[-4000:])
    return r.stderr if verbose else LIB_PATH


def declared_symbols():
    """Every function include/zfb.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zfb_[a-z0-9_]+)\s*\(", text)))


def write_ps3d(path, k, pk, count):
    """The spectrum as a text file PAT can plot: one row per shell, space separated, k in column 2 and P(k) in column 3
    (Analysis/pat/nyx/cinema.py:103-105 reads exactly those two with extract_csv_col(path, ' ', 2 / 3))."""
    with open(path, "w") as f:
        for b in range(len(k)):
            if float(count[b]) > 0:
                f.write("%d %d %.17g %.17g\n" % (b, int(count[b]), float(k[b]), float(pk[b])))


_lib = None


def lib():
    """Load libzfb.so.  Never falls back to anything else: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZfbError("CUDA extension %s is missing: run __graft_entry__.build() (no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    sz, u, i, vp, d = C.c_size_t, C.c_uint, C.c_int, C.c_void_p, C.c_double
    PP = C.POINTER(Partials)
    MP = C.POINTER(Mode)
    sig = {
        "zfb_maxbits": (u, [i, i, d, i]),
        "zfb_host_cache_invalidate": (i, [vp]),
        "zfb_host_pair_serial": (C.c_ulonglong, [vp]),
        "zfb_stream_bytes": (sz, [i, sz, sz, sz, u]),
        "zfb_stream_capacity": (sz, [i, sz, sz, sz, u]),
        "zfb_partition": (i, [i, i, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
        "zfb_create": (i, [C.POINTER(vp), i]),
        "zfb_destroy": (i, [vp]),
        "zfb_set_stream": (i, [vp, vp]),
        "zfb_reset_stream": (i, [vp]),
        "zfb_synchronize": (i, [vp]),
        "zfb_last_error": (C.c_char_p, [vp]),
        "zfb_launch_count": (C.c_uint64, [vp]),
        "zfb_encode_dev": (i, [vp, i, i, sz, sz, sz, u, vp, vp]),
        "zfb_decode_dev": (i, [vp, i, i, sz, sz, sz, u, vp, vp, vp]),
        "zfb_metrics_dev": (i, [vp, i, vp, vp, sz]),
        "zfb_power_spectrum_bin_dev": (i, [vp, vp, sz, sz, sz, u, vp]),
        "zfb_comm_unique_id": (i, [vp]),
        "zfb_comm_init": (i, [vp, i, i, vp]),
        "zfb_comm_adopt": (i, [vp, vp, i, i]),
        "zfb_comm_size": (i, [vp]),
        "zfb_comm_destroy": (i, [vp]),
        "zfb_allreduce_partials": (i, [vp]),
        "zfb_allreduce_counts": (i, [vp, vp, sz]),
        "zfb_encode_range_dev": (i, [vp, i, i, sz, sz, sz, u, vp, vp]),
        "zfb_allreduce_range": (i, [vp]),
        "zfb_range_get": (i, [vp, C.POINTER(d), C.POINTER(d)]),
        "zfb_decode_hist_dev": (i, [vp, i, i, sz, sz, sz, u, vp, vp, vp, d, d, vp]),
        "zfb_gio_open": (i, [C.c_char_p, C.POINTER(vp), C.c_char_p, sz]),
        "zfb_gio_close": (i, [vp]),
        "zfb_gio_counts": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(C.c_uint64), C.POINTER(i)]),
        "zfb_gio_var": (i, [vp, i, C.c_char_p, C.POINTER(i), C.POINTER(u)]),
        "zfb_gio_find": (i, [vp, C.c_char_p]),
        "zfb_gio_rank_elems": (C.c_uint64, [vp, i]),
        "zfb_gio_phys": (i, [vp, C.POINTER(d), C.POINTER(d)]),
        "zfb_gio_rank_range": (None, [i, i, i, C.POINTER(i), C.POINTER(i)]),
        "zfb_gio_read": (i, [vp, vp, i, i, i, vp, sz, C.POINTER(C.c_uint64), i]),
        "zfb_gio_write_host": (i, [C.c_char_p, i, C.POINTER(C.c_char_p), C.POINTER(u), C.POINTER(i), C.POINTER(vp), C.c_uint64,
                                   C.POINTER(d), C.POINTER(d), C.c_char_p, sz]),
        "zfb_h5_dataset": (i, [C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(i),
                               C.POINTER(C.c_uint64), C.POINTER(i), C.POINTER(i), C.c_char_p, sz]),
        "zfb_file_read_block_dev": (i, [vp, C.c_char_p, C.c_uint64, i, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), vp]),
        "zfb_file_write_block_dev": (i, [vp, C.c_char_p, C.c_uint64, i, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), vp]),
        "zfb_value_psnr_range": (d, [PP]),
        "zfb_value_nrmse": (d, [PP]),
        "zfb_partials_dev": (vp, [vp]),
        "zfb_partials_get": (i, [vp, PP]),
        "zfb_partials_set": (i, [vp, PP]),
        "zfb_histogram_dev": (i, [vp, i, i, d, d, vp, vp, sz, vp, vp]),
        "zfb_compress_host": (i, [vp, i, i, sz, sz, sz, u, vp, vp, C.POINTER(sz)]),
        "zfb_decompress_host": (i, [vp, i, i, sz, sz, sz, u, vp, vp, vp]),
        "zfb_metrics_host": (i, [vp, i, vp, vp, sz, PP]),
        "zfb_histogram_host": (i, [vp, i, i, d, d, vp, vp, sz, vp, C.POINTER(C.c_uint64)]),
        "zfb_release_resident": (i, [vp]),
        "zfb_value_absolute_error": (d, [PP]),
        "zfb_value_relative_error": (d, [PP]),
        "zfb_value_mse": (d, [PP]),
        "zfb_value_psnr": (d, [PP]),
        "zfb_value_minmax": (d, [PP]),
        "zfb_partials_combine": (None, [PP, PP, i]),
        "zfb_mode_rate": (i, [MP, i, i, d, i]),
        "zfb_mode_accuracy": (i, [MP, d]),
        "zfb_mode_precision": (i, [MP, u]),
        "zfb_maximum_size": (sz, [i, i, sz, sz, sz, MP]),
        "zfb_block_count": (sz, [i, sz, sz, sz]),
        "zfb_index_unit": (sz, [i]),
        "zfb_index_count": (sz, [i, sz, sz, sz]),
        "zfb_encode_var_dev": (i, [vp, i, i, sz, sz, sz, MP, vp, vp, sz, vp, C.POINTER(sz)]),
        "zfb_index_expand_dev": (i, [vp, i, i, sz, sz, sz, MP, vp, sz, vp, sz, vp]),
        "zfb_decode_var_dev": (i, [vp, i, i, sz, sz, sz, MP, vp, sz, vp, vp, vp]),
        "zfb_compress_var_host": (i, [vp, i, i, sz, sz, sz, MP, vp, vp, sz, C.POINTER(sz)]),
        "zfb_decompress_var_host": (i, [vp, i, i, sz, sz, sz, MP, vp, sz, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def dims_of(shape):
    """CBench order (n[0] slowest ... n[d-1] fastest) -> (dims, nx, ny, nz); zfpCompressorGpu.hpp:110-118."""
    shape = tuple(int(s) for s in shape)
    if len(shape) == 1:
        return 1, shape[0], 1, 1
    if len(shape) == 2:
        return 2, shape[1], shape[0], 1
    if len(shape) == 3:
        return 3, shape[2], shape[1], shape[0]
    raise ValueError("1 to 3 dimensions")


def _dtype_code(dt):
    import numpy as np

    if isinstance(dt, type) and issubclass(dt, np.generic):
        dt = np.dtype(dt)
    else:
        dt = np.dtype(str(dt).replace("torch.", ""))
    if dt == np.float32:
        return FLOAT
    if dt == np.float64:
        return DOUBLE
    raise TypeError("float32 / float64 only: %s" % dt)


def maxbits(dtype, dims, rate, wra=1):
    return int(lib().zfb_maxbits(_dtype_code(dtype), dims, float(rate), int(wra)))


def stream_bytes(shape, mb):
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfb_stream_bytes(d, nx, ny, nz, mb))


def stream_capacity(shape, mb):
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfb_stream_capacity(d, nx, ny, nz, mb))


def mode_rate(dtype, dims, rate, wra=1):
    m = Mode()
    if lib().zfb_mode_rate(C.byref(m), _dtype_code(dtype), dims, float(rate), int(wra)):
        raise ZfbError("zfb_mode_rate: bad arguments")
    return m


def mode_accuracy(tolerance):
    """zfp_stream_set_accuracy: CBench's "abs" parameter (zfp/zfpCompressor.hpp:118)."""
    m = Mode()
    if lib().zfb_mode_accuracy(C.byref(m), float(tolerance)):
        raise ZfbError("zfb_mode_accuracy: bad tolerance")
    return m


def mode_precision(precision):
    """zfp_stream_set_precision: CBench's "rel" parameter (zfp/zfpCompressor.hpp:120)."""
    m = Mode()
    if lib().zfb_mode_precision(C.byref(m), int(precision)):
        raise ZfbError("zfb_mode_precision: bad precision")
    return m


def maximum_size(shape, dtype, mode):
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfb_maximum_size(_dtype_code(dtype), d, nx, ny, nz, C.byref(mode)))


def block_count(shape):
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfb_block_count(d, nx, ny, nz))


def index_unit(dims):
    """Blocks per index entry of a variable-rate stream (1 for 2D/3D, 8 for 1D)."""
    return int(lib().zfb_index_unit(int(dims)))


def index_count(shape):
    """Index entries of a variable-rate stream; the index array holds index_count() + 1 uint64 bit offsets."""
    d, nx, ny, nz = dims_of(shape)
    return int(lib().zfb_index_count(d, nx, ny, nz))


def partition(rank, nranks, shape3):
    """CBench's getPartition (dataLoaderInterface.hpp:123-215): -> (count[3], offset[3]), slowest axis first."""
    n = (C.c_size_t * 3)(*shape3)
    cnt, off = (C.c_size_t * 3)(), (C.c_size_t * 3)()
    rc = lib().zfb_partition(rank, nranks, n, cnt, off)
    if rc:
        raise ZfbError("zfb_partition: %d" % rc)
    return tuple(cnt), tuple(off)


def metric_values(p):
    """The five CSV values (getGlobalValue of each metric, main.cpp:340) from (combined) partials."""
    L = lib()
    r = C.byref(p)
    return {"absolute_error": L.zfb_value_absolute_error(r), "relative_error": L.zfb_value_relative_error(r),
            "mse": L.zfb_value_mse(r), "psnr": L.zfb_value_psnr(r), "minmax": L.zfb_value_minmax(r),
            "min": -p.neg_min_orig, "max": p.max_orig, "sum_abs": p.sum_abs, "sum_rel": p.sum_rel,
            "sum_sq": p.sum_sq, "n": p.n,
            # extra metrics out of the same partials: PSNR over the value range, normalised RMSE
            "psnr_range": L.zfb_value_psnr_range(r), "nrmse": L.zfb_value_nrmse(r)}


def combine_partials(parts):
    """Allreduce semantics of the reference metrics (SUM / MAX / MIN) over a list of Partials."""
    acc = Partials(0.0, 0.0, 0.0, 0.0, 0.0, -1.7976931348623157e308, -1.7976931348623157e308, 0)
    arr = (Partials * len(parts))(*parts)
    lib().zfb_partials_combine(C.byref(acc), arr, len(parts))
    return acc


class GenericIOFile:
    """A GenericIO particle file (HACC): header tables parsed by the C ABI (csrc/zfb_io.cuh); columns are read into
    host arrays (read) or straight into device memory (Codec.gio_read_dev)."""

    def __init__(self, path):
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib().zfb_gio_open(os.fsencode(path), C.byref(h), err, 256)
        if rc != 0:
            raise ZfbError("zfb_gio_open(%s): %s" % (path, err.value.decode()))
        self._h = h
        nv, nr, ne, ok = C.c_int(), C.c_int(), C.c_uint64(), C.c_int()
        lib().zfb_gio_counts(h, C.byref(nv), C.byref(nr), C.byref(ne), C.byref(ok))
        self.nvars, self.nranks, self.nelems, self.header_crc_ok = nv.value, nr.value, ne.value, bool(ok.value)
        self.vars = []
        for k in range(self.nvars):
            name = C.create_string_buffer(256)
            es, fl = C.c_int(), C.c_uint()
            lib().zfb_gio_var(h, k, name, C.byref(es), C.byref(fl))
            self.vars.append({"name": name.value.decode(), "size": es.value, "flags": fl.value, "is_float": bool(fl.value & 1),
                              "is_signed": bool(fl.value & 2)})
        self.rank_elems = [int(lib().zfb_gio_rank_elems(h, r)) for r in range(self.nranks)]
        o, sc = (C.c_double * 3)(), (C.c_double * 3)()
        lib().zfb_gio_phys(h, o, sc)
        self.phys_origin, self.phys_scale = list(o), list(sc)

    def find(self, name):
        return int(lib().zfb_gio_find(self._h, name.encode()))

    def rank_range(self, mpi_rank, mpi_size):
        """File ranks CBench's MPI rank loads (HACCDataLoader.hpp:189-194)."""
        a, b = C.c_int(), C.c_int()
        lib().zfb_gio_rank_range(self.nranks, mpi_rank, mpi_size, C.byref(a), C.byref(b))
        return a.value, b.value

    def dtype_of(self, var):
        import numpy as np

        v = self.vars[var]
        if v["is_float"]:
            return np.dtype("<f%d" % v["size"])
        return np.dtype("<%s%d" % ("i" if v["is_signed"] else "u", v["size"]))

    def read(self, name, r0=0, r1=None, check_crc=True):
        """Column `name` of file ranks [r0, r1) as a host array (what HACCDataLoader::loadData allocates)."""
        import numpy as np

        var = self.find(name)
        if var < 0:
            raise ZfbError("no variable %r" % name)
        r1 = self.nranks if r1 is None else r1
        n = sum(self.rank_elems[r0:r1])
        out = np.empty(n, dtype=self.dtype_of(var))
        got = C.c_uint64()
        rc = lib().zfb_gio_read(None, self._h, var, r0, r1, out.ctypes.data, out.nbytes, C.byref(got), 1 if check_crc else 0)
        if rc != 0 or got.value != n:
            raise ZfbError("zfb_gio_read failed (%d)" % rc)
        return out

    def close(self):
        if self._h:
            lib().zfb_gio_close(self._h)
            self._h = None


def gio_write(path, columns, origin=None, scale=None):
    """One rank's GenericIO file from {name: array} (HACCDataLoader::writeData for a single rank), CRCs included.
    x / y / z get the physical-coordinate flags the HACC loader sets (HACCDataLoader.hpp:385-420)."""
    import numpy as np

    names = list(columns)
    arrs = [np.ascontiguousarray(columns[k]) for k in names]
    n = arrs[0].size
    assert all(a.size == n for a in arrs)
    cn = (C.c_char_p * len(names))(*[k.encode() for k in names])
    fl = (C.c_uint * len(names))()
    szs = (C.c_int * len(names))()
    ptr = (C.c_void_p * len(names))()
    for k, a in enumerate(arrs):
        f = (1 if a.dtype.kind == "f" else 0) | (2 if a.dtype.kind in "fi" else 0)
        f |= {"x": 4, "y": 8, "z": 16}.get(names[k], 0)
        fl[k], szs[k], ptr[k] = f, a.dtype.itemsize, a.ctypes.data
    o = (C.c_double * 3)(*(origin or [0, 0, 0]))
    sc = (C.c_double * 3)(*(scale or [0, 0, 0]))
    err = C.create_string_buffer(256)
    rc = lib().zfb_gio_write_host(os.fsencode(path), len(names), cn, fl, szs, ptr, n, o, sc, err, 256)
    if rc != 0:
        raise ZfbError("zfb_gio_write_host: " + err.value.decode())


def h5_dataset(path, dataset):
    """Location and shape of a contiguous HDF5 dataset: dict(offset, bytes, shape, elem_size, is_float)."""
    off, nb, nd, es, fl = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int(), C.c_int()
    dims = (C.c_uint64 * 8)()
    err = C.create_string_buffer(256)
    rc = lib().zfb_h5_dataset(os.fsencode(path), dataset.encode(), C.byref(off), C.byref(nb), C.byref(nd), dims, C.byref(es),
                              C.byref(fl), err, 256)
    if rc != 0:
        raise ZfbError("zfb_h5_dataset(%s, %s): %s" % (path, dataset, err.value.decode()))
    return {"offset": off.value, "bytes": nb.value, "shape": tuple(int(dims[k]) for k in range(nd.value)),
            "elem_size": es.value, "is_float": bool(fl.value)}


class Codec:
    """One context = one GPU (zfb_create).  Device methods take torch CUDA tensors, host methods numpy arrays."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().zfb_create(C.byref(self._h), int(device))
        if rc:
            self._h = None
            raise ZfbError("zfb_create(device=%d) failed with %d: no usable sm_100 device (there is no CPU fallback)"
                           % (device, rc))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib().zfb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise ZfbError("%s failed (%d): %s" % (what, rc, lib().zfb_last_error(self._h).decode()))

    def use_stream(self, cuda_stream_ptr):
        self._check(lib().zfb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "zfb_set_stream")

    def use_torch_stream(self):
        """Order the codec's kernels with torch ops: enqueue on torch's current stream of this device."""
        import torch

        self.use_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def use_own_stream(self):
        self._check(lib().zfb_reset_stream(self._h), "zfb_reset_stream")

    def synchronize(self):
        self._check(lib().zfb_synchronize(self._h), "zfb_synchronize")

    @property
    def launches(self):
        return int(lib().zfb_launch_count(self._h))

    # ---- device path -------------------------------------------------------------------------
    def encode(self, field, mb, out=None):
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(field.shape)
        nbytes = int(lib().zfb_stream_bytes(d, nx, ny, nz, mb))
        if out is None:
            out = torch.empty(nbytes + 16, dtype=torch.uint8, device=field.device)
        assert field.is_contiguous() and out.numel() >= nbytes
        self._check(lib().zfb_encode_dev(self._h, _dtype_code(field.dtype), d, nx, ny, nz, mb, field.data_ptr(),
                                         out.data_ptr()), "zfb_encode_dev")
        return out[:nbytes]

    def encode_range(self, field, mb, out=None):
        """encode() that also harvests the field's min / max on the way (zfb_encode_range_dev); read them with
        range() -- after allreduce_range() in a multi-GPU job."""
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(field.shape)
        nbytes = int(lib().zfb_stream_bytes(d, nx, ny, nz, mb))
        if out is None:
            out = torch.empty(nbytes + 16, dtype=torch.uint8, device=field.device)
        assert field.is_contiguous() and out.numel() >= nbytes
        self._check(lib().zfb_encode_range_dev(self._h, _dtype_code(field.dtype), d, nx, ny, nz, mb, field.data_ptr(),
                                               out.data_ptr()), "zfb_encode_range_dev")
        return out[:nbytes]

    def allreduce_range(self):
        self._check(lib().zfb_allreduce_range(self._h), "zfb_allreduce_range")

    def range(self):
        lo, hi = C.c_double(), C.c_double()
        self._check(lib().zfb_range_get(self._h, C.byref(lo), C.byref(hi)), "zfb_range_get")
        return lo.value, hi.value

    def decode_hist(self, stream, shape, dtype, mb, lo, hi, orig=None, out=None, counts=None):
        """decode() that also bins the decoded values over [lo, hi] (the minmax metric's histogram) in the same pass.
        Returns (out, counts[:1024], counts[1024] = out-of-range values)."""
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(shape)
        if out is None:
            out = torch.empty(tuple(shape), dtype=dtype, device=stream.device)
        if counts is None:
            counts = torch.empty(NBINS + 1, dtype=torch.int64, device=stream.device)
        self._check(lib().zfb_decode_hist_dev(self._h, _dtype_code(dtype), d, nx, ny, nz, mb, stream.data_ptr(),
                                              out.data_ptr(), orig.data_ptr() if orig is not None else None,
                                              float(lo), float(hi), counts.data_ptr()), "zfb_decode_hist_dev")
        return out, counts[:NBINS], counts[NBINS]

    def decode(self, stream, shape, dtype, mb, orig=None, out=None):
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(shape)
        if out is None:
            out = torch.empty(tuple(shape), dtype=dtype, device=stream.device)
        self._check(lib().zfb_decode_dev(self._h, _dtype_code(dtype), d, nx, ny, nz, mb, stream.data_ptr(),
                                         out.data_ptr(), orig.data_ptr() if orig is not None else None),
                    "zfb_decode_dev")
        return out

    def metrics(self, orig, approx):
        self.use_torch_stream()
        self._check(lib().zfb_metrics_dev(self._h, _dtype_code(orig.dtype), orig.data_ptr(), approx.data_ptr(),
                                          orig.numel()), "zfb_metrics_dev")
        return self.partials()

    def power_spectrum(self, field, nbins=None):
        """Shell-binned power spectrum of a resident 3D float32 field (what PAT's NYX analysis compares, see
        include/zfb.h): returns (k, pk, count) as float64 tensors; k = mean |k| of shell b = round(|k|), pk = mean
        |F|^2 / N^2 of the shell.  The transform is torch.fft.rfftn (cuFFT), the binning is zfb_power_spectrum_bin_dev."""
        import torch

        assert field.dim() == 3 and field.dtype == torch.float32 and field.is_cuda and field.is_contiguous()
        nz, ny, nx = field.shape
        if nbins is None:
            nbins = min(nx, ny, nz) // 2 + 1
        self.use_torch_stream()
        fhat = torch.fft.rfftn(field, dim=(0, 1, 2)).contiguous()
        out = torch.empty(3 * nbins, dtype=torch.float64, device=field.device)
        self._check(lib().zfb_power_spectrum_bin_dev(self._h, fhat.data_ptr(), nx, ny, nz, nbins, out.data_ptr()),
                    "zfb_power_spectrum_bin_dev")
        cnt, sk, sp = out[:nbins], out[nbins:2 * nbins], out[2 * nbins:]
        n = float(nx) * float(ny) * float(nz)
        safe = torch.clamp(cnt, min=1.0)
        return sk / safe, sp / safe / (n * n), cnt

    def partials(self):
        p = Partials()
        self._check(lib().zfb_partials_get(self._h, C.byref(p)), "zfb_partials_get")
        return p

    def set_partials(self, p):
        self._check(lib().zfb_partials_set(self._h, C.byref(p)), "zfb_partials_set")

    def comm_init(self, dist, device):
        """The context's own NCCL communicator over the ranks of a torch.distributed job: rank 0 draws the unique id,
        the 128 bytes travel through the job's existing broadcast (CBench would use MPI_Bcast), every rank joins."""
        import torch

        world, rank = dist.get_world_size(), dist.get_rank()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(lib().zfb_comm_unique_id(buf), "zfb_comm_unique_id")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        self._check(lib().zfb_comm_init(self._h, world, rank, raw), "zfb_comm_init")

    def allreduce_partials(self):
        """One ncclAllGather + rank-ordered combine of the device partials, on the context's stream (no-op on one rank)."""
        self._check(lib().zfb_allreduce_partials(self._h), "zfb_allreduce_partials")

    def allreduce_counts(self, counts):
        self._check(lib().zfb_allreduce_counts(self._h, counts.data_ptr(), counts.numel()), "zfb_allreduce_counts")

    def gio_read_dev(self, gio, name, r0=0, r1=None, check_crc=True):
        """Column `name` of file ranks [r0, r1) of a GenericIOFile straight into a device tensor."""
        import torch

        self.use_torch_stream()
        var = gio.find(name)
        if var < 0:
            raise ZfbError("no variable %r" % name)
        r1 = gio.nranks if r1 is None else r1
        n = sum(gio.rank_elems[r0:r1])
        tdt = {"f4": torch.float32, "f8": torch.float64, "i8": torch.int64, "i4": torch.int32, "u2": torch.int16,
               "i2": torch.int16, "u1": torch.uint8}[gio.dtype_of(var).str[1:]]
        out = torch.empty(n, dtype=tdt, device="cuda:%d" % self.device)
        got = C.c_uint64()
        self._check(lib().zfb_gio_read(self._h, gio._h, var, r0, r1, out.data_ptr(), out.numel() * out.element_size(),
                                       C.byref(got), 1 if check_crc else 0), "zfb_gio_read")
        return out

    def read_block_dev(self, path, base_offset, dtype, total, off, cnt):
        """3D hyperslab (slowest axis first) of a contiguous array in a file -> packed device tensor of shape cnt."""
        import torch

        self.use_torch_stream()
        out = torch.empty(tuple(cnt), dtype=dtype, device="cuda:%d" % self.device)
        t3, o3, c3 = (C.c_size_t * 3)(*total), (C.c_size_t * 3)(*off), (C.c_size_t * 3)(*cnt)
        self._check(lib().zfb_file_read_block_dev(self._h, os.fsencode(path), base_offset, out.element_size(), t3, o3, c3,
                                                  out.data_ptr()), "zfb_file_read_block_dev")
        return out

    def write_block_dev(self, path, base_offset, tensor, total, off):
        """Packed device tensor -> 3D hyperslab of a contiguous array in a file (decompressed write-back)."""
        self.use_torch_stream()
        assert tensor.is_contiguous() and tensor.dim() == 3
        t3, o3, c3 = (C.c_size_t * 3)(*total), (C.c_size_t * 3)(*off), (C.c_size_t * 3)(*tensor.shape)
        self._check(lib().zfb_file_write_block_dev(self._h, os.fsencode(path), base_offset, tensor.element_size(), t3, o3, c3,
                                                   tensor.data_ptr()), "zfb_file_write_block_dev")

    def partials_tensor(self):
        """The context's device-resident zfb_partials viewed as 8 x int64 (raw bits; reinterpret with .view)."""
        import torch

        ptr = lib().zfb_partials_dev(self._h)
        n = C.sizeof(Partials) // 8
        iface = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        holder = type("_P", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device="cuda:%d" % self.device)

    def histogram(self, which, lo, hi, orig, approx):
        import torch

        self.use_torch_stream()
        counts = torch.zeros(NBINS + 1, dtype=torch.int64, device=approx.device)
        self._check(lib().zfb_histogram_dev(self._h, _dtype_code(approx.dtype), which, float(lo), float(hi),
                                            orig.data_ptr() if orig is not None else None, approx.data_ptr(),
                                            approx.numel(), counts.data_ptr(), counts.data_ptr() + 8 * NBINS),
                    "zfb_histogram_dev")
        return counts[:NBINS], counts[NBINS:]

    # ---- variable-rate modes (fixed accuracy / fixed precision) --------------------------------
    def encode_var(self, field, mode, out=None, index=None):
        """Returns (stream[:nbytes], nbytes).  index: optional uint64 tensor of index_count()+1 entries that
        receives the bit offsets; without it the context keeps the index for decode_var of this stream."""
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(field.shape)
        cap = maximum_size(field.shape, field.dtype, mode)
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=field.device)
        assert field.is_contiguous() and out.numel() >= cap
        nbytes = C.c_size_t()
        self._check(lib().zfb_encode_var_dev(self._h, _dtype_code(field.dtype), d, nx, ny, nz, C.byref(mode),
                                             field.data_ptr(), out.data_ptr(), out.numel(),
                                             index.data_ptr() if index is not None else None, C.byref(nbytes)),
                    "zfb_encode_var_dev")
        return out[: nbytes.value], nbytes.value

    def decode_var(self, stream, shape, dtype, mode, index=None, orig=None, out=None):
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(shape)
        if out is None:
            out = torch.empty(tuple(shape), dtype=dtype, device=stream.device)
        self._check(lib().zfb_decode_var_dev(self._h, _dtype_code(dtype), d, nx, ny, nz, C.byref(mode), stream.data_ptr(),
                                             stream.numel(), index.data_ptr() if index is not None else None,
                                             out.data_ptr(), orig.data_ptr() if orig is not None else None),
                    "zfb_decode_var_dev")
        return out

    def index_expand(self, stream, shape, dtype, mode, coarse, stride, out=None):
        """coarse: every stride-th index entry plus the total (uint64 / int64 tensor) -> the full index tensor."""
        import torch

        self.use_torch_stream()
        d, nx, ny, nz = dims_of(shape)
        if out is None:
            out = torch.empty(index_count(shape) + 1, dtype=torch.int64, device=stream.device)
        self._check(lib().zfb_index_expand_dev(self._h, _dtype_code(dtype), d, nx, ny, nz, C.byref(mode), stream.data_ptr(),
                                               stream.numel(), coarse.data_ptr(), int(stride), out.data_ptr()),
                    "zfb_index_expand_dev")
        return out

    def compress_var_host(self, a, mode):
        import numpy as np

        d, nx, ny, nz = dims_of(a.shape)
        cap = maximum_size(a.shape, a.dtype, mode)
        out = np.empty(cap, dtype=np.uint8)
        cb = C.c_size_t()
        self._check(lib().zfb_compress_var_host(self._h, _dtype_code(a.dtype), d, nx, ny, nz, C.byref(mode), a.ctypes.data,
                                                out.ctypes.data, cap, C.byref(cb)), "zfb_compress_var_host")
        return out[: cb.value]

    def decompress_var_host(self, stream, shape, dtype, mode, orig=None, out=None):
        import numpy as np

        d, nx, ny, nz = dims_of(shape)
        if out is None:
            out = np.empty(tuple(shape), dtype=dtype)
        self._check(lib().zfb_decompress_var_host(self._h, _dtype_code(dtype), d, nx, ny, nz, C.byref(mode),
                                                  stream.ctypes.data, stream.size, out.ctypes.data,
                                                  orig.ctypes.data if orig is not None else None),
                    "zfb_decompress_var_host")
        return out

    # ---- host path (what the CBench plugin calls) -------------------------------------------
    def compress_host(self, a, mb, out=None):
        import numpy as np

        d, nx, ny, nz = dims_of(a.shape)
        nbytes = int(lib().zfb_stream_bytes(d, nx, ny, nz, mb))
        if out is None:
            out = np.empty(nbytes, dtype=np.uint8)
        cb = C.c_size_t()
        self._check(lib().zfb_compress_host(self._h, _dtype_code(a.dtype), d, nx, ny, nz, mb, a.ctypes.data,
                                            out.ctypes.data, C.byref(cb)), "zfb_compress_host")
        return out[: cb.value]

    def decompress_host(self, stream, shape, dtype, mb, orig=None, out=None):
        import numpy as np

        d, nx, ny, nz = dims_of(shape)
        if out is None:
            out = np.empty(tuple(shape), dtype=dtype)
        self._check(lib().zfb_decompress_host(self._h, _dtype_code(dtype), d, nx, ny, nz, mb, stream.ctypes.data,
                                              out.ctypes.data, orig.ctypes.data if orig is not None else None),
                    "zfb_decompress_host")
        return out

    def compress_host_ptr(self, dtype_code, shape, mb, in_ptr, out_ptr):
        d, nx, ny, nz = dims_of(shape)
        cb = C.c_size_t()
        self._check(lib().zfb_compress_host(self._h, dtype_code, d, nx, ny, nz, mb, in_ptr, out_ptr, C.byref(cb)),
                    "zfb_compress_host")
        return cb.value

    def decompress_host_ptr(self, dtype_code, shape, mb, stream_ptr, out_ptr, orig_ptr=None):
        d, nx, ny, nz = dims_of(shape)
        self._check(lib().zfb_decompress_host(self._h, dtype_code, d, nx, ny, nz, mb, stream_ptr, out_ptr, orig_ptr),
                    "zfb_decompress_host")

    def metrics_host(self, orig, approx, reuse=None):
        """Metric partials of two host arrays.  The C ABI recognises buffers it already holds by address and size only,
        and a numpy array can be rewritten in place: unless `reuse` is True, the device copies are dropped first --
        except when decompress_host registered the pair since the previous call (then the copies are this very data
        and the fused partials are returned without any upload)."""
        serial = lib().zfb_host_pair_serial(self._h)
        if reuse is None:
            reuse = serial != getattr(self, "_pair_serial_seen", 0)
        self._pair_serial_seen = serial
        if not reuse:
            self._check(lib().zfb_host_cache_invalidate(self._h), "zfb_host_cache_invalidate")
        p = Partials()
        self._check(lib().zfb_metrics_host(self._h, _dtype_code(orig.dtype), orig.ctypes.data, approx.ctypes.data,
                                           orig.size, C.byref(p)), "zfb_metrics_host")
        return p

    def histogram_host(self, which, lo, hi, orig, approx):
        import numpy as np

        counts = np.zeros(NBINS, dtype=np.uint64)
        oob = C.c_uint64()
        self._check(lib().zfb_histogram_host(self._h, _dtype_code(approx.dtype), which, float(lo), float(hi),
                                             orig.ctypes.data if orig is not None else None, approx.ctypes.data,
                                             approx.size, counts.ctypes.data, C.byref(oob)), "zfb_histogram_host")
        return counts, oob.value

    def release(self):
        self._check(lib().zfb_release_resident(self._h), "zfb_release_resident")
