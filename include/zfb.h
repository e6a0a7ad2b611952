/*
 * include/zfb.h -- C ABI of the B200-native zfp fixed-rate codec + fused CBench error metrics.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point names
 * the reference interface it replaces (paths relative to the lanl/VizAly-Foresight tree).  The CBench
 * plugin shim that binds these (vizaly-foresight_b200/cbench/zfpB200Compressor.hpp and
 * gpuMetrics.hpp) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - dtype: ZFB_FLOAT (1) / ZFB_DOUBLE (2)        (zfp_type_float / zfp_type_double,
 *                                                    zfpCompressorGpu.hpp:56-66)
 *   - dims 1..3 and nx,ny,nz with x the FASTEST axis (zfp_field_3d(input, type, n[2], n[1], n[0]),
 *     zfpCompressorGpu.hpp:110-118).  Unused extents are 1.
 *   - maxbits: bits per 4^dims block, from zfb_maxbits() (zfp_stream_set_rate, :129)
 *   - every function returns 0 on success or a negative ZFB_E* code; nothing throws across the ABI.
 *   - "dev" entry points take DEVICE pointers and enqueue on the context's stream (or the stream given
 *     to zfb_set_stream) without synchronising; "host" entry points take HOST pointers, do the
 *     host<->device copies themselves and return when the result is in host memory.
 *   - There is no CPU fallback: without a usable CUDA device every compute call returns ZFB_ENODEV.
 *   - A context is not thread-safe: use it from one host thread at a time (CBench is single-threaded per rank,
 *     one context per rank / GPU); distinct contexts are independent.
 */
#ifndef ZFB_H
#define ZFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZFB_FLOAT 1
#define ZFB_DOUBLE 2

#define ZFB_OK 0
#define ZFB_EINVAL (-1)   /* bad argument (dims, dtype, maxbits below 1+EBITS, null pointer)          */
#define ZFB_ENODEV (-2)   /* no CUDA device / CUDA extension unusable: there is no CPU fallback        */
#define ZFB_ECUDA (-3)    /* a CUDA call failed; zfb_last_error() has the text                          */
#define ZFB_ENOMEM (-4)
#define ZFB_ESTATE (-5)   /* call order violated (e.g. fused metrics requested with no resident field)  */

#define ZFB_NBINS 1024    /* histogram bins (absoluteError.hpp:111, relativeError.hpp:124, minmaxMetric.hpp:98) */

typedef struct zfb_ctx zfb_ctx;

/* Everything the five CBench metrics need, for one rank's sub-field.  The SUM / MAX / MIN members are
 * exactly the quantities the reference MPI_Allreduces (absoluteError.hpp:80-92, relativeError.hpp:93-105,
 * meansquareError.hpp:70-71, psnrError.hpp:75-83, minmaxMetric.hpp:77-81), laid out so that a multi-GPU
 * combine is three small all-reduces: SUM over sum[0..2] and n, MAX over max[0..3] (min is stored
 * negated so that MAX works for it too). */
typedef struct zfb_partials {
  double sum_sq;    /* sum (o-a)^2          -> mse, psnr                                   */
  double sum_abs;   /* sum |o-a|            -> absolute_error log lines                     */
  double sum_rel;   /* sum relError(o,a,1)  -> relative_error log lines                     */
  double max_abs;   /* max |o-a|            -> absolute_error value                         */
  double max_rel;   /* max relError         -> relative_error value                         */
  double max_orig;  /* max o                -> psnr, minmax value                           */
  double neg_min_orig; /* -(min o)          -> minmax log                                   */
  uint64_t n;       /* number of values                                                     */
} zfb_partials;

/* ---- parameters (pure host arithmetic, usable without a GPU) ------------------------------------ */

/* zfp_stream_set_rate(zfp, rate, type, dims, wra): bits per block.
 * Replaces the call at zfpCompressorGpu.hpp:129,232 (CBench passes wra = 8, i.e. non-zero). */
unsigned zfb_maxbits(int dtype, int dims, double rate, int wra);

/* Bytes zfp_compress() returns (= CompressorInterface::cbytes, zfpCompressorGpu.hpp:145,156-157). */
size_t zfb_stream_bytes(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits);

/* Bytes zfp_stream_maximum_size() returns (buffer CBench mallocs, zfpCompressorGpu.hpp:134-135). */
size_t zfb_stream_capacity(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits);

/* CBench's recursive-bisection domain split (dataLoaderInterface.hpp:123-215): extents n[3] (slowest
 * first) over nranks ranks; writes this rank's count[3] and offset[3]. */
int zfb_partition(int rank, int nranks, const size_t n[3], size_t count[3], size_t offset[3]);

/* ---- context ------------------------------------------------------------------------------------- */

/* Binds a context to CUDA device `device` (CompressorInterface::init, zfpCompressorGpu.hpp:50-53, plus
 * zfp_stream_set_execution(zfp, zfp_exec_cuda), :143).  Returns ZFB_ENODEV when no device is usable. */
int zfb_create(zfb_ctx** ctx, int device);
int zfb_destroy(zfb_ctx* ctx);                       /* CompressorInterface::close, :269-272 */
/* Enqueue all "dev" calls on an existing CUDA stream (a cudaStream_t passed as void*; NULL is the legacy
 * default stream, as in the CUDA runtime).  zfb_reset_stream goes back to the context's own stream. */
int zfb_set_stream(zfb_ctx* ctx, void* cuda_stream);
int zfb_reset_stream(zfb_ctx* ctx);
int zfb_synchronize(zfb_ctx* ctx);
const char* zfb_last_error(const zfb_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t zfb_launch_count(const zfb_ctx* ctx);

/* ---- device-pointer hot path ----------------------------------------------------------------------- */

/* zfp_compress (zfpCompressorGpu.hpp:145): d_in = field, d_stream receives zfb_stream_bytes() bytes. */
int zfb_encode_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                   const void* d_in, void* d_stream);

/* zfp_decompress (zfpCompressorGpu.hpp:249) fused with the five metric passes of main.cpp:306-352:
 * d_out receives the decompressed field; when d_orig != NULL the metric partials of (d_orig, d_out)
 * are accumulated in the same pass and left in the context (zfb_partials_dev / zfb_partials_get). */
int zfb_decode_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                   const void* d_stream, void* d_out, const void* d_orig);

/* Stand-alone streaming metric pass over two device arrays (any compressor's output):
 * absoluteError/relativeError/meansquareError/psnrError/minmaxMetric::execute first loops.
 * Both pointers must be 16-byte aligned (any cudaMalloc'ed array is). */
int zfb_metrics_dev(zfb_ctx* ctx, int dtype, const void* d_orig, const void* d_approx, size_t n);

/* Device address of the context's zfb_partials (valid after the stream reaches the producing call);
 * multi-GPU callers all-reduce it in place (NCCL) before reading. */
void* zfb_partials_dev(zfb_ctx* ctx);
/* Synchronise and copy the partials to the host. */
int zfb_partials_get(zfb_ctx* ctx, zfb_partials* out);
/* Overwrite the device partials (used after a host-side / MPI combine so histograms see global values). */
int zfb_partials_set(zfb_ctx* ctx, const zfb_partials* in);

/* ---- multi-GPU exchange: the metric partials of all ranks over NCCL / NVLink --------------------------
 * Replaces the three blocking MPI_Allreduce calls every metric object makes (absoluteError.hpp:80-92,
 * relativeError.hpp:93-105, meansquareError.hpp:70-71, psnrError.hpp:75-83, minmaxMetric.hpp:77-81): one
 * ncclAllGather of the 64-byte zfb_partials and a rank-ordered combine on the device -- the result is bitwise the same
 * on every rank and equal to zfb_partials_combine over the ranks in order.  NCCL is loaded at run time (libnccl.so.2,
 * or ZFB_NCCL_LIB); there is no link-time dependency.
 *   rank 0: zfb_comm_unique_id(id) -> broadcast the 128 bytes (MPI_Bcast in CBench) -> every rank:
 *   zfb_comm_init(ctx, nranks, rank, id).  zfb_comm_adopt takes an ncclComm_t the caller already owns (same NCCL
 *   library).  zfb_allreduce_partials / zfb_allreduce_counts run on the context's stream and are no-ops on a
 *   context without a communicator (single rank). */
int zfb_comm_unique_id(void* out128);
int zfb_comm_init(zfb_ctx* ctx, int nranks, int rank, const void* id128);
int zfb_comm_adopt(zfb_ctx* ctx, void* nccl_comm, int nranks, int rank);
int zfb_comm_size(const zfb_ctx* ctx);
int zfb_comm_destroy(zfb_ctx* ctx);
int zfb_allreduce_partials(zfb_ctx* ctx);
int zfb_allreduce_counts(zfb_ctx* ctx, uint64_t* d_counts, size_t count);

/* ---- value-range (minmax) histogram without a second pass over memory ----------------------------------
 * minmaxMetric.hpp:62-133 bins the DECOMPRESSED values over [global min, global max] of the ORIGINAL field.
 * zfb_encode_range_dev = zfb_encode_dev that also harvests the field's min / max (fused into the float 3D tile kernel;
 * a streaming pass for the other kernels); zfb_allreduce_range combines the ranges of all ranks (same single
 * all-gather as the partials); zfb_range_get reads the range; zfb_decode_hist_dev = zfb_decode_dev that also bins the
 * values it decodes into d_counts[0 .. ZFB_NBINS-1] (d_counts[ZFB_NBINS] = out-of-range values, clamped to the end
 * bins), fused into the float 3D tile kernel -- no extra memory traffic -- and a stand-alone pass otherwise.  Counts
 * equal the reference's loop (minmaxMetric.hpp:108-117) bin for bin. */
int zfb_encode_range_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                         const void* d_in, void* d_stream);
int zfb_allreduce_range(zfb_ctx* ctx);
int zfb_range_get(zfb_ctx* ctx, double* lo, double* hi);
int zfb_decode_hist_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                        const void* d_stream, void* d_out, const void* d_orig, double lo, double hi, uint64_t* d_counts);

/* Second-pass histograms (absoluteError.hpp:104-139, relativeError.hpp:117-153, minmaxMetric.hpp:92-133).
 * which: 0 = abs error over [0, hi], 1 = rel error over [0, hi], 2 = approx values over [lo, hi].
 * d_counts: ZFB_NBINS uint64 device counters, zero-initialised by the call.  Values outside the range
 * are clamped to the end bins (the reference indexes out of bounds there) and counted in *d_oob
 * (nullable device uint64). */
int zfb_histogram_dev(zfb_ctx* ctx, int dtype, int which, double lo, double hi, const void* d_orig,
                      const void* d_approx, size_t n, uint64_t* d_counts, uint64_t* d_oob);

/* ---- host-pointer entry points (what the CBench plugin binds) -------------------------------------- */

/* ZFPCompressorGpu::compress body (zfpCompressorGpu.hpp:104-157): h_in host field -> h_stream (caller
 * allocated, >= zfb_stream_bytes()).  The uploaded field stays resident on the device, keyed by h_in,
 * so that the following zfb_decompress_host can fuse the metrics without a second upload. */
int zfb_compress_host(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                      const void* h_in, void* h_stream, size_t* cbytes);

/* ZFPCompressorGpu::decompress body (zfpCompressorGpu.hpp:205-262): h_stream -> h_out (caller allocated,
 * nx*ny*nz*sizeof(T)).  If h_orig != NULL the metrics of (h_orig, h_out) are fused into the pass; the
 * resident copy from zfb_compress_host is used when h_orig is the pointer it was given. */
int zfb_decompress_host(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                        const void* h_stream, void* h_out, const void* h_orig);

/* The host-pointer entry points recognise "the buffers I already hold on the device" by host address and byte size
 * only.  They cannot notice a buffer that was rewritten in place, or freed and re-allocated at the same address --
 * CBench does the latter between scalars (main.cpp:395-398: std::free(decompdata), ioMgr->close()).  A caller whose
 * buffers may have changed since the last call drops the residency first; zfb_host_pair_serial counts the
 * zfb_decompress*_host calls that registered an (original, decompressed) pair, so a caller can tell whether the pair
 * it is about to ask metrics for was produced by this library since it last looked (cbench/gpuMetrics.hpp does). */
int zfb_host_cache_invalidate(zfb_ctx* ctx);
unsigned long long zfb_host_pair_serial(const zfb_ctx* ctx);

/* MetricInterface::execute(original, approx, n) for host arrays (main.cpp:336): reuses the fused
 * partials when (h_orig, h_approx) are the buffers of the last zfb_decompress_host, else uploads both
 * (see zfb_host_cache_invalidate for when the caller must drop the residency). */
int zfb_metrics_host(zfb_ctx* ctx, int dtype, const void* h_orig, const void* h_approx, size_t n,
                     zfb_partials* out);

/* Histogram of host arrays; counts[ZFB_NBINS] host uint64.  Uses resident device copies when the
 * pointers match the last compress/decompress pair. */
int zfb_histogram_host(zfb_ctx* ctx, int dtype, int which, double lo, double hi, const void* h_orig,
                       const void* h_approx, size_t n, uint64_t* counts, uint64_t* oob);

/* Drop the resident device copies (called from CompressorInterface::close / between scalars). */
int zfb_release_resident(zfb_ctx* ctx);

/* ---- variable-rate streams: zfp's fixed-accuracy / fixed-precision modes ----------------------------
 * What the CPU zfp plugin exposes as "abs" and "rel" (CBench/compressors/zfp/zfpCompressor.hpp:83-93,
 * 117-120) and every shipped input uses (e.g. inputs/hacc/hacc_cbench_test.json).  Blocks have different
 * lengths and are concatenated bit-tight, so a decoder needs the bit offset of every block (the index):
 * zfb keeps it from the encode (as ZFPCompressor::decompress relies on state left by compress,
 * zfpCompressor.hpp:22,160) or hands it to the caller; a bare stream is indexed by a sequential walk. */

/* zfp_stream's four parameters */
typedef struct zfb_mode {
  unsigned minbits, maxbits, maxprec;
  int minexp;
} zfb_mode;

int zfb_mode_rate(zfb_mode* m, int dtype, int dims, double rate, int wra); /* zfp_stream_set_rate      */
int zfb_mode_accuracy(zfb_mode* m, double tolerance);                      /* zfp_stream_set_accuracy  (:118) */
int zfb_mode_precision(zfb_mode* m, unsigned precision);                   /* zfp_stream_set_precision (:120) */

/* Bytes zfp_stream_maximum_size() returns for this mode (buffer CBench mallocs, zfpCompressor.hpp:123-124). */
size_t zfb_maximum_size(int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m);
/* Number of 4^d blocks. */
size_t zfb_block_count(int dims, size_t nx, size_t ny, size_t nz);
/* The index: zfb_index_count() + 1 uint64 bit offsets, one per zfb_index_unit(dims) consecutive blocks
 * (2D/3D: every block; 1D: every 8th block -- a 1D block is 16 bytes of data, finer bookkeeping would cost
 * more than the data), the total number of bits last. */
size_t zfb_index_unit(int dims);
size_t zfb_index_count(int dims, size_t nx, size_t ny, size_t nz);

/* zfp_compress() in any mode on device buffers (zfpCompressor.hpp:132).  d_stream holds cap_bytes >=
 * zfb_maximum_size().  d_index: nullable device array of zfb_index_count()+1 uint64 that receives the
 * index; with NULL the context keeps the index for the next
 * zfb_decode_var_dev of the same d_stream.  *stream_bytes = what zfp_compress returns (the call
 * synchronises: the size is data dependent). */
int zfb_encode_var_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                       const void* d_in, void* d_stream, size_t cap_bytes, uint64_t* d_index, size_t* stream_bytes);

/* zfp_decompress() in any mode (zfpCompressor.hpp:224).  d_index: the index of this stream, or NULL: the
 * context's index is used if it was built for this d_stream (same address and length: if the buffer has
 * been overwritten with another stream since, pass that stream's index or call zfb_release_resident
 * first), else the stream is walked block by block to rebuild it (one thread: ~3 us per block, i.e. seconds per million
 * blocks).  d_orig nullable: fuses the metrics as zfb_decode_dev does. */
int zfb_decode_var_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                       const void* d_stream, size_t stream_bytes, const uint64_t* d_index, void* d_out,
                       const void* d_orig);

/* A stream kept for later needs its index, but not all of it: keep every `stride`-th entry (entries 0, stride,
 * 2*stride, ... of the index, then the total number of bits: ceil(zfb_index_count()/stride) + 1 values, e.g. 1 MB
 * for the 16 Mi blocks of the benchmark slab at stride 128) and expand it again here: the segments between two
 * kept entries are walked in parallel (milliseconds instead of the sequential walk's minutes).  Returns ZFB_ESTATE
 * if a segment does not end where the next kept entry says (wrong parameters or corrupt stream).  Synchronises. */
int zfb_index_expand_dev(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                         const void* d_stream, size_t stream_bytes, const uint64_t* d_coarse, size_t stride,
                         uint64_t* d_index);

/* Host-pointer forms (ZFPCompressor::compress / decompress bodies, zfpCompressor.hpp:96-141,186-234).
 * The uploaded field, the stream and its index stay resident, keyed by h_in / h_stream. */
int zfb_compress_var_host(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                          const void* h_in, void* h_stream, size_t cap_bytes, size_t* cbytes);
int zfb_decompress_var_host(zfb_ctx* ctx, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                            const void* h_stream, size_t stream_bytes, void* h_out, const void* h_orig);

/* ---- scalar metric values from (combined) partials: pure host arithmetic --------------------------- */
double zfb_value_absolute_error(const zfb_partials* p); /* absoluteError.hpp:78-82   */
double zfb_value_relative_error(const zfb_partials* p); /* relativeError.hpp:91-95   */
double zfb_value_mse(const zfb_partials* p);            /* meansquareError.hpp:64-72 */
double zfb_value_psnr(const zfb_partials* p);           /* psnrError.hpp:85-88       */
double zfb_value_minmax(const zfb_partials* p);         /* minmaxMetric.hpp:89-90    */
/* extra metrics from the same partials: PSNR over the value range (max - min) and the normalised RMSE */
double zfb_value_psnr_range(const zfb_partials* p);
double zfb_value_nrmse(const zfb_partials* p);
/* Shell-binned power spectrum of a resident 3D float field, the quantity Foresight's NYX analysis compares between the
 * original and every decompressed field: PAT plots the ratio of columns 2 and 3 (k, P(k)) of gimlet's *_ps3d.txt files
 * (Analysis/pat/nyx/cinema.py:84-130, Analysis/pat/nyx/README.md:19-25; gimlet itself is not part of the reference, and
 * the ratio needs no particular normalisation).  d_fhat: the real-to-complex transform of the field, nz x ny x (nx/2+1)
 * complex float values, x fastest (the cufftExecR2C / numpy rfftn layout).  d_out receives 3 x nbins doubles:
 * [0] number of modes in shell b = round(|k|), [1] sum of |k| over them, [2] sum of |F|^2; nbins <= 2048, extents < 4096.
 * The mirror modes the half-spectrum leaves out are counted with weight 2. */
int zfb_power_spectrum_bin_dev(zfb_ctx* ctx, const void* d_fhat, size_t nx, size_t ny, size_t nz, unsigned nbins,
                               double* d_out);
/* Combine `count` ranks' partials the way the reference's Allreduces do (SUM / MAX / MIN). */
void zfb_partials_combine(zfb_partials* acc, const zfb_partials* in, int count);

/* ---- device-resident loaders / writers (what hands the hot path its input, and takes its output) ----------
 * The three file kinds CBench loads from, read straight into device memory (file -> pinned double buffer -> device, no
 * pageable host array in between) so that zfb_encode_dev can start from there, and the decompressed write-back of
 * main.cpp:440-491.  What lands on the device is byte for byte what the reference loader hands compress().
 *
 * GenericIO (HACC; format GenericIO.cxx:373-441, loader HACCDataLoader.hpp:143-330): */
typedef struct zfb_gio zfb_gio;
int zfb_gio_open(const char* path, zfb_gio** out, char* errbuf, size_t errlen);
int zfb_gio_close(zfb_gio* g);
int zfb_gio_counts(const zfb_gio* g, int* nvars, int* nranks, uint64_t* nelems_total, int* header_crc_ok);
int zfb_gio_var(const zfb_gio* g, int i, char* name256, int* elem_size, unsigned* flags);
int zfb_gio_find(const zfb_gio* g, const char* name);                 /* index of a variable, -1 if absent */
uint64_t zfb_gio_rank_elems(const zfb_gio* g, int rank);              /* gioReader->readNumElems(rank) */
int zfb_gio_phys(const zfb_gio* g, double* origin3, double* scale3);
/* the file ranks MPI rank `mpi_rank` of `mpi_size` loads (HACCDataLoader.hpp:189-194) */
void zfb_gio_rank_range(int nfile_ranks, int mpi_rank, int mpi_size, int* r0, int* r1);
/* variable `var` of file ranks [r0, r1), concatenated: into device memory (ctx != NULL) or host memory (ctx == NULL) */
int zfb_gio_read(zfb_ctx* ctx, zfb_gio* g, int var, int r0, int r1, void* dst, size_t dst_bytes, uint64_t* nelems,
                 int check_crc);
/* one rank's file from host arrays (HACCDataLoader::writeData, HACCDataLoader.hpp:358-429), CRCs included */
int zfb_gio_write_host(const char* path, int nvars, const char* const* names, const unsigned* flags, const int* sizes,
                       const void* const* data, uint64_t nelems, const double* origin3, const double* scale3, char* errbuf,
                       size_t errlen);
/* HDF5 (NYX; NYXDataLoader.hpp:143-251): where a CONTIGUOUS dataset lives (no libhdf5 needed for that layout) */
int zfb_h5_dataset(const char* path, const char* dataset, uint64_t* offset, uint64_t* bytes, int* ndims, uint64_t* dims8,
                   int* elem_size, int* is_float, char* errbuf, size_t errlen);
/* a 3D hyperslab (slowest axis first, like sizePerDim / rankOffset) of a contiguous array in a file <-> device memory:
 * the HDF5 dataset above at its offset, or a raw .gda array at offset 0 (GDADataLoader.hpp:118-171); the write goes
 * into an existing copy of the file (NYX write-back) or creates a raw file */
int zfb_file_read_block_dev(zfb_ctx* ctx, const char* path, uint64_t base_offset, int elem_size, const size_t* total3,
                            const size_t* off3, const size_t* cnt3, void* d_dst);
int zfb_file_write_block_dev(zfb_ctx* ctx, const char* path, uint64_t base_offset, int elem_size, const size_t* total3,
                             const size_t* off3, const size_t* cnt3, const void* d_src);

#ifdef __cplusplus
}
#endif
#endif /* ZFB_H */
