// zfb_api_var.cuh -- C ABI of the variable-rate (fixed-accuracy / fixed-precision) path; included at the end of
// zfb_api.cu (shares its context and helpers).  Kernels: zfb_var.cuh.
#pragma once
#include "zfb_var.cuh"

// ------------------------------------------------------------------------------------------------
// pure host arithmetic [zfp-recall: zfp 0.5.5 src/zfp.c]
// ------------------------------------------------------------------------------------------------
static const unsigned ZFB_MIN_BITS = 1u, ZFB_MAX_BITS = 16657u, ZFB_MAX_PREC = 64u;
static const int ZFB_MIN_EXP = -1074;

extern "C" int zfb_mode_rate(zfb_mode* m, int dtype, int dims, double rate, int wra)
{
  if (!m) return ZFB_EINVAL;
  const unsigned mb = zfb_maxbits(dtype, dims, rate, wra);
  if (!mb) return ZFB_EINVAL;
  m->minbits = m->maxbits = mb;
  m->maxprec = ZFB_MAX_PREC;
  m->minexp = ZFB_MIN_EXP;
  return ZFB_OK;
}

extern "C" int zfb_mode_accuracy(zfb_mode* m, double tolerance)
{
  if (!m || !(tolerance >= 0)) return ZFB_EINVAL;
  int emin = ZFB_MIN_EXP;
  if (tolerance > 0) {
    frexp(tolerance, &emin);  // tolerance = x * 2^emin, 0.5 <= x < 1
    emin--;                   // 2^emin <= tolerance < 2^(emin + 1)
  }
  m->minbits = ZFB_MIN_BITS;
  m->maxbits = ZFB_MAX_BITS;
  m->maxprec = ZFB_MAX_PREC;
  m->minexp = emin;
  return ZFB_OK;
}

extern "C" int zfb_mode_precision(zfb_mode* m, unsigned precision)
{
  if (!m) return ZFB_EINVAL;
  m->minbits = ZFB_MIN_BITS;
  m->maxbits = ZFB_MAX_BITS;
  m->maxprec = precision ? (precision < ZFB_MAX_PREC ? precision : ZFB_MAX_PREC) : ZFB_MAX_PREC;
  m->minexp = ZFB_MIN_EXP;
  return ZFB_OK;
}

extern "C" size_t zfb_block_count(int dims, size_t nx, size_t ny, size_t nz)
{
  if (dims < 1 || dims > 3) return 0;
  return count_blocks(dims, nx, ny, nz);
}

static size_t index_unit(int dims) { return dims == 1 ? (size_t)VAR1_GROUP : 1; }

extern "C" size_t zfb_index_unit(int dims) { return dims >= 1 && dims <= 3 ? index_unit(dims) : 0; }

extern "C" size_t zfb_index_count(int dims, size_t nx, size_t ny, size_t nz)
{
  if (dims < 1 || dims > 3) return 0;
  const size_t u = index_unit(dims);
  return (count_blocks(dims, nx, ny, nz) + u - 1) / u;
}

extern "C" size_t zfb_maximum_size(int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m)
{
  if (!m || dims < 1 || dims > 3 || (dtype != ZFB_FLOAT && dtype != ZFB_DOUBLE)) return 0;
  const unsigned values = 1u << (2 * dims);
  const unsigned tprec = dtype == ZFB_FLOAT ? 32u : 64u;
  unsigned maxbits = 1u + (dtype == ZFB_FLOAT ? 8u : 11u);
  maxbits += values - 1u + values * (m->maxprec < tprec ? m->maxprec : tprec);
  if (maxbits > m->maxbits) maxbits = m->maxbits;
  if (maxbits < m->minbits) maxbits = m->minbits;
  return ((148 + count_blocks(dims, nx, ny, nz) * (size_t)maxbits + 63) & ~(size_t)63) / 8;
}

// ------------------------------------------------------------------------------------------------
// device path
// ------------------------------------------------------------------------------------------------
static int check_var_args(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m)
{
  if (!c) return ZFB_EINVAL;
  if (dims < 1 || dims > 3 || !nx || (dims >= 2 && !ny) || (dims >= 3 && !nz)) { c->err = "bad dims"; return ZFB_EINVAL; }
  if (dtype != ZFB_FLOAT && dtype != ZFB_DOUBLE) { c->err = "bad dtype"; return ZFB_EINVAL; }
  if (!m) { c->err = "null mode"; return ZFB_EINVAL; }
  const unsigned hdr = dtype == ZFB_FLOAT ? 9u : 12u;
  if (m->maxbits < hdr || m->minbits > m->maxbits || m->maxbits > ZFB_MAX_BITS || !m->maxprec || m->maxprec > 64u ||
      m->minexp < ZFB_MIN_EXP || m->minexp > 2048) {
    c->err = "bad mode parameters";
    return ZFB_EINVAL;
  }
  return ZFB_OK;
}

static stream_params to_params(const zfb_mode* m)
{
  stream_params sp;
  sp.minbits = m->minbits; sp.maxbits = m->maxbits; sp.maxprec = m->maxprec; sp.minexp = m->minexp;
  return sp;
}

template <typename Scalar> static unsigned var_words(int dims)  // staging words per index unit
{
  return dims == 1 ? var1_caps<Scalar>::WORDS : dims == 2 ? var_caps<Scalar, 2>::WORDS : var_caps<Scalar, 3>::WORDS;
}

static int var_grid(const zfb_ctx* c, size_t nblocks)
{
  const size_t want = (nblocks + 127) / 128, cap = (size_t)c->sm_count * 32;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

template <typename Scalar>
static int launch_enc_var(zfb_ctx* c, int dims, const geom& g, const stream_params& sp, const void* d_in, uint64_t* stage,
                          uint32_t* lens)
{
  const int grid = var_grid(c, (g.nblocks + index_unit(dims) - 1) / index_unit(dims));
  const Scalar* in = (const Scalar*)d_in;
  if (dims == 1) enc_var1_kernel<Scalar><<<grid, 128, 0, c->stream>>>(in, stage, lens, g, sp);
  else if (dims == 2) enc_var_kernel<Scalar, 2><<<grid, 128, 0, c->stream>>>(in, stage, lens, g, sp);
  else if constexpr (sizeof(Scalar) == 8) {
    const size_t want = (g.nblocks + VAR64_THREADS - 1) / VAR64_THREADS, cap = (size_t)c->sm_count * 32;
    enc_var_f64_kernel<<<(unsigned)(want < cap ? (want ? want : 1) : cap), VAR64_THREADS, 0, c->stream>>>((const double*)d_in, stage, lens, g, sp);
  } else enc_var_kernel<Scalar, 3><<<grid, 128, 0, c->stream>>>(in, stage, lens, g, sp);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

template <typename Scalar>
static int launch_compact_var(zfb_ctx* c, int dims, const geom& g, const uint64_t* stage, const uint64_t* index,
                              uint64_t* stream, size_t cap_words)
{
  const size_t units = (g.nblocks + index_unit(dims) - 1) / index_unit(dims);
  const int grid = var_grid(c, units);
  constexpr bool COL32 = true;  // 3D blocks are staged by the 32-bit table coders
  if (dims == 1) compact_var_kernel<var1_caps<Scalar>::WORDS, false><<<grid, 128, 0, c->stream>>>(stage, index, units, stream, cap_words);
  else if (dims == 2) compact_var_kernel<var_caps<Scalar, 2>::WORDS, false><<<grid, 128, 0, c->stream>>>(stage, index, units, stream, cap_words);
  else compact_var_kernel<var_caps<Scalar, 3>::WORDS, COL32><<<grid, 128, 0, c->stream>>>(stage, index, units, stream, cap_words);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

// One (sub-)field -> its part of the stream, no host synchronisation.  `index` points at this part's first entry of
// the (global) index; with `chained` the bit count already sitting there (the total of the parts before) is the
// base of this part, so consecutive slabs of a field can be queued back to back.
static int encode_var_part(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                           const void* d_in, void* d_stream, size_t cap_words, uint64_t* index, bool chained)
{
  int rc;
  geom g = make_geom(dims, nx, ny, nz, 64);
  g.stream_words = cap_words;
  const stream_params sp = to_params(m);
  const unsigned words = dtype == ZFB_FLOAT ? var_words<float>(dims) : var_words<double>(dims);
  const size_t units = (g.nblocks + index_unit(dims) - 1) / index_unit(dims);  // index entries (1D: groups of blocks)
  const size_t groups = (units + 31) / 32;
  if ((rc = ensure(c, c->var_stage, groups * 32 * (size_t)words * 8))) return rc;
  if ((rc = ensure(c, c->var_lens, units * 4))) return rc;
  const size_t nchunks = (units + SCAN_CHUNK - 1) / SCAN_CHUNK;
  if ((rc = ensure(c, c->var_chunks, (nchunks + 1) * 8))) return rc;
  uint64_t* stage = (uint64_t*)c->var_stage.p;
  uint32_t* lens = (uint32_t*)c->var_lens.p;
  uint64_t* chunks = (uint64_t*)c->var_chunks.p;
  const uint64_t* base = chained ? index : nullptr;

  rc = dtype == ZFB_FLOAT ? launch_enc_var<float>(c, dims, g, sp, d_in, stage, lens)
                          : launch_enc_var<double>(c, dims, g, sp, d_in, stage, lens);
  if (rc) return rc;
  scan_sums_kernel<<<(unsigned)nchunks, SCAN_THREADS, 0, c->stream>>>(lens, units, chunks);
  scan_chunks_kernel<<<1, SCAN_THREADS, 0, c->stream>>>(chunks, nchunks);
  scan_offsets_kernel<<<(unsigned)nchunks, SCAN_THREADS, 0, c->stream>>>(lens, units, chunks, nchunks, index, base);
  zero_stream_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>((uint64_t*)d_stream, base, index + units, cap_words);
  c->launches += 4;
  CU_TRY(c, cudaGetLastError());
  return dtype == ZFB_FLOAT ? launch_compact_var<float>(c, dims, g, stage, index, (uint64_t*)d_stream, cap_words)
                            : launch_compact_var<double>(c, dims, g, stage, index, (uint64_t*)d_stream, cap_words);
}

static int check_var_minbits(zfb_ctx* c, int dtype, int dims, const zfb_mode* m)
{
  if (dims == 1 && m->minbits > (dtype == ZFB_FLOAT ? var_caps<float, 1>::BITS : var_caps<double, 1>::BITS)) {
    c->err = "1D: minbits above the longest codable block is not supported";
    return ZFB_EINVAL;
  }
  return ZFB_OK;
}

extern "C" int zfb_encode_var_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                                  const void* d_in, void* d_stream, size_t cap_bytes, uint64_t* d_index,
                                  size_t* stream_bytes)
{
  int rc = check_var_args(c, dtype, dims, nx, ny, nz, m);
  if (rc) return rc;
  if (!d_in || !d_stream) { c->err = "null pointer"; return ZFB_EINVAL; }
  if (cap_bytes < zfb_maximum_size(dtype, dims, nx, ny, nz, m)) { c->err = "stream buffer smaller than zfb_maximum_size()"; return ZFB_EINVAL; }
  if ((rc = check_var_minbits(c, dtype, dims, m))) return rc;
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t units = zfb_index_count(dims, nx, ny, nz);
  uint64_t* index = d_index;
  c->var_index_key = nullptr;
  if (!index) {
    if ((rc = ensure(c, c->var_index, (units + 1) * 8))) return rc;
    index = (uint64_t*)c->var_index.p;
  }
  if ((rc = encode_var_part(c, dtype, dims, nx, ny, nz, m, d_in, d_stream, cap_bytes / 8, index, false))) return rc;
  uint64_t total_bits = 0;
  CU_TRY(c, cudaMemcpyAsync(&total_bits, index + units, 8, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  const size_t bytes = (size_t)((total_bits + 63) / 64) * 8;  // stream_flush pads to a word
  if (bytes > cap_bytes) { c->err = "internal: stream longer than its bound"; return ZFB_ESTATE; }
  if (stream_bytes) *stream_bytes = bytes;
  if (!d_index) {
    c->var_index_key = d_stream;
    c->var_index_blocks = units;
    c->var_index_bits = total_bits;
  }
  return ZFB_OK;
}

template <typename Scalar>
static int launch_index_var(zfb_ctx* c, int dims, const void* d_stream, size_t stream_words, size_t nblocks,
                            const stream_params& sp, uint64_t* index)
{
  const uint64_t* st = (const uint64_t*)d_stream;
  if (dims == 1) index_var_kernel<Scalar, 1><<<1, 32, 0, c->stream>>>(st, stream_words, nblocks, sp, index);
  else if (dims == 2) index_var_kernel<Scalar, 2><<<1, 32, 0, c->stream>>>(st, stream_words, nblocks, sp, index);
  else index_var_kernel<Scalar, 3><<<1, 32, 0, c->stream>>>(st, stream_words, nblocks, sp, index);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

template <typename Scalar, bool METRICS>
static int launch_dec_var(zfb_ctx* c, int dims, const geom& g, const stream_params& sp, const void* d_stream,
                          const uint64_t* index, void* d_out, const void* d_orig, int* rows)
{
  int grid = var_grid(c, (g.nblocks + index_unit(dims) - 1) / index_unit(dims));
  if (grid > c->part_rows) grid = c->part_rows;
  const uint64_t* st = (const uint64_t*)d_stream;
  Scalar* out = (Scalar*)d_out;
  const Scalar* orig = (const Scalar*)d_orig;
  if (dims == 1) dec_var1_kernel<Scalar, METRICS><<<grid, 128, 0, c->stream>>>(st, index, out, orig, g, sp, c->part);
  else if (dims == 2) dec_var_kernel<Scalar, 2, METRICS><<<grid, 128, 0, c->stream>>>(st, index, out, orig, g, sp, c->part);
  else if constexpr (sizeof(Scalar) == 4) {
    const int persistent = c->sm_count * 4;  // 4 resident CTAs per SM, tiles of 128 blocks taken round-robin
    if (grid > persistent) grid = persistent;
    dec_var_f32_kernel<METRICS><<<grid, 128, 0, c->stream>>>(st, index, (float*)d_out, (const float*)d_orig, g, sp, c->part);
  } else dec_var_kernel<Scalar, 3, METRICS><<<grid, 128, 0, c->stream>>>(st, index, out, orig, g, sp, c->part);
  c->launches++;
  *rows = grid;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

template <typename Scalar>
static void launch_index_expand(zfb_ctx* c, int dims, const void* d_stream, size_t stream_words, size_t nblocks,
                                const stream_params& sp, const uint64_t* coarse, size_t stride, uint64_t* index, unsigned* bad,
                                int grid)
{
  const uint64_t* st = (const uint64_t*)d_stream;
  if (dims == 1) index_expand_kernel<Scalar, 1><<<grid, 64, 0, c->stream>>>(st, stream_words, nblocks, sp, coarse, stride, index, bad);
  else if (dims == 2) index_expand_kernel<Scalar, 2><<<grid, 64, 0, c->stream>>>(st, stream_words, nblocks, sp, coarse, stride, index, bad);
  else index_expand_kernel<Scalar, 3><<<grid, 64, 0, c->stream>>>(st, stream_words, nblocks, sp, coarse, stride, index, bad);
}

extern "C" int zfb_index_expand_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                                    const void* d_stream, size_t stream_bytes, const uint64_t* d_coarse, size_t stride,
                                    uint64_t* d_index)
{
  int rc = check_var_args(c, dtype, dims, nx, ny, nz, m);
  if (rc) return rc;
  if (!d_stream || !d_coarse || !d_index || !stride || !stream_bytes) { c->err = "null pointer / zero stride"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t nblocks = count_blocks(dims, nx, ny, nz);
  const size_t units = (nblocks + index_unit(dims) - 1) / index_unit(dims);
  const size_t nseg = (units + stride - 1) / stride;
  if ((rc = ensure(c, c->var_chunks, 16))) return rc;
  unsigned* bad = (unsigned*)c->var_chunks.p;
  CU_TRY(c, cudaMemsetAsync(bad, 0, 4, c->stream));
  const size_t want = (nseg + 63) / 64, cap = (size_t)c->sm_count * 32;
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  const stream_params sp = to_params(m);
  if (dtype == ZFB_FLOAT) launch_index_expand<float>(c, dims, d_stream, (stream_bytes + 7) / 8, nblocks, sp, d_coarse, stride, d_index, bad, grid);
  else launch_index_expand<double>(c, dims, d_stream, (stream_bytes + 7) / 8, nblocks, sp, d_coarse, stride, d_index, bad, grid);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  unsigned h_bad = 0;
  CU_TRY(c, cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  if (h_bad) { c->err = "coarse index does not match the stream (wrong mode parameters or corrupt data)"; return ZFB_ESTATE; }
  return ZFB_OK;
}

// Decode one (sub-)field whose index entries start at `index`; metrics (d_orig != NULL) end in the context's partials.
static int decode_var_part(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                           const void* d_stream, size_t stream_words, const uint64_t* index, void* d_out, const void* d_orig)
{
  int rc;
  geom g = make_geom(dims, nx, ny, nz, 64);
  g.stream_words = stream_words;
  const stream_params sp = to_params(m);
  int rows = 0;
  const bool metrics = d_orig != nullptr;
  if (dtype == ZFB_DOUBLE && dims == 3 && (!metrics || ((((uintptr_t)d_orig | (uintptr_t)d_out) & 15) == 0))) {
    const size_t want = (g.nblocks + VAR64_THREADS - 1) / VAR64_THREADS, cap = (size_t)c->sm_count * 32;
    dec_var_f64_kernel<<<(unsigned)(want < cap ? (want ? want : 1) : cap), VAR64_THREADS, 0, c->stream>>>(
        (const uint64_t*)d_stream, index, (double*)d_out, g, sp);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return metrics ? zfb_metrics_dev(c, dtype, d_orig, d_out, g.nx * g.ny * g.nz) : ZFB_OK;
  }
  if (dtype == ZFB_FLOAT) {
    rc = metrics ? launch_dec_var<float, true>(c, dims, g, sp, d_stream, index, d_out, d_orig, &rows)
                 : launch_dec_var<float, false>(c, dims, g, sp, d_stream, index, d_out, d_orig, &rows);
  } else {
    rc = metrics ? launch_dec_var<double, true>(c, dims, g, sp, d_stream, index, d_out, d_orig, &rows)
                 : launch_dec_var<double, false>(c, dims, g, sp, d_stream, index, d_out, d_orig, &rows);
  }
  if (rc) return rc;
  if (metrics) return finalize(c, rows, g.nx * g.ny * g.nz);
  return ZFB_OK;
}

extern "C" int zfb_decode_var_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                                  const void* d_stream, size_t stream_bytes, const uint64_t* d_index, void* d_out,
                                  const void* d_orig)
{
  int rc = check_var_args(c, dtype, dims, nx, ny, nz, m);
  if (rc) return rc;
  if (!d_stream || !d_out || !stream_bytes) { c->err = "null pointer"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t stream_words = (stream_bytes + 7) / 8;
  const size_t nblocks = count_blocks(dims, nx, ny, nz);
  const uint64_t* index = d_index;
  const size_t units = (nblocks + index_unit(dims) - 1) / index_unit(dims);
  if (!index) {
    if (c->var_index_key == d_stream && c->var_index_blocks == units && c->var_index.p &&
        (size_t)((c->var_index_bits + 63) / 64) == stream_words) {
      index = (const uint64_t*)c->var_index.p;
    } else {
      c->var_index_key = nullptr;
      const stream_params sp = to_params(m);
      if ((rc = ensure(c, c->var_index, (units + 1) * 8))) return rc;
      rc = dtype == ZFB_FLOAT ? launch_index_var<float>(c, dims, d_stream, stream_words, nblocks, sp, (uint64_t*)c->var_index.p)
                              : launch_index_var<double>(c, dims, d_stream, stream_words, nblocks, sp, (uint64_t*)c->var_index.p);
      if (rc) return rc;
      index = (const uint64_t*)c->var_index.p;
    }
  }
  return decode_var_part(c, dtype, dims, nx, ny, nz, m, d_stream, stream_words, index, d_out, d_orig);
}

// ------------------------------------------------------------------------------------------------
// host-pointer forms: slabs along the slowest axis, pipelined like the fixed-rate entry points.  The slabs of a
// variable-rate stream join at arbitrary bits, so slab s+1 is queued with the running bit count that slab s leaves
// in the index as its base (no host round trip), and the host only learns the slab boundaries one slab late, which
// is when it queues that slab's share of the download.
// ------------------------------------------------------------------------------------------------
static slab_plan plan_var_slabs(int dims, size_t nx, size_t ny, size_t nz, size_t esz)
{
  slab_plan sp = plan_slabs(dims, nx, ny, nz, 128, esz);  // thickness a multiple of 4
  if (dims == 1 && sp.nslabs > 1) {  // 1D: whole index units (8 blocks = 32 values)
    size_t t = sp.thick / 32 * 32;
    if (t < 32) t = 32;
    sp.thick = t >= sp.slowest ? sp.slowest : t;
    sp.nslabs = (sp.slowest + sp.thick - 1) / sp.thick;
  }
  return sp;
}

extern "C" int zfb_compress_var_host(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                                     const void* h_in, void* h_stream, size_t cap_bytes, size_t* cbytes)
{
  int rc = check_var_args(c, dtype, dims, nx, ny, nz, m);
  if (rc) return rc;
  if (!h_in || !h_stream) { c->err = "null pointer"; return ZFB_EINVAL; }
  const size_t need = zfb_maximum_size(dtype, dims, nx, ny, nz, m);
  if (cap_bytes < need) { c->err = "stream buffer smaller than zfb_maximum_size()"; return ZFB_EINVAL; }
  if ((rc = check_var_minbits(c, dtype, dims, m))) return rc;
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ZFB_FLOAT ? 4 : 8;
  const size_t n = nx * (dims >= 2 ? ny : 1) * (dims >= 3 ? nz : 1);
  const size_t units = zfb_index_count(dims, nx, ny, nz);
  if ((rc = ensure(c, c->d_orig, n * esz))) return rc;
  if ((rc = ensure(c, c->d_stream, need + 16))) return rc;
  if ((rc = ensure(c, c->var_index, (units + 1) * 8))) return rc;
  c->fused_valid = false;
  c->h_stream_key = nullptr;  // d_stream is about to hold a variable-rate stream
  c->h_orig_key = nullptr;
  c->var_host_key = nullptr;
  c->var_index_key = nullptr;
  c->var_slab_bits.clear();

  const slab_plan sp = plan_var_slabs(dims, nx, ny, nz, esz);
  uint64_t* index = (uint64_t*)c->var_index.p;
  uint64_t* h_tot = nullptr;  // pinned: running bit count after every slab
  if (cudaMallocHost(&h_tot, (sp.nslabs + 1) * 8) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMallocHost failed"; return ZFB_ENOMEM; }
  cudaStream_t up, down;
  std::vector<cudaEvent_t> ev_up(sp.nslabs), ev_k(sp.nslabs);
  cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking);
  for (size_t i = 0; i < sp.nslabs; i++) { cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming); }
  int status = ZFB_OK;
  size_t row0 = 0, u0 = 0, sent_words = 0;
  c->var_slab_bits.push_back(0);
  auto download_upto = [&](size_t slab) {  // queue the words that slab `slab` (already finished on the device) completed
    if (cudaEventSynchronize(ev_k[slab]) != cudaSuccess) { status = ZFB_ECUDA; return; }
    const uint64_t bits = h_tot[slab];
    c->var_slab_bits.push_back(bits);
    const bool last = slab + 1 == sp.nslabs;
    const size_t upto = last ? (size_t)((bits + 63) / 64) : (size_t)(bits / 64);  // the word a later slab still ORs into stays
    if (upto > sent_words) {
      if (cudaMemcpyAsync((char*)h_stream + sent_words * 8, (char*)c->d_stream.p + sent_words * 8, (upto - sent_words) * 8,
                          cudaMemcpyDeviceToHost, down) != cudaSuccess) status = ZFB_ECUDA;
      sent_words = upto;
    }
  };
  for (size_t s = 0; s < sp.nslabs && status == ZFB_OK; s++) {
    const size_t rows = (row0 + sp.thick <= sp.slowest) ? sp.thick : sp.slowest - row0;
    size_t sx, sy, sz;
    slab_dims(dims, nx, ny, nz, rows, sx, sy, sz);
    const size_t off_e = row0 * sp.row_elems;
    char* dsrc = (char*)c->d_orig.p + off_e * esz;
    if (cudaMemcpyAsync(dsrc, (const char*)h_in + off_e * esz, rows * sp.row_elems * esz, cudaMemcpyHostToDevice, up) != cudaSuccess) { status = ZFB_ECUDA; break; }
    cudaEventRecord(ev_up[s], up);
    cudaStreamWaitEvent(c->stream, ev_up[s], 0);
    status = encode_var_part(c, dtype, dims, sx, sy, sz, m, dsrc, c->d_stream.p, need / 8, index + u0, s > 0);
    if (status) break;
    u0 += zfb_index_count(dims, sx, sy, sz);
    cudaMemcpyAsync(&h_tot[s], index + u0, 8, cudaMemcpyDeviceToHost, c->stream);
    cudaEventRecord(ev_k[s], c->stream);
    if (s > 0) download_upto(s - 1);  // one slab late: the upload and the kernels of slab s are already queued
    row0 += rows;
  }
  if (status == ZFB_OK) download_upto(sp.nslabs - 1);
  cudaError_t e1 = cudaStreamSynchronize(up), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(down);
  for (size_t i = 0; i < sp.nslabs; i++) { cudaEventDestroy(ev_up[i]); cudaEventDestroy(ev_k[i]); }
  cudaStreamDestroy(up);
  cudaStreamDestroy(down);
  const uint64_t total_bits = status == ZFB_OK ? h_tot[sp.nslabs - 1] : 0;
  cudaFreeHost(h_tot);
  if (status == ZFB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
    c->err = std::string("compress_var_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3);
    status = ZFB_ECUDA;
  }
  if (status) { c->var_slab_bits.clear(); return status; }
  if (u0 != units) { c->err = "internal: slabs do not tile the index"; return ZFB_ESTATE; }
  const size_t bytes = (size_t)((total_bits + 63) / 64) * 8;
  if (bytes > cap_bytes) { c->err = "internal: stream longer than its bound"; return ZFB_ESTATE; }
  if (cbytes) *cbytes = bytes;
  c->h_orig_key = h_in;
  c->orig_bytes = n * esz;
  c->var_index_key = c->d_stream.p;
  c->var_index_blocks = units;
  c->var_index_bits = total_bits;
  c->var_host_key = h_stream;   // the resident stream + index describe this host buffer ...
  c->var_host_bytes = bytes;
  sample_of(h_stream, bytes, c->stream_sample);  // ... as long as it still holds these bytes, coded with these parameters
  c->var_host_mode = *m;
  return ZFB_OK;
}

extern "C" int zfb_decompress_var_host(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, const zfb_mode* m,
                                       const void* h_stream, size_t stream_bytes, void* h_out, const void* h_orig)
{
  int rc = check_var_args(c, dtype, dims, nx, ny, nz, m);
  if (rc) return rc;
  if (!h_out || !h_stream || !stream_bytes) { c->err = "null pointer"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ZFB_FLOAT ? 4 : 8;
  const size_t n = nx * (dims >= 2 ? ny : 1) * (dims >= 3 ? nz : 1);
  const size_t units = zfb_index_count(dims, nx, ny, nz);
  if ((rc = ensure(c, c->d_out, n * esz))) return rc;
  const bool metrics = h_orig != nullptr;
  const bool orig_resident = metrics && h_orig == c->h_orig_key && c->orig_bytes == n * esz && c->d_orig.p;
  if (metrics && !orig_resident) {
    if ((rc = ensure(c, c->d_orig, n * esz))) return rc;
  }
  c->fused_valid = false;
  c->h_stream_key = nullptr;  // d_stream is about to hold a variable-rate stream
  const slab_plan sp = plan_var_slabs(dims, nx, ny, nz, esz);
  // the index kept by zfb_compress_var_host is only trusted together with the bytes it was built from, so the
  // stream is uploaded again (cheap next to the field) and the index reused when the host buffer is the same
  bool index_resident = c->var_host_key == h_stream && c->var_host_bytes == stream_bytes &&
                        c->var_index_key == c->d_stream.p && c->d_stream.p && c->var_index_blocks == units &&
                        c->var_slab_bits.size() == sp.nslabs + 1 && memcmp(&c->var_host_mode, m, sizeof *m) == 0;
  if (index_resident) {  // same address and size is not enough: the buffer must still hold the stream the index describes
    unsigned char now[128];
    sample_of(h_stream, stream_bytes, now);
    index_resident = memcmp(now, c->stream_sample, sizeof now) == 0;
  }
  if (!index_resident) {
    // a stream from elsewhere: upload it whole, rebuild the index on the device, decode in one piece
    if ((rc = ensure(c, c->d_stream, stream_bytes + 16))) return rc;
    c->var_index_key = nullptr;
    c->var_host_key = nullptr;
    CU_TRY(c, cudaMemcpyAsync(c->d_stream.p, h_stream, stream_bytes, cudaMemcpyHostToDevice, c->stream));
    if (metrics && !orig_resident) CU_TRY(c, cudaMemcpyAsync(c->d_orig.p, h_orig, n * esz, cudaMemcpyHostToDevice, c->stream));
    if ((rc = zfb_decode_var_dev(c, dtype, dims, nx, ny, nz, m, c->d_stream.p, stream_bytes, nullptr, c->d_out.p,
                                 metrics ? c->d_orig.p : nullptr)))
      return rc;
    CU_TRY(c, cudaMemcpyAsync(h_out, c->d_out.p, n * esz, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
  } else {
    const size_t stream_words = (stream_bytes + 7) / 8;
    const uint64_t* index = (const uint64_t*)c->var_index.p;
    cudaStream_t up, down;
    std::vector<cudaEvent_t> ev_up(sp.nslabs), ev_k(sp.nslabs);
    cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking);
    for (size_t i = 0; i < sp.nslabs; i++) { cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming); }
    std::vector<zfb_partials> slab_parts(metrics ? sp.nslabs : 0);
    partials_dev* d_slab_parts = nullptr;
    int status = ZFB_OK;
    if (metrics && cudaMalloc(&d_slab_parts, sp.nslabs * sizeof(partials_dev)) != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc failed"; status = ZFB_ENOMEM; }
    size_t row0 = 0, u0 = 0, sent_words = 0;
    for (size_t s = 0; s < sp.nslabs && status == ZFB_OK; s++) {
      const size_t rows = (row0 + sp.thick <= sp.slowest) ? sp.thick : sp.slowest - row0;
      size_t sx, sy, sz;
      slab_dims(dims, nx, ny, nz, rows, sx, sy, sz);
      const size_t off_e = row0 * sp.row_elems, slab_bytes = rows * sp.row_elems * esz;
      // the words this slab's blocks read: up to its last bit, plus the decoders' look-ahead
      size_t upto = (size_t)((c->var_slab_bits[s + 1] + 63) / 64) + 2;
      if (upto > stream_words || s + 1 == sp.nslabs) upto = stream_words;
      if (upto > sent_words) {
        if (cudaMemcpyAsync((char*)c->d_stream.p + sent_words * 8, (const char*)h_stream + sent_words * 8, (upto - sent_words) * 8,
                            cudaMemcpyHostToDevice, up) != cudaSuccess) { status = ZFB_ECUDA; break; }
        sent_words = upto;
      }
      char* dorig = metrics ? (char*)c->d_orig.p + off_e * esz : nullptr;
      if (metrics && !orig_resident &&
          cudaMemcpyAsync(dorig, (const char*)h_orig + off_e * esz, slab_bytes, cudaMemcpyHostToDevice, up) != cudaSuccess) { status = ZFB_ECUDA; break; }
      cudaEventRecord(ev_up[s], up);
      cudaStreamWaitEvent(c->stream, ev_up[s], 0);
      char* dout = (char*)c->d_out.p + off_e * esz;
      status = decode_var_part(c, dtype, dims, sx, sy, sz, m, c->d_stream.p, stream_words, index + u0, dout, dorig);
      if (status) break;
      if (metrics) cudaMemcpyAsync(d_slab_parts + s, c->result, sizeof(partials_dev), cudaMemcpyDeviceToDevice, c->stream);
      cudaEventRecord(ev_k[s], c->stream);
      cudaStreamWaitEvent(down, ev_k[s], 0);
      if (cudaMemcpyAsync((char*)h_out + off_e * esz, dout, slab_bytes, cudaMemcpyDeviceToHost, down) != cudaSuccess) { status = ZFB_ECUDA; break; }
      u0 += zfb_index_count(dims, sx, sy, sz);
      row0 += rows;
    }
    cudaError_t e1 = cudaStreamSynchronize(up), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(down);
    if (status == ZFB_OK && metrics &&
        cudaMemcpy(slab_parts.data(), d_slab_parts, sp.nslabs * sizeof(partials_dev), cudaMemcpyDeviceToHost) != cudaSuccess) status = ZFB_ECUDA;
    if (d_slab_parts) cudaFree(d_slab_parts);
    for (size_t i = 0; i < sp.nslabs; i++) { cudaEventDestroy(ev_up[i]); cudaEventDestroy(ev_k[i]); }
    cudaStreamDestroy(up);
    cudaStreamDestroy(down);
    if (status == ZFB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
      c->err = std::string("decompress_var_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3);
      status = ZFB_ECUDA;
    }
    if (status) return status;
    if (metrics) {
      zfb_partials total;
      memset(&total, 0, sizeof total);
      total.max_orig = -1.7976931348623157e308;
      total.neg_min_orig = -1.7976931348623157e308;
      zfb_partials_combine(&total, slab_parts.data(), (int)sp.nslabs);
      if ((rc = zfb_partials_set(c, &total))) return rc;
    }
  }
  if (metrics) {
    c->h_orig_key = h_orig;
    c->orig_bytes = n * esz;
    c->fused_valid = true;
  }
  c->h_out_key = h_out;
  c->out_bytes = n * esz;
  c->pair_serial++;
  return ZFB_OK;
}
