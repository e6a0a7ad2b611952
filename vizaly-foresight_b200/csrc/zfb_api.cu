// zfb_api.cu -- the C ABI of include/zfb.h: context, kernel dispatch, TMA descriptors, host-pointer
// wrappers.  No CPU fallback anywhere: every compute entry point needs a live CUDA context.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <deque>

#include "../../include/zfb.h"
#include "zfb_kernels.cuh"
#include "zfb_ws.cuh"
#include "zfb_spectrum.cuh"
#include "zfb_hostio.cuh"
#include "zfb_comm.cuh"

using namespace zfb;

static_assert(sizeof(zfb_partials) == sizeof(partials_dev), "zfb_partials layout");

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct devbuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct zfb_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  encode_tiled_fn encode_tiled = nullptr;
  double* part = nullptr;       // per-CTA partial rows
  int part_rows = 0;
  partials_dev* result = nullptr;
  partials_dev* enc_result = nullptr;   // min / max of the last field encoded by zfb_encode_range_dev (other members unused)
  uint64_t* hist_counts = nullptr;      // zfb_decode_dev bins the decoded values into these counters on the way (zfb_decode_hist_dev)
  double hist_lo = 0, hist_hi = 0;
  bool want_range = false;              // zfb_encode_dev harvests min / max on the way (set by zfb_encode_range_dev)
  uint64_t launches = 0;
  std::string err;
  bool force_generic = false;
  bool dec_ws = true;   // warp-specialised float decode kernel (ZFB_DEC_WS=0 switches back to the tile kernel)
  bool enc_warp = true; // warp-autonomous float encode kernel (ZFB_ENC_WARP=0 switches back to the tile kernel)
  int occ_encw[64] = {0};
  // host-pointer path: resident device copies
  devbuf d_orig, d_stream, d_out, d_hist, d_pk;
  const void* h_orig_key = nullptr;   // host pointer the resident original came from
  size_t orig_bytes = 0;
  const void* h_out_key = nullptr;    // host pointer the last decompressed field went to
  size_t out_bytes = 0;
  bool fused_valid = false;           // result holds metrics of (h_orig_key, h_out_key)
  unsigned long long pair_serial = 0; // zfb_decompress*_host calls that registered an (original, decompressed) pair
  const void* h_stream_key = nullptr; // host buffer whose fixed-rate stream is still in d_stream (zfb_compress_host)
  size_t stream_key_bytes = 0;
  unsigned stream_key_maxbits = 0;
  unsigned char stream_sample[128] = {0};
  // multi-GPU exchange (zfb_comm.cuh)
  ncclComm_t comm = nullptr;
  bool comm_owned = false;
  int comm_ranks = 1, comm_rank = 0;
  partials_dev* gathered = nullptr;   // comm_ranks entries
  copy_pool* pool = nullptr;          // copy threads for pageable caller buffers (zfb_hostio.cuh)
  pinned_pair pin_in, pin_out, pin_st;  // pinned double buffers: field slabs in / out, stream pieces
  int occ_enc[80] = {0}, occ_dec[80] = {0}, occ_decm[80] = {0};  // resident CTAs per SM, by slot words
  std::vector<uint64_t> var_slab_bits;  // host copy: bit offset at which every slab of the resident stream starts (+ total)
  // variable-rate path (zfb_api_var.cuh): scratch + the block index of the last encoded stream
  devbuf var_stage, var_lens, var_chunks, var_index;
  const void* var_index_key = nullptr;  // device stream the index in var_index describes
  size_t var_index_blocks = 0;
  uint64_t var_index_bits = 0;
  const void* var_host_key = nullptr;   // host stream buffer the resident d_stream + index came from
  size_t var_host_bytes = 0;
  zfb_mode var_host_mode = {};          // parameters the resident variable-rate stream was coded with
};

#define CU_TRY(ctx, call)                                                                   \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return ZFB_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// pure host arithmetic
// ------------------------------------------------------------------------------------------------
extern "C" unsigned zfb_maxbits(int dtype, int dims, double rate, int wra)
{
  if (dims < 1 || dims > 3 || (dtype != ZFB_FLOAT && dtype != ZFB_DOUBLE)) return 0;
  const unsigned n = 1u << (2 * dims);
  unsigned bits = (unsigned)std::floor(n * rate + 0.5);
  const unsigned lo = dtype == ZFB_FLOAT ? 9u : 12u;   // 1 + EBITS
  if (bits < lo) bits = lo;
  if (wra) bits = (bits + 63u) & ~63u;                 // stream_word_bits = 64
  return bits;
}

static size_t count_blocks(int dims, size_t nx, size_t ny, size_t nz)
{
  size_t b = (nx + 3) / 4;
  if (dims >= 2) b *= (ny + 3) / 4;
  if (dims >= 3) b *= (nz + 3) / 4;
  return b;
}

extern "C" size_t zfb_stream_bytes(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits)
{
  const size_t bits = count_blocks(dims, nx, ny, nz) * (size_t)maxbits;
  return ((bits + 63) / 64) * 8;
}

extern "C" size_t zfb_stream_capacity(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits)
{
  const size_t bits = 148 + count_blocks(dims, nx, ny, nz) * (size_t)maxbits;  // ZFP_HEADER_MAX_BITS
  return ((bits + 63) & ~(size_t)63) / 8;
}

extern "C" int zfb_partition(int rank, int nranks, const size_t n[3], size_t count[3], size_t offset[3])
{
  if (!n || !count || !offset || nranks < 1 || rank < 0 || rank >= nranks) return ZFB_EINVAL;
  struct box { size_t lo[3], hi[3]; };
  std::deque<box> q;
  q.push_back(box{{0, 0, 0}, {n[0], n[1], n[2]}});
  int axis = 0;
  int stuck = 0;
  while ((int)q.size() < nranks && stuck < 3) {
    const size_t cur = q.size();
    bool split_any = false;
    for (size_t i = 0; i < cur; i++) {
      box parent = q.front();
      q.pop_front();
      if (parent.hi[axis] - parent.lo[axis] <= 1) { q.push_back(parent); continue; }
      box a = parent, b = parent;
      a.hi[axis] = (parent.lo[axis] + parent.hi[axis]) / 2;
      b.lo[axis] = a.hi[axis];
      q.push_back(a);
      q.push_back(b);
      split_any = true;
      if ((int)q.size() >= nranks) break;
    }
    stuck = split_any ? 0 : stuck + 1;
    axis = (axis + 1) % 3;
  }
  if (rank >= (int)q.size()) return ZFB_EINVAL;
  const box& me = q[rank];
  for (int d = 0; d < 3; d++) { count[d] = me.hi[d] - me.lo[d]; offset[d] = me.lo[d]; }
  return ZFB_OK;
}

extern "C" double zfb_value_absolute_error(const zfb_partials* p) { return p->max_abs; }
extern "C" double zfb_value_relative_error(const zfb_partials* p) { return p->max_rel; }
extern "C" double zfb_value_mse(const zfb_partials* p) { return p->sum_sq / (double)p->n; }
extern "C" double zfb_value_psnr(const zfb_partials* p)
{
  const double mse = p->sum_sq / (double)p->n;
  const double mx = p->max_orig > -999999999.0 ? p->max_orig : -999999999.0;  // the reference's running maximum starts there (psnrError.hpp:58)
  return 10 * std::log10(std::pow(mx, 2.0) / mse);
}
extern "C" double zfb_value_minmax(const zfb_partials* p) { return p->max_orig; }
// Extra metrics from the same partials (no further pass): PSNR over the value RANGE (max - min), the usual definition
// in the compression literature -- CBench's "psnr" uses max(original) alone (psnrError.hpp:85-88) -- and the
// normalised root-mean-square error.
extern "C" double zfb_value_psnr_range(const zfb_partials* p)
{
  const double mse = p->sum_sq / (double)p->n, range = p->max_orig + p->neg_min_orig;
  return 10 * std::log10(range * range / mse);
}
extern "C" double zfb_value_nrmse(const zfb_partials* p)
{
  const double range = p->max_orig + p->neg_min_orig;
  return std::sqrt(p->sum_sq / (double)p->n) / range;
}

extern "C" void zfb_partials_combine(zfb_partials* acc, const zfb_partials* in, int count)
{
  for (int i = 0; i < count; i++) {
    const zfb_partials& s = in[i];
    acc->sum_sq += s.sum_sq; acc->sum_abs += s.sum_abs; acc->sum_rel += s.sum_rel;
    acc->max_abs = std::fmax(acc->max_abs, s.max_abs); acc->max_rel = std::fmax(acc->max_rel, s.max_rel);
    acc->max_orig = std::fmax(acc->max_orig, s.max_orig);
    acc->neg_min_orig = std::fmax(acc->neg_min_orig, s.neg_min_orig);
    acc->n += s.n;
  }
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
static const size_t ENC_TILE_SMEM_BASE = BOXES_PER_TILE * BOX_FLOATS * sizeof(float);

// [tile / plane scratch 32 KiB][2 x slot staging][mbarriers]  (layouts documented at the kernels)
// padded slot staging: (2W + 1) 32-bit words per thread, rounded to 16 bytes
static size_t padded_slots(unsigned W) { return (((size_t)TILE_THREADS * (2 * W + 2) * 4 + 15) & ~(size_t)15); }
static size_t enc_tile_smem(unsigned W) { return ENC_TILE_SMEM_BASE + (size_t)TILE_THREADS * W * 8 + 3 * 8; }
static size_t dec_tile_smem(unsigned W, bool /*metrics*/)
{
  return ENC_TILE_SMEM_BASE + DEC_SLOT_BUFS * (size_t)TILE_THREADS * W * 8 + padded_slots(W) + 16 + 4 * 8;
}

extern "C" int zfb_create(zfb_ctx** out, int device)
{
  if (!out) return ZFB_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return ZFB_ENODEV;  // no CPU fallback
  }
  zfb_ctx* c = new (std::nothrow) zfb_ctx();
  if (!c) return ZFB_ENOMEM;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return ZFB_ENODEV; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return ZFB_ENODEV; }
  if (prop.major < 10) {  // kernels are built for sm_100a only
    delete c;
    return ZFB_ENODEV;
  }
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ZFB_ECUDA; }
  c->stream = c->own_stream;
  c->part_rows = c->sm_count * 32;
  if (cudaMalloc(&c->part, (size_t)c->part_rows * PART_STRIDE * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&c->result, sizeof(partials_dev)) != cudaSuccess ||
      cudaMalloc(&c->enc_result, sizeof(partials_dev)) != cudaSuccess) {
    delete c;
    return ZFB_ENOMEM;
  }
  cudaMemset(c->result, 0, sizeof(partials_dev));
  cudaMemset(c->enc_result, 0, sizeof(partials_dev));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    c->encode_tiled = (encode_tiled_fn)fn;
  const char* fg = std::getenv("ZFB_FORCE_GENERIC");
  c->force_generic = fg && fg[0] == '1';
  // opt in to large dynamic shared memory for the tile kernels
  // (the limit applies to static + dynamic, and the decode kernels keep a small static reduction buffer)
  const int max_smem = 100 * 1024 + 1024;
  if (cudaFuncSetAttribute(enc3_f32_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(enc3_f32_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(dec3_f32_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(dec3_f32_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(dec3_f32_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
    cudaGetLastError();
    c->encode_tiled = nullptr;  // tile kernels unusable: everything goes through the generic kernels
  }
  if (cudaFuncSetAttribute(enc3_f64_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(dec3_f64_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(dec3_f64_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
    cudaGetLastError();
    c->encode_tiled = nullptr;
  }
  // warp-specialised float decode (zfb_ws.cuh): one 640-thread CTA per SM with up to 227 KiB of shared memory
  {
    const char* e = std::getenv("ZFB_DEC_WS");
    c->dec_ws = !(e && e[0] == '0');
    const int ws_smem = 227 * 1024 - 8 * 1024;
    {
      const char* ew = std::getenv("ZFB_ENC_WARP");
      c->enc_warp = !(ew && ew[0] == '0');
      if (cudaFuncSetAttribute(ws::enc3_f32_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
          cudaFuncSetAttribute(ws::enc3_f32_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
        cudaGetLastError();
        c->enc_warp = false;
      }
    }
    if (cudaFuncSetAttribute(ws::dec3_f64_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_smem) != cudaSuccess ||
        cudaFuncSetAttribute(ws::dec3_f32_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_smem) != cudaSuccess ||
        cudaFuncSetAttribute(ws::dec3_f32_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_smem) != cudaSuccess) {
      cudaGetLastError();
      c->dec_ws = false;
    }
  }
  *out = c;
  return ZFB_OK;
}

static void free_buf(devbuf& b)
{
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

extern "C" int zfb_release_resident(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_buf(c->d_orig); free_buf(c->d_stream); free_buf(c->d_out); free_buf(c->d_hist); free_buf(c->d_pk);
  free_buf(c->var_stage); free_buf(c->var_lens); free_buf(c->var_chunks); free_buf(c->var_index);
  c->var_index_key = c->var_host_key = nullptr;
  c->h_orig_key = c->h_out_key = nullptr;
  c->orig_bytes = c->out_bytes = 0;
  c->fused_valid = false;
  c->h_stream_key = nullptr;
  c->pin_in.release(); c->pin_out.release(); c->pin_st.release();
  return ZFB_OK;
}

// The host-pointer entry points recognise "the same buffers as before" by address and size only: they cannot see a
// buffer being rewritten or freed and re-allocated at the same address (CBench does exactly that between scalars,
// main.cpp:395-398).  Callers drop the residency whenever the contents may have changed.
extern "C" int zfb_host_cache_invalidate(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  c->h_orig_key = c->h_out_key = nullptr;
  c->orig_bytes = c->out_bytes = 0;
  c->fused_valid = false;
  c->var_index_key = c->var_host_key = nullptr;
  c->h_stream_key = nullptr;
  return ZFB_OK;
}

extern "C" unsigned long long zfb_host_pair_serial(const zfb_ctx* c) { return c ? c->pair_serial : 0ull; }

extern "C" int zfb_comm_destroy(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  if (c->comm && c->comm_owned && nccl().ok) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    nccl().CommDestroy(c->comm);
  }
  c->comm = nullptr;
  c->comm_owned = false;
  c->comm_ranks = 1; c->comm_rank = 0;
  if (c->gathered) cudaFree(c->gathered);
  c->gathered = nullptr;
  return ZFB_OK;
}

extern "C" int zfb_comm_unique_id(void* out128)
{
  if (!out128) return ZFB_EINVAL;
  if (!nccl().ok) return ZFB_ESTATE;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return ZFB_ECUDA;
  memcpy(out128, &id, sizeof id);
  return ZFB_OK;
}

static int comm_attach(zfb_ctx* c, ncclComm_t comm, bool owned, int nranks, int rank)
{
  zfb_comm_destroy(c);
  if (cudaMalloc(&c->gathered, (size_t)nranks * sizeof(partials_dev)) != cudaSuccess) {
    cudaGetLastError();
    if (owned) nccl().CommDestroy(comm);
    c->err = "cudaMalloc failed";
    return ZFB_ENOMEM;
  }
  c->comm = comm; c->comm_owned = owned; c->comm_ranks = nranks; c->comm_rank = rank;
  return ZFB_OK;
}

extern "C" int zfb_comm_init(zfb_ctx* c, int nranks, int rank, const void* id128)
{
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return ZFB_EINVAL;
  if (!nccl().ok) { c->err = "NCCL unavailable: " + nccl().why; return ZFB_ESTATE; }
  CU_TRY(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = nccl().CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) { c->err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r); return ZFB_ECUDA; }
  return comm_attach(c, comm, true, nranks, rank);
}

extern "C" int zfb_comm_adopt(zfb_ctx* c, void* nccl_comm, int nranks, int rank)
{
  if (!c || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return ZFB_EINVAL;
  if (!nccl().ok) { c->err = "NCCL unavailable: " + nccl().why; return ZFB_ESTATE; }
  CU_TRY(c, cudaSetDevice(c->device));
  return comm_attach(c, (ncclComm_t)nccl_comm, false, nranks, rank);
}

// In place on the context's device partials, asynchronous on the context's stream.  Without a communicator (one
// rank) it is a no-op, so single-GPU callers need no special case.
extern "C" int zfb_allreduce_partials(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  if (!c->comm || c->comm_ranks == 1) return ZFB_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  const ncclResult_t r = nccl().AllGather(c->result, c->gathered, sizeof(partials_dev), ncclUint8, c->comm, c->stream);
  if (r != ncclSuccess) { c->err = std::string("ncclAllGather: ") + nccl().GetErrorString(r); return ZFB_ECUDA; }
  combine_partials_kernel<<<1, 32, 0, c->stream>>>(c->gathered, c->comm_ranks, c->result);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

// SUM of histogram counters (absoluteError.hpp:129, relativeError.hpp:141, minmaxMetric.hpp:122), in place
extern "C" int zfb_allreduce_counts(zfb_ctx* c, uint64_t* d_counts, size_t count)
{
  if (!c || !d_counts) return ZFB_EINVAL;
  if (!c->comm || c->comm_ranks == 1) return ZFB_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  const ncclResult_t r = nccl().AllReduce(d_counts, d_counts, count, ncclUint64, ncclSum, c->comm, c->stream);
  if (r != ncclSuccess) { c->err = std::string("ncclAllReduce: ") + nccl().GetErrorString(r); return ZFB_ECUDA; }
  return ZFB_OK;
}

extern "C" int zfb_comm_size(const zfb_ctx* c) { return c ? c->comm_ranks : 0; }

extern "C" int zfb_destroy(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  zfb_comm_destroy(c);
  zfb_release_resident(c);
  if (c->part) cudaFree(c->part);
  if (c->result) cudaFree(c->result);
  if (c->enc_result) cudaFree(c->enc_result);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c->pool;
  delete c;
  return ZFB_OK;
}

extern "C" int zfb_set_stream(zfb_ctx* c, void* s)
{
  if (!c) return ZFB_EINVAL;
  c->stream = (cudaStream_t)s;
  return ZFB_OK;
}

extern "C" int zfb_reset_stream(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  c->stream = c->own_stream;
  return ZFB_OK;
}

extern "C" int zfb_synchronize(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return ZFB_OK;
}

extern "C" const char* zfb_last_error(const zfb_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" uint64_t zfb_launch_count(const zfb_ctx* c) { return c ? c->launches : 0; }
extern "C" void* zfb_partials_dev(zfb_ctx* c) { return c ? (void*)c->result : nullptr; }

extern "C" int zfb_partials_get(zfb_ctx* c, zfb_partials* out)
{
  if (!c || !out) return ZFB_EINVAL;
  CU_TRY(c, cudaMemcpyAsync(out, c->result, sizeof(partials_dev), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return ZFB_OK;
}

extern "C" int zfb_partials_set(zfb_ctx* c, const zfb_partials* in)
{
  if (!c || !in) return ZFB_EINVAL;
  CU_TRY(c, cudaMemcpyAsync(c->result, in, sizeof(partials_dev), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return ZFB_OK;
}

// ------------------------------------------------------------------------------------------------
// dispatch helpers
// ------------------------------------------------------------------------------------------------
static int check_args(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits)
{
  if (!c) return ZFB_EINVAL;
  if (dims < 1 || dims > 3 || !nx || (dims >= 2 && !ny) || (dims >= 3 && !nz)) { c->err = "bad dims"; return ZFB_EINVAL; }
  if (dtype != ZFB_FLOAT && dtype != ZFB_DOUBLE) { c->err = "bad dtype"; return ZFB_EINVAL; }
  if (maxbits < (dtype == ZFB_FLOAT ? 9u : 12u) || maxbits > 4160u) { c->err = "bad maxbits"; return ZFB_EINVAL; }
  return ZFB_OK;
}

static geom make_geom(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits)
{
  geom g;
  g.nx = nx; g.ny = dims >= 2 ? ny : 1; g.nz = dims >= 3 ? nz : 1;
  g.bx = (g.nx + 3) / 4; g.by = (g.ny + 3) / 4; g.bz = (g.nz + 3) / 4;
  g.nblocks = g.bx * g.by * g.bz;
  g.stream_words = (g.nblocks * (size_t)maxbits + 63) / 64;
  return g;
}

static bool tile_path_ok(const zfb_ctx* c, int dtype, int dims, const geom& g, unsigned maxbits, bool metrics)
{
  if (c->force_generic || !c->encode_tiled) return false;
  if (dims != 3) return false;
  if ((g.nx & 3) || g.ny < 4 || g.nz < 4) return false;  // TMA needs 16-byte row strides; partial y/z rows go to the edge kernel
  if (maxbits & 63u) return false;
  const unsigned W = maxbits >> 6;
  if (!(W % 2 == 0 || g.bx % 2 == 0)) return false;  // bulk copies need 16-byte granularity
  if (g.nx >= (1ull << 31) || g.ny >= (1ull << 31) || g.nz >= (1ull << 31)) return false;
  const size_t bpr = (g.bx + BOX_BLOCKS - 1) / BOX_BLOCKS;
  if (bpr * g.by * g.bz >= (1ull << 31)) return false;
  if (dtype == ZFB_DOUBLE) return W <= 40;  // 32 KiB tile + 2 x 64*W*8 slot staging
  const size_t smem = metrics ? dec_tile_smem(W, true) : enc_tile_smem(W);
  if (smem > 100 * 1024 || enc_tile_smem(W) > 100 * 1024) return false;
  return true;
}

// 1D float fast path: slots that are whole 32-bit words, at most 128 bits
static bool line_path_ok(const zfb_ctx* c, int dtype, int dims, const geom& g, unsigned maxbits)
{
  if (c->force_generic) return false;
  return dtype == ZFB_FLOAT && dims == 1 && (maxbits & 31u) == 0 && maxbits <= 128u && g.nx >= 4;
}

static tile_geom make_tile_geom(const geom& g, unsigned maxbits)
{
  tile_geom t;
  t.nbx = (unsigned)g.bx; t.nby = (unsigned)(g.ny / 4); t.nbz = (unsigned)(g.nz / 4);
  t.nby_t = (unsigned)g.by;
  t.bpr = (t.nbx + BOX_BLOCKS - 1) / BOX_BLOCKS;
  t.nboxes = t.bpr * t.nby * t.nbz;
  t.ntiles = (t.nboxes + BOXES_PER_TILE - 1) / BOXES_PER_TILE;
  t.nx = (unsigned)g.nx; t.ny = (unsigned)g.ny; t.nz = (unsigned)g.nz;
  t.W = maxbits >> 6;
  t.div_wpr = make_fastdiv((t.nbx + 31) / 32 ? (t.nbx + 31) / 32 : 1);
  t.div_nby = make_fastdiv(t.nby ? t.nby : 1);
  return t;
}

static int make_tmap_3d(zfb_ctx* c, CUtensorMap* tm, const void* base, const geom& g, int dtype, unsigned box_x = 256)
{
  const cuuint64_t es = dtype == ZFB_FLOAT ? 4 : 8;
  cuuint64_t gdim[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)g.nz};
  cuuint64_t gstr[2] = {(cuuint64_t)g.nx * es, (cuuint64_t)g.nx * g.ny * es};
  cuuint32_t box[3] = {box_x, 4, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode_tiled(tm, dtype == ZFB_FLOAT ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                               const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    c->err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
    return ZFB_ECUDA;
  }
  return ZFB_OK;
}
static int make_tmap_3d_f32(zfb_ctx* c, CUtensorMap* tm, const void* base, const geom& g)
{
  return make_tmap_3d(c, tm, base, g, ZFB_FLOAT);
}

static int generic_grid(const zfb_ctx* c, size_t nblocks)
{
  size_t want = (nblocks + 127) / 128;
  size_t cap = (size_t)c->sm_count * 16;
  if (cap > (size_t)c->part_rows) cap = c->part_rows;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

static int finalize(zfb_ctx* c, int rows, size_t n)
{
  finalize_partials_kernel<<<1, 256, 0, c->stream>>>(c->part, rows, c->result, (unsigned long long)n);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

template <typename Scalar>
static int launch_enc_generic(zfb_ctx* c, int dims, const geom& g, unsigned maxbits, const void* d_in, void* d_stream,
                              int edge_only, size_t b_begin = 0)
{
  const int grid = generic_grid(c, g.nblocks - b_begin);
  const Scalar* in = (const Scalar*)d_in;
  uint64_t* st = (uint64_t*)d_stream;
  if (dims == 1) enc_generic_kernel<Scalar, 1><<<grid, 128, 0, c->stream>>>(in, st, g, maxbits, edge_only, b_begin);
  else if (dims == 2) enc_generic_kernel<Scalar, 2><<<grid, 128, 0, c->stream>>>(in, st, g, maxbits, edge_only, b_begin);
  else enc_generic_kernel<Scalar, 3><<<grid, 128, 0, c->stream>>>(in, st, g, maxbits, edge_only, b_begin);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

template <typename Scalar, bool METRICS>
static int launch_dec_generic(zfb_ctx* c, int dims, const geom& g, unsigned maxbits, const void* d_stream, void* d_out,
                              const void* d_orig, int edge_only, int* rows, size_t b_begin = 0, int row0 = 0)
{
  int grid = generic_grid(c, g.nblocks - b_begin);
  if (grid > c->part_rows - row0) grid = c->part_rows - row0;
  const uint64_t* st = (const uint64_t*)d_stream;
  Scalar* out = (Scalar*)d_out;
  const Scalar* orig = (const Scalar*)d_orig;
  double* part = c->part + (size_t)row0 * PART_STRIDE;
  if (dims == 1) dec_generic_kernel<Scalar, 1, METRICS><<<grid, 128, 0, c->stream>>>(st, out, orig, g, maxbits, edge_only, part, b_begin);
  else if (dims == 2) dec_generic_kernel<Scalar, 2, METRICS><<<grid, 128, 0, c->stream>>>(st, out, orig, g, maxbits, edge_only, part, b_begin);
  else dec_generic_kernel<Scalar, 3, METRICS><<<grid, 128, 0, c->stream>>>(st, out, orig, g, maxbits, edge_only, part, b_begin);
  c->launches++;
  *rows = grid;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

// ------------------------------------------------------------------------------------------------
// device-pointer hot path
// ------------------------------------------------------------------------------------------------
extern "C" int zfb_encode_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                              const void* d_in, void* d_stream)
{
  int rc = check_args(c, dtype, dims, nx, ny, nz, maxbits);
  if (rc) return rc;
  if (!d_in || !d_stream) { c->err = "null pointer"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const geom g = make_geom(dims, nx, ny, nz, maxbits);
  if (dtype == ZFB_DOUBLE && tile_path_ok(c, dtype, dims, g, maxbits, false) && ((uintptr_t)d_in & 15) == 0 &&
      ((uintptr_t)d_stream & 15) == 0) {
    CUtensorMap tm;
    rc = make_tmap_3d(c, &tm, d_in, g, ZFB_DOUBLE);
    if (rc) return rc;
    const tile_geom tg = make_tile_geom(g, maxbits);
    const size_t smem = DBOX_DOUBLES * sizeof(double) + (size_t)DTILE_THREADS * tg.W * 8 + 16;
    unsigned grid = (unsigned)c->sm_count * 4u;
    if (grid > tg.nboxes) grid = tg.nboxes;
    enc3_f64_tile_kernel<<<grid, DTILE_THREADS, smem, c->stream>>>(tm, (uint64_t*)d_stream, tg, maxbits);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    if ((g.ny | g.nz) & 3) return launch_enc_generic<double>(c, dims, g, maxbits, d_in, d_stream, 1);
    return ZFB_OK;
  }
  if (dtype == ZFB_FLOAT && tile_path_ok(c, dtype, dims, g, maxbits, false) && ((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_stream & 15) == 0) {
    CUtensorMap tm;
    const tile_geom tg0 = make_tile_geom(g, maxbits);
    if (c->enc_warp && tg0.W < 64 && ws::enc_warp_smem(tg0.W) <= 100 * 1024 &&
        (size_t)((tg0.nbx + 31) / 32) * tg0.nby * tg0.nbz < (1ull << 31)) {
      // warp-autonomous tiles (zfb_ws.cuh): a box of 128 x 4 x 4 values per warp
      rc = make_tmap_3d(c, &tm, d_in, g, ZFB_FLOAT, 128);
      if (rc) return rc;
      const size_t wsmem = ws::enc_warp_smem(tg0.W);
      int nb = c->occ_encw[tg0.W];
      if (!nb) {
        CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ws::enc3_f32_warp_kernel<false>, 32 * ws::EW_WARPS, wsmem));
        if (nb < 1) nb = 1;
        c->occ_encw[tg0.W] = nb;
      }
      const unsigned nwt = ((tg0.nbx + 31) / 32) * tg0.nby * tg0.nbz;
      unsigned grid = (unsigned)c->sm_count * (unsigned)nb;
      const unsigned want = (nwt + ws::EW_WARPS - 1) / ws::EW_WARPS;
      if (grid > want) grid = want;
      if (c->want_range && !((g.ny | g.nz) & 3) && (int)grid <= c->part_rows) {
        ws::enc3_f32_warp_kernel<true><<<grid, 32 * ws::EW_WARPS, wsmem, c->stream>>>(tm, (uint64_t*)d_stream, tg0, maxbits, c->part);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        finalize_partials_kernel<<<1, 256, 0, c->stream>>>(c->part, (int)grid, c->enc_result, (unsigned long long)g.nx * g.ny * g.nz);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        c->want_range = false;  // harvested
        return ZFB_OK;
      }
      ws::enc3_f32_warp_kernel<false><<<grid, 32 * ws::EW_WARPS, wsmem, c->stream>>>(tm, (uint64_t*)d_stream, tg0, maxbits, c->part);
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      if ((g.ny | g.nz) & 3) return launch_enc_generic<float>(c, dims, g, maxbits, d_in, d_stream, 1);
      return ZFB_OK;
    }
    rc = make_tmap_3d_f32(c, &tm, d_in, g);
    if (rc) return rc;
    const tile_geom tg = make_tile_geom(g, maxbits);
    const size_t smem = enc_tile_smem(tg.W);
    int nb = c->occ_enc[tg.W];
    if (!nb) {
      CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, enc3_f32_tile_kernel<false>, TILE_THREADS, smem));
      if (nb < 1) nb = 1;
      c->occ_enc[tg.W] = nb;
    }
    unsigned grid = (unsigned)c->sm_count * (unsigned)nb;
    if (grid > tg.ntiles) grid = tg.ntiles;
    if (c->want_range && !((g.ny | g.nz) & 3) && (int)grid <= c->part_rows) {
      // min / max of the field harvested by the same pass (zfb_encode_range_dev)
      enc3_f32_tile_kernel<true><<<grid, TILE_THREADS, smem, c->stream>>>(tm, (uint64_t*)d_stream, tg, maxbits, c->part);
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      finalize_partials_kernel<<<1, 256, 0, c->stream>>>(c->part, (int)grid, c->enc_result, (unsigned long long)(g.nx * g.ny * g.nz));
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      c->want_range = false;  // done: the caller does not need the stand-alone pass
      return ZFB_OK;
    }
    enc3_f32_tile_kernel<false><<<grid, TILE_THREADS, smem, c->stream>>>(tm, (uint64_t*)d_stream, tg, maxbits, c->part);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    if ((g.ny | g.nz) & 3) return launch_enc_generic<float>(c, dims, g, maxbits, d_in, d_stream, 1);  // partial blocks
    return ZFB_OK;
  }
  if (line_path_ok(c, dtype, dims, g, maxbits) && ((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_stream & 15) == 0) {
    // a trailing partial block is OR-ed into the stream by the generic kernel: everything behind the full blocks'
    // slots (the partial slot can span two 64-bit words, e.g. 96-bit slots) is cleared FIRST
    const size_t nfull = g.nx / 4;
    const unsigned nw = maxbits >> 5;
    if ((g.nx & 3) && (maxbits & 63u)) {
      const size_t written = nfull * nw * 4;
      CU_TRY(c, cudaMemsetAsync((char*)d_stream + written, 0, g.stream_words * 8 - written, c->stream));
    }
    size_t want = (nfull + 255) / 256;
#ifndef ZFB_ENC1_GRIDM
#define ZFB_ENC1_GRIDM 16  // CTAs per SM of the grid-stride launch: two waves of the 8 resident CTAs (one wave: 0.87 ms, two or four: 0.84, sixteen: 1.02)
#endif
    const int grid = (int)(want < (size_t)c->sm_count * ZFB_ENC1_GRIDM ? (want ? want : 1) : (size_t)c->sm_count * ZFB_ENC1_GRIDM);
    const float* in = (const float*)d_in;
    uint32_t* st = (uint32_t*)d_stream;
    if (nw == 1) enc1_f32_kernel<1><<<grid, 256, 0, c->stream>>>(in, st, nfull, maxbits);
    else if (nw == 2) enc1_f32_kernel<2><<<grid, 256, 0, c->stream>>>(in, st, nfull, maxbits);
    else enc1_f32_kernel<4><<<grid, 256, 0, c->stream>>>(in, st, nfull, maxbits);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    if (g.nx & 3) return launch_enc_generic<float>(c, dims, g, maxbits, d_in, d_stream, 1, g.nblocks - 1);
    return ZFB_OK;
  }
  if (maxbits & 63u) {  // unaligned slots are OR-ed into a zeroed stream
    CU_TRY(c, cudaMemsetAsync(d_stream, 0, g.stream_words * 8, c->stream));
  }
  if (dtype == ZFB_FLOAT) return launch_enc_generic<float>(c, dims, g, maxbits, d_in, d_stream, 0);
  return launch_enc_generic<double>(c, dims, g, maxbits, d_in, d_stream, 0);
}

extern "C" int zfb_decode_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                              const void* d_stream, void* d_out, const void* d_orig)
{
  int rc = check_args(c, dtype, dims, nx, ny, nz, maxbits);
  if (rc) return rc;
  if (!d_out || !d_stream) { c->err = "null pointer"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const geom g = make_geom(dims, nx, ny, nz, maxbits);
  const size_t n = g.nx * g.ny * g.nz;
  const bool metrics = d_orig != nullptr;
  int rows = 0;
  if (dtype == ZFB_DOUBLE && tile_path_ok(c, dtype, dims, g, maxbits, metrics) && ((uintptr_t)d_out & 15) == 0 &&
      ((uintptr_t)d_stream & 15) == 0 && (!metrics || ((uintptr_t)d_orig & 15) == 0)) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (metrics) {
      rc = make_tmap_3d(c, &tm, d_orig, g, ZFB_DOUBLE);
      if (rc) return rc;
    }
    const tile_geom tg = make_tile_geom(g, maxbits);
    const size_t smem = DBOX_DOUBLES * sizeof(double) + 2 * (size_t)DTILE_THREADS * tg.W * 8 + 32;
    unsigned grid = (unsigned)c->sm_count * 4u;
    if (grid > tg.nboxes) grid = tg.nboxes;
    if ((int)grid > c->part_rows / 2) grid = c->part_rows / 2;
    // float64: the metric arithmetic in double, unrolled over 64 values, makes the fused kernel 118 KiB of SASS
    // and instruction-fetch bound (5.4 ms vs 3.2 ms for the 1 GiB-block benchmark slab), so the default is the
    // lean decode kernel followed by the streaming metric kernel over (orig, out).  ZFB_F64_FUSED=1 fuses.
    static const bool unfused = [] { const char* e = std::getenv("ZFB_F64_FUSED"); return !(e && e[0] == '1'); }();
    // warp-specialised decode (zfb_ws.cuh): parser warps and transform warps, no fused metrics
    const bool ws64 = c->dec_ws && (unfused || !metrics) && ws::d::smem_bytes(tg.W) <= (size_t)(227 * 1024 - 8 * 1024) &&
                      (size_t)((tg.nbx + 31) / 32) * tg.nby * tg.nbz < (1ull << 31);
    if (ws64) {
      const unsigned nwt = ((tg.nbx + 31) / 32) * tg.nby * tg.nbz;
      unsigned wgrid = (unsigned)c->sm_count;
      const unsigned want = (nwt + ws::d::DP_WARPS - 1) / ws::d::DP_WARPS;
      if (wgrid > want) wgrid = want;
      ws::dec3_f64_ws_kernel<<<wgrid, ws::d::DTHREADS, ws::d::smem_bytes(tg.W), c->stream>>>((const uint64_t*)d_stream, (double*)d_out, tg, maxbits);
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      if ((g.ny | g.nz) & 3) {
        int er = 0;
        rc = launch_dec_generic<double, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &er, 0, 0);
        if (rc) return rc;
      }
      if (metrics) return zfb_metrics_dev(c, dtype, d_orig, d_out, n);
      return ZFB_OK;
    }
    if (metrics && unfused) {
      dec3_f64_tile_kernel<false><<<grid, DTILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (double*)d_out, tg, maxbits, c->part);
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      if ((g.ny | g.nz) & 3) {
        int er = 0;
        rc = launch_dec_generic<double, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &er, 0, 0);
        if (rc) return rc;
      }
      return zfb_metrics_dev(c, dtype, d_orig, d_out, n);
    }
    if (metrics)
      dec3_f64_tile_kernel<true><<<grid, DTILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (double*)d_out, tg, maxbits, c->part);
    else
      dec3_f64_tile_kernel<false><<<grid, DTILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (double*)d_out, tg, maxbits, c->part);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    rows = (int)grid;
    if ((g.ny | g.nz) & 3) {
      int erows = 0;
      rc = metrics ? launch_dec_generic<double, true>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, 0, rows)
                   : launch_dec_generic<double, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, 0, rows);
      if (rc) return rc;
      rows += erows;
    }
  } else if (dtype == ZFB_FLOAT && tile_path_ok(c, dtype, dims, g, maxbits, metrics) && ((uintptr_t)d_out & 15) == 0 &&
      ((uintptr_t)d_stream & 15) == 0 && (!metrics || ((uintptr_t)d_orig & 15) == 0)) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (metrics) {
      rc = make_tmap_3d_f32(c, &tm, d_orig, g);
      if (rc) return rc;
    }
    const tile_geom tg = make_tile_geom(g, maxbits);
    const char* fh0 = std::getenv("ZFB_FUSED_HIST");
    const bool want_fused_hist = metrics && c->hist_counts && fh0 && fh0[0] == '1';
    if (c->dec_ws && !want_fused_hist && ws::smem_bytes(tg.W) <= (size_t)(227 * 1024 - 8 * 1024) &&
        (size_t)((tg.nbx + 31) / 32) * tg.nby * tg.nbz < (1ull << 31)) {
      // warp-specialised kernel: parser warps and transform warps of one CTA per SM (zfb_ws.cuh)
      const unsigned nwt = ((tg.nbx + 31) / 32) * tg.nby * tg.nbz;
      unsigned grid = (unsigned)c->sm_count;
      const unsigned want = (nwt + ws::P_WARPS - 1) / ws::P_WARPS;
      if (grid > want) grid = want;
      if ((int)grid > c->part_rows / 2) grid = c->part_rows / 2;
      const size_t smem = ws::smem_bytes(tg.W);
      if (metrics)
        ws::dec3_f32_ws_kernel<true><<<grid, ws::THREADS, smem, c->stream>>>((const float*)d_orig, (const uint64_t*)d_stream, (float*)d_out, tg, maxbits, c->part);
      else
        ws::dec3_f32_ws_kernel<false><<<grid, ws::THREADS, smem, c->stream>>>((const float*)d_orig, (const uint64_t*)d_stream, (float*)d_out, tg, maxbits, c->part);
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      rows = (int)grid;
      if ((g.ny | g.nz) & 3) {  // partial blocks
        int erows = 0;
        rc = metrics ? launch_dec_generic<float, true>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, 0, rows)
                     : launch_dec_generic<float, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, 0, rows);
        if (rc) return rc;
        rows += erows;
      }
    } else {
    const size_t smem = dec_tile_smem(tg.W, metrics);
    int nb = metrics ? c->occ_decm[tg.W] : c->occ_dec[tg.W];
    if (!nb) {
      if (metrics) CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dec3_f32_tile_kernel<true, false>, TILE_THREADS, smem));
      else CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dec3_f32_tile_kernel<false, false>, TILE_THREADS, smem));
      if (nb < 1) nb = 1;
      (metrics ? c->occ_decm[tg.W] : c->occ_dec[tg.W]) = nb;
    }
    unsigned grid = (unsigned)c->sm_count * (unsigned)nb;
    if (grid > tg.ntiles) grid = tg.ntiles;
    if ((int)grid > c->part_rows / 2) grid = c->part_rows / 2;
    static const bool f32_unfused = [] { const char* e = std::getenv("ZFB_F32_UNFUSED"); return e && e[0] == '1'; }();
    if (metrics && f32_unfused && !((g.ny | g.nz) & 3)) {  // experiment switch (fused is faster for float)
      dec3_f32_tile_kernel<false, false><<<grid, TILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (float*)d_out, tg, maxbits, c->part, 0.0, 0.0, nullptr);
      c->launches++;
      CU_TRY(c, cudaGetLastError());
      return zfb_metrics_dev(c, dtype, d_orig, d_out, n);
    }
    // The in-kernel binning leaves DRAM traffic unchanged but, as measured (profiles/r02_notes.md), costs more time than
    // the stand-alone pass over the decoded field (16 ms against 4.0 + 2.9 ms on the benchmark slab: 64 shared-memory
    // updates per block next to 64 live values), so it is opt-in: ZFB_FUSED_HIST=1.
    const char* fh = std::getenv("ZFB_FUSED_HIST");
    const bool fused_hist = fh && fh[0] == '1';
    if (metrics && c->hist_counts && fused_hist && !((g.ny | g.nz) & 3)) {
      // value-range histogram of the decoded values fused into the pass (zfb_decode_hist_dev); a CTA more of shared
      // memory for the counters still leaves the same number of CTAs per SM
      dec3_f32_tile_kernel<true, true><<<grid, TILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (float*)d_out, tg, maxbits, c->part,
                                                                            c->hist_lo, c->hist_hi, (unsigned long long*)c->hist_counts);
      c->hist_counts = nullptr;  // done: the caller does not need the stand-alone pass
    } else if (metrics)
      dec3_f32_tile_kernel<true, false><<<grid, TILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (float*)d_out, tg, maxbits, c->part, 0.0, 0.0, nullptr);
    else
      dec3_f32_tile_kernel<false, false><<<grid, TILE_THREADS, smem, c->stream>>>(tm, (const uint64_t*)d_stream, (float*)d_out, tg, maxbits, c->part, 0.0, 0.0, nullptr);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    rows = (int)grid;
    if ((g.ny | g.nz) & 3) {  // partial blocks
      int erows = 0;
      rc = metrics ? launch_dec_generic<float, true>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, 0, rows)
                   : launch_dec_generic<float, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, 0, rows);
      if (rc) return rc;
      rows += erows;
    }
    }  // tile kernel (not warp-specialised)
  } else if (line_path_ok(c, dtype, dims, g, maxbits) && ((uintptr_t)d_out & 15) == 0 && ((uintptr_t)d_stream & 15) == 0 &&
             (!metrics || ((uintptr_t)d_orig & 15) == 0)) {
    const size_t nfull = g.nx / 4;
    const unsigned nw = maxbits >> 5;
    size_t want = (nfull + 255) / 256;
#ifndef ZFB_DEC1_GRIDM
#define ZFB_DEC1_GRIDM 24  // CTAs per SM of the grid-stride launch: four waves of the 6 resident CTAs (one wave: 1.29 ms, two: 1.19, four: 1.16)
#endif
    int grid = (int)(want < (size_t)c->sm_count * ZFB_DEC1_GRIDM ? (want ? want : 1) : (size_t)c->sm_count * ZFB_DEC1_GRIDM);
    if (grid > c->part_rows - 1) grid = c->part_rows - 1;
    const uint32_t* st = (const uint32_t*)d_stream;
    float* o = (float*)d_out;
    const float* og = (const float*)d_orig;
#define ZFB_DEC1(W)                                                                                              \
    do {                                                                                                         \
      if (metrics) dec1_f32_kernel<W, true><<<grid, 256, 0, c->stream>>>(st, o, og, nfull, maxbits, c->part);    \
      else dec1_f32_kernel<W, false><<<grid, 256, 0, c->stream>>>(st, o, og, nfull, maxbits, c->part);           \
    } while (0)
    if (nw == 1) ZFB_DEC1(1);
    else if (nw == 2) ZFB_DEC1(2);
    else ZFB_DEC1(4);
#undef ZFB_DEC1
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    rows = grid;
    if (g.nx & 3) {
      int erows = 0;
      rc = metrics ? launch_dec_generic<float, true>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, g.nblocks - 1, rows)
                   : launch_dec_generic<float, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 1, &erows, g.nblocks - 1, rows);
      if (rc) return rc;
      rows += erows;
    }
  } else {
    if (dtype == ZFB_FLOAT) {
      rc = metrics ? launch_dec_generic<float, true>(c, dims, g, maxbits, d_stream, d_out, d_orig, 0, &rows)
                   : launch_dec_generic<float, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 0, &rows);
    } else {
      rc = metrics ? launch_dec_generic<double, true>(c, dims, g, maxbits, d_stream, d_out, d_orig, 0, &rows)
                   : launch_dec_generic<double, false>(c, dims, g, maxbits, d_stream, d_out, d_orig, 0, &rows);
    }
    if (rc) return rc;
  }
  if (metrics) return finalize(c, rows, n);
  return ZFB_OK;
}

extern "C" int zfb_metrics_dev(zfb_ctx* c, int dtype, const void* d_orig, const void* d_approx, size_t n)
{
  if (!c || !d_orig || !d_approx || !n || (dtype != ZFB_FLOAT && dtype != ZFB_DOUBLE)) return ZFB_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  size_t want = (n / 4 + 255) / 256;
  int grid = (int)(want < (size_t)c->sm_count * 8 ? (want ? want : 1) : (size_t)c->sm_count * 8);
  if (grid > c->part_rows) grid = c->part_rows;
  if (((uintptr_t)d_orig | (uintptr_t)d_approx) & 15) { c->err = "metrics: pointers must be 16-byte aligned"; return ZFB_EINVAL; }
  if (dtype == ZFB_FLOAT) {
    metrics_kernel<float><<<grid, 256, 0, c->stream>>>((const float*)d_orig, (const float*)d_approx, n, c->part);
  } else {
    metrics_kernel<double><<<grid, 256, 0, c->stream>>>((const double*)d_orig, (const double*)d_approx, n, c->part);
  }
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return finalize(c, grid, n);
}

// ---- value-range (minmax) histogram without a second pass ------------------------------------------------
// minmaxMetric.hpp:62-133 bins the DECOMPRESSED values over [global min, global max] of the ORIGINAL field.  The
// range is known once the field has been read -- which the encode pass does anyway -- so the encode kernel harvests
// it and the decode kernel bins what it has just decoded: the histogram costs no memory traffic at all.
extern "C" int zfb_encode_range_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                                    const void* d_in, void* d_stream)
{
  if (!c) return ZFB_EINVAL;
  c->want_range = true;
  int rc = zfb_encode_dev(c, dtype, dims, nx, ny, nz, maxbits, d_in, d_stream);
  const bool fused = !c->want_range;
  c->want_range = false;
  if (rc || fused) return rc;
  // kernels without the fused harvest (anything but whole-block float 3D): one streaming pass over the field
  const size_t n = nx * (dims > 1 ? ny : 1) * (dims > 2 ? nz : 1);
  partials_dev keep;
  CU_TRY(c, cudaMemcpyAsync(&keep, c->result, sizeof keep, cudaMemcpyDeviceToHost, c->stream));
  if ((rc = zfb_metrics_dev(c, dtype, d_in, d_in, n))) return rc;
  CU_TRY(c, cudaMemcpyAsync(c->enc_result, c->result, sizeof(partials_dev), cudaMemcpyDeviceToDevice, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->result, &keep, sizeof keep, cudaMemcpyHostToDevice, c->stream));  // leave the metric partials as they were
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return ZFB_OK;
}

extern "C" int zfb_range_get(zfb_ctx* c, double* lo, double* hi)
{
  if (!c || !lo || !hi) return ZFB_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  partials_dev h;
  CU_TRY(c, cudaMemcpyAsync(&h, c->enc_result, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  *lo = -h.neg_min_orig;
  *hi = h.max_orig;
  return ZFB_OK;
}

// all ranks' ranges combined (max of the maxima, min of the minima): the same single all-gather as the partials
extern "C" int zfb_allreduce_range(zfb_ctx* c)
{
  if (!c) return ZFB_EINVAL;
  if (!c->comm || c->comm_ranks == 1) return ZFB_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  const ncclResult_t r = nccl().AllGather(c->enc_result, c->gathered, sizeof(partials_dev), ncclUint8, c->comm, c->stream);
  if (r != ncclSuccess) { c->err = std::string("ncclAllGather: ") + nccl().GetErrorString(r); return ZFB_ECUDA; }
  combine_partials_kernel<<<1, 32, 0, c->stream>>>(c->gathered, c->comm_ranks, c->enc_result);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

// zfb_decode_dev + the value-range histogram of the decoded values over [lo, hi] in the same pass.
// d_counts: ZFB_NBINS + 1 uint64 device counters (the last one counts out-of-range values), zeroed by the call.
extern "C" int zfb_decode_hist_dev(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                                   const void* d_stream, void* d_out, const void* d_orig, double lo, double hi,
                                   uint64_t* d_counts)
{
  if (!c || !d_counts || !(hi > lo)) { if (c) c->err = "null counters or empty range"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaMemsetAsync(d_counts, 0, (ZFB_NBINS + 1) * sizeof(uint64_t), c->stream));
  c->hist_counts = d_counts; c->hist_lo = lo; c->hist_hi = hi;
  int rc = zfb_decode_dev(c, dtype, dims, nx, ny, nz, maxbits, d_stream, d_out, d_orig);
  const bool fused = c->hist_counts == nullptr;
  c->hist_counts = nullptr;
  if (rc || fused) return rc;
  // kernels without the fused binning: the stand-alone pass over the decoded field
  const size_t n = nx * (dims > 1 ? ny : 1) * (dims > 2 ? nz : 1);
  return zfb_histogram_dev(c, dtype, 2, lo, hi, nullptr, d_out, n, d_counts, d_counts + ZFB_NBINS);
}

extern "C" int zfb_histogram_dev(zfb_ctx* c, int dtype, int which, double lo, double hi, const void* d_orig,
                                 const void* d_approx, size_t n, uint64_t* d_counts, uint64_t* d_oob)
{
  if (!c || !d_approx || !d_counts || which < 0 || which > 2 || (which != 2 && !d_orig)) return ZFB_EINVAL;
  if (dtype != ZFB_FLOAT && dtype != ZFB_DOUBLE) return ZFB_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaMemsetAsync(d_counts, 0, ZFB_NBINS * sizeof(uint64_t), c->stream));
  if (d_oob) CU_TRY(c, cudaMemsetAsync(d_oob, 0, sizeof(uint64_t), c->stream));
  size_t want = (n + 255) / 256;
  int grid = (int)(want < (size_t)c->sm_count * 8 ? (want ? want : 1) : (size_t)c->sm_count * 8);
  if (dtype == ZFB_FLOAT)
    hist_kernel<float><<<grid, 256, 0, c->stream>>>(which, lo, hi, (const float*)d_orig, (const float*)d_approx, n,
                                                    (unsigned long long*)d_counts, (unsigned long long*)d_oob);
  else
    hist_kernel<double><<<grid, 256, 0, c->stream>>>(which, lo, hi, (const double*)d_orig, (const double*)d_approx, n,
                                                     (unsigned long long*)d_counts, (unsigned long long*)d_oob);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

static int ensure(zfb_ctx* c, devbuf& b, size_t bytes);

// Shell-binned power spectrum of a resident field from its real-to-complex transform (zfb_spectrum.cuh)
extern "C" int zfb_power_spectrum_bin_dev(zfb_ctx* c, const void* d_fhat, size_t nx, size_t ny, size_t nz, unsigned nbins,
                                          double* d_out)
{
  if (!c || !d_fhat || !d_out || !nx || !ny || !nz || !nbins || nbins > (unsigned)PK_MAX_BINS) return ZFB_EINVAL;
  if (nx >= (1ull << 12) || ny >= (1ull << 12) || nz >= (1ull << 12)) { c->err = "power spectrum: extents above 4095 are not supported"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t nlines = ny * nz;
  unsigned grid = (unsigned)c->sm_count * 4u;
  if ((size_t)grid > nlines) grid = (unsigned)nlines;
  int rc = ensure(c, c->d_pk, (size_t)grid * 3u * nbins * sizeof(double));
  if (rc) return rc;
  pk_bin_kernel<<<grid, 256, 3u * nbins * sizeof(double), c->stream>>>((const float2*)d_fhat, (unsigned)nx, (unsigned)ny, (unsigned)nz,
                                                                      nbins, (double*)c->d_pk.p);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  pk_finalize_kernel<<<(3u * nbins + 255u) / 256u, 256, 0, c->stream>>>((const double*)c->d_pk.p, grid, nbins, d_out);
  c->launches++;
  CU_TRY(c, cudaGetLastError());
  return ZFB_OK;
}

// ------------------------------------------------------------------------------------------------
// host-pointer entry points
// ------------------------------------------------------------------------------------------------
static int ensure(zfb_ctx* c, devbuf& b, size_t bytes)
{
  if (b.cap >= bytes) return ZFB_OK;
  if (b.p) { cudaStreamSynchronize(c->stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  if (cudaMalloc(&b.p, bytes) != cudaSuccess) {
    cudaGetLastError();
    c->err = "cudaMalloc failed";
    return ZFB_ENOMEM;
  }
  b.cap = bytes;
  return ZFB_OK;
}

// Chunked, double-buffered pipeline along the slowest axis.  In fixed-rate mode with whole-word slots a
// slab whose thickness is a multiple of 4 compresses to a contiguous piece of the whole-field stream
// (SURVEY 8e), so the upload of slab i+1, the kernel on slab i and the download of slab i-1 overlap.
struct slab_plan {
  size_t nslabs, thick, slowest;      // slabs, thickness (in rows of the slowest axis), extent
  size_t row_elems;                   // elements per unit of the slowest axis
  size_t blocks_per_4rows;            // blocks per 4 rows of the slowest axis (dims > 1) or per 4 elements
};

static slab_plan plan_slabs(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, size_t esz)
{
  slab_plan p;
  p.slowest = dims == 1 ? nx : dims == 2 ? ny : nz;
  p.row_elems = dims == 1 ? 1 : dims == 2 ? nx : nx * ny;
  p.blocks_per_4rows = dims == 1 ? 1 : dims == 2 ? (nx + 3) / 4 : ((nx + 3) / 4) * ((ny + 3) / 4);
  static const size_t target_mb = [] {  // ~64 MiB of field per slab; ZFB_SLAB_MB overrides (tests use small slabs)
    const char* e = std::getenv("ZFB_SLAB_MB");
    const long v = e ? std::atol(e) : 0;
    return (size_t)(v > 0 ? v : 64);
  }();
  const size_t target = target_mb << 20;
  size_t thick = target / (p.row_elems * esz);
  // a slab must be a whole number of block rows (4) and its piece of the stream a whole number of
  // 16-byte units, so that every slab keeps the fast path's alignment: q*blocks*maxbits % 128 == 0
  size_t q = 1;
  while (((q * p.blocks_per_4rows * (size_t)maxbits) & 127u) != 0) q++;
  const size_t quantum = 4 * q;
  thick = thick / quantum * quantum;
  if (thick < quantum) thick = quantum;
  if (thick >= p.slowest) { p.thick = p.slowest; p.nslabs = 1; }
  else { p.thick = thick; p.nslabs = (p.slowest + thick - 1) / thick; }
  return p;
}

static void slab_dims(int dims, size_t nx, size_t ny, size_t nz, size_t rows, size_t& sx, size_t& sy, size_t& sz)
{
  sx = nx; sy = ny; sz = nz;
  if (dims == 1) sx = rows;
  else if (dims == 2) sy = rows;
  else sz = rows;
}

// Streams and events of one host-pointer call (RAII: every exit path releases them).
struct host_call {
  cudaStream_t up = nullptr, down = nullptr;
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_dn[2] = {nullptr, nullptr};
  bool ok = false;
  host_call()
  {
    ok = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
      ok = cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&ev_dn[i], cudaEventDisableTiming) == cudaSuccess;
  }
  ~host_call()
  {
    for (int i = 0; i < 2; i++) {
      if (ev_up[i]) cudaEventDestroy(ev_up[i]);
      if (ev_k[i]) cudaEventDestroy(ev_k[i]);
      if (ev_dn[i]) cudaEventDestroy(ev_dn[i]);
    }
    if (up) cudaStreamDestroy(up);
    if (down) cudaStreamDestroy(down);
  }
};

static copy_pool& pool_of(zfb_ctx* c)
{
  if (!c->pool) c->pool = new copy_pool(default_copy_threads());
  return *c->pool;
}

// 64 bytes from each end of a host buffer: enough to notice that a buffer at a known address holds something else now
static void sample_of(const void* p, size_t bytes, unsigned char (&out)[128])
{
  memset(out, 0, sizeof out);
  const size_t k = bytes < 64 ? bytes : 64;
  memcpy(out, p, k);
  memcpy(out + 64, (const char*)p + bytes - k, k);
}

extern "C" int zfb_compress_host(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                                 const void* h_in, void* h_stream, size_t* cbytes)
{
  int rc = check_args(c, dtype, dims, nx, ny, nz, maxbits);
  if (rc) return rc;
  if (!h_in || !h_stream) { c->err = "null pointer"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ZFB_FLOAT ? 4 : 8;
  const geom g = make_geom(dims, nx, ny, nz, maxbits);
  const size_t n = g.nx * g.ny * g.nz;
  const size_t sbytes = g.stream_words * 8;
  if ((rc = ensure(c, c->d_orig, n * esz))) return rc;
  if ((rc = ensure(c, c->d_stream, sbytes + 16))) return rc;
  c->fused_valid = false;
  c->h_orig_key = nullptr;
  c->h_stream_key = nullptr;

  const slab_plan sp = plan_slabs(dims, nx, ny, nz, maxbits, esz);
  // pageable buffers go through pinned double buffers filled / drained by the copy threads (zfb_hostio.cuh)
  const bool in_direct = is_pinned(h_in), out_direct = is_pinned(h_stream);
  size_t max_piece = 0;
  {
    size_t sx, sy, sz;
    slab_dims(dims, nx, ny, nz, sp.thick, sx, sy, sz);
    max_piece = make_geom(dims, sx, sy, sz, maxbits).stream_words * 8;
  }
  if (!in_direct && c->pin_in.ensure(sp.thick * sp.row_elems * esz)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  if (!out_direct && c->pin_st.ensure(max_piece)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  host_call hc;
  if (!hc.ok) { c->err = "cannot create streams / events"; cudaGetLastError(); return ZFB_ECUDA; }
  int status = ZFB_OK;
  size_t row0 = 0, word0 = 0;
  size_t prev_off = 0, prev_bytes = 0;  // the stream piece still in its pinned buffer
  for (size_t s = 0; s < sp.nslabs && status == ZFB_OK; s++) {
    const size_t rows = (row0 + sp.thick <= sp.slowest) ? sp.thick : sp.slowest - row0;
    size_t sx, sy, sz;
    slab_dims(dims, nx, ny, nz, rows, sx, sy, sz);
    const geom sg = make_geom(dims, sx, sy, sz, maxbits);
    const size_t off_e = row0 * sp.row_elems, slab_bytes = rows * sp.row_elems * esz;
    const char* hsrc = (const char*)h_in + off_e * esz;
    char* dsrc = (char*)c->d_orig.p + off_e * esz;
    char* dst = (char*)c->d_stream.p + word0 * 8;
    const int b = (int)(s & 1);
    if (!in_direct) {
      if (s >= 2 && cudaEventSynchronize(hc.ev_up[b]) != cudaSuccess) { status = ZFB_ECUDA; break; }  // buffer b is free again
      pool_of(c).copy(c->pin_in.p[b], hsrc, slab_bytes);
      hsrc = (const char*)c->pin_in.p[b];
    }
    if (cudaMemcpyAsync(dsrc, hsrc, slab_bytes, cudaMemcpyHostToDevice, hc.up) != cudaSuccess) { status = ZFB_ECUDA; break; }
    cudaEventRecord(hc.ev_up[b], hc.up);
    cudaStreamWaitEvent(c->stream, hc.ev_up[b], 0);
    status = zfb_encode_dev(c, dtype, dims, sx, sy, sz, maxbits, dsrc, dst);
    if (status) break;
    cudaEventRecord(hc.ev_k[b], c->stream);
    cudaStreamWaitEvent(hc.down, hc.ev_k[b], 0);
    if (out_direct) {
      if (cudaMemcpyAsync((char*)h_stream + word0 * 8, dst, sg.stream_words * 8, cudaMemcpyDeviceToHost, hc.down) != cudaSuccess) { status = ZFB_ECUDA; break; }
    } else {
      if (cudaMemcpyAsync(c->pin_st.p[b], dst, sg.stream_words * 8, cudaMemcpyDeviceToHost, hc.down) != cudaSuccess) { status = ZFB_ECUDA; break; }
      cudaEventRecord(hc.ev_dn[b], hc.down);
      if (prev_bytes) {  // the previous piece has landed by now (or will in a moment): hand it to the caller
        if (cudaEventSynchronize(hc.ev_dn[b ^ 1]) != cudaSuccess) { status = ZFB_ECUDA; break; }
        pool_of(c).copy((char*)h_stream + prev_off, c->pin_st.p[b ^ 1], prev_bytes);
      }
      prev_off = word0 * 8; prev_bytes = sg.stream_words * 8;
    }
    row0 += rows;
    word0 += sg.stream_words;
  }
  cudaError_t e1 = cudaStreamSynchronize(hc.up), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(hc.down);
  if (status == ZFB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
    c->err = std::string("compress_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3);
    status = ZFB_ECUDA;
  }
  if (status) return status;
  if (!out_direct && prev_bytes) pool_of(c).copy((char*)h_stream + prev_off, c->pin_st.p[(sp.nslabs - 1) & 1], prev_bytes);
  if (word0 != g.stream_words) { c->err = "internal: slab streams do not tile the stream"; return ZFB_ESTATE; }
  if (cbytes) *cbytes = sbytes;
  c->h_orig_key = h_in;
  c->orig_bytes = n * esz;
  // the stream stays on the device: zfb_decompress_host of this very buffer needs no upload
  c->h_stream_key = h_stream;
  c->stream_key_bytes = sbytes;
  c->stream_key_maxbits = maxbits;
  sample_of(h_stream, sbytes, c->stream_sample);
  return ZFB_OK;
}

extern "C" int zfb_decompress_host(zfb_ctx* c, int dtype, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                                   const void* h_stream, void* h_out, const void* h_orig)
{
  int rc = check_args(c, dtype, dims, nx, ny, nz, maxbits);
  if (rc) return rc;
  if (!h_out || !h_stream) { c->err = "null pointer"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ZFB_FLOAT ? 4 : 8;
  const geom g = make_geom(dims, nx, ny, nz, maxbits);
  const size_t n = g.nx * g.ny * g.nz;
  const size_t sbytes = g.stream_words * 8;
  // the stream zfb_compress_host just produced is still on the device (same buffer, same size, same ends)
  bool stream_resident = h_stream == c->h_stream_key && c->stream_key_bytes == sbytes && c->stream_key_maxbits == maxbits &&
                         c->d_stream.p && c->d_stream.cap >= sbytes;
  if (stream_resident) {
    unsigned char now[128];
    sample_of(h_stream, sbytes, now);
    stream_resident = memcmp(now, c->stream_sample, sizeof now) == 0;
  }
  if (!stream_resident) {
    c->h_stream_key = nullptr;
    if ((rc = ensure(c, c->d_stream, sbytes + 16))) return rc;
  }
  if ((rc = ensure(c, c->d_out, n * esz))) return rc;
  const bool metrics = h_orig != nullptr;
  bool orig_resident = metrics && h_orig == c->h_orig_key && c->orig_bytes == n * esz && c->d_orig.p;
  if (metrics && !orig_resident) {
    if ((rc = ensure(c, c->d_orig, n * esz))) return rc;
  }
  c->fused_valid = false;

  // The fused partials are per-launch, so slabs are combined on the host in slab order.
  const slab_plan sp = plan_slabs(dims, nx, ny, nz, maxbits, esz);
  const bool st_direct = stream_resident || is_pinned(h_stream), out_direct = is_pinned(h_out);
  const bool or_direct = !metrics || orig_resident || is_pinned(h_orig);
  size_t max_piece = 0;
  {
    size_t sx, sy, sz;
    slab_dims(dims, nx, ny, nz, sp.thick, sx, sy, sz);
    max_piece = make_geom(dims, sx, sy, sz, maxbits).stream_words * 8;
  }
  const size_t max_slab = sp.thick * sp.row_elems * esz;
  if (!st_direct && c->pin_st.ensure(max_piece)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  if (!or_direct && c->pin_in.ensure(max_slab)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  if (!out_direct && c->pin_out.ensure(max_slab)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  host_call hc;
  if (!hc.ok) { c->err = "cannot create streams / events"; cudaGetLastError(); return ZFB_ECUDA; }
  std::vector<zfb_partials> slab_parts(metrics ? sp.nslabs : 0);
  partials_dev* d_slab_parts = nullptr;
  if (metrics && cudaMalloc(&d_slab_parts, sp.nslabs * sizeof(partials_dev)) != cudaSuccess) {
    cudaGetLastError();
    c->err = "cudaMalloc failed";
    return ZFB_ENOMEM;
  }
  int status = ZFB_OK;
  size_t row0 = 0, word0 = 0;
  size_t prev_off = 0, prev_bytes = 0;  // the decoded slab still in its pinned buffer
  for (size_t s = 0; s < sp.nslabs && status == ZFB_OK; s++) {
    const size_t rows = (row0 + sp.thick <= sp.slowest) ? sp.thick : sp.slowest - row0;
    size_t sx, sy, sz;
    slab_dims(dims, nx, ny, nz, rows, sx, sy, sz);
    const geom sg = make_geom(dims, sx, sy, sz, maxbits);
    const size_t off_e = row0 * sp.row_elems;
    const size_t slab_bytes = rows * sp.row_elems * esz;
    char* dstream = (char*)c->d_stream.p + word0 * 8;
    char* dout = (char*)c->d_out.p + off_e * esz;
    char* dorig = metrics ? (char*)c->d_orig.p + off_e * esz : nullptr;
    const int b = (int)(s & 1);
    if ((!st_direct || !or_direct) && s >= 2 && cudaEventSynchronize(hc.ev_up[b]) != cudaSuccess) { status = ZFB_ECUDA; break; }
    if (!stream_resident) {
      const char* src = (const char*)h_stream + word0 * 8;
      if (!st_direct) {
        pool_of(c).copy(c->pin_st.p[b], src, sg.stream_words * 8);
        src = (const char*)c->pin_st.p[b];
      }
      if (cudaMemcpyAsync(dstream, src, sg.stream_words * 8, cudaMemcpyHostToDevice, hc.up) != cudaSuccess) { status = ZFB_ECUDA; break; }
    }
    if (metrics && !orig_resident) {
      const char* src = (const char*)h_orig + off_e * esz;
      if (!or_direct) {
        pool_of(c).copy(c->pin_in.p[b], src, slab_bytes);
        src = (const char*)c->pin_in.p[b];
      }
      if (cudaMemcpyAsync(dorig, src, slab_bytes, cudaMemcpyHostToDevice, hc.up) != cudaSuccess) { status = ZFB_ECUDA; break; }
    }
    cudaEventRecord(hc.ev_up[b], hc.up);
    cudaStreamWaitEvent(c->stream, hc.ev_up[b], 0);
    status = zfb_decode_dev(c, dtype, dims, sx, sy, sz, maxbits, dstream, dout, dorig);
    if (status) break;
    if (metrics) cudaMemcpyAsync(d_slab_parts + s, c->result, sizeof(partials_dev), cudaMemcpyDeviceToDevice, c->stream);
    cudaEventRecord(hc.ev_k[b], c->stream);
    cudaStreamWaitEvent(hc.down, hc.ev_k[b], 0);
    if (out_direct) {
      if (cudaMemcpyAsync((char*)h_out + off_e * esz, dout, slab_bytes, cudaMemcpyDeviceToHost, hc.down) != cudaSuccess) { status = ZFB_ECUDA; break; }
    } else {
      if (cudaMemcpyAsync(c->pin_out.p[b], dout, slab_bytes, cudaMemcpyDeviceToHost, hc.down) != cudaSuccess) { status = ZFB_ECUDA; break; }
      cudaEventRecord(hc.ev_dn[b], hc.down);
      if (prev_bytes) {  // hand the previous slab to the caller while this one is in flight
        if (cudaEventSynchronize(hc.ev_dn[b ^ 1]) != cudaSuccess) { status = ZFB_ECUDA; break; }
        pool_of(c).copy((char*)h_out + prev_off, c->pin_out.p[b ^ 1], prev_bytes);
      }
      prev_off = off_e * esz; prev_bytes = slab_bytes;
    }
    row0 += rows;
    word0 += sg.stream_words;
  }
  cudaError_t e1 = cudaStreamSynchronize(hc.up), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(hc.down);
  if (status == ZFB_OK && metrics) {
    if (cudaMemcpy(slab_parts.data(), d_slab_parts, sp.nslabs * sizeof(partials_dev), cudaMemcpyDeviceToHost) != cudaSuccess) status = ZFB_ECUDA;
  }
  if (d_slab_parts) cudaFree(d_slab_parts);
  if (status == ZFB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
    c->err = std::string("decompress_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3);
    status = ZFB_ECUDA;
  }
  if (status) return status;
  if (!out_direct && prev_bytes) pool_of(c).copy((char*)h_out + prev_off, c->pin_out.p[(sp.nslabs - 1) & 1], prev_bytes);
  if (metrics) {
    zfb_partials total;
    memset(&total, 0, sizeof total);
    total.max_orig = -1.7976931348623157e308;
    total.neg_min_orig = -1.7976931348623157e308;
    zfb_partials_combine(&total, slab_parts.data(), (int)sp.nslabs);
    if ((rc = zfb_partials_set(c, &total))) return rc;
    c->h_orig_key = h_orig;
    c->orig_bytes = n * esz;
    c->fused_valid = true;
  }
  c->h_out_key = h_out;
  c->out_bytes = n * esz;
  c->pair_serial++;
  return ZFB_OK;
}

extern "C" int zfb_metrics_host(zfb_ctx* c, int dtype, const void* h_orig, const void* h_approx, size_t n,
                                zfb_partials* out)
{
  if (!c || !h_orig || !h_approx || !n || !out) return ZFB_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ZFB_FLOAT ? 4 : 8;
  if (c->fused_valid && h_orig == c->h_orig_key && h_approx == c->h_out_key && n * esz == c->out_bytes)
    return zfb_partials_get(c, out);
  int rc;
  const bool orig_res = h_orig == c->h_orig_key && c->orig_bytes == n * esz && c->d_orig.p;
  const bool out_res = h_approx == c->h_out_key && c->out_bytes == n * esz && c->d_out.p;
  if (!orig_res) {
    if ((rc = ensure(c, c->d_orig, n * esz))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_orig.p, h_orig, n * esz, cudaMemcpyHostToDevice, c->stream));
    c->h_orig_key = h_orig; c->orig_bytes = n * esz;
  }
  if (!out_res) {
    if ((rc = ensure(c, c->d_out, n * esz))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_out.p, h_approx, n * esz, cudaMemcpyHostToDevice, c->stream));
    c->h_out_key = h_approx; c->out_bytes = n * esz;
  }
  if ((rc = zfb_metrics_dev(c, dtype, c->d_orig.p, c->d_out.p, n))) return rc;
  c->fused_valid = true;
  return zfb_partials_get(c, out);
}

extern "C" int zfb_histogram_host(zfb_ctx* c, int dtype, int which, double lo, double hi, const void* h_orig,
                                  const void* h_approx, size_t n, uint64_t* counts, uint64_t* oob)
{
  if (!c || !h_approx || !counts || !n) return ZFB_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ZFB_FLOAT ? 4 : 8;
  int rc;
  const bool orig_res = h_orig && h_orig == c->h_orig_key && c->orig_bytes == n * esz && c->d_orig.p;
  const bool out_res = h_approx == c->h_out_key && c->out_bytes == n * esz && c->d_out.p;
  if (h_orig && !orig_res) {
    if ((rc = ensure(c, c->d_orig, n * esz))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_orig.p, h_orig, n * esz, cudaMemcpyHostToDevice, c->stream));
    c->h_orig_key = h_orig; c->orig_bytes = n * esz; c->fused_valid = false;
  }
  if (!out_res) {
    if ((rc = ensure(c, c->d_out, n * esz))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_out.p, h_approx, n * esz, cudaMemcpyHostToDevice, c->stream));
    c->h_out_key = h_approx; c->out_bytes = n * esz; c->fused_valid = false;
  }
  if ((rc = ensure(c, c->d_hist, (ZFB_NBINS + 1) * sizeof(uint64_t)))) return rc;
  uint64_t* dh = (uint64_t*)c->d_hist.p;
  if ((rc = zfb_histogram_dev(c, dtype, which, lo, hi, h_orig ? c->d_orig.p : nullptr, c->d_out.p, n, dh, dh + ZFB_NBINS))) return rc;
  std::vector<uint64_t> tmp(ZFB_NBINS + 1);
  CU_TRY(c, cudaMemcpyAsync(tmp.data(), dh, (ZFB_NBINS + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  memcpy(counts, tmp.data(), ZFB_NBINS * sizeof(uint64_t));
  if (oob) *oob = tmp[ZFB_NBINS];
  return ZFB_OK;
}

#include "zfb_api_var.cuh"
#include "zfb_api_io.cuh"
