// zfb_kernels.cuh -- sm_100a kernels for the CBench hot path:
//   compress   = ZFPCompressorGpu::compress   -> zfp_compress    (zfpCompressorGpu.hpp:145)
//   decompress = ZFPCompressorGpu::decompress -> zfp_decompress  (zfpCompressorGpu.hpp:249)
//   metrics    = the five MetricInterface::execute passes of main.cpp:306-352
//                (absoluteError.hpp:65-143, relativeError.hpp:66-155, meansquareError.hpp:55-78,
//                 psnrError.hpp:56-96, minmaxMetric.hpp:60-135), fused into the decode pass.
//
// Work decomposition (B200-first, not cuZFP's): one 4^d block per thread with all 4^d values in
// registers (zfb_codec.cuh).  What differs between kernels is only how bytes move:
//   * enc3/dec3 "tile" kernels (3D, dims % 4 == 0): persistent CTAs; a CTA owns two boxes of
//     256x4x4 values (= 64 x-consecutive blocks each); boxes arrive in shared memory through TMA
//     (cp.async.bulk.tensor.3d + mbarrier) one tile ahead of the math, so each thread's 16 row loads
//     are conflict-free 128-bit LDS; compressed slots of a tile are contiguous in the stream and move
//     with one bulk copy (cp.async.bulk) in each direction; decoded rows leave as coalesced 128-bit
//     stores (a warp writes 512 contiguous bytes per row).
//   * enc3/dec3 f64 tile kernels: the same design with 64 threads and one 32 KiB box per CTA (a block of
//     doubles needs ~128 registers for its values alone); metrics run as the streaming pass by default.
//   * enc1/dec1 kernels (1D float, HACC): one 4-value block per thread, one 128-bit load / store per block.
//   * generic kernels (2D, partial blocks, unaligned slots, anything else): same codec, direct global access.
//   * variable-rate kernels (fixed accuracy / precision): zfb_var.cuh.
// Metric partials: per-thread -> warp shuffle -> CTA -> one row of `part` per CTA -> fixed-order
// final reduction (deterministic for a given grid).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "zfb_codec.cuh"
#include "zfb_codec_f3.cuh"
#include "zfb_codec_f1.cuh"
#include "zfb_codec_d3.cuh"

namespace zfb {

// ---------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a): mbarrier, TMA tensor / bulk copies, proxy fences
// ---------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 3D tiled TMA load: box at element coordinates (x, y, z) -> shared memory, completion on mbarrier
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int x, int y, int z, uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
// the same box requested into L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tmap, int x, int y, int z)
{
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(x), "r"(y), "r"(z) : "memory");
}
// 1D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 1D bulk copy shared -> global
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy writes to shared memory visible to the async proxy (TMA / bulk store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* t)
{
  asm volatile("prefetch.tensormap [%0];" ::"l"(t) : "memory");
}
}  // namespace ptx

// ---------------------------------------------------------------------------------------------
// geometry shared by host and device
// ---------------------------------------------------------------------------------------------
struct geom {
  size_t nx, ny, nz;     // extents, x fastest; unused = 1
  size_t bx, by, bz;     // blocks per axis
  size_t nblocks;
  size_t stream_words;   // 64-bit words in the stream
};

constexpr int PART_STRIDE = 8;  // doubles per CTA row in the partials scratch

// float/3D coder tables in shared memory (zfb_codec_f3.cuh); every thread of the CTA must call build_coder_tables.
// An encoding kernel carries only ENC + NEE (2.3 KiB), a decoding kernel only DEC + NE (4.5 KiB).
template <bool ENCODER> struct coder_smem_t {
  uint32_t tab[ENCODER ? f3::ENC_WORDS : f3::DEC_WORDS];
};
typedef coder_smem_t<true> enc_coder_smem;
typedef coder_smem_t<false> dec_coder_smem;
template <bool ENCODER> __device__ __forceinline__ f3::tables build_coder_tables(coder_smem_t<ENCODER>& cs)
{
  for (unsigned i = threadIdx.x; i < (ENCODER ? f3::ENC_WORDS : f3::DEC_WORDS); i += blockDim.x)
    cs.tab[i] = ENCODER ? f3::enc_table_word(i) : f3::dec_table_word(i);
  __syncthreads();
  f3::tables t;
  t.enc = cs.tab; t.dec = cs.tab;
  t.enc_s = (uint32_t)__cvta_generic_to_shared(cs.tab);
  t.dec_s = t.enc_s;
  // opaque to the compiler: otherwise it rebuilds the shared window address (S2UR SR_CgaCtaId, ULEA, ...) in front
  // of every table load inside the coder loops instead of keeping it in a register
  asm volatile("mov.u32 %0, %0;" : "+r"(t.enc_s));
  asm volatile("mov.u32 %0, %0;" : "+r"(t.dec_s));
  return t;
}

// plane scratch of the float/3D coder for kernels without a shared-memory tile (generic path):
// 16 rows x blockDim (128) x 16 B; only the float/3D instantiations carry it
template <bool FAST> struct fast_scratch {
  __device__ __forceinline__ f3::plane_rows rows() { f3::plane_rows r; r.p = nullptr; r.stride = 0; return r; }
};
template <> struct fast_scratch<true> {
  f3::u32x4 buf[16 * 128];
  __device__ __forceinline__ f3::plane_rows rows() { f3::plane_rows r; r.p = buf + threadIdx.x; r.stride = 128; return r; }
};

// ---------------------------------------------------------------------------------------------
// metric accumulation (one thread)
// ---------------------------------------------------------------------------------------------
template <typename MaxT> struct macc_t {
  double s_sq, s_abs, s_rel;
  MaxT m_abs, m_rel, m_o, m_no;  // maxima (float data keeps them in float: exact, and 4 registers fewer); m_no = max(-o)
  __device__ __forceinline__ void init()
  {
    s_sq = s_abs = s_rel = 0.0;
    m_abs = 0;  // the reference's error vectors start with n zeros (absoluteError.hpp:66)
    m_rel = 0;
    m_o = sizeof(MaxT) == 4 ? (MaxT)-3.402823466e38f : (MaxT)-1.7976931348623157e308;
    m_no = m_o;
  }
};
typedef macc_t<double> macc;    // double fields, final reduction
typedef macc_t<float> maccf;    // float fields

// float: accumulates one row of 4 values into float partial sums (flushed to double by the caller)
struct frow {
  float ssq, sab, srl, mab, mrl, mo, mno;
  __device__ __forceinline__ void init()
  {
    ssq = sab = srl = 0.f;
    mab = mrl = 0.f;
    mo = -3.402823466e38f;
    mno = -3.402823466e38f;
  }
  __device__ __forceinline__ void add(float o, float a)
  {
    const float d = o - a;            // float subtraction, as in the reference (static_cast<float*>)
    const float ad = fabsf(d);
    const float ao = fabsf(o);
    // relError(o, a, 1): |o| < 1 -> absolute error, else |d|/|o| (relativeError.hpp:66-75)
    // div.approx.ftz (MUFU.RCP + FMUL, <= 2 ulp, no denormal fix-up code) returns 0 for divisors above
    // 2^126, so both operands are scaled by 1/4 first (exact); q is only used when |o| >= 1, and the
    // metric tolerance is 1e-6 relative
    float q;
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(ad * 0.25f), "f"(ao * 0.25f));
    const float rl = ao < 1.0f ? ad : q;
    ssq = fmaf(d, d, ssq);
    sab += ad;
    srl += rl;
    mab = fmaxf(mab, ad);
    mrl = fmaxf(mrl, rl);
    mo = fmaxf(mo, o);
    mno = fmaxf(mno, -o);
  }
  // one row of four values; nested fmaxf lets the compiler use the 3-input FMNMX3
  __device__ __forceinline__ void add4(const float4& o, float a0, float a1, float a2, float a3)
  {
    const float d0 = o.x - a0, d1 = o.y - a1, d2 = o.z - a2, d3 = o.w - a3;
    const float e0 = fabsf(d0), e1 = fabsf(d1), e2 = fabsf(d2), e3 = fabsf(d3);
    const float b0 = fabsf(o.x), b1 = fabsf(o.y), b2 = fabsf(o.z), b3 = fabsf(o.w);
    // relError(o, a, 1) = |d| / max(|o|, 1): one reciprocal estimate (MUFU.RCP, <= 2 ulp; the metric tolerance is 1e-6
    // relative).  Divisor and dividend are scaled by 1/4 (exact) because rcp.approx.ftz returns 0 above 2^126.
    float c0, c1, c2, c3;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(fmaxf(b0, 1.0f) * 0.25f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(fmaxf(b1, 1.0f) * 0.25f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c2) : "f"(fmaxf(b2, 1.0f) * 0.25f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c3) : "f"(fmaxf(b3, 1.0f) * 0.25f));
    const float r0 = (e0 * 0.25f) * c0, r1 = (e1 * 0.25f) * c1, r2 = (e2 * 0.25f) * c2, r3 = (e3 * 0.25f) * c3;
    ssq = fmaf(d0, d0, ssq); ssq = fmaf(d1, d1, ssq); ssq = fmaf(d2, d2, ssq); ssq = fmaf(d3, d3, ssq);
    sab += (e0 + e1) + (e2 + e3);
    srl += (r0 + r1) + (r2 + r3);
    mab = fmaxf(fmaxf(mab, e0), e1); mab = fmaxf(fmaxf(mab, e2), e3);
    mrl = fmaxf(fmaxf(mrl, r0), r1); mrl = fmaxf(fmaxf(mrl, r2), r3);
    mo = fmaxf(fmaxf(mo, o.x), o.y); mo = fmaxf(fmaxf(mo, o.z), o.w);
    mno = fmaxf(fmaxf(mno, -o.x), -o.y); mno = fmaxf(fmaxf(mno, -o.z), -o.w);
  }
  // sums go to double; the maxima simply keep running in float (no reset needed)
  __device__ __forceinline__ void flush(maccf& m)
  {
    m.s_sq += (double)ssq; m.s_abs += (double)sab; m.s_rel += (double)srl;
    m.m_abs = fmaxf(m.m_abs, mab); m.m_rel = fmaxf(m.m_rel, mrl); m.m_o = fmaxf(m.m_o, mo); m.m_no = fmaxf(m.m_no, mno);
    ssq = sab = srl = 0.f;
  }
  __device__ __forceinline__ void flush(macc& m)
  {
    m.s_sq += (double)ssq; m.s_abs += (double)sab; m.s_rel += (double)srl;
    m.m_abs = fmax(m.m_abs, (double)mab); m.m_rel = fmax(m.m_rel, (double)mrl);
    m.m_o = fmax(m.m_o, (double)mo); m.m_no = fmax(m.m_no, (double)mno);
    init();
  }
};

// |d| / |o| for |o| >= 1 to ~1e-7 relative: both operands are scaled by the exact power of two that brings |o|
// into [1, 2), then divided in float (MUFU.RCP).  An IEEE double division costs ~30 instructions per value and,
// unrolled 64 times, pushed the double kernels out of the instruction cache.
__device__ __forceinline__ double rel_quotient(double ad, double ao)
{
  int e = ((__double2hiint(ao) >> 20) & 0x7ff) - 1023;  // ao >= 1 -> e >= 0
  e = e > 1022 ? 1022 : e;
  const double sc = __hiloint2double((1023 - e) << 20, 0);  // 2^-e
  const float aof = (float)(ao * sc), adf = (float)(ad * sc);
  float q;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(adf), "f"(aof));
  return (double)q;
}

__device__ __forceinline__ void macc_add(macc& m, double o, double a)
{
  // double fields: same formulas carried out in double (the reference reinterprets doubles as floats,
  // which is meaningless).  Maxima by compare-select (inputs are finite; fmax's NaN handling
  // costs ~14 instructions each in double).
  const double d = o - a, ad = fabs(d), ao = fabs(o);
  const double rl = ao < 1.0 ? ad : rel_quotient(ad, ao);
  m.s_sq = fma(d, d, m.s_sq); m.s_abs += ad; m.s_rel += rl;
  m.m_abs = ad > m.m_abs ? ad : m.m_abs; m.m_rel = rl > m.m_rel ? rl : m.m_rel;
  m.m_o = o > m.m_o ? o : m.m_o; m.m_no = -o > m.m_no ? -o : m.m_no;
}

__device__ __forceinline__ double shfl_xor_d(double v, int lane)
{
  return __shfl_xor_sync(0xffffffffu, v, lane);
}

// CTA-wide reduction of the per-thread accumulators; thread 0 writes one row of `part`.
// red: shared scratch of 7 * (blockDim.x / 32) doubles.
template <typename MaxT>
__device__ __forceinline__ void macc_block_store(const macc_t<MaxT>& mi, double* red, double* part_row)
{
  macc m;
  m.s_sq = mi.s_sq; m.s_abs = mi.s_abs; m.s_rel = mi.s_rel;
  m.m_abs = (double)mi.m_abs; m.m_rel = (double)mi.m_rel; m.m_o = (double)mi.m_o; m.m_no = (double)mi.m_no;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    m.s_sq += shfl_xor_d(m.s_sq, o);
    m.s_abs += shfl_xor_d(m.s_abs, o);
    m.s_rel += shfl_xor_d(m.s_rel, o);
    m.m_abs = fmax(m.m_abs, shfl_xor_d(m.m_abs, o));
    m.m_rel = fmax(m.m_rel, shfl_xor_d(m.m_rel, o));
    m.m_o = fmax(m.m_o, shfl_xor_d(m.m_o, o));
    m.m_no = fmax(m.m_no, shfl_xor_d(m.m_no, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    double* r = red + 7 * warp;
    r[0] = m.s_sq; r[1] = m.s_abs; r[2] = m.s_rel; r[3] = m.m_abs; r[4] = m.m_rel; r[5] = m.m_o; r[6] = m.m_no;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = -1.7976931348623157e308, a6 = -1.7976931348623157e308;
    for (int w = 0; w < nw; w++) {  // fixed order
      const double* r = red + 7 * w;
      a0 += r[0]; a1 += r[1]; a2 += r[2];
      a3 = fmax(a3, r[3]); a4 = fmax(a4, r[4]); a5 = fmax(a5, r[5]); a6 = fmax(a6, r[6]);
    }
    part_row[0] = a0; part_row[1] = a1; part_row[2] = a2; part_row[3] = a3; part_row[4] = a4; part_row[5] = a5;
    part_row[6] = a6; part_row[7] = 0.0;
  }
}

// Value-range (minmax) histogram of decoded float values, fused into the decode kernels: bin of v over [lo, lo + range]
// exactly as minmaxMetric.hpp:112 computes it -- int(((range * ((v - lo) / range)) / (range / 1024)), `pos - 1` when
// pos >= 1024 -- with out-of-range positions clamped and counted.  The quotient is first estimated in float (error
// below 2e-4 bins: v - lo, the scale and the product each round once, 3 * 2^-24 * 1025); only when the estimate lies
// within 1e-3 of an integer, where that error could change the truncation, is the reference's double expression
// evaluated (0.2 % of the values: a wider margin makes nearly every warp take the slow path for some lane).
struct value_binner {
  double lo, range, bin;
  float lof, scale;
  bool exact_lo;  // lo is a float value (always, for float fields): the float estimate starts from the same number
  __device__ __forceinline__ void init(double lo_, double hi_)
  {
    lo = lo_; range = hi_ - lo_; bin = range / 1024.0;
    lof = (float)lo_;
    scale = (float)(1024.0 / range);
    exact_lo = (double)lof == lo_;
  }
  // the reference's expression in double, range checks included; out of line: inlined 64 times (one per value of a
  // block) its two divisions alone are 100 KB of code and the kernel stalls on instruction fetch
  __device__ __noinline__ int bin_exact(float v, unsigned int* bad) const
  {
    const double q = (range * (((double)v - lo) / range)) / bin;
    int pos;
    if (!(q >= 0.0)) { pos = 0; atomicAdd(bad, 1u); }
    else if (q >= 1025.0) { pos = 1023; atomicAdd(bad, 1u); }
    else {
      pos = (int)q;
      if (pos >= 1024) pos = pos - 1;
      if (pos >= 1024) { pos = 1023; atomicAdd(bad, 1u); }
    }
    return pos;
  }
  __device__ __forceinline__ int bin_of(float v, unsigned int* bad) const
  {
    const float qf = (v - lof) * scale;
    // fast path, float only: strictly inside the range and not within 1e-3 of an integer
    if (fabsf(qf - rintf(qf)) > 1e-3f && qf > 0.5f && qf < 1023.5f && exact_lo) return (int)qf;
    return bin_exact(v, bad);
  }
};

struct partials_dev {  // mirrors zfb_partials (include/zfb.h)
  double sum_sq, sum_abs, sum_rel, max_abs, max_rel, max_orig, neg_min_orig;
  unsigned long long n;
};

// fixed-order reduction of the per-CTA rows into the context's partials
__global__ void __launch_bounds__(256) finalize_partials_kernel(const double* __restrict__ part, int nrows,
                                                                partials_dev* __restrict__ out, unsigned long long n)
{
  __shared__ double red[7 * 8];
  macc m;
  m.init();
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
    const double* p = part + (size_t)r * PART_STRIDE;
    m.s_sq += p[0]; m.s_abs += p[1]; m.s_rel += p[2];
    m.m_abs = fmax(m.m_abs, p[3]); m.m_rel = fmax(m.m_rel, p[4]); m.m_o = fmax(m.m_o, p[5]); m.m_no = fmax(m.m_no, p[6]);
  }
  __shared__ double row[PART_STRIDE];
  macc_block_store(m, red, row);
  if (threadIdx.x == 0) {
    out->sum_sq = row[0]; out->sum_abs = row[1]; out->sum_rel = row[2]; out->max_abs = row[3]; out->max_rel = row[4];
    out->max_orig = row[5]; out->neg_min_orig = row[6]; out->n = n;
  }
}

// ---------------------------------------------------------------------------------------------
// generic kernels: any dims / dtype / maxbits, partial blocks, optional "edge only" mode
// ---------------------------------------------------------------------------------------------
template <typename Scalar> struct vec4;
template <> struct vec4<float> { typedef float4 type; };
template <> struct vec4<double> { typedef double4 type; };

template <typename Scalar, int DIMS>
__global__ void __launch_bounds__(128) enc_generic_kernel(const Scalar* __restrict__ in, uint64_t* __restrict__ stream,
                                                          geom g, unsigned maxbits, int edge_only, size_t b_begin)
{
  constexpr int BS = 1 << (2 * DIMS);
  constexpr bool FAST = sizeof(Scalar) == 4 && DIMS == 3;
  __shared__ enc_coder_smem cs;
  __shared__ fast_scratch<FAST> fsc;
  f3::tables tb;
  if (FAST) tb = build_coder_tables(cs);
  const f3::plane_rows ps = fsc.rows();
  const bool aligned = (maxbits & 63u) == 0;
  const bool vec = (reinterpret_cast<size_t>(in) & 15) == 0;  // 1D: 16-byte loads only from a 16-byte aligned array
  for (size_t b = b_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < g.nblocks; b += (size_t)gridDim.x * blockDim.x) {
    const size_t ix = b % g.bx, iy = (b / g.bx) % g.by, iz = b / (g.bx * g.by);
    const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
    const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
    const unsigned my = DIMS >= 2 ? (unsigned)(g.ny - y0 < 4 ? g.ny - y0 : 4) : 1u;
    const unsigned mz = DIMS >= 3 ? (unsigned)(g.nz - z0 < 4 ? g.nz - z0 : 4) : 1u;
    const bool full = mx == 4 && (DIMS < 2 || my == 4) && (DIMS < 3 || mz == 4);
    if (edge_only && full) continue;
    Scalar f[BS];
    if (DIMS == 1 && full && vec) {
      if constexpr (sizeof(Scalar) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + x0));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
      } else {
        const double2 v0 = __ldg(reinterpret_cast<const double2*>(in + x0));
        const double2 v1 = __ldg(reinterpret_cast<const double2*>(in + x0 + 2));
        f[0] = v0.x; f[1] = v0.y; f[2] = v1.x; f[3] = v1.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < (DIMS >= 3 ? 4 : 1); k++)
#pragma unroll
        for (int j = 0; j < (DIMS >= 2 ? 4 : 1); j++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const bool ok = (unsigned)i < mx && (unsigned)j < my && (unsigned)k < mz;
            f[(16 * k + 4 * j + i) & (BS - 1)] = ok ? __ldg(in + (x0 + i) + g.nx * ((y0 + j) + g.ny * (z0 + k))) : (Scalar)0;
          }
      if (!full) pad_block<Scalar, DIMS>(f, mx, my, mz);
    }
    if constexpr (FAST) {
      uint32_t* s32 = reinterpret_cast<uint32_t*>(stream);
      if (aligned) {
        f3::sink32 sink(s32 + b * (size_t)(maxbits >> 5), maxbits >> 5);
        f3::encode_block(f, maxbits, sink, tb, ps);
      } else {
        f3::or_sink32 sink(s32, b * (size_t)maxbits, maxbits);
        f3::encode_block(f, maxbits, sink, tb, ps);
      }
    } else {
      if (aligned) {
        aligned_sink sink(stream + b * (size_t)(maxbits >> 6), maxbits >> 6);
        encode_block<Scalar, DIMS>(f, maxbits, sink);
      } else {
        or_sink sink(stream, b * (size_t)maxbits, maxbits);
        encode_block<Scalar, DIMS>(f, maxbits, sink);
      }
    }
  }
}

template <typename Scalar, int DIMS, bool METRICS>
__global__ void __launch_bounds__(128) dec_generic_kernel(const uint64_t* __restrict__ stream, Scalar* __restrict__ out,
                                                          const Scalar* __restrict__ orig, geom g, unsigned maxbits,
                                                          int edge_only, double* __restrict__ part, size_t b_begin)
{
  constexpr int BS = 1 << (2 * DIMS);
  constexpr bool FAST = sizeof(Scalar) == 4 && DIMS == 3;
  __shared__ double red[7 * 4];
  __shared__ dec_coder_smem cs;
  __shared__ fast_scratch<FAST> fsc;
  f3::tables tb;
  if (FAST) tb = build_coder_tables(cs);
  const f3::plane_rows ps = fsc.rows();
  macc_t<Scalar> m;
  m.init();
  // 1D: 16-byte accesses only when both arrays are 16-byte aligned (a slice like field[1:] is not)
  const bool vec = (reinterpret_cast<size_t>(out) & 15) == 0 && (!METRICS || (reinterpret_cast<size_t>(orig) & 15) == 0);
  for (size_t b = b_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < g.nblocks; b += (size_t)gridDim.x * blockDim.x) {
    const size_t ix = b % g.bx, iy = (b / g.bx) % g.by, iz = b / (g.bx * g.by);
    const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
    const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
    const unsigned my = DIMS >= 2 ? (unsigned)(g.ny - y0 < 4 ? g.ny - y0 : 4) : 1u;
    const unsigned mz = DIMS >= 3 ? (unsigned)(g.nz - z0 < 4 ? g.nz - z0 : 4) : 1u;
    const bool full = mx == 4 && (DIMS < 2 || my == 4) && (DIMS < 3 || mz == 4);
    if (edge_only && full) continue;
    const size_t startbit = b * (size_t)maxbits;
    const size_t w0 = startbit >> 6;
    const size_t left = g.stream_words - w0;
    Scalar f[BS];
    if constexpr (FAST) {
      const size_t v0 = startbit >> 5;
      const size_t left32 = 2 * g.stream_words - v0;
      f3::reader32 r(reinterpret_cast<const uint32_t*>(stream) + v0, left32 > 0xffffffffull ? 0xffffffffu : (unsigned)left32,
                     (unsigned)(startbit & 31));
      f3::decode_block(f, maxbits, r, tb, ps);
    } else {
      slot_reader r(stream + w0, left > 0xffffffffull ? 0xffffffffu : (unsigned)left, (unsigned)(startbit & 63));
      decode_block<Scalar, DIMS>(f, maxbits, r);
    }
    if (DIMS == 1 && full && sizeof(Scalar) == 4 && vec) {
      if constexpr (sizeof(Scalar) == 4) {
        float4 v;
        v.x = f[0]; v.y = f[1]; v.z = f[2]; v.w = f[3];
        if (METRICS) {
          const float4 o = __ldg(reinterpret_cast<const float4*>(orig + x0));
          frow fr;
          fr.init();
          fr.add(o.x, v.x); fr.add(o.y, v.y); fr.add(o.z, v.z); fr.add(o.w, v.w);
          fr.flush(m);
        }
        *reinterpret_cast<float4*>(out + x0) = v;
      }
    } else {
#pragma unroll
      for (int k = 0; k < (DIMS >= 3 ? 4 : 1); k++)
#pragma unroll
        for (int j = 0; j < (DIMS >= 2 ? 4 : 1); j++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const bool ok = (unsigned)i < mx && (unsigned)j < my && (unsigned)k < mz;
            if (ok) {
              const size_t idx = (x0 + i) + g.nx * ((y0 + j) + g.ny * (z0 + k));
              const Scalar a = f[(16 * k + 4 * j + i) & (BS - 1)];
              out[idx] = a;
              if (METRICS) {
                if constexpr (sizeof(Scalar) == 4) {
                  frow fr;
                  fr.init();
                  fr.add(__ldg(orig + idx), a);
                  fr.flush(m);
                } else {
                  macc_add(m, (double)__ldg(orig + idx), (double)a);
                }
              }
            }
          }
    }
  }
  if (METRICS) macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
}

// ---------------------------------------------------------------------------------------------
// 3D float tile kernels (dims % 4 == 0; slots whole words; see file header)
// ---------------------------------------------------------------------------------------------
constexpr int TILE_THREADS = 128;              // 4 warps, one 4x4x4 block per thread
constexpr int BOX_BLOCKS = 64;                 // blocks per TMA box (256 x-values)
constexpr int BOX_FLOATS = 256 * 4 * 4;        // 16 KiB
constexpr int BOXES_PER_TILE = TILE_THREADS / BOX_BLOCKS;

// Division of a 32-bit index by a divisor fixed per launch: one multiply-high and two shifts (Granlund-Montgomery) in
// place of the ~20-instruction sequence the compiler emits for `/` -- the warp-tile coordinates are derived once per tile
// in the transform warps of the decode kernel, where every instruction is on the critical path.
struct fastdiv {
  unsigned d, m, s1, s2;
  __host__ __device__ unsigned div(unsigned n) const
  {
#if defined(__CUDA_ARCH__)
    const unsigned t = __umulhi(m, n);
#else
    const unsigned t = (unsigned)(((unsigned long long)m * n) >> 32);
#endif
    return (t + ((n - t) >> s1)) >> s2;
  }
};
inline fastdiv make_fastdiv(unsigned d)
{
  fastdiv f;
  f.d = d;
  unsigned L = 0;
  while ((1ull << L) < d) L++;
  f.m = (unsigned)((((1ull << L) - d) << 32) / d + 1ull);
  f.s1 = L < 1 ? L : 1u;
  f.s2 = L < 1 ? 0u : L - 1u;
  return f;
}

struct tile_geom {
  fastdiv div_wpr, div_nby;  // warp-tiles per block row ((nbx + 31) / 32) and nby, for the warp-tile kernels
  unsigned nbx, nby, nbz;    // FULL blocks per axis handled by the tile kernels (ny/4, nz/4; nx % 4 == 0)
  unsigned nby_t;            // block rows per z-slab in the stream, partial row included ((ny+3)/4)
  unsigned bpr;              // boxes per block-row = ceil(nbx / 64)
  unsigned nboxes;           // bpr * nby * nbz
  unsigned ntiles;           // ceil(nboxes / 2)
  unsigned nx, ny, nz;
  unsigned W;                // 64-bit words per slot
};

struct box_info {
  unsigned x0, y0, z0;   // element coordinates of the box origin
  size_t blk0;           // first block index (stream order)
  unsigned nvalid;       // blocks of the box that exist (<= 64); 0 if the box does not exist
};

__device__ __forceinline__ box_info box_of(const tile_geom& tg, unsigned q)
{
  box_info b;
  if (q >= tg.nboxes) { b.x0 = b.y0 = b.z0 = 0; b.blk0 = 0; b.nvalid = 0; return b; }
  const unsigned row = q / tg.bpr, xb = q - row * tg.bpr;
  const unsigned iy = row % tg.nby, iz = row / tg.nby;
  b.x0 = xb * 256u; b.y0 = iy * 4u; b.z0 = iz * 4u;
  b.blk0 = ((size_t)iz * tg.nby_t + iy) * tg.nbx + (size_t)xb * BOX_BLOCKS;
  const unsigned rem = tg.nbx - xb * BOX_BLOCKS;
  b.nvalid = rem < (unsigned)BOX_BLOCKS ? rem : (unsigned)BOX_BLOCKS;
  return b;
}

// dynamic shared memory layout (encode): [tile: 2 boxes x 16 KiB][out: 128*W words][3 mbarriers]
// Per tile: wait for the TMA boxes -> 16 LDS.128 per thread -> transform in registers -> bit planes parked
// in the thread's own 16-byte columns of the tile (in place) -> plane loop emits the slot into `out`
// -> one __syncthreads -> thread 0 issues this tile's bulk store, waits until the store has read `out`,
// and only then issues the next tile's TMA loads: every thread waits for those loads before it touches
// `out` again, so one staging buffer is enough (43 KiB per CTA -> 5 CTAs per SM at 95 registers).
// MINMAX: the minimum and maximum of the field are harvested on the way (the values are in registers anyway) into one
// row of `part` per CTA, for the value-range histogram that the decode pass fuses (minmaxMetric.hpp:62-81 needs them
// before any value can be binned).
template <bool MINMAX>
#ifndef ZFB_ENC_CTAS
#define ZFB_ENC_CTAS 5
#endif
__global__ void __launch_bounds__(TILE_THREADS, ZFB_ENC_CTAS)
enc3_f32_tile_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t* __restrict__ stream, tile_geom tg,
                     unsigned maxbits, double* __restrict__ part)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  // slot staging in the stream's own layout (one bulk store per tile).  The 64-byte slot stride makes the lanes of a warp
  // hit two banks when they write word j of their slots; a padded layout with the warp copying its piece to the stream
  // itself removes the conflicts but costs more than it saves here (2.21 against 2.16 ms, profiles/r02_notes.md).
  uint64_t* outw = reinterpret_cast<uint64_t*>(smem_raw + BOXES_PER_TILE * BOX_FLOATS * sizeof(float));
  const unsigned out_words = TILE_THREADS * tg.W;
  uint64_t* bars = outw + (size_t)out_words;
  __shared__ enc_coder_smem cs;

  const unsigned tid = threadIdx.x;
  // the box index through a shuffle: warp-uniform to the compiler, so what derives from it stays in uniform registers
  const unsigned box = __shfl_sync(0xffffffffu, tid >> 6, 0), lb = tid & 63;

  // box descriptors of the current / next tile, computed once by thread 0 (integer divisions) and read by all
  __shared__ box_info sb[2][BOXES_PER_TILE];
  unsigned t = blockIdx.x;
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::mbar_init(&bars[2], 1);  // `out` free again: the previous tile's bulk store has read it
    ptx::fence_barrier_init();
    sb[0][0] = box_of(tg, 2 * t);
    sb[0][1] = box_of(tg, 2 * t + 1);
  }
  const f3::tables tb = build_coder_tables(cs);  // includes the __syncthreads that publishes barriers and sb[0]

  auto issue_loads = [&](const box_info* bi) {
#pragma unroll
    for (int q = 0; q < BOXES_PER_TILE; q++) {
      if (bi[q].nvalid) {
        ptx::mbar_expect_tx(&bars[q], BOX_FLOATS * sizeof(float));
        ptx::tma_load_3d(tile + q * BOX_FLOATS, &tmap, (int)bi[q].x0, (int)bi[q].y0, (int)bi[q].z0, &bars[q]);
      }
    }
  };

  if (tid == 0 && t < tg.ntiles) issue_loads(sb[0]);

  f3::smem_rows_t<0> ps;  // the thread's 16-byte column of its box: row r at r * 1 KiB
  ps.base = ptx::smem_u32(tile + box * BOX_FLOATS) + 16u * lb;
  ps.stride = 1024u;
  float vmax = -3.402823466e38f, vnmin = -3.402823466e38f;  // MINMAX: max(f), max(-f)

  for (unsigned it = 0; t < tg.ntiles; t += gridDim.x, it++) {
    const box_info b0 = sb[it & 1][0], b1 = sb[it & 1][1];
    const box_info mine = box ? b1 : b0;
    const bool active = lb < mine.nvalid;
    if (tid == 0) {
      // The tile buffer is busy until this tile is coded (the plane rows live in it), so the next tile's TMA load can
      // only be issued at the end; its boxes are pulled into L2 now, which takes DRAM out of that load's latency
      // (ncu: 11 % of the warps' time was the wait for the tile).
      box_info* nb = sb[(it + 1) & 1];
      nb[0] = box_of(tg, 2 * (t + gridDim.x));
      nb[1] = box_of(tg, 2 * (t + gridDim.x) + 1);
#pragma unroll
      for (int q = 0; q < BOXES_PER_TILE; q++)
        if (nb[q].nvalid) ptx::tma_prefetch_3d(&tmap, (int)nb[q].x0, (int)nb[q].y0, (int)nb[q].z0);
    }
    if (mine.nvalid) {
      ptx::mbar_wait(&bars[box], it & 1);
      if (active) {
        float f[64];
        const float4* src = reinterpret_cast<const float4*>(tile + box * BOX_FLOATS) + lb;
#pragma unroll
        for (int r = 0; r < 16; r++) {
          const float4 v = src[r * 64];  // row r of the box is 256 floats = 64 float4
          f[4 * r] = v.x; f[4 * r + 1] = v.y; f[4 * r + 2] = v.z; f[4 * r + 3] = v.w;
          if (MINMAX) {
            vmax = fmaxf(fmaxf(vmax, v.x), v.y); vmax = fmaxf(fmaxf(vmax, v.z), v.w);
            vnmin = fmaxf(fmaxf(vnmin, -v.x), -v.y); vnmin = fmaxf(fmaxf(vnmin, -v.z), -v.w);
          }
        }
        const unsigned p = (box ? b0.nvalid : 0u) + lb;  // slot position inside the tile's contiguous span
        f3::smem_sink32 sink(ptx::smem_u32(outw + (size_t)p * tg.W), 2 * tg.W);
        ptx::mbar_wait(&bars[2], (it & 1) ^ 1);  // (long done by now: the store was issued before this tile's loads landed)
        f3::encode_block(f, maxbits, sink, tb, ps);
      }
    }
    ptx::fence_proxy_async();              // slot words and the in-place plane rows -> async proxy
    __syncthreads();                       // tile buffer consumed, `out` complete, next descriptors published
    if (tid == 0) {
      // The next tile's loads go first: the tile buffer is free, and the store below then reads `out` while they are
      // in flight (they used to wait for it).  `out` is handed back through its own barrier.
      if (t + gridDim.x < tg.ntiles) issue_loads(sb[(it + 1) & 1]);
      // the two boxes' slots are adjacent in the stream except across a partial block row
      if (b1.nvalid && b1.blk0 != b0.blk0 + b0.nvalid) {
        ptx::bulk_store(stream + b0.blk0 * tg.W, outw, b0.nvalid * tg.W * 8u);
        ptx::bulk_store(stream + b1.blk0 * tg.W, outw + (size_t)b0.nvalid * tg.W, b1.nvalid * tg.W * 8u);
      } else {
        ptx::bulk_store(stream + b0.blk0 * tg.W, outw, (b0.nvalid + b1.nvalid) * tg.W * 8u);
      }
      ptx::bulk_commit();
      ptx::bulk_wait_read0();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(&bars[2])) : "memory");
    }
  }
  if (tid == 0) ptx::bulk_wait0();
  if (MINMAX) {
    __shared__ double red[7 * 4];
    maccf m;
    m.init();
    m.m_o = vmax; m.m_no = vnmin;
    macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
  }
}

// dynamic shared memory layout (decode): [tile: 2 x 16 KiB plane scratch, then the original boxes]
//                                        [slots: 2 x 128*W words][4 mbarriers]
// Per tile: wait for the slot span (bulk copy, issued two tiles ahead into a double buffer) -> plane loop
// parses the slot into the thread's plane rows -> rows back into registers -> __syncthreads -> thread 0
// refills the slot buffer just consumed and starts the TMA load of this tile's ORIGINAL boxes into the
// plane-scratch buffer while all threads run the inverse transform -> metrics against the original rows
// (LDS.128) -> 16 coalesced 128-bit stores of the decoded rows -> __syncthreads.
// 4 CTAs/SM at 128 registers: a 5-CTA / 96-register build runs at the same speed but its spills add
// ~50 % DRAM write traffic (profiles/r01_notes.md), so the spill-free configuration is kept.
// HIST: the decoded values are also binned into the 1024-bin value-range histogram of the minmax metric
// (minmaxMetric.hpp:92-133) over [hlo, hhi] = the field's global minimum / maximum harvested by the encode pass:
// shared-memory counters per CTA, flushed once; no extra pass over memory.
#ifndef ZFB_DEC_SLOT_BUFS
#define ZFB_DEC_SLOT_BUFS 1  // landing buffers for the slot span per CTA (1: the next tile's span arrives during the transform; 2: two tiles ahead)
#endif
#ifndef ZFB_DEC_CTAS
#define ZFB_DEC_CTAS 4       // resident CTAs per SM the register budget is cut for
#endif
constexpr unsigned DEC_SLOT_BUFS = ZFB_DEC_SLOT_BUFS;
template <bool METRICS, bool HIST>
__global__ void __launch_bounds__(TILE_THREADS, ZFB_DEC_CTAS)
dec3_f32_tile_kernel(const __grid_constant__ CUtensorMap tmap_orig, const uint64_t* __restrict__ stream,
                     float* __restrict__ out, tile_geom tg, unsigned maxbits, double* __restrict__ part, double hlo,
                     double hhi, unsigned long long* __restrict__ hcounts)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  uint64_t* slots = reinterpret_cast<uint64_t*>(smem_raw + BOXES_PER_TILE * BOX_FLOATS * sizeof(float));
  const unsigned slot_words = TILE_THREADS * tg.W;
  // The slot span lands contiguously (bulk copy) and is re-laid by its warp into a PADDED area: slot p at 32-bit word
  // p * (2W + 1).  With the natural 64-byte slot stride the 32 lanes of a warp hit two banks whenever they fetch the
  // same word of their slots (ncu: 5.8 wavefronts per shared load, shared pipe 51 % busy); the odd stride is conflict
  // free.
  const unsigned W2 = 2u * tg.W, pstride = W2 + 1u;
  const uint32_t land_s = ptx::smem_u32(slots);
  const uint32_t pad_s = land_s + DEC_SLOT_BUFS * slot_words * 8u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(slots + DEC_SLOT_BUFS * (size_t)slot_words) +
                                               (((size_t)TILE_THREADS * pstride * 4u + 16u + 15u) & ~(size_t)15));  // [0,1] slot buffers full, [2,3] orig boxes full
  __shared__ double red[7 * 4];
  __shared__ dec_coder_smem cs;
  __shared__ box_info sb[3][BOXES_PER_TILE];  // ring of box descriptors: current, next, next-next tile

  const unsigned tid = threadIdx.x;
  const unsigned box = tid >> 6, lb = tid & 63;
  unsigned t = blockIdx.x;
  if (tid == 0) {
    if (METRICS) ptx::prefetch_tmap(&tmap_orig);
    for (int i = 0; i < 4; i++) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
    for (int j = 0; j < 2; j++) {
      sb[j][0] = box_of(tg, 2 * (t + j * gridDim.x));
      sb[j][1] = box_of(tg, 2 * (t + j * gridDim.x) + 1);
    }
  }
  const f3::tables tb = build_coder_tables(cs);  // includes the __syncthreads that publishes barriers and sb

  auto issue_slots = [&](const box_info* bi, unsigned buf) {
    uint64_t* dst = slots + (size_t)buf * slot_words;
    const unsigned bytes = (bi[0].nvalid + bi[1].nvalid) * tg.W * 8u;
    ptx::mbar_expect_tx(&bars[buf], bytes);
    if (bi[1].nvalid && (bi[1].blk0 != bi[0].blk0 + bi[0].nvalid || bi[0].nvalid != (unsigned)BOX_BLOCKS)) {
      // box 1's slots always land at slot 64 of the buffer, so that thread tid's slot is slot tid and a warp's slots
      // are its own 32
      ptx::bulk_load(dst, stream + bi[0].blk0 * tg.W, bi[0].nvalid * tg.W * 8u, &bars[buf]);
      ptx::bulk_load(dst + (size_t)BOX_BLOCKS * tg.W, stream + bi[1].blk0 * tg.W, bi[1].nvalid * tg.W * 8u, &bars[buf]);
    } else {
      ptx::bulk_load(dst, stream + bi[0].blk0 * tg.W, bytes, &bars[buf]);
    }
  };
  auto issue_orig = [&](const box_info* bi) {
#pragma unroll
    for (int q = 0; q < BOXES_PER_TILE; q++) {
      if (bi[q].nvalid) {
        ptx::mbar_expect_tx(&bars[2 + q], BOX_FLOATS * sizeof(float));
        ptx::tma_load_3d(tile + q * BOX_FLOATS, &tmap_orig, (int)bi[q].x0, (int)bi[q].y0, (int)bi[q].z0, &bars[2 + q]);
      }
    }
  };

  if (tid == 0) {
    if (t < tg.ntiles) issue_slots(sb[0], 0);
    if (DEC_SLOT_BUFS == 2 && t + gridDim.x < tg.ntiles) issue_slots(sb[1], 1);
  }

  f3::smem_rows_t<0> ps;  // the thread's 16-byte column of its box: row r at r * 1 KiB
  ps.base = ptx::smem_u32(tile + box * BOX_FLOATS) + 16u * lb;
  ps.stride = 1024u;

  maccf m;
  m.init();
  __shared__ unsigned int hbins[HIST ? 1024 : 1];
  __shared__ unsigned int hbad;
  value_binner vb;
  if (HIST) {
    vb.init(hlo, hhi);
    for (unsigned i = tid; i < 1024u; i += TILE_THREADS) hbins[i] = 0u;
    if (tid == 0) hbad = 0u;
    __syncthreads();
  }
  unsigned cur = 0;  // it % 3
  for (unsigned it = 0; t < tg.ntiles; t += gridDim.x, it++, cur = cur == 2 ? 0 : cur + 1) {
    const unsigned buf = DEC_SLOT_BUFS == 2 ? (it & 1) : 0u, nn = cur == 0 ? 2 : cur - 1;  // nn = (it + 2) % 3
    const box_info b0 = sb[cur][0], b1 = sb[cur][1];
    const box_info mine = box ? b1 : b0;
    const bool active = lb < mine.nvalid;
    const unsigned nslots = b0.nvalid + b1.nvalid;
    ptx::mbar_wait(&bars[buf], (DEC_SLOT_BUFS == 2 ? (it >> 1) : it) & 1);
    f3::planes64 P;
    int emax = 0;
    bool nonzero = false, low_used = false;
    {
      // re-lay the warp's 32 slots (words [32 w W2, 32 (w + 1) W2) of the landing buffer) into the padded area
      const uint32_t src = land_s + buf * slot_words * 8u + 4u * (tid & ~31u) * W2;
      const uint32_t dstp = pad_s + 4u * (tid & ~31u) * pstride;
      const unsigned lane = tid & 31u;
      if (W2 == 16u) {  // 8 bits per value: all 16 loads in flight, then 16 stores
        uint32_t v[16];
#pragma unroll
        for (int m = 0; m < 16; m++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[m]) : "r"(src + 4u * (lane + 32u * m)));
#pragma unroll
        for (int m = 0; m < 16; m++) {
          const unsigned i = lane + 32u * m;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(dstp + 4u * ((i >> 4) * 17u + (i & 15u))), "r"(v[m]) : "memory");
        }
      } else if ((W2 & (W2 - 1u)) == 0u) {
        const unsigned sh = 31u - (unsigned)__clz((int)W2);
#pragma unroll 4
        for (unsigned i = lane; i < 32u * W2; i += 32u) {
          uint32_t v;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(src + 4u * i));
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(dstp + 4u * ((i >> sh) * pstride + (i & (W2 - 1u)))), "r"(v) : "memory");
        }
      } else {
        for (unsigned i = lane; i < 32u * W2; i += 32u) {
          uint32_t v;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(src + 4u * i));
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(dstp + 4u * ((i / W2) * pstride + i % W2)), "r"(v) : "memory");
        }
      }
      __syncwarp();
    }
    if (active) {
      f3::smem_words r;
      r.base = pad_s + 4u * tid * pstride;
      nonzero = f3::decode_planes(P, emax, low_used, maxbits, r, tb, ps);
    }
    ptx::fence_proxy_async();  // plane-row writes (generic proxy) before the TMA load overwrites the buffer
    if (tid == 0) {
      sb[nn][0] = box_of(tg, 2 * (t + 2 * gridDim.x));
      sb[nn][1] = box_of(tg, 2 * (t + 2 * gridDim.x) + 1);
    }
    __syncthreads();           // slots[buf] parsed; planes back in registers; descriptors of tile it+2 published
    if (tid == 0) {
      if (DEC_SLOT_BUFS == 2) {
        if (t + 2 * gridDim.x < tg.ntiles) issue_slots(sb[nn], buf);
      } else if (t + gridDim.x < tg.ntiles) {
        issue_slots(sb[cur == 2 ? 0 : cur + 1], 0);  // the one buffer has been parsed: the next tile's slots land during the transform
      }
      if (METRICS) issue_orig(sb[cur]);
    }
    if (active) {
      float f[64];
      f3::decode_finish(f, P, emax, nonzero, low_used);
      float* dst = out + ((size_t)mine.z0 * tg.ny + mine.y0) * tg.nx + mine.x0 + 4 * lb;
      if (METRICS) {
        ptx::mbar_wait(&bars[2 + box], it & 1);
        const float4* src = reinterpret_cast<const float4*>(tile + box * BOX_FLOATS) + lb;
        frow fr;
        fr.init();
#pragma unroll
        for (int r16 = 0; r16 < 16; r16++) {
          const float4 o = src[r16 * 64];
          fr.add4(o, f[4 * r16], f[4 * r16 + 1], f[4 * r16 + 2], f[4 * r16 + 3]);
          if ((r16 & 3) == 3) fr.flush(m);
        }
      }
      if (HIST) {
        // The 64 values of a block lie close together: they are counted in sixteen 8-bit counters (two registers)
        // covering the 16 bins around the first value's, and only the non-empty counters go to the CTA's shared
        // histogram -- one atomic per distinct bin instead of one per value (shared atomics at one per value cost
        // three times the rest of the kernel).  Values outside the window take the direct atomic.
        const int b0 = vb.bin_of(f[0], &hbad);
        int base = b0 - 8;
        base = base < 0 ? 0 : (base > 1008 ? 1008 : base);
        uint64_t c_lo = 0, c_hi = 0;
#pragma unroll
        for (int i = 0; i < 64; i++) {  // fully unrolled: f[] must stay in registers
          const int d = (i ? vb.bin_of(f[i], &hbad) : b0) - base;
          const uint64_t one = (uint64_t)1 << ((d & 7) * 8);
          if ((unsigned)d < 8u) c_lo += one;
          else if ((unsigned)d < 16u) c_hi += one;
          else atomicAdd(&hbins[base + d], 1u);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const unsigned n0 = (unsigned)(c_lo >> (8 * j)) & 0xffu, n1 = (unsigned)(c_hi >> (8 * j)) & 0xffu;
          if (n0) atomicAdd(&hbins[base + j], n0);
          if (n1) atomicAdd(&hbins[base + 8 + j], n1);
        }
      }
      const size_t rs = tg.nx, ps2 = (size_t)tg.ny * tg.nx;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float* rowp = dst + (size_t)k * ps2;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float4 v;
          v.x = f[16 * k + 4 * j]; v.y = f[16 * k + 4 * j + 1]; v.z = f[16 * k + 4 * j + 2]; v.w = f[16 * k + 4 * j + 3];
          __stcs(reinterpret_cast<float4*>(rowp + (size_t)j * rs), v);
        }
      }
    } else if (METRICS && mine.nvalid) {
      ptx::mbar_wait(&bars[2 + box], it & 1);  // keep every thread's view of the barrier phase in step
    }
    if (METRICS) __syncthreads();  // the original rows are consumed before the next tile's planes land there
  }
  if (METRICS) macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
  if (HIST) {
    __syncthreads();
    for (unsigned i = tid; i < 1024u; i += TILE_THREADS)
      if (hbins[i]) atomicAdd(&hcounts[i], (unsigned long long)hbins[i]);
    if (tid == 0 && hbad) atomicAdd(&hcounts[1024], (unsigned long long)hbad);
  }
}

// ---------------------------------------------------------------------------------------------
// 3D double tile kernels (nx % 4 == 0, whole-word slots): the float design with 64 threads per CTA and ONE
// TMA box of 256x4x4 doubles (32 KiB, 64 x-consecutive blocks) per tile.  A block needs ~128 registers for
// its 64 doubles alone, so CTAs are small and 4 fit per SM.  Plane rows are parked in place: a thread's 16
// rows of 32 bytes hold its 32 plane rows of 16 bytes.
// ---------------------------------------------------------------------------------------------
constexpr int DTILE_THREADS = 64;
constexpr int DBOX_DOUBLES = 256 * 4 * 4;  // 32 KiB

// dynamic shared memory: [tile 32 KiB][out: 64*W words][1 mbarrier]
__global__ void __launch_bounds__(DTILE_THREADS, 4)
enc3_f64_tile_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t* __restrict__ stream, tile_geom tg,
                     unsigned maxbits)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);
  uint64_t* outw = reinterpret_cast<uint64_t*>(smem_raw + DBOX_DOUBLES * sizeof(double));
  uint64_t* bars = outw + (size_t)DTILE_THREADS * tg.W;
  __shared__ enc_coder_smem cs;
  __shared__ box_info sb[2];

  const unsigned tid = threadIdx.x;
  unsigned t = blockIdx.x;  // tile = box index
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap);
    ptx::mbar_init(&bars[0], 1);
    ptx::fence_barrier_init();
    sb[0] = box_of(tg, t);
  }
  const f3::tables tb = build_coder_tables(cs);

  auto issue_load = [&](const box_info& bi) {
    if (bi.nvalid) {
      ptx::mbar_expect_tx(&bars[0], DBOX_DOUBLES * sizeof(double));
      ptx::tma_load_3d(tile, &tmap, (int)bi.x0, (int)bi.y0, (int)bi.z0, &bars[0]);
    }
  };
  if (tid == 0 && t < tg.nboxes) issue_load(sb[0]);

  f3::smem_rows_t<1> ps;  // the thread's 32-byte column: two plane rows per 2 KiB row of the box
  ps.base = ptx::smem_u32(tile) + 32u * tid;
  ps.stride = 2048u;

  for (unsigned it = 0; t < tg.nboxes; t += gridDim.x, it++) {
    const box_info b0 = sb[it & 1];
    const bool active = tid < b0.nvalid;
    if (tid == 0) {
      // the next box into L2 while this one is coded (its TMA load can only be issued when the tile buffer is free)
      sb[(it + 1) & 1] = box_of(tg, t + gridDim.x);
      if (sb[(it + 1) & 1].nvalid) ptx::tma_prefetch_3d(&tmap, (int)sb[(it + 1) & 1].x0, (int)sb[(it + 1) & 1].y0, (int)sb[(it + 1) & 1].z0);
    }
    ptx::mbar_wait(&bars[0], it & 1);
    if (active) {
      double f[64];
      const double2* src = reinterpret_cast<const double2*>(tile) + 2 * tid;
#pragma unroll
      for (int r = 0; r < 16; r++) {
        const double2 v0 = src[r * 128], v1 = src[r * 128 + 1];
        f[4 * r] = v0.x; f[4 * r + 1] = v0.y; f[4 * r + 2] = v1.x; f[4 * r + 3] = v1.y;
      }
      f3::smem_sink32 sink(ptx::smem_u32(outw + (size_t)tid * tg.W), 2 * tg.W);
      d3::encode_block(f, maxbits, sink, tb, ps);
    }
    ptx::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      ptx::bulk_store(stream + b0.blk0 * tg.W, outw, b0.nvalid * tg.W * 8u);
      ptx::bulk_commit();
      ptx::bulk_wait_read0();
      if (t + gridDim.x < tg.nboxes) issue_load(sb[(it + 1) & 1]);
    }
  }
  if (tid == 0) ptx::bulk_wait0();
}

// dynamic shared memory: [tile 32 KiB: plane scratch, then the original box][slots: 2 x 64*W words][3 mbarriers]
template <bool METRICS>
__global__ void __launch_bounds__(DTILE_THREADS, 4)
dec3_f64_tile_kernel(const __grid_constant__ CUtensorMap tmap_orig, const uint64_t* __restrict__ stream,
                     double* __restrict__ out, tile_geom tg, unsigned maxbits, double* __restrict__ part)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);
  uint64_t* slots = reinterpret_cast<uint64_t*>(smem_raw + DBOX_DOUBLES * sizeof(double));
  const unsigned slot_words = DTILE_THREADS * tg.W;
  uint64_t* bars = slots + 2 * (size_t)slot_words;  // [0,1] slot buffers, [2] orig box
  __shared__ double red[7 * 2];
  __shared__ dec_coder_smem cs;
  __shared__ box_info sb[3];

  const unsigned tid = threadIdx.x;
  unsigned t = blockIdx.x;
  if (tid == 0) {
    if (METRICS) ptx::prefetch_tmap(&tmap_orig);
    for (int i = 0; i < 3; i++) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
    sb[0] = box_of(tg, t);
    sb[1] = box_of(tg, t + gridDim.x);
  }
  const f3::tables tb = build_coder_tables(cs);

  auto issue_slots = [&](const box_info& bi, unsigned buf) {
    const unsigned bytes = bi.nvalid * tg.W * 8u;
    ptx::mbar_expect_tx(&bars[buf], bytes);
    ptx::bulk_load(slots + (size_t)buf * slot_words, stream + bi.blk0 * tg.W, bytes, &bars[buf]);
  };
  if (tid == 0) {
    if (t < tg.nboxes) issue_slots(sb[0], 0);
    if (t + gridDim.x < tg.nboxes) issue_slots(sb[1], 1);
  }

  f3::smem_rows_t<1> ps;
  ps.base = ptx::smem_u32(tile) + 32u * tid;
  ps.stride = 2048u;

  macc m;
  m.init();
  unsigned cur = 0;
  for (unsigned it = 0; t < tg.nboxes; t += gridDim.x, it++, cur = cur == 2 ? 0 : cur + 1) {
    const unsigned buf = it & 1, nn = cur == 0 ? 2 : cur - 1;
    const box_info b0 = sb[cur];
    const bool active = tid < b0.nvalid;
    ptx::mbar_wait(&bars[buf], (it >> 1) & 1);
    f3::planes64 PH, PL;
    int emax = 0, kend = 63;
    bool nonzero = false;
    if (active) {
      f3::smem_words r;
      r.base = ptx::smem_u32(slots + (size_t)buf * slot_words + (size_t)tid * tg.W);
      nonzero = d3::decode_planes(PH, PL, emax, kend, maxbits, r, tb, ps);
    }
    ptx::fence_proxy_async();
    if (tid == 0) sb[nn] = box_of(tg, t + 2 * gridDim.x);
    __syncthreads();
    if (tid == 0) {
      if (t + 2 * gridDim.x < tg.nboxes) issue_slots(sb[nn], buf);
      if (METRICS && b0.nvalid) {
        ptx::mbar_expect_tx(&bars[2], DBOX_DOUBLES * sizeof(double));
        ptx::tma_load_3d(tile, &tmap_orig, (int)b0.x0, (int)b0.y0, (int)b0.z0, &bars[2]);
      }
    }
    if (active) {
      double f[64];
      d3::decode_finish(f, PH, PL, emax, nonzero, kend);
      double* dst = out + ((size_t)b0.z0 * tg.ny + b0.y0) * tg.nx + b0.x0 + 4 * tid;
      if (METRICS) {
        ptx::mbar_wait(&bars[2], it & 1);
        const double2* src = reinterpret_cast<const double2*>(tile) + 2 * tid;
#pragma unroll
        for (int r = 0; r < 16; r++) {
          const double2 o0 = src[r * 128], o1 = src[r * 128 + 1];
          macc_add(m, o0.x, f[4 * r]); macc_add(m, o0.y, f[4 * r + 1]); macc_add(m, o1.x, f[4 * r + 2]); macc_add(m, o1.y, f[4 * r + 3]);
        }
      }
      const size_t rs = tg.nx, ps2 = (size_t)tg.ny * tg.nx;
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double* rowp = dst + (size_t)k * ps2 + (size_t)j * rs;
          __stcs(reinterpret_cast<double2*>(rowp), make_double2(f[16 * k + 4 * j], f[16 * k + 4 * j + 1]));
          __stcs(reinterpret_cast<double2*>(rowp) + 1, make_double2(f[16 * k + 4 * j + 2], f[16 * k + 4 * j + 3]));
        }
    } else if (METRICS && b0.nvalid) {
      ptx::mbar_wait(&bars[2], it & 1);
    }
    if (METRICS) __syncthreads();
  }
  if (METRICS) macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
}

// ---------------------------------------------------------------------------------------------
// 1D float kernels (HACC particle arrays): one 4-value block per thread, one 128-bit load / store per
// block, slot words written with one 4/8/16-byte store per thread (coalesced across the warp).
// Only FULL blocks (b < nfull); a trailing partial block goes through the generic kernel.
// ---------------------------------------------------------------------------------------------
struct coder1_smem {
  uint32_t spread[256];
  uint32_t unspread[256];
  uint16_t dec[512];
  uint16_t enc[64];
};
// ENC / DEC: which half of the tables the kernel needs (the decode table alone costs up to 16 plane codes per entry)
template <bool ENC = true, bool DEC = true>
__device__ __forceinline__ f1::tables build_coder1_tables(coder1_smem& cs)
{
  if (ENC) {
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) cs.spread[i] = f1::spread_entry(i);
    for (unsigned i = threadIdx.x; i < 64; i += blockDim.x) cs.enc[i] = f1::enc_entry(i);
  }
  if (DEC) {
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) cs.unspread[i] = f1::unspread_entry(i);
    for (unsigned i = threadIdx.x; i < 512; i += blockDim.x) cs.dec[i] = f1::dec_entry(i);
  }
  __syncthreads();
  f1::tables t;
  t.enc = cs.enc; t.dec = cs.dec; t.spread = cs.spread; t.unspread = cs.unspread;
  return t;
}

template <int MAXW>
__global__ void __launch_bounds__(256) enc1_f32_kernel(const float* __restrict__ in, uint32_t* __restrict__ stream,
                                                       size_t nfull, unsigned maxbits)
{
  __shared__ coder1_smem cs;
  const f1::tables tb = build_coder1_tables<true, false>(cs);
  const unsigned nw = maxbits >> 5;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nfull; b += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(in) + b);
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[MAXW];
    f1::encode_block<MAXW>(f, maxbits, o, tb);
    if constexpr (MAXW == 1) stream[b] = o[0];
    else if constexpr (MAXW == 2) reinterpret_cast<uint2*>(stream)[b] = make_uint2(o[0], o[1]);
    else {
      if (nw == 4) reinterpret_cast<uint4*>(stream)[b] = make_uint4(o[0], o[1], o[2], o[3]);
      else { stream[3 * b] = o[0]; stream[3 * b + 1] = o[1]; stream[3 * b + 2] = o[2]; }
    }
  }
}

#ifndef ZFB_DEC1_MINB
#define ZFB_DEC1_MINB 6  // resident CTAs per SM the register allocation must allow for 32/64-bit slots (40 registers; 4 CTAs at 50:
                         // decode 1.27 -> 1.19 ms); wider slots keep 4 (at 40 registers the 128-bit instantiation spills)
#endif
template <int MAXW, bool METRICS>
__global__ void __launch_bounds__(256, MAXW <= 2 ? ZFB_DEC1_MINB : 4) dec1_f32_kernel(const uint32_t* __restrict__ stream, float* __restrict__ out,
                                                       const float* __restrict__ orig, size_t nfull, unsigned maxbits,
                                                       double* __restrict__ part)
{
  __shared__ coder1_smem cs;
  __shared__ double red[7 * 8];
  const f1::tables tb = build_coder1_tables<false, true>(cs);
  const unsigned nw = maxbits >> 5;
  maccf m;
  m.init();
  frow fr;
  fr.init();
  int pending = 0;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nfull; b += (size_t)gridDim.x * blockDim.x) {
    uint32_t w[MAXW];
    if constexpr (MAXW == 1) w[0] = __ldcs(stream + b);
    else if constexpr (MAXW == 2) { const uint2 t = __ldcs(reinterpret_cast<const uint2*>(stream) + b); w[0] = t.x; w[1] = t.y; }
    else {
      if (nw == 4) { const uint4 t = __ldcs(reinterpret_cast<const uint4*>(stream) + b); w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w; }
      else { w[0] = stream[3 * b]; w[1] = stream[3 * b + 1]; w[2] = stream[3 * b + 2]; w[3] = 0u; }
    }
    float f[4];
    f1::decode_block<MAXW>(f, maxbits, w, tb);
    if (METRICS) {
      const float4 o = __ldcs(reinterpret_cast<const float4*>(orig) + b);
      fr.add4(o, f[0], f[1], f[2], f[3]);
      if (++pending == 4) { fr.flush(m); pending = 0; }
    }
    __stcs(reinterpret_cast<float4*>(out) + b, make_float4(f[0], f[1], f[2], f[3]));
  }
  if (METRICS) {
    fr.flush(m);
    macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
  }
}

// ---------------------------------------------------------------------------------------------
// stand-alone streaming metrics and histograms
// ---------------------------------------------------------------------------------------------
template <typename Scalar>
__global__ void __launch_bounds__(256) metrics_kernel(const Scalar* __restrict__ orig, const Scalar* __restrict__ approx,
                                                      size_t n, double* __restrict__ part)
{
  __shared__ double red[7 * 8];
  macc_t<Scalar> m;
  m.init();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (sizeof(Scalar) == 4) {
    const size_t n4 = n / 4;
    const float4* o4 = reinterpret_cast<const float4*>(orig);
    const float4* a4 = reinterpret_cast<const float4*>(approx);
    frow fr;
    fr.init();
    int cnt = 0;
    for (; i + stride < n4; i += 2 * stride) {  // two independent 16-byte pairs in flight
      const float4 o0 = __ldcs(o4 + i), a0 = __ldcs(a4 + i);
      const float4 o1 = __ldcs(o4 + i + stride), a1 = __ldcs(a4 + i + stride);
      fr.add4(o0, a0.x, a0.y, a0.z, a0.w);
      fr.add4(o1, a1.x, a1.y, a1.z, a1.w);
      if (++cnt == 2) { fr.flush(m); cnt = 0; }
    }
    if (i < n4) {
      const float4 o = __ldcs(o4 + i), a = __ldcs(a4 + i);
      fr.add4(o, a.x, a.y, a.z, a.w);
    }
    const size_t tail = n4 * 4 + ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (tail < n) fr.add(orig[tail], approx[tail]);
    fr.flush(m);
  } else {
    // 16-byte loads, two independent pairs in flight per thread and iteration
    const size_t n2 = n / 2;
    const double2* o2 = reinterpret_cast<const double2*>(orig);
    const double2* a2 = reinterpret_cast<const double2*>(approx);
    for (; i + stride < n2; i += 2 * stride) {
      const double2 o0 = __ldcs(o2 + i), a0 = __ldcs(a2 + i);
      const double2 o1 = __ldcs(o2 + i + stride), a1 = __ldcs(a2 + i + stride);
      macc_add(m, o0.x, a0.x); macc_add(m, o0.y, a0.y);
      macc_add(m, o1.x, a1.x); macc_add(m, o1.y, a1.y);
    }
    if (i < n2) {
      const double2 o0 = __ldcs(o2 + i), a0 = __ldcs(a2 + i);
      macc_add(m, o0.x, a0.x); macc_add(m, o0.y, a0.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) macc_add(m, (double)orig[n - 1], (double)approx[n - 1]);
  }
  macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
}

// which: 0 abs error, 1 rel error, 2 approx value.  Binning follows the reference expression by
// expression (double arithmetic, int truncation, `if (pos >= numBins) pos = pos - 1`), with out-of-range
// positions clamped and counted instead of indexing out of bounds.
// One value -> its bin, exactly as the reference computes it (counts out-of-range values in `bad`).
template <typename Scalar>
__device__ __forceinline__ int hist_bin(int which, Scalar ov, Scalar av, double lo, double range, double bin, double rbin,
                                        double rrange, unsigned int* bad)
{
  // The reference divides in double and truncates.  An IEEE division is ~40 instructions, so the quotient is first
  // estimated with reciprocal multiplies (a few ulp off); only when the estimate lies within 1e-9 (relative) of an
  // integer -- where a few ulp could change the truncation -- is the reference's exact expression evaluated.
  double q;
  if (which == 2) {
    const double v = (double)av;
    q = (range * ((v - lo) * rrange)) * rbin;
    const double k = rint(q);
    if (fabs(q - k) <= 1e-9 * (k + 1.0)) q = (range * ((v - lo) / range)) / bin;  // minmaxMetric.hpp:112
  } else {
    double err;
    if constexpr (sizeof(Scalar) == 4) {
      const float o = ov, a = av;
      const float ad = fabsf(o - a);
      if (which == 0) err = (double)ad;                                                    // absoluteError.hpp:119
      else err = (double)(fabsf(o) < 1.0f ? ad : (float)((double)ad / (double)fabsf(o)));  // relativeError.hpp:131
    } else {
      const double o = (double)ov, a = (double)av;
      const double ad = fabs(o - a);
      err = (which == 0 || fabs(o) < 1.0) ? ad : ad / fabs(o);
    }
    q = err * rbin;
    const double k = rint(q);
    if (fabs(q - k) <= 1e-9 * (k + 1.0)) q = err / bin;  // absoluteError.hpp:119-121, relativeError.hpp:131-133
  }
  int pos;
  if (!(q >= 0.0)) { pos = 0; atomicAdd(bad, 1u); }
  else if (q >= 1025.0) { pos = 1023; atomicAdd(bad, 1u); }
  else {
    pos = (int)q;
    if (pos >= 1024) pos = pos - 1;
    if (pos >= 1024) { pos = 1023; atomicAdd(bad, 1u); }
  }
  return pos;
}

template <typename Scalar>
__global__ void __launch_bounds__(256) hist_kernel(int which, double lo, double hi, const Scalar* __restrict__ orig,
                                                   const Scalar* __restrict__ approx, size_t n,
                                                   unsigned long long* __restrict__ counts,
                                                   unsigned long long* __restrict__ oob)
{
  __shared__ unsigned int h[1024];
  __shared__ unsigned int bad;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) h[i] = 0;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  const double range = hi - lo;
  const double bin = range / 1024.0;
  const double rbin = 1.0 / bin, rrange = 1.0 / range;
  // a CTA handles at most 2^31 values between flushes, so 32-bit shared counters cannot overflow
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if constexpr (sizeof(Scalar) == 4) {
    // 16-byte loads, four values per thread and step (the loop is otherwise latency bound on 4-byte loads)
    const bool vec = ((reinterpret_cast<size_t>(approx) | (which == 2 ? 0 : reinterpret_cast<size_t>(orig))) & 15) == 0;
    if (vec) {
      const size_t n4 = n / 4;
      const float4* o4 = reinterpret_cast<const float4*>(orig);
      const float4* a4 = reinterpret_cast<const float4*>(approx);
      for (size_t i = tid; i < n4; i += stride) {
        const float4 a = __ldcs(a4 + i);
        float4 o = a;
        if (which != 2) o = __ldcs(o4 + i);
        const int p0 = hist_bin<float>(which, o.x, a.x, lo, range, bin, rbin, rrange, &bad);
        const int p1 = hist_bin<float>(which, o.y, a.y, lo, range, bin, rbin, rrange, &bad);
        const int p2 = hist_bin<float>(which, o.z, a.z, lo, range, bin, rbin, rrange, &bad);
        const int p3 = hist_bin<float>(which, o.w, a.w, lo, range, bin, rbin, rrange, &bad);
        atomicAdd(&h[p0], 1u); atomicAdd(&h[p1], 1u); atomicAdd(&h[p2], 1u); atomicAdd(&h[p3], 1u);
      }
      done = n4 * 4;
    }
  }
  for (size_t i = done + tid; i < n; i += stride) {
    const Scalar a = approx[i];
    const Scalar o = which == 2 ? a : orig[i];
    atomicAdd(&h[hist_bin<Scalar>(which, o, a, lo, range, bin, rbin, rrange, &bad)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x)
    if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
  if (threadIdx.x == 0 && bad && oob) atomicAdd(oob, (unsigned long long)bad);
}

}  // namespace zfb
