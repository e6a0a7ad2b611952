// zfb_var.cuh -- variable-rate streams: zfp's fixed-accuracy and fixed-precision modes, the ones the CPU zfp
// plugin exposes as "abs" and "rel" (CBench/compressors/zfp/zfpCompressor.hpp:83-93,117-120).
//
// A variable-rate stream is the bit-tight concatenation of blocks of different lengths, so the encoder is three
// steps instead of one:
//   1. enc_var_kernel      every thread encodes one block into a private staging column (words of the 32 blocks of
//                          a warp interleaved, so staging stores and loads are coalesced) and records its length
//   2. scan_*_kernel       exclusive prefix sum of the lengths -> absolute bit offset of every block (the index)
//   3. compact_var_kernel  every thread shifts its staged words to its bit offset in the (zeroed) stream: interior
//                          words are plain stores, the two boundary words shared with the neighbours are OR-ed in
// The decoder needs the index: dec_var_kernel takes the offsets array (kept by the caller from the encode, exactly
// as CBench's decompress relies on state left by compress, zfpCompressor.hpp:22,160), and index_var_kernel rebuilds
// it from a bare stream by walking the blocks in order (slow: the format has no resynchronisation points).
#pragma once
#include "zfb_kernels.cuh"

namespace zfb {

template <typename Scalar, int DIMS> struct var_caps {
  static constexpr unsigned BS = 1u << (2 * DIMS);
  // zfp_stream_maximum_size's per-block bound: 1 + EBITS + (4^d - 1) + 4^d * precision
  static constexpr unsigned BITS = 1u + traits<Scalar>::EBITS + (BS - 1u) + BS * (unsigned)traits<Scalar>::PBITS;
  static constexpr unsigned WORDS = (BITS + 63u) / 64u + 1u;
};

// 1D: a block is four values, so one thread strings VAR1_GROUP consecutive blocks together and the index has one
// entry per group (per-block lengths, offsets and boundary atomics would cost more than the data itself)
constexpr unsigned VAR1_GROUP = 8;
template <typename Scalar> struct var1_caps {
  static constexpr unsigned WORDS = (VAR1_GROUP * var_caps<Scalar, 1>::BITS + 63u) / 64u + 1u;
};

// staging column of one block: word j of block b lives at stage[(b / 32) * 32 * WORDS + j * 32 + b % 32]
// (a minbits floor above the longest codable block is padding only: the column holds `cap` words, the compaction
// copies at most that many and the rest of the block stays zero in the pre-zeroed stream)
struct stage_col_sink {
  uint64_t* dst;
  unsigned widx, cap;
  ZFB_HD stage_col_sink(uint64_t* d, unsigned cap_words) : dst(d), widx(0), cap(cap_words) {}
  ZFB_HD void word(uint64_t w)
  {
    if (widx < cap) dst[32u * widx] = w;
    widx++;
  }
  ZFB_HD void finish() {}
  ZFB_HD void pad_to(unsigned bits) { while (64u * widx < bits && widx < cap) word(0); }
};

// the same column seen as 32-bit words (fast float/3D coder): 32-bit word j at stage32[2 * base + j * 32]
struct stage_col_sink32 {
  uint32_t* dst;
  unsigned widx, cap;
  ZFB_HD stage_col_sink32(uint32_t* d, unsigned cap_words32) : dst(d), widx(0), cap(cap_words32) {}
  ZFB_HD void word(uint32_t w) { dst[32u * widx++] = w; }
  ZFB_HD void word_if(bool on, uint32_t w)
  {
    if (on) dst[32u * widx] = w;
    widx += on ? 1u : 0u;
  }
  ZFB_HD void word2(uint32_t w0, uint32_t w1) { word(w0); word(w1); }
  ZFB_HD void finish() {}
  ZFB_HD unsigned words() const { return widx; }
  ZFB_HD void pad_to(unsigned bits) { while (32u * widx < bits && widx < cap) word(0u); }
};

template <unsigned WORDS> __device__ __forceinline__ size_t stage_base(size_t b)
{
  return (b >> 5) * (size_t)(32u * WORDS) + (b & 31u);
}

// block b of the field -> 4^d values, partial blocks padded (A.3)
template <typename Scalar, int DIMS>
__device__ __forceinline__ void gather_block(const Scalar* __restrict__ in, const geom& g, size_t b,
                                             Scalar (&f)[1 << (2 * DIMS)])
{
  constexpr int BS = 1 << (2 * DIMS);
  const size_t ix = b % g.bx, iy = (b / g.bx) % g.by, iz = b / (g.bx * g.by);
  const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
  const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
  const unsigned my = DIMS >= 2 ? (unsigned)(g.ny - y0 < 4 ? g.ny - y0 : 4) : 1u;
  const unsigned mz = DIMS >= 3 ? (unsigned)(g.nz - z0 < 4 ? g.nz - z0 : 4) : 1u;
  const bool full = mx == 4 && (DIMS < 2 || my == 4) && (DIMS < 3 || mz == 4);
  const bool vec = ((g.nx & 3) == 0) && ((reinterpret_cast<size_t>(in) & 15) == 0) && sizeof(Scalar) == 4;
  if (full && vec) {
#pragma unroll
    for (int k = 0; k < (DIMS >= 3 ? 4 : 1); k++)
#pragma unroll
      for (int j = 0; j < (DIMS >= 2 ? 4 : 1); j++) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + x0 + g.nx * ((y0 + j) + g.ny * (z0 + k))));
        f[(16 * k + 4 * j + 0) & (BS - 1)] = (Scalar)v.x;
        f[(16 * k + 4 * j + 1) & (BS - 1)] = (Scalar)v.y;
        f[(16 * k + 4 * j + 2) & (BS - 1)] = (Scalar)v.z;
        f[(16 * k + 4 * j + 3) & (BS - 1)] = (Scalar)v.w;
      }
    return;
  }
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

// ---- step 1: encode into staging, record lengths ----------------------------------------------
template <typename Scalar, int DIMS>
__global__ void __launch_bounds__(128) enc_var_kernel(const Scalar* __restrict__ in, uint64_t* __restrict__ stage,
                                                      uint32_t* __restrict__ lens, geom g, stream_params sp)
{
  constexpr int BS = 1 << (2 * DIMS);
  constexpr unsigned WORDS = var_caps<Scalar, DIMS>::WORDS;
  constexpr bool FAST = sizeof(Scalar) == 4 && DIMS == 3;  // table coder with the planes parked in shared memory
  __shared__ enc_coder_smem cs;
  __shared__ fast_scratch<FAST> fsc;
  f3::tables tb;
  if (FAST) tb = build_coder_tables(cs);
  const f3::plane_rows ps = fsc.rows();
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < g.nblocks; b += (size_t)gridDim.x * blockDim.x) {
    Scalar f[BS];
    gather_block<Scalar, DIMS>(in, g, b, f);
    if constexpr (FAST) {
      // 32-bit word j of the column: the two halves of 64-bit word j/2 sit 32 words apart, see compact_var_kernel
      stage_col_sink32 sink(reinterpret_cast<uint32_t*>(stage) + 2 * ((b >> 5) * (size_t)(32u * WORDS)) + (b & 31u), 2u * WORDS);
      lens[b] = f3::encode_block(f, sp, sink, tb, ps);
    } else {
      stage_col_sink sink(stage + stage_base<WORDS>(b), WORDS);
      lens[b] = encode_block<Scalar, DIMS>(f, sp, sink);
    }
  }
}

template <typename Scalar, int DIMS, bool METRICS>
__device__ __forceinline__ void scatter_block(const Scalar (&f)[1 << (2 * DIMS)], Scalar* __restrict__ out,
                                              const Scalar* __restrict__ orig, const geom& g, size_t x0, size_t y0,
                                              size_t z0, unsigned mx, unsigned my, unsigned mz, macc_t<Scalar>& m);

// 1D: one thread = one group of VAR1_GROUP blocks = one staging column = one index entry
template <typename Scalar>
__global__ void __launch_bounds__(128) enc_var1_kernel(const Scalar* __restrict__ in, uint64_t* __restrict__ stage,
                                                       uint32_t* __restrict__ lens, geom g, stream_params sp)
{
  constexpr unsigned WORDS = var1_caps<Scalar>::WORDS;
  constexpr bool F32 = sizeof(Scalar) == 4;
  __shared__ coder1_smem cs;
  f1::tables tb;
  if (F32) tb = build_coder1_tables<true, false>(cs);
  const bool fast = F32 && f1::var_params_ok(sp);  // per-plane table coder (accuracy / precision modes)
  const size_t ngroups = (g.nblocks + VAR1_GROUP - 1) / VAR1_GROUP;
  const bool vec = (reinterpret_cast<size_t>(in) & 15) == 0;
  for (size_t grp = (size_t)blockIdx.x * blockDim.x + threadIdx.x; grp < ngroups; grp += (size_t)gridDim.x * blockDim.x) {
    stage_col_sink sink(stage + stage_base<WORDS>(grp), WORDS);
    bit_writer<stage_col_sink> bw(sink);
    unsigned total = 0;
    for (unsigned q = 0; q < VAR1_GROUP; q++) {
      const size_t b = grp * VAR1_GROUP + q;
      if (b >= g.nblocks) break;
      const size_t x0 = 4 * b;
      Scalar f[4];
      if (x0 + 4 <= g.nx && vec) {
        if constexpr (sizeof(Scalar) == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(in + x0));
          f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        } else {
          const double2 v0 = __ldg(reinterpret_cast<const double2*>(in + x0));
          const double2 v1 = __ldg(reinterpret_cast<const double2*>(in + x0 + 2));
          f[0] = v0.x; f[1] = v0.y; f[2] = v1.x; f[3] = v1.y;
        }
      } else {
        const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
#pragma unroll
        for (int i = 0; i < 4; i++) f[i] = (unsigned)i < mx ? __ldg(in + x0 + i) : (Scalar)0;
        if (mx < 4) pad_block<Scalar, 1>(f, mx, 1u, 1u);
      }
      if constexpr (F32) {
        if (fast) { total += f1::encode_block_var(f, sp, bw, tb); continue; }
      }
      total += encode_block_append<Scalar, 1>(f, sp, bw);
    }
    bw.flush();
    lens[grp] = total;
  }
}

constexpr unsigned VAR1_STAGE_WORDS = 3072;  // 32-bit words of stream staged per CTA tile (128 groups)

template <typename Scalar, bool METRICS>
__global__ void __launch_bounds__(128) dec_var1_kernel(const uint64_t* __restrict__ stream, const uint64_t* __restrict__ offsets,
                                                       Scalar* __restrict__ out, const Scalar* __restrict__ orig, geom g,
                                                       stream_params sp, double* __restrict__ part)
{
  __shared__ double red[7 * 4];
  constexpr bool F32 = sizeof(Scalar) == 4;
  __shared__ coder1_smem cs;
  __shared__ uint32_t piece[F32 ? VAR1_STAGE_WORDS : 1];
  __shared__ uint64_t range[2];
  f1::tables tb;
  if (F32) tb = build_coder1_tables<false, true>(cs);
  const bool fast = F32 && f1::var_params_ok(sp);
  macc_t<Scalar> m;
  m.init();
  const size_t ngroups = (g.nblocks + VAR1_GROUP - 1) / VAR1_GROUP;
  const bool vec = ((reinterpret_cast<size_t>(out) & 15) == 0) && (!METRICS || (reinterpret_cast<size_t>(orig) & 15) == 0);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
  const size_t total32 = 2 * g.stream_words;
  const size_t ntiles = (ngroups + 127) / 128;
  // a CTA takes 128 consecutive groups: their bits are one contiguous piece of the stream, staged in shared memory
  // with coalesced loads when it fits (a thread reading its own words from global memory waits a full DRAM
  // latency at every 32 bits)
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t grp = tile * 128 + threadIdx.x;
    const bool active = grp < ngroups;
    uint64_t pos = active ? offsets[grp] : 0;
    bool staged = false;
    size_t v_begin = 0, v_end = 0;
    if (fast) {
      if (threadIdx.x == 0) range[0] = pos;
      if (active && (threadIdx.x == 127 || grp + 1 == ngroups)) range[1] = offsets[grp + 1];
      __syncthreads();
      v_begin = (size_t)(range[0] >> 5);
      v_end = (size_t)((range[1] + 31) >> 5) + 2;  // + the reader's two look-ahead words
      staged = v_end - v_begin <= VAR1_STAGE_WORDS;
      if (staged)
        for (size_t i = threadIdx.x; i < v_end - v_begin; i += 128) piece[i] = v_begin + i < total32 ? __ldg(s32 + v_begin + i) : 0u;
      __syncthreads();
    }
    if (active) {
      for (unsigned q = 0; q < VAR1_GROUP; q++) {
        const size_t b = grp * VAR1_GROUP + q;
        if (b >= g.nblocks) break;
        Scalar f[4];
        bool done = false;
        if constexpr (F32) {
          if (fast) {
            const size_t v0 = (size_t)(pos >> 5);
            const size_t left32 = total32 > v0 ? total32 - v0 : 0;
            const uint32_t* src = staged ? piece + (v0 - v_begin) : s32 + v0;
            const unsigned navail = staged ? (unsigned)(v_end - v0) : (left32 > 0xffffffffull ? 0xffffffffu : (unsigned)left32);
            f3::reader32 r(src, navail, (unsigned)(pos & 31));
            pos += f1::decode_block_var(f, sp, r, tb);
            done = true;
          }
        }
        if (!done) {
          const size_t w0 = (size_t)(pos >> 6);
          const size_t left = g.stream_words > w0 ? g.stream_words - w0 : 0;
          slot_reader r(stream + w0, left > 0xffffffffull ? 0xffffffffu : (unsigned)left, (unsigned)(pos & 63));
          pos += decode_block<Scalar, 1, false>(f, sp, r);
        }
        const size_t x0 = 4 * b;
        if (x0 + 4 <= g.nx && vec) {
          if constexpr (sizeof(Scalar) == 4) {
            float4 v;
            v.x = f[0]; v.y = f[1]; v.z = f[2]; v.w = f[3];
            if (METRICS) {
              frow fr;
              fr.init();
              fr.add4(__ldg(reinterpret_cast<const float4*>(orig + x0)), v.x, v.y, v.z, v.w);
              fr.flush(m);
            }
            *reinterpret_cast<float4*>(out + x0) = v;
          } else {
            if (METRICS) {
              const double2 o0 = __ldg(reinterpret_cast<const double2*>(orig + x0));
              const double2 o1 = __ldg(reinterpret_cast<const double2*>(orig + x0 + 2));
              macc_add(m, o0.x, (double)f[0]); macc_add(m, o0.y, (double)f[1]);
              macc_add(m, o1.x, (double)f[2]); macc_add(m, o1.y, (double)f[3]);
            }
            *reinterpret_cast<double2*>(out + x0) = make_double2((double)f[0], (double)f[1]);
            *reinterpret_cast<double2*>(out + x0 + 2) = make_double2((double)f[2], (double)f[3]);
          }
        } else {
          const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
          scatter_block<Scalar, 1, METRICS>(f, out, orig, g, x0, 0, 0, mx, 1u, 1u, m);
        }
      }
    }
    if (fast) __syncthreads();  // the piece and the range are rewritten by the next tile
  }
  if (METRICS) macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
}

// ---- step 2: exclusive scan of the lengths (three small kernels, fixed order, no atomics) ----------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;  // lengths per CTA

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* sh, uint64_t& total)
{
  // sh: 32 + 1 words of shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < nw ? sh[lane] : 0;
    uint64_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) sh[lane] = winc - w;
    if (lane == nw - 1) sh[32] = winc;
  }
  __syncthreads();
  const uint64_t r = sh[warp] + inc - v;
  total = sh[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(const uint32_t* __restrict__ lens, size_t n,
                                                                 uint64_t* __restrict__ chunk_sums)
{
  __shared__ uint64_t sh[33];
  const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++)
    if (base + i < n) s += lens[base + i];
  uint64_t total;
  block_exclusive_scan(s, sh, total);
  if (threadIdx.x == 0) chunk_sums[blockIdx.x] = total;
}

// one CTA: chunk sums -> exclusive chunk offsets (in place), grand total to chunk_sums[nchunks]
__global__ void __launch_bounds__(SCAN_THREADS) scan_chunks_kernel(uint64_t* __restrict__ chunk_sums, size_t nchunks)
{
  __shared__ uint64_t sh[33];
  uint64_t carry = 0;
  for (size_t c0 = 0; c0 < nchunks; c0 += SCAN_THREADS) {
    const size_t c = c0 + threadIdx.x;
    const uint64_t v = c < nchunks ? chunk_sums[c] : 0;
    uint64_t total;
    const uint64_t ex = block_exclusive_scan(v, sh, total);
    if (c < nchunks) chunk_sums[c] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) chunk_sums[nchunks] = carry;
}

// offsets[b] = first bit of block b, offsets[n] = total bits
// base (nullable): device scalar added to every offset -- the bits of the slabs encoded before this one.  It may be
// offsets[0] itself (the total the previous slab's scan left there): every thread then writes back what it read.
__global__ void __launch_bounds__(SCAN_THREADS) scan_offsets_kernel(const uint32_t* __restrict__ lens, size_t n,
                                                                    const uint64_t* __restrict__ chunk_sums,
                                                                    size_t nchunks, uint64_t* offsets,
                                                                    const uint64_t* base_ptr)
{
  __shared__ uint64_t sh[33];
  const uint64_t bit0 = base_ptr ? *base_ptr : 0;
  const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t l[SCAN_ITEMS];
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    l[i] = base + i < n ? lens[base + i] : 0u;
    s += l[i];
  }
  uint64_t total;
  uint64_t ex = bit0 + chunk_sums[blockIdx.x] + block_exclusive_scan(s, sh, total);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    if (base + i < n) offsets[base + i] = ex;
    ex += l[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = bit0 + chunk_sums[nchunks];
}

// zero the words the compaction will OR into: ceil(total_bits / 64) words, total read from the device
// (a later slab starts behind the word that holds the previous slab's last bits: that word is already in use)
__global__ void __launch_bounds__(256) zero_stream_kernel(uint64_t* __restrict__ stream, const uint64_t* base_bits,
                                                          const uint64_t* __restrict__ total_bits, size_t cap_words)
{
  const size_t first = base_bits ? (size_t)((*base_bits + 63) >> 6) : 0;
  size_t nw = (size_t)((*total_bits + 63) >> 6);
  if (nw > cap_words) nw = cap_words;
  for (size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (size_t)gridDim.x * blockDim.x) stream[i] = 0;
}

// ---- step 3: staged words -> bit-tight stream ---------------------------------------------------
template <unsigned WORDS, bool COL32>
__global__ void __launch_bounds__(128) compact_var_kernel(const uint64_t* __restrict__ stage, const uint64_t* __restrict__ offsets,
                                                          size_t nblocks, uint64_t* __restrict__ stream, size_t cap_words)
{
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks; b += (size_t)gridDim.x * blockDim.x) {
    const uint64_t off = offsets[b], end = offsets[b + 1];
    const unsigned len = (unsigned)(end - off);
    const unsigned nw_all = (len + 63u) >> 6;
    const unsigned nw = nw_all < WORDS ? nw_all : WORDS;  // beyond the column there is only minbits padding (zeros)
    const uint64_t* col = stage + stage_base<WORDS>(b);
    const uint32_t* col32 = reinterpret_cast<const uint32_t*>(stage) + 2 * ((b >> 5) * (size_t)(32u * WORDS)) + (b & 31u);
    const unsigned s = (unsigned)(off & 63u);
    const size_t wi = (size_t)(off >> 6);
    uint64_t prev = 0;  // bits carried into the next output word
    // four staged words per step: the loads of a step are independent, so their latencies overlap
    for (unsigned j0 = 0; j0 <= nw; j0 += 4) {
      uint64_t w[4];
#pragma unroll
      for (unsigned q = 0; q < 4; q++) {
        const unsigned j = j0 + q;
        w[q] = 0;
        if (j < nw) {
          if (COL32) w[q] = (uint64_t)col32[64u * j] | ((uint64_t)col32[64u * j + 32u] << 32);
          else w[q] = col[32u * j];
        }
      }
#pragma unroll
      for (unsigned q = 0; q < 4; q++) {
        const unsigned j = j0 + q;
        if (j > nw) break;
        uint64_t v = w[q];
        if (j < nw) {
          const unsigned valid = len - 64u * j;
          if (valid < 64u) v &= lowmask64(valid);
        }
        const uint64_t o = s ? ((v << s) | prev) : v;
        prev = s ? (v >> (64u - s)) : 0;
        const size_t i = wi + j;
        if (i < cap_words && o) {  // a zero contribution needs no write: the words were zeroed up front
          // words wholly inside [off, end) belong to this block alone; the first and the last are shared
          const bool whole = (uint64_t)i * 64u >= off && (uint64_t)(i + 1) * 64u <= end;
          if (whole) stream[i] = o;
          else atomicOr(reinterpret_cast<unsigned long long*>(stream + i), (unsigned long long)o);
        }
      }
    }
  }
}

// ---- decode with the index ---------------------------------------------------------------------
// scatter one decoded block (partial blocks: only the values inside the field), metrics fused
template <typename Scalar, int DIMS, bool METRICS>
__device__ __forceinline__ void scatter_block(const Scalar (&f)[1 << (2 * DIMS)], Scalar* __restrict__ out,
                                              const Scalar* __restrict__ orig, const geom& g, size_t x0, size_t y0,
                                              size_t z0, unsigned mx, unsigned my, unsigned mz, macc_t<Scalar>& m)
{
  constexpr int BS = 1 << (2 * DIMS);
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

// any type / dims: thread per block, the block's bits read straight from the stream
template <typename Scalar, int DIMS, bool METRICS>
__global__ void __launch_bounds__(128) dec_var_kernel(const uint64_t* __restrict__ stream, const uint64_t* __restrict__ offsets,
                                                      Scalar* __restrict__ out, const Scalar* __restrict__ orig, geom g,
                                                      stream_params sp, double* __restrict__ part)
{
  constexpr int BS = 1 << (2 * DIMS);
  __shared__ double red[7 * 4];
  macc_t<Scalar> m;
  m.init();
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < g.nblocks; b += (size_t)gridDim.x * blockDim.x) {
    const size_t ix = b % g.bx, iy = (b / g.bx) % g.by, iz = b / (g.bx * g.by);
    const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
    const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
    const unsigned my = DIMS >= 2 ? (unsigned)(g.ny - y0 < 4 ? g.ny - y0 : 4) : 1u;
    const unsigned mz = DIMS >= 3 ? (unsigned)(g.nz - z0 < 4 ? g.nz - z0 : 4) : 1u;
    const uint64_t startbit = offsets[b];
    const size_t w0 = (size_t)(startbit >> 6);
    const size_t left = g.stream_words > w0 ? g.stream_words - w0 : 0;
    Scalar f[BS];
    slot_reader r(stream + w0, left > 0xffffffffull ? 0xffffffffu : (unsigned)left, (unsigned)(startbit & 63));
    decode_block<Scalar, DIMS, false>(f, sp, r);
    scatter_block<Scalar, DIMS, METRICS>(f, out, orig, g, x0, y0, z0, mx, my, mz, m);
  }
  if (METRICS) macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
}

// float / 3D: a CTA takes 128 consecutive blocks, whose bits are one contiguous piece of the stream: the piece is
// staged in shared memory with coalesced loads (when it fits: VAR_STAGE_WORDS covers ~12 bits per value on average)
// and parsed from there by the table coder; whole blocks go out as 16-byte rows with the metrics fused.
constexpr unsigned VAR_STAGE_WORDS = 3072;  // 32-bit words; static shared memory stays under 48 KiB

template <bool METRICS>
__global__ void __launch_bounds__(128, 4) dec_var_f32_kernel(const uint64_t* __restrict__ stream,
                                                             const uint64_t* __restrict__ offsets, float* __restrict__ out,
                                                             const float* __restrict__ orig, geom g, stream_params sp,
                                                             double* __restrict__ part)
{
  __shared__ double red[7 * 4];
  __shared__ dec_coder_smem cs;
  __shared__ fast_scratch<true> fsc;
  __shared__ uint32_t piece[VAR_STAGE_WORDS];
  __shared__ uint64_t range[2];
  const f3::tables tb = build_coder_tables(cs);
  const f3::plane_rows ps = fsc.rows();
  maccf m;
  m.init();
  const bool vec = ((g.nx & 3) == 0) && ((reinterpret_cast<size_t>(out) & 15) == 0) &&
                   (!METRICS || (reinterpret_cast<size_t>(orig) & 15) == 0);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
  const size_t total32 = 2 * g.stream_words;
  const size_t ntiles = (g.nblocks + 127) / 128;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t b = tile * 128 + threadIdx.x;
    const bool active = b < g.nblocks;
    const uint64_t startbit = active ? offsets[b] : 0;
    if (threadIdx.x == 0) range[0] = startbit;
    if (active && (threadIdx.x == 127 || b + 1 == g.nblocks)) range[1] = offsets[b + 1];
    __syncthreads();
    const size_t v_begin = (size_t)(range[0] >> 5);
    const size_t v_end = (size_t)((range[1] + 31) >> 5) + 2;  // + the reader's two look-ahead words
    const bool staged = v_end - v_begin <= VAR_STAGE_WORDS;
    if (staged) {
      for (size_t i = threadIdx.x; i < v_end - v_begin; i += 128) piece[i] = v_begin + i < total32 ? __ldg(s32 + v_begin + i) : 0u;
    }
    __syncthreads();
    if (active) {
      const size_t ix = b % g.bx, iy = (b / g.bx) % g.by, iz = b / (g.bx * g.by);
      const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
      const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
      const unsigned my = (unsigned)(g.ny - y0 < 4 ? g.ny - y0 : 4);
      const unsigned mz = (unsigned)(g.nz - z0 < 4 ? g.nz - z0 : 4);
      const size_t v0 = (size_t)(startbit >> 5);
      float f[64];
      {
        // one call site (generic loads): the staged piece, or the stream itself when the piece was too long
        const size_t left32 = total32 > v0 ? total32 - v0 : 0;
        const uint32_t* src = staged ? piece + (v0 - v_begin) : s32 + v0;
        const unsigned navail = staged ? (unsigned)(v_end - v0) : (left32 > 0xffffffffull ? 0xffffffffu : (unsigned)left32);
        f3::reader32 r(src, navail, (unsigned)(startbit & 31));
        f3::decode_block(f, sp, r, tb, ps);
      }
      if (vec && my == 4 && mz == 4) {  // whole block, 16-byte rows: vector stores, four values per metric step
        frow fr;
        fr.init();
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const size_t idx = x0 + g.nx * ((y0 + j) + g.ny * (z0 + k));
            const int q = 16 * k + 4 * j;
            float4 v;
            v.x = f[q]; v.y = f[q + 1]; v.z = f[q + 2]; v.w = f[q + 3];
            if (METRICS) fr.add4(__ldg(reinterpret_cast<const float4*>(orig + idx)), v.x, v.y, v.z, v.w);
            __stcs(reinterpret_cast<float4*>(out + idx), v);
          }
          if (METRICS) fr.flush(m);
        }
      } else {
        scatter_block<float, 3, METRICS>(f, out, orig, g, x0, y0, z0, mx, my, mz, m);
      }
    }
    __syncthreads();  // the piece and the range are rewritten by the next tile
  }
  if (METRICS) macc_block_store(m, red, part + (size_t)blockIdx.x * PART_STRIDE);
}

// double / 3D: the table coder with 64 planes parked in shared memory (32 rows x 16 bytes per block, so 64 threads
// per CTA).  Decode leaves the metrics to the streaming pass (zfb_api_var.cuh), as the fixed-rate double path does.
constexpr int VAR64_THREADS = 64;

__global__ void __launch_bounds__(VAR64_THREADS) enc_var_f64_kernel(const double* __restrict__ in, uint64_t* __restrict__ stage,
                                                                    uint32_t* __restrict__ lens, geom g, stream_params sp)
{
  constexpr unsigned WORDS = var_caps<double, 3>::WORDS;
  __shared__ enc_coder_smem cs;
  __shared__ f3::u32x4 rows[32 * VAR64_THREADS];
  const f3::tables tb = build_coder_tables(cs);
  f3::plane_rows ps;
  ps.p = rows + threadIdx.x;
  ps.stride = VAR64_THREADS;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < g.nblocks; b += (size_t)gridDim.x * blockDim.x) {
    double f[64];
    gather_block<double, 3>(in, g, b, f);
    stage_col_sink32 sink(reinterpret_cast<uint32_t*>(stage) + 2 * ((b >> 5) * (size_t)(32u * WORDS)) + (b & 31u), 2u * WORDS);
    lens[b] = d3::encode_block(f, sp, sink, tb, ps);
  }
}

__global__ void __launch_bounds__(VAR64_THREADS) dec_var_f64_kernel(const uint64_t* __restrict__ stream,
                                                                    const uint64_t* __restrict__ offsets, double* __restrict__ out,
                                                                    geom g, stream_params sp)
{
  __shared__ dec_coder_smem cs;
  __shared__ f3::u32x4 rows[32 * VAR64_THREADS];
  const f3::tables tb = build_coder_tables(cs);
  f3::plane_rows ps;
  ps.p = rows + threadIdx.x;
  ps.stride = VAR64_THREADS;
  macc m;  // unused (METRICS = false), scatter_block's signature
  m.init();
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
  const size_t total32 = 2 * g.stream_words;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < g.nblocks; b += (size_t)gridDim.x * blockDim.x) {
    const size_t ix = b % g.bx, iy = (b / g.bx) % g.by, iz = b / (g.bx * g.by);
    const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
    const unsigned mx = (unsigned)(g.nx - x0 < 4 ? g.nx - x0 : 4);
    const unsigned my = (unsigned)(g.ny - y0 < 4 ? g.ny - y0 : 4);
    const unsigned mz = (unsigned)(g.nz - z0 < 4 ? g.nz - z0 : 4);
    const uint64_t startbit = offsets[b];
    const size_t v0 = (size_t)(startbit >> 5);
    const size_t left32 = total32 > v0 ? total32 - v0 : 0;
    f3::reader32 r(s32 + v0, left32 > 0xffffffffull ? 0xffffffffu : (unsigned)left32, (unsigned)(startbit & 31));
    double f[64];
    d3::decode_block(f, sp, r, tb, ps);
    scatter_block<double, 3, false>(f, out, nullptr, g, x0, y0, z0, mx, my, mz, m);
  }
}

// ---- index rebuild: walk a bare stream block by block (one thread; the format is strictly sequential) ----
template <typename Scalar, int DIMS>
__global__ void index_var_kernel(const uint64_t* __restrict__ stream, size_t stream_words, size_t nblocks, stream_params sp,
                                 uint64_t* __restrict__ offsets)
{
  if (blockIdx.x || threadIdx.x) return;
  constexpr int BS = 1 << (2 * DIMS);
  constexpr size_t UNIT = DIMS == 1 ? VAR1_GROUP : 1;  // blocks per index entry
  uint64_t pos = 0;
  Scalar f[BS];
  for (size_t b = 0; b < nblocks; b++) {
    if (b % UNIT == 0) offsets[b / UNIT] = pos;
    const size_t w0 = (size_t)(pos >> 6);
    const size_t left = stream_words > w0 ? stream_words - w0 : 0;
    slot_reader r(stream + w0, left > 0xffffffffull ? 0xffffffffu : (unsigned)left, (unsigned)(pos & 63));
    pos += decode_block<Scalar, DIMS, true>(f, sp, r);
  }
  offsets[(nblocks + UNIT - 1) / UNIT] = pos;
}

// ---- index expansion: a coarse index (every `stride`-th entry) -> the full index, segments walked in parallel ----
// Each thread walks one segment with the parse-only decoder; `bad` is raised if a walk does not end where the next
// coarse entry says it should (wrong parameters, wrong stream, corrupt bits).
template <typename Scalar, int DIMS>
__global__ void __launch_bounds__(64) index_expand_kernel(const uint64_t* __restrict__ stream, size_t stream_words, size_t nblocks,
                                                          stream_params sp, const uint64_t* __restrict__ coarse, size_t stride,
                                                          uint64_t* __restrict__ offsets, unsigned* __restrict__ bad)
{
  constexpr int BS = 1 << (2 * DIMS);
  constexpr size_t UNIT = DIMS == 1 ? VAR1_GROUP : 1;
  const size_t units = (nblocks + UNIT - 1) / UNIT;
  const size_t nseg = (units + stride - 1) / stride;
  Scalar f[BS];
  for (size_t seg = (size_t)blockIdx.x * blockDim.x + threadIdx.x; seg < nseg; seg += (size_t)gridDim.x * blockDim.x) {
    uint64_t pos = coarse[seg];
    const size_t u_end = (seg + 1) * stride < units ? (seg + 1) * stride : units;
    for (size_t u = seg * stride; u < u_end; u++) {
      offsets[u] = pos;
      for (size_t q = 0; q < UNIT; q++) {
        const size_t b = u * UNIT + q;
        if (b >= nblocks) break;
        const size_t w0 = (size_t)(pos >> 6);
        const size_t left = stream_words > w0 ? stream_words - w0 : 0;
        slot_reader r(stream + w0, left > 0xffffffffull ? 0xffffffffu : (unsigned)left, (unsigned)(pos & 63));
        pos += decode_block<Scalar, DIMS, true>(f, sp, r);
      }
    }
    if (pos != coarse[seg + 1]) atomicOr(bad, 1u);
    if (seg + 1 == nseg) offsets[units] = pos;
  }
}

}  // namespace zfb
