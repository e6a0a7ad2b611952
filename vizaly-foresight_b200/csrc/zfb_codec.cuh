// zfb_codec.cuh -- per-block fixed-rate zfp codec, one 4^d block per thread, everything in registers.
//
// This is the arithmetic that CBench reaches through zfp_compress / zfp_decompress
// (reference call sites: CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp:145 and :249; codec =
// LLNL zfp 0.5.5 fixed-rate mode, restated in SURVEY.md Appendix A).  It is written from scratch for
// sm_100a: all loops over block elements are fully unrolled so that every coefficient lives in a
// register, the sequency permutation is a compile-time renaming, bit planes are produced by a SWAR
// bit-matrix transpose (not per-bit extraction), and the embedded coder runs in O(planes + ones) with
// ffs/clz run-length arithmetic instead of one iteration per emitted bit.
//
// The functions are __host__ __device__ so that tests/host_emul.cpp can run the very same code on
// the CPU against the oracle; nothing on the product path ever executes them on the host.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define ZFB_HD __host__ __device__ __forceinline__
#define ZFB_UNROLL _Pragma("unroll")
#define ZFB_ROLLED _Pragma("unroll 1")
#else
#define ZFB_HD inline
#define ZFB_UNROLL
#define ZFB_ROLLED
#endif

namespace zfb {

// ---------------------------------------------------------------------------------------------
// small bit helpers
// ---------------------------------------------------------------------------------------------
ZFB_HD int ctz64(uint64_t x)  // x != 0
{
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}

ZFB_HD uint64_t lowmask64(unsigned m)  // m in [0, 64]
{
  return m >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << m) - 1);
}

ZFB_HD uint32_t f32_bits(float f)
{
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
ZFB_HD float bits_f32(uint32_t u)
{
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
ZFB_HD uint64_t f64_bits(double f)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(f);
#else
  union { double f; uint64_t u; } c; c.f = f; return c.u;
#endif
}
ZFB_HD double bits_f64(uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { double f; uint64_t u; } c; c.u = u; return c.f;
#endif
}

// ---------------------------------------------------------------------------------------------
// scalar traits (float -> int32, EBITS 8, EBIAS 127; double -> int64, 11, 1023)
// ---------------------------------------------------------------------------------------------
template <typename Scalar> struct traits;

template <> struct traits<float> {
  typedef int32_t Int;
  typedef uint32_t UInt;
  static constexpr int EBITS = 8, EBIAS = 127, PBITS = 32;
  static constexpr uint32_t NBMASK = 0xaaaaaaaau;
  // biased IEEE exponent field of a non-negative value
  static ZFB_HD int expfield(float m) { return (int)(f32_bits(m) >> 23) & 0xff; }
  // 2^e as Scalar, with ldexpf's overflow (inf) and gradual underflow behaviour
  static ZFB_HD float pow2(int e)
  {
    if (e > 127) return bits_f32(0x7f800000u);
    if (e >= -126) return bits_f32((uint32_t)(e + 127) << 23);
    if (e >= -149) return bits_f32((uint32_t)1 << (e + 149));
    return 0.0f;
  }
  static ZFB_HD Int trunc_cast(float p)  // C cast (Int)p for |p| < 2^31
  {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(p);
#else
    return (Int)p;
#endif
  }
  static ZFB_HD float to_scalar(Int i)
  {
#if defined(__CUDA_ARCH__)
    return __int2float_rn(i);
#else
    return (float)i;
#endif
  }
  static ZFB_HD Int int_min() { return (Int)0x80000000u; }
};

template <> struct traits<double> {
  typedef int64_t Int;
  typedef uint64_t UInt;
  static constexpr int EBITS = 11, EBIAS = 1023, PBITS = 64;
  static constexpr uint64_t NBMASK = 0xaaaaaaaaaaaaaaaaull;
  static ZFB_HD int expfield(double m) { return (int)(f64_bits(m) >> 52) & 0x7ff; }
  static ZFB_HD double pow2(int e)
  {
    if (e > 1023) return bits_f64(0x7ff0000000000000ull);
    if (e >= -1022) return bits_f64((uint64_t)(e + 1023) << 52);
    if (e >= -1074) return bits_f64((uint64_t)1 << (e + 1074));
    return 0.0;
  }
  static ZFB_HD Int trunc_cast(double p)
  {
#if defined(__CUDA_ARCH__)
    return __double2ll_rz(p);
#else
    return (Int)p;
#endif
  }
  static ZFB_HD double to_scalar(Int i)
  {
#if defined(__CUDA_ARCH__)
    return __ll2double_rn(i);
#else
    return (double)i;
#endif
  }
  static ZFB_HD Int int_min() { return (Int)0x8000000000000000ull; }
};

// ---------------------------------------------------------------------------------------------
// sequency order (A.7): coefficient i of the coded sequence is transform output perm<D>(i)
// ---------------------------------------------------------------------------------------------
template <int DIMS> struct order;
template <> struct order<1> {
  static ZFB_HD constexpr int at(int i) { return i; }
};
template <> struct order<2> {
  static ZFB_HD constexpr int at(int i)
  {
    constexpr unsigned char t[16] = {0, 1, 4, 5, 2, 8, 6, 9, 3, 12, 10, 7, 13, 11, 14, 15};
    return t[i];
  }
};
template <> struct order<3> {
  static ZFB_HD constexpr int at(int i)
  {
    constexpr unsigned char t[64] = {
        0,  1,  4,  16, 20, 17, 5,  2,  8,  32, 21, 6,  18, 24, 9,  33, 36, 3,  12, 48, 22, 25,
        37, 40, 34, 10, 7,  19, 28, 13, 49, 52, 41, 38, 26, 23, 29, 53, 11, 35, 44, 14, 50, 56,
        42, 27, 39, 45, 30, 54, 57, 60, 51, 15, 43, 46, 58, 61, 55, 31, 62, 59, 47, 63};
    return t[i];
  }
};

// ---------------------------------------------------------------------------------------------
// lifting (A.6).  Arithmetic is two's-complement wraparound: done on UInt for adds, Int for >>.
// ---------------------------------------------------------------------------------------------
template <typename Int> struct unsigned_of;
template <> struct unsigned_of<int32_t> { typedef uint32_t type; };
template <> struct unsigned_of<int64_t> { typedef uint64_t type; };

// On the device the 32-bit adds and subtractions of the FORWARD lifting steps are written as multiply-adds with a factor of
// one that the compiler cannot see through (a __constant__ operand): IMAD runs on the FMA pipe, IADD3 / LEA on the ALU
// pipe, and the ALU pipe -- shifts, logic, compares, min/max of the coder and the bit transposes -- is what bounds the
// float kernels (ncu: ALU 70-74 % busy, FMA 17-25 %).  ptxas balances the two pipes per basic block, not per kernel.
#if defined(__CUDACC__)
__constant__ int32_t zfb_k_one = 1;
#endif
#ifndef ZFB_FMA_ADDS
#define ZFB_FMA_ADDS 1
#endif
ZFB_HD uint32_t lift_add(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__) && ZFB_FMA_ADDS
  return (uint32_t)((int32_t)a * zfb_k_one + (int32_t)b);
#else
  return a + b;
#endif
}
ZFB_HD uint32_t lift_sub(uint32_t a, uint32_t b)  // a - b
{
#if defined(__CUDA_ARCH__) && ZFB_FMA_ADDS
  return (uint32_t)((int32_t)a * zfb_k_one - (int32_t)b);
#else
  return a - b;
#endif
}
ZFB_HD uint64_t lift_add(uint64_t a, uint64_t b) { return a + b; }
ZFB_HD uint64_t lift_sub(uint64_t a, uint64_t b) { return a - b; }

template <typename Int> ZFB_HD void fwd_lift(Int& x_, Int& y_, Int& z_, Int& w_)
{
  typedef typename unsigned_of<Int>::type U;
  U x = (U)x_, y = (U)y_, z = (U)z_, w = (U)w_;
#define ZFB_SAR1(v) ((U)((Int)(v) >> 1))
  x = lift_add(x, w); x = ZFB_SAR1(x); w = lift_sub(w, x);
  z = lift_add(z, y); z = ZFB_SAR1(z); y = lift_sub(y, z);
  x = lift_add(x, z); x = ZFB_SAR1(x); z = lift_sub(z, x);
  w = lift_add(w, y); w = ZFB_SAR1(w); y = lift_sub(y, w);
  w = lift_add(w, ZFB_SAR1(y)); y = lift_sub(y, ZFB_SAR1(w));
  x_ = (Int)x; y_ = (Int)y; z_ = (Int)z; w_ = (Int)w;
}

template <typename Int> ZFB_HD void inv_lift(Int& x_, Int& y_, Int& z_, Int& w_)
{
  typedef typename unsigned_of<Int>::type U;
  U x = (U)x_, y = (U)y_, z = (U)z_, w = (U)w_;
  // plain adds here: the decode kernels are not helped by the IMAD form (measured: 2.76 -> 2.87 ms on the slab)
  y += ZFB_SAR1(w); w -= ZFB_SAR1(y);
  y += w; w <<= 1; w -= y;
  z += x; x <<= 1; x -= z;
  y += z; z <<= 1; z -= y;
  w += x; x <<= 1; x -= w;
#undef ZFB_SAR1
  x_ = (Int)x; y_ = (Int)y; z_ = (Int)z; w_ = (Int)w;
}

template <typename Int, int DIMS> ZFB_HD void fwd_xform(Int (&p)[1 << (2 * DIMS)])
{
  if constexpr (DIMS == 1) {
    fwd_lift(p[0], p[1], p[2], p[3]);
  } else if constexpr (DIMS == 2) {
    ZFB_UNROLL for (int y = 0; y < 4; y++) fwd_lift(p[4 * y], p[4 * y + 1], p[4 * y + 2], p[4 * y + 3]);
    ZFB_UNROLL for (int x = 0; x < 4; x++) fwd_lift(p[x], p[x + 4], p[x + 8], p[x + 12]);
  } else {
    ZFB_UNROLL for (int r = 0; r < 16; r++) fwd_lift(p[4 * r], p[4 * r + 1], p[4 * r + 2], p[4 * r + 3]);
    ZFB_UNROLL for (int z = 0; z < 4; z++)
      ZFB_UNROLL for (int x = 0; x < 4; x++)
        fwd_lift(p[16 * z + x], p[16 * z + x + 4], p[16 * z + x + 8], p[16 * z + x + 12]);
    ZFB_UNROLL for (int q = 0; q < 16; q++) fwd_lift(p[q], p[q + 16], p[q + 32], p[q + 48]);
  }
}

template <typename Int, int DIMS> ZFB_HD void inv_xform(Int (&p)[1 << (2 * DIMS)])
{
  if constexpr (DIMS == 1) {
    inv_lift(p[0], p[1], p[2], p[3]);
  } else if constexpr (DIMS == 2) {
    ZFB_UNROLL for (int x = 0; x < 4; x++) inv_lift(p[x], p[x + 4], p[x + 8], p[x + 12]);
    ZFB_UNROLL for (int y = 0; y < 4; y++) inv_lift(p[4 * y], p[4 * y + 1], p[4 * y + 2], p[4 * y + 3]);
  } else {
    ZFB_UNROLL for (int q = 0; q < 16; q++) inv_lift(p[q], p[q + 16], p[q + 32], p[q + 48]);
    ZFB_UNROLL for (int z = 0; z < 4; z++)
      ZFB_UNROLL for (int x = 0; x < 4; x++)
        inv_lift(p[16 * z + x], p[16 * z + x + 4], p[16 * z + x + 8], p[16 * z + x + 12]);
    ZFB_UNROLL for (int r = 0; r < 16; r++) inv_lift(p[4 * r], p[4 * r + 1], p[4 * r + 2], p[4 * r + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// bit-matrix transpose, LSB convention: after the call word k bit i == (before) word i bit k.
// N words of N bits each, log2(N) butterfly stages, fully unrolled -> pure register code.
// It is an involution, so the same routine turns coefficients into planes and planes back.
// ---------------------------------------------------------------------------------------------
ZFB_HD void transpose32(uint32_t (&a)[32])
{
  ZFB_UNROLL for (int k = 0; k < 16; k++) {
    uint32_t t = ((a[k] >> 16) ^ a[k + 16]) & 0x0000ffffu;
    a[k] ^= t << 16; a[k + 16] ^= t;
  }
  ZFB_UNROLL for (int g = 0; g < 32; g += 16)
    ZFB_UNROLL for (int k = 0; k < 8; k++) {
      uint32_t t = ((a[g + k] >> 8) ^ a[g + k + 8]) & 0x00ff00ffu;
      a[g + k] ^= t << 8; a[g + k + 8] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 32; g += 8)
    ZFB_UNROLL for (int k = 0; k < 4; k++) {
      uint32_t t = ((a[g + k] >> 4) ^ a[g + k + 4]) & 0x0f0f0f0fu;
      a[g + k] ^= t << 4; a[g + k + 4] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 32; g += 4)
    ZFB_UNROLL for (int k = 0; k < 2; k++) {
      uint32_t t = ((a[g + k] >> 2) ^ a[g + k + 2]) & 0x33333333u;
      a[g + k] ^= t << 2; a[g + k + 2] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 32; g += 2) {
    uint32_t t = ((a[g] >> 1) ^ a[g + 1]) & 0x55555555u;
    a[g] ^= t << 1; a[g + 1] ^= t;
  }
}

ZFB_HD void transpose64(uint64_t (&a)[64])
{
  ZFB_UNROLL for (int k = 0; k < 32; k++) {
    uint64_t t = ((a[k] >> 32) ^ a[k + 32]) & 0x00000000ffffffffull;
    a[k] ^= t << 32; a[k + 32] ^= t;
  }
  ZFB_UNROLL for (int g = 0; g < 64; g += 32)
    ZFB_UNROLL for (int k = 0; k < 16; k++) {
      uint64_t t = ((a[g + k] >> 16) ^ a[g + k + 16]) & 0x0000ffff0000ffffull;
      a[g + k] ^= t << 16; a[g + k + 16] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 64; g += 16)
    ZFB_UNROLL for (int k = 0; k < 8; k++) {
      uint64_t t = ((a[g + k] >> 8) ^ a[g + k + 8]) & 0x00ff00ff00ff00ffull;
      a[g + k] ^= t << 8; a[g + k + 8] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 64; g += 8)
    ZFB_UNROLL for (int k = 0; k < 4; k++) {
      uint64_t t = ((a[g + k] >> 4) ^ a[g + k + 4]) & 0x0f0f0f0f0f0f0f0full;
      a[g + k] ^= t << 4; a[g + k + 4] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 64; g += 4)
    ZFB_UNROLL for (int k = 0; k < 2; k++) {
      uint64_t t = ((a[g + k] >> 2) ^ a[g + k + 2]) & 0x3333333333333333ull;
      a[g + k] ^= t << 2; a[g + k + 2] ^= t;
    }
  ZFB_UNROLL for (int g = 0; g < 64; g += 2) {
    uint64_t t = ((a[g] >> 1) ^ a[g + 1]) & 0x5555555555555555ull;
    a[g] ^= t << 1; a[g + 1] ^= t;
  }
}

// Plane container: P.get(k) is the 64-bit word whose bit i is bit k of coefficient i (sequency order).
template <typename UInt, int BS> struct planes;

// float, 64 coefficients: two 32x32 transposes (coefficients 0..31 and 32..63)
template <> struct planes<uint32_t, 64> {
  uint32_t lo[32], hi[32];
  ZFB_HD void from_coeffs(const uint32_t (&u)[64])
  {
    ZFB_UNROLL for (int i = 0; i < 32; i++) { lo[i] = u[i]; hi[i] = u[i + 32]; }
    transpose32(lo); transpose32(hi);
  }
  ZFB_HD void to_coeffs(uint32_t (&u)[64])
  {
    transpose32(lo); transpose32(hi);
    ZFB_UNROLL for (int i = 0; i < 32; i++) { u[i] = lo[i]; u[i + 32] = hi[i]; }
  }
  ZFB_HD uint64_t get(int k) const { return (uint64_t)lo[k] | ((uint64_t)hi[k] << 32); }
  ZFB_HD void set(int k, uint64_t x) { lo[k] = (uint32_t)x; hi[k] = (uint32_t)(x >> 32); }
  ZFB_HD void clear() { ZFB_UNROLL for (int i = 0; i < 32; i++) { lo[i] = 0; hi[i] = 0; } }
};

// float, 16 coefficients
template <> struct planes<uint32_t, 16> {
  uint32_t a[32];
  ZFB_HD void from_coeffs(const uint32_t (&u)[16])
  {
    ZFB_UNROLL for (int i = 0; i < 16; i++) { a[i] = u[i]; a[i + 16] = 0; }
    transpose32(a);
  }
  ZFB_HD void to_coeffs(uint32_t (&u)[16])
  {
    transpose32(a);
    ZFB_UNROLL for (int i = 0; i < 16; i++) u[i] = a[i];
  }
  ZFB_HD uint64_t get(int k) const { return a[k]; }
  ZFB_HD void set(int k, uint64_t x) { a[k] = (uint32_t)x; }
  ZFB_HD void clear() { ZFB_UNROLL for (int i = 0; i < 32; i++) a[i] = 0; }
};

// 4 coefficients (1D): planes are nibbles; extract / deposit directly
template <typename UInt> struct planes<UInt, 4> {
  UInt c[4];
  ZFB_HD void from_coeffs(const UInt (&u)[4]) { ZFB_UNROLL for (int i = 0; i < 4; i++) c[i] = u[i]; }
  ZFB_HD void to_coeffs(UInt (&u)[4]) { ZFB_UNROLL for (int i = 0; i < 4; i++) u[i] = c[i]; }
  ZFB_HD uint64_t get(int k) const
  {
    return (uint64_t)(((c[0] >> k) & 1u) | (((c[1] >> k) & 1u) << 1) | (((c[2] >> k) & 1u) << 2) |
                      (((c[3] >> k) & 1u) << 3));
  }
  ZFB_HD void set(int k, uint64_t x)
  {
    ZFB_UNROLL for (int i = 0; i < 4; i++) c[i] |= (UInt)((x >> i) & 1u) << k;
  }
  ZFB_HD void clear() { ZFB_UNROLL for (int i = 0; i < 4; i++) c[i] = 0; }
};

// double, 64 coefficients: one 64x64 transpose
template <> struct planes<uint64_t, 64> {
  uint64_t a[64];
  ZFB_HD void from_coeffs(const uint64_t (&u)[64])
  {
    ZFB_UNROLL for (int i = 0; i < 64; i++) a[i] = u[i];
    transpose64(a);
  }
  ZFB_HD void to_coeffs(uint64_t (&u)[64])
  {
    transpose64(a);
    ZFB_UNROLL for (int i = 0; i < 64; i++) u[i] = a[i];
  }
  ZFB_HD uint64_t get(int k) const { return a[k]; }
  ZFB_HD void set(int k, uint64_t x) { a[k] = x; }
  ZFB_HD void clear() { ZFB_UNROLL for (int i = 0; i < 64; i++) a[i] = 0; }
};

// double, 16 coefficients
template <> struct planes<uint64_t, 16> {
  uint64_t a[64];
  ZFB_HD void from_coeffs(const uint64_t (&u)[16])
  {
    ZFB_UNROLL for (int i = 0; i < 64; i++) a[i] = i < 16 ? u[i & 15] : 0;
    transpose64(a);
  }
  ZFB_HD void to_coeffs(uint64_t (&u)[16])
  {
    transpose64(a);
    ZFB_UNROLL for (int i = 0; i < 16; i++) u[i] = a[i];
  }
  ZFB_HD uint64_t get(int k) const { return a[k]; }
  ZFB_HD void set(int k, uint64_t x) { a[k] = x; }
  ZFB_HD void clear() { ZFB_UNROLL for (int i = 0; i < 64; i++) a[i] = 0; }
};

// ---------------------------------------------------------------------------------------------
// bit sinks / sources.  A "slot" is one block's maxbits-long bit string (A.9: 64-bit little-endian
// words, LSB first).
// ---------------------------------------------------------------------------------------------

// Slot that starts on a word boundary and is a whole number of words (CBench's wra != 0 case,
// zfpCompressorGpu.hpp:129): plain stores, every word of the slot written exactly once.
struct aligned_sink {
  uint64_t* dst;
  unsigned widx, nwords;
  ZFB_HD aligned_sink(uint64_t* d, unsigned n) : dst(d), widx(0), nwords(n) {}
  ZFB_HD void word(uint64_t w)
  {
    if (widx < nwords) dst[widx] = w;
    widx++;
  }
  ZFB_HD void finish()
  {
    for (; widx < nwords; widx++) dst[widx] = 0;
  }
  ZFB_HD void pad_to(unsigned) {}  // finish() already zero-filled the slot
};

// Slot at an arbitrary bit offset of a pre-zeroed stream (maxbits not a multiple of 64, i.e. the
// wra == 0 extension): neighbouring blocks share words, so words are OR-ed in atomically.
struct or_sink {
  uint64_t* base;     // stream word 0
  size_t startbit;    // absolute bit position of this slot
  unsigned maxbits, j;
  ZFB_HD or_sink(uint64_t* b, size_t sb, unsigned mb) : base(b), startbit(sb), maxbits(mb), j(0) {}
  ZFB_HD static void or64(uint64_t* p, uint64_t v)
  {
#if defined(__CUDA_ARCH__)
    atomicOr((unsigned long long*)p, (unsigned long long)v);
#else
    *p |= v;
#endif
  }
  ZFB_HD void word(uint64_t w)
  {
    const unsigned off = 64u * j++;
    if (off >= maxbits) return;
    const unsigned valid = maxbits - off;
    if (valid < 64) w &= lowmask64(valid);
    if (!w) return;
    const size_t pos = startbit + off;
    const unsigned s = (unsigned)(pos & 63);
    or64(base + (pos >> 6), w << s);
    if (s && (w >> (64 - s))) or64(base + (pos >> 6) + 1, w >> (64 - s));
  }
  ZFB_HD void finish() {}
  ZFB_HD void pad_to(unsigned) {}  // the stream is pre-zeroed
};

// Accumulates emitted bit strings into 64-bit words and hands completed words to the sink.
template <class Sink> struct bit_writer {
  Sink& sink;
  uint64_t acc;
  unsigned cnt;     // bits currently held in acc, < 64
  unsigned total;   // bits emitted so far
  ZFB_HD explicit bit_writer(Sink& s) : sink(s), acc(0), cnt(0), total(0) {}
  // append the `len` (0..64) low bits of v; bits of v above len must be zero
  ZFB_HD void put(uint64_t v, unsigned len)
  {
    acc |= v << cnt;
    unsigned t = cnt + len;
    if (t >= 64) {
      sink.word(acc);
      acc = cnt ? (v >> (64 - cnt)) : 0;
      t -= 64;
    }
    cnt = t;
    total += len;
  }
  ZFB_HD void flush()
  {
    if (cnt) sink.word(acc);
    acc = 0; cnt = 0;
  }
};

// Reads a slot given a word pointer and a starting bit offset (< 64).  `avail` = number of words that
// may legally be dereferenced from `w` (guards the look-ahead word at the end of the stream).
struct slot_reader {
  const uint64_t* w;
  unsigned avail;
  unsigned pos;
  ZFB_HD slot_reader(const uint64_t* p, unsigned avail_words, unsigned startbit) : w(p), avail(avail_words), pos(startbit) {}
  ZFB_HD uint64_t peek() const
  {
    unsigned i = pos >> 6, s = pos & 63;
    uint64_t a = i < avail ? w[i] : 0;
    if (!s) return a;
    uint64_t b = (i + 1 < avail) ? w[i + 1] : 0;
    return (a >> s) | (b << (64 - s));
  }
};

// ---------------------------------------------------------------------------------------------
// zfp_stream's four parameters.  Fixed rate: minbits = maxbits, maxprec = 64, minexp = -1074
// (zfp_stream_set_rate); fixed accuracy / precision (the CPU zfp plugin's "abs" / "rel",
// zfp/zfpCompressor.hpp:117-120): minbits = 1, maxbits = 16657 and a per-block plane count
// precision(emax) = min(maxprec, max(0, emax - minexp + 2 (dims + 1))).
// ---------------------------------------------------------------------------------------------
struct stream_params {
  unsigned minbits, maxbits, maxprec;
  int minexp;
};

ZFB_HD stream_params fixed_rate_params(unsigned maxbits)
{
  stream_params sp;
  sp.minbits = maxbits; sp.maxbits = maxbits; sp.maxprec = 64u; sp.minexp = -1074;
  return sp;
}

template <int DIMS> ZFB_HD int block_precision(int emax, const stream_params& sp)
{
  int p = emax - sp.minexp + 2 * (DIMS + 1);
  p = p < 0 ? 0 : p;
  return p > (int)sp.maxprec ? (int)sp.maxprec : p;
}

// Sink of the variable-rate path: whole words into a private staging area that is large enough for the longest
// possible block (1 + EBITS + 4^d - 1 + 4^d * PBITS bits, zfp_stream_maximum_size's bound).
struct stage_sink {
  uint64_t* dst;
  unsigned widx;
  ZFB_HD explicit stage_sink(uint64_t* d) : dst(d), widx(0) {}
  ZFB_HD void word(uint64_t w) { dst[widx++] = w; }
  ZFB_HD void finish() {}
  ZFB_HD void pad_to(unsigned bits) { while (64u * widx < bits) word(0); }  // minbits padding is part of the block
};

// ---------------------------------------------------------------------------------------------
// encode one block (A.4, A.5, A.8).  Emits exactly maxbits bits into the sink (as whole words; the
// sink drops anything past the slot and zero-fills the rest).  The fixed-rate stream is the first
// maxbits bits of the unbounded embedded stream, so the coder runs without per-bit budget checks and
// stops as soon as the budget is covered.
// ---------------------------------------------------------------------------------------------
// General form: returns the number of bits the block occupies (>= minbits, <= maxbits).  With minbits ==
// maxbits the sink pads / truncates to the slot; otherwise the caller places the `bits` emitted bits.
template <typename Scalar, int DIMS, class Sink>
ZFB_HD unsigned encode_block(const Scalar (&f)[1 << (2 * DIMS)], const stream_params& sp, Sink& sink)
{
  const unsigned maxbits = sp.maxbits;
  typedef traits<Scalar> T;
  typedef typename T::Int Int;
  typedef typename T::UInt UInt;
  constexpr int BS = 1 << (2 * DIMS);

  // block maximum; NaNs lose every comparison exactly as in `if (max < f) max = f`
  Scalar mx = 0;
  ZFB_UNROLL for (int i = 0; i < BS; i++) {
    Scalar a = f[i] < 0 ? -f[i] : f[i];
    mx = (mx < a) ? a : mx;
  }
  const int E = T::expfield(mx);               // emax = E - (EBIAS - 1); biased e = E + 1
  const int emax = E - (T::EBIAS - 1);
  const int prec = block_precision<DIMS>(emax, sp);  // > 0 for any finite non-zero input in fixed-rate mode
  if (!(mx > 0) || prec == 0) {  // all-zero block, or nothing above the accuracy floor: a single 0 bit and padding
    sink.word(0);
    sink.finish();
    sink.pad_to(sp.minbits);
    return sp.minbits > 1u ? sp.minbits : 1u;
  }
  const int kmin = T::PBITS > prec ? T::PBITS - prec : 0;

  // block-floating-point cast: i = (Int)(2^(p-2-emax) * f), truncation toward zero
  Int ib[BS];
  {
    const int se = (T::PBITS - 2) - emax;
    const Scalar s = T::pow2(se);
    ZFB_UNROLL for (int i = 0; i < BS; i++) ib[i] = T::trunc_cast(s * f[i]);
    if (se > T::EBIAS) {
      // scale overflowed to +inf (block max < 2^-97 for float): the reference's cast of inf/NaN is
      // undefined; x86 (cvttss2si) yields the integer-indefinite value for every element.
      ZFB_UNROLL for (int i = 0; i < BS; i++) ib[i] = T::int_min();
    }
  }
  fwd_xform<Int, DIMS>(ib);

  // negabinary + sequency order (compile-time renaming)
  UInt u[BS];
  ZFB_UNROLL for (int i = 0; i < BS; i++) u[i] = ((UInt)ib[order<DIMS>::at(i)] + T::NBMASK) ^ T::NBMASK;

  planes<UInt, BS> P;
  P.from_coeffs(u);

  bit_writer<Sink> bw(sink);
  bw.put(2 * (uint64_t)(E + 1) + 1, 1 + T::EBITS);

  unsigned n = 0;
  // The plane loop stays rolled: fully unrolled it is hundreds of KB of SASS per kernel and the kernels stall
  // on instruction fetch.  P is then indexed dynamically (thread-local memory, L1 resident).
  ZFB_ROLLED for (int k = T::PBITS - 1; k >= 0; k--) {
    if (bw.total >= maxbits || k < kmin) break;
    uint64_t x = P.get(k);
    // first n coefficients are already significant: their bits go out verbatim
    bw.put(x & lowmask64(n), n);
    x = n < 64 ? x >> n : 0;
    // group tests + unary run lengths for the rest of the plane
    while (n < (unsigned)BS && bw.total < maxbits) {
      if (!x) { bw.put(0, 1); break; }
      const unsigned z = (unsigned)ctz64(x);              // zeros before the next one; n + z <= BS-1
      const unsigned one = (n + z < (unsigned)(BS - 1));  // the last coefficient's one is implicit
      bw.put((uint64_t)1 | ((uint64_t)one << (z + 1)), z + 1 + one);
      x = (z + 1 < 64) ? x >> (z + 1) : 0;
      n += z + 1;
    }
  }
  bw.flush();
  sink.finish();
  const unsigned used = bw.total < maxbits ? bw.total : maxbits;
  sink.pad_to(sp.minbits);
  return used < sp.minbits ? sp.minbits : used;
}

// Append form: the block's bits (exactly the returned count, padding and truncation included) go to a bit writer
// the caller owns, behind whatever it already holds.  Used where one thread strings several small blocks
// together (1D variable rate: a block is 16 bytes of data, so per-block bookkeeping would dominate).
template <typename Scalar, int DIMS, class Writer>
ZFB_HD unsigned encode_block_append(const Scalar (&f)[1 << (2 * DIMS)], const stream_params& sp, Writer& bw)
{
  typedef traits<Scalar> T;
  typedef typename T::Int Int;
  typedef typename T::UInt UInt;
  constexpr int BS = 1 << (2 * DIMS);
  const unsigned maxbits = sp.maxbits;
  unsigned used = 0;
  auto put = [&](uint64_t v, unsigned len) {  // never beyond the block's budget
    const unsigned room = maxbits - used;
    if (len > room) { v &= lowmask64(room); len = room; }
    bw.put(v, len);
    used += len;
  };
  auto pad = [&]() {
    while (used < sp.minbits) {
      const unsigned l = sp.minbits - used < 64u ? sp.minbits - used : 64u;
      bw.put(0, l);
      used += l;
    }
  };

  Scalar mx = 0;
  ZFB_UNROLL for (int i = 0; i < BS; i++) {
    Scalar a = f[i] < 0 ? -f[i] : f[i];
    mx = (mx < a) ? a : mx;
  }
  const int E = T::expfield(mx);
  const int emax = E - (T::EBIAS - 1);
  const int prec = block_precision<DIMS>(emax, sp);
  if (!(mx > 0) || prec == 0) {
    put(0, 1);
    pad();
    return used;
  }
  const int kmin = T::PBITS > prec ? T::PBITS - prec : 0;
  Int ib[BS];
  {
    const int se = (T::PBITS - 2) - emax;
    const Scalar s = T::pow2(se);
    ZFB_UNROLL for (int i = 0; i < BS; i++) ib[i] = T::trunc_cast(s * f[i]);
    if (se > T::EBIAS) { ZFB_UNROLL for (int i = 0; i < BS; i++) ib[i] = T::int_min(); }  // see encode_block
  }
  fwd_xform<Int, DIMS>(ib);
  UInt u[BS];
  ZFB_UNROLL for (int i = 0; i < BS; i++) u[i] = ((UInt)ib[order<DIMS>::at(i)] + T::NBMASK) ^ T::NBMASK;
  planes<UInt, BS> P;
  P.from_coeffs(u);
  put(2 * (uint64_t)(E + 1) + 1, 1 + T::EBITS);
  unsigned n = 0;
  ZFB_ROLLED for (int k = T::PBITS - 1; k >= 0; k--) {
    if (used >= maxbits || k < kmin) break;
    uint64_t x = P.get(k);
    put(x & lowmask64(n), n);
    x = n < 64 ? x >> n : 0;
    while (n < (unsigned)BS && used < maxbits) {
      if (!x) { put(0, 1); break; }
      const unsigned z = (unsigned)ctz64(x);
      const unsigned one = (n + z < (unsigned)(BS - 1));
      put((uint64_t)1 | ((uint64_t)one << (z + 1)), z + 1 + one);
      x = (z + 1 < 64) ? x >> (z + 1) : 0;
      n += z + 1;
    }
  }
  pad();
  return used;
}

// Fixed-rate form: exactly maxbits bits per block.
template <typename Scalar, int DIMS, class Sink>
ZFB_HD void encode_block(const Scalar (&f)[1 << (2 * DIMS)], unsigned maxbits, Sink& sink)
{
  encode_block<Scalar, DIMS>(f, fixed_rate_params(maxbits), sink);
}

// ---------------------------------------------------------------------------------------------
// decode one block (A.5, A.8).  `r` is positioned at the first bit of the slot.
// Budget bookkeeping is exact: zfp's decoder deposits a one when the budget ends inside a zero run.
// ---------------------------------------------------------------------------------------------
// General form: returns the number of bits the block occupied.  PARSE_ONLY skips the reconstruction (used to
// rebuild the block index of a variable-rate stream).
template <typename Scalar, int DIMS, bool PARSE_ONLY, class Reader>
ZFB_HD unsigned decode_block(Scalar (&f)[1 << (2 * DIMS)], const stream_params& sp, Reader& r)
{
  const unsigned maxbits = sp.maxbits;
  typedef traits<Scalar> T;
  typedef typename T::Int Int;
  typedef typename T::UInt UInt;
  constexpr int BS = 1 << (2 * DIMS);

  uint64_t win = r.peek();
  if (!(win & 1)) {
    if (!PARSE_ONLY) { ZFB_UNROLL for (int i = 0; i < BS; i++) f[i] = 0; }
    return sp.minbits > 1u ? sp.minbits : 1u;
  }
  const int e = (int)((win >> 1) & (((uint64_t)1 << T::EBITS) - 1));
  const int emax = e - T::EBIAS;
  r.pos += 1 + T::EBITS;
  const int prec = block_precision<DIMS>(emax, sp);
  const int kmin = T::PBITS > prec ? T::PBITS - prec : 0;

  planes<UInt, BS> P;
  P.clear();
  unsigned remaining = maxbits - (1 + T::EBITS);
  unsigned n = 0;
  ZFB_UNROLL for (int k = T::PBITS - 1; k >= 0; k--) {  // unrolled: P stays in registers (a rolled loop puts it in local memory and is slower here)
    if (!remaining || k < kmin) break;
    const unsigned m = n < remaining ? n : remaining;
    remaining -= m;
    uint64_t x = r.peek() & lowmask64(m);
    r.pos += m;
    while (n < (unsigned)BS && remaining) {
      win = r.peek();
      remaining--; r.pos++;
      if (!(win & 1)) break;
      win >>= 1;
      const unsigned room = (unsigned)(BS - 1) - n;
      const unsigned maxrun = room < remaining ? room : remaining;
      const unsigned z = win ? (unsigned)ctz64(win) : 64u;
      if (z < maxrun) { r.pos += z + 1; remaining -= z + 1; n += z; }
      else            { r.pos += maxrun; remaining -= maxrun; n += maxrun; }
      x |= (uint64_t)1 << n;
      n++;
    }
    if (!PARSE_ONLY) P.set(k, x);
  }
  const unsigned used = maxbits - remaining;
  if (PARSE_ONLY) return used < sp.minbits ? sp.minbits : used;

  UInt u[BS];
  P.to_coeffs(u);
  Int ib[BS];
  ZFB_UNROLL for (int i = 0; i < BS; i++) ib[order<DIMS>::at(i)] = (Int)((u[i] ^ T::NBMASK) - T::NBMASK);
  inv_xform<Int, DIMS>(ib);
  const Scalar s = T::pow2(emax - (T::PBITS - 2));
  ZFB_UNROLL for (int i = 0; i < BS; i++) f[i] = s * T::to_scalar(ib[i]);
  return used < sp.minbits ? sp.minbits : used;
}

// Fixed-rate form.
template <typename Scalar, int DIMS, class Reader>
ZFB_HD void decode_block(Scalar (&f)[1 << (2 * DIMS)], unsigned maxbits, Reader& r)
{
  decode_block<Scalar, DIMS, false>(f, fixed_rate_params(maxbits), r);
}

// ---------------------------------------------------------------------------------------------
// partial-block padding (A.3): n valid values out of 4 at stride s
// ---------------------------------------------------------------------------------------------
template <typename Scalar> ZFB_HD void pad4(Scalar& p0, Scalar& p1, Scalar& p2, Scalar& p3, unsigned n)
{
  if (n < 1) p0 = 0;
  if (n < 2) p1 = p0;
  if (n < 3) p2 = p1;
  if (n < 4) p3 = p0;
}

template <typename Scalar, int DIMS>
ZFB_HD void pad_block(Scalar (&q)[1 << (2 * DIMS)], unsigned mx, unsigned my, unsigned mz)
{
  constexpr int BS = 1 << (2 * DIMS);
  // along x for every row; rows that are not present are overwritten by the y/z passes below,
  // so the result equals padding only the present rows (A.3)
  ZFB_UNROLL for (int r = 0; r < BS / 4; r++) pad4(q[4 * r], q[4 * r + 1], q[4 * r + 2], q[4 * r + 3], mx);
  if constexpr (DIMS == 2) {
    ZFB_UNROLL for (int x = 0; x < 4; x++) pad4(q[x], q[x + 4], q[x + 8], q[x + 12], my);
  }
  if constexpr (DIMS == 3) {
    ZFB_UNROLL for (int z = 0; z < 4; z++)
      ZFB_UNROLL for (int x = 0; x < 4; x++)
        pad4(q[16 * z + x], q[16 * z + x + 4], q[16 * z + x + 8], q[16 * z + x + 12], my);
    ZFB_UNROLL for (int i = 0; i < 16; i++) pad4(q[i], q[i + 16], q[i + 32], q[i + 48], mz);
  }
  (void)my; (void)mz;
}

}  // namespace zfb
