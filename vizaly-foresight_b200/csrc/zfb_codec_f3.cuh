// zfb_codec_f3.cuh -- the hot instantiation of the block codec: float32, 3D (64 coefficients, 32 planes).
//
// Same bit stream as encode_block<float,3> / decode_block<float,3> in zfb_codec.cuh (and therefore as
// zfp 0.5.5 fixed-rate; reference call sites zfpCompressorGpu.hpp:145,249),
// but organised for SIMT execution with one block per lane:
//
//   * no per-bit and no per-coefficient loops in the embedded coder.  zfp's group-test / unary code for
//     the not-yet-significant part of a bit plane is the substitution  0 -> "0", 1 -> "1 t"  (t = "more
//     ones follow") behind a leading test bit; the encoder expands a byte at a time through a 256-entry
//     table, the decoder parses a byte at a time through a 2x256-entry finite-state table.  Work per
//     plane is ceil(region/8) table steps instead of one divergent iteration per coefficient.
//   * all bit I/O in 32-bit words with funnel shifts (no emulated 64-bit shifts); the decoder keeps a
//     64-bit register window over the slot.
//   * bit planes come from a 32x32 SWAR transpose whose 16- and 8-bit stages are byte permutes (PRMT);
//     the halves holding planes 31..16 and 15..0 are finished lazily, so blocks whose budget ends in the
//     upper planes (the common case at <= 8 bits/value) never pay for the lower ones.
//
// Host-device like zfb_codec.cuh so tests/host_emul.cpp can bit-compare it with the oracle on the CPU.
#pragma once
#include "zfb_codec.cuh"

#if defined(__CUDACC__)
#define ZFB_NOUNROLL _Pragma("unroll 1")
#else
#define ZFB_NOUNROLL
#endif

namespace zfb {
namespace f3 {

// Plane scratch of one block: 16 rows of four words; row r = { plane 2r of coefficients 0..31, plane 2r of
// coefficients 32..63, plane 2r+1 (0..31), plane 2r+1 (32..63) }.  In the kernels the rows live in the
// thread's own 16-byte columns of the shared-memory tile (row r of lane l at 16 B * (r * stride + l)), so
// every access is a conflict-free 128-bit LDS/STS.  Parking the planes in memory is what lets the coder
// run as a real loop over planes: fully unrolled it is > 200 KB of SASS and the kernel stalls on
// instruction fetch (profiles/r01_notes.md).
struct alignas(16) u32x4 {
  uint32_t x, y, z, w;
};

// Row q lives at p[(q >> SH) * stride + (q & (2^SH - 1))]: SH = 0 for 16-byte columns (float tiles), SH = 1 for
// the 32-byte columns of a double tile (two consecutive rows share one 32-byte row segment).
template <int SH> struct plane_rows_t {
  u32x4* p;
  unsigned stride;
  ZFB_HD unsigned at(int r) const { return SH ? ((unsigned)r >> SH) * stride + ((unsigned)r & ((1u << SH) - 1u)) : (unsigned)r * stride; }
  ZFB_HD u32x4 ld(int r) const { return p[at(r)]; }
  ZFB_HD void st(int r, const u32x4& v) const { p[at(r)] = v; }
  // one plane: plane k is the (k & 1) half of row k >> 1
  ZFB_HD void ld64(int k, uint32_t& lo, uint32_t& hi) const
  {
#if defined(__CUDA_ARCH__)
    const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(p + at(k >> 1)) + 2 * (k & 1));
    lo = v.x; hi = v.y;
#else
    const u32x4& v = p[at(k >> 1)];
    lo = (k & 1) ? v.z : v.x; hi = (k & 1) ? v.w : v.y;
#endif
  }
  ZFB_HD void st64(int k, uint32_t lo, uint32_t hi) const
  {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(p + at(k >> 1)) + 2 * (k & 1)) = make_uint2(lo, hi);
#else
    u32x4& v = p[at(k >> 1)];
    if (k & 1) { v.z = lo; v.w = hi; } else { v.x = lo; v.y = hi; }
#endif
  }
};
typedef plane_rows_t<0> plane_rows;

#if defined(__CUDACC__)
// The same rows addressed through a 32-bit shared-memory address (tile kernels): a generic pointer costs 64-bit address
// arithmetic and a generic-to-shared conversion on every access, ~14 instructions per plane instead of 3.
template <int SH> struct smem_rows_t {
  uint32_t base;    // shared address of row 0 of this thread's column
  unsigned stride;  // bytes between consecutive row groups (rows q and q + 2^SH ... )
  __device__ __forceinline__ uint32_t at(int r) const
  {
    return SH ? base + ((unsigned)r >> SH) * stride + ((unsigned)r & ((1u << SH) - 1u)) * 16u : base + (unsigned)r * stride;
  }
  __device__ __forceinline__ u32x4 ld(int r) const
  {
    u32x4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(at(r)));
    return v;
  }
  __device__ __forceinline__ void st(int r, const u32x4& v) const
  {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(at(r)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  __device__ __forceinline__ void ld64(int k, uint32_t& lo, uint32_t& hi) const
  {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(at(k >> 1) + 8u * (unsigned)(k & 1)));
  }
  __device__ __forceinline__ void st64(int k, uint32_t lo, uint32_t hi) const
  {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(at(k >> 1) + 8u * (unsigned)(k & 1)), "r"(lo), "r"(hi) : "memory");
  }
};
#endif

// ---------------------------------------------------------------------------------------------
// small 32-bit helpers
// ---------------------------------------------------------------------------------------------
ZFB_HD uint32_t mask32(unsigned m)  // m in [0, 32]
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(0xffffffffu, 0u, 32u - m);  // (0:ffffffff) >> (32 - m), shift clamped to 32
#else
  return m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
#endif
}
ZFB_HD uint32_t fsr(uint32_t lo, uint32_t hi, unsigned s)  // low word of (hi:lo) >> s, s in [0, 32]
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, s);
#else
  return s == 0 ? lo : s >= 32 ? hi : (lo >> s) | (hi << (32 - s));
#endif
}
ZFB_HD uint32_t carry_out(uint32_t v, unsigned s)  // bits of v that fall off the top of a word when shifted left by s < 32
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(v, 0u, s);
#else
  return s ? v >> (32 - s) : 0u;
#endif
}
ZFB_HD uint32_t shl_pair(uint32_t lo, uint32_t hi, unsigned s)  // high word of (hi:lo) << s, s in [0, 32)
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
ZFB_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
ZFB_HD int msb32(uint32_t v)  // v != 0
{
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}
ZFB_HD uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m)  // (a & ~m) | (b & m) as one LOP3
{
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
  return d;
#else
  return (a & ~m) | (b & m);
#endif
}
// Other three-input functions of (a, b, m) as one LOP3 (device) -- used to fold the negabinary XOR into the last
// transpose stage.  LUT = f(0xF0, 0xCC, 0xAA).
template <int LUT> ZFB_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(m), "n"(LUT));
  return d;
#else
  uint32_t d = 0;
  for (int i = 0; i < 32; i++) {
    const unsigned idx = (((a >> i) & 1u) << 2) | (((b >> i) & 1u) << 1) | ((m >> i) & 1u);
    d |= (uint32_t)((LUT >> idx) & 1) << i;
  }
  return d;
#endif
}
ZFB_HD uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  const uint64_t t = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((t >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}

// ---------------------------------------------------------------------------------------------
// coder tables (built once per CTA into shared memory; 512 B + 2 KiB)
// ---------------------------------------------------------------------------------------------
// ENC[b], b in 0..255: byte b of the not yet significant part of a plane (LSB = first coefficient) expanded by
// 0 -> "0", 1 -> "11" (a one is followed by the group test "more ones follow"); 8 + popc(b) code bits.
// ENC[256 + b]: the same for the byte that holds the plane's LAST one: the code stops behind that one and the group
// test that follows it is 0; msb(b) + 1 + popc(b) code bits (ENC[256] = nothing at all: a byte behind the last one).
// Entry = code | length << 16.
ZFB_HD uint32_t enc_entry(unsigned i)
{
  const unsigned b = i & 255u;
  const bool last = i >= 256u;
  unsigned e = 0, pos = 0;
  int top = -1;
  for (int j = 0; j < 8; j++)
    if ((b >> j) & 1u) top = j;
  const int nbits = last ? top + 1 : 8;
  for (int j = 0; j < nbits; j++) {
    if ((b >> j) & 1u) {
      e |= ((last && j == top) ? 1u : 3u) << pos;
      pos += 2;
    } else pos += 1;
  }
  return e | (pos << 16) | ((unsigned)(top + 1) << 24);  // bits 24-27: coefficients up to and including the last one
}

// Near-end tables: the last row = 64 - n <= 4 coefficients of a plane (n >= 60) are coded in at most 7 bits, so one
// lookup does the whole group-test part of such a plane, end rules included (the one of coefficient 63 is implicit).
// NEE[row * 16 + y] (encoder, y = the row's coefficient bits): code | length << 8 | coefficients advanced << 12.
// NE[row * 128 + v] (decoder, v = next 7 stream bits): bits consumed | coefficients advanced << 4 | ones << 8.
// Row 0 (n == 64) is all zero: nothing is coded.  Both follow zfp's loops literally.
constexpr unsigned NEE_SIZE = 5 * 16, NE_SIZE = 5 * 128;
ZFB_HD uint32_t nee_entry(unsigned i)
{
  const unsigned row = i >> 4;
  unsigned x = i & 15u, code = 0, len = 0, dn = 0;
  if (row > 4u || x >> row) return 0u;
  while (dn < row) {
    const unsigned t = x ? 1u : 0u;
    code |= t << len; len++;
    if (!t) break;
    while (dn < row - 1u) {
      const unsigned b = x & 1u;
      code |= b << len; len++;
      if (b) break;
      x >>= 1; dn++;
    }
    x >>= 1; dn++;
  }
  return code | (len << 8) | (dn << 12);
}
ZFB_HD uint32_t ne_entry(unsigned i)
{
  const unsigned row = i >> 7, v = i & 127u;
  unsigned nbits = 0, dn = 0, pat = 0;
  if (row > 4u) return 0u;
  while (dn < row) {
    const unsigned t = (v >> nbits) & 1u;
    nbits++;
    if (!t) break;
    while (dn < row - 1u) {
      const unsigned b = (v >> nbits) & 1u;
      nbits++;
      if (b) break;
      dn++;
    }
    pat |= 1u << dn; dn++;
  }
  return nbits | (dn << 4) | (pat << 8);
}

// DEC[state][v]: run zfp's group-test parser over the 8 stream bits v (LSB first) from `state`
//   state 0 (A) = expecting a test bit, state 1 (B) = inside a zero run (next bit is a coefficient bit).
// Entry, laid out for the parse loop's critical path (table load -> window shift -> next table load):
//   bits 0-4   stream bits consumed: the entry itself is the shift amount of wrap-mode funnel shifts
//   bits 8-15  ones deposited (relative to the current coefficient index), bits 16-19 coefficients advanced
//   bits 30-31 next state (0 A, 1 B, 2 END = a zero test bit was consumed): e >> 20 is the byte offset of the
//              next state's half of the table, bit 31 alone says END
ZFB_HD uint32_t byte_of(uint32_t v, int i)  // byte i of v, zero extended (one PRMT on the device)
{
#if defined(__CUDA_ARCH__)
  return __byte_perm(v, 0u, 0x4440u + (unsigned)i);
#else
  return (v >> (8 * i)) & 0xffu;
#endif
}
ZFB_HD unsigned dec_nbits(uint32_t e) { return e & 31u; }
ZFB_HD uint32_t dec_pat(uint32_t e) { return byte_of(e, 1); }
ZFB_HD unsigned dec_npos(uint32_t e) { return byte_of(e, 2); }  // bits 20-23 are zero
ZFB_HD unsigned dec_state(uint32_t e) { return e >> 30; }
ZFB_HD uint32_t dec_entry(unsigned state, unsigned v)
{
  unsigned pattern = 0, npos = 0, nbits = 0;
  while (nbits < 8 && state != 2) {
    const unsigned bit = (v >> nbits) & 1u;
    nbits++;
    if (state == 0) state = bit ? 1u : 2u;
    else {
      if (bit) { pattern |= 1u << npos; state = 0; }
      npos++;
    }
  }
  return nbits | (pattern << 8) | (npos << 16) | (state << 30);
}

// Table of the flat parser (parse_planes_flat): zfp's group-test parser run over exactly b <= 8 stream bits.
//   FLAT[state * 512 + (1 << b) + v], state 0 (A) = expecting a test bit, 1 (B) = inside a zero run, v = the next b bits.
//   A step parses fewer than 8 bits at the end of the budget and where more bits could carry the coefficient index past
//   63 (whose one is implicit); b = 0 (entry 1) parses nothing.
// Entry: bits 0-4 stream bits consumed, bit 5 END (a zero test bit was consumed: the plane is done), bit 11 next state
// is B (= the byte offset of B's half of the table), bits 16-23 ones deposited (relative to the current coefficient),
// bits 24-31 coefficients advanced.
constexpr unsigned FLAT_WORDS = 1024, FLAT_B = 0x800u, FLAT_END = 0x20u;
ZFB_HD uint32_t flat_entry(unsigned i)
{
  unsigned state = i >> 9;
  const unsigned j = i & 511u;
  if (!j) return 0u;
  unsigned nb = 0;
  while ((2u << nb) <= j) nb++;  // j = (1 << nb) + v
  const unsigned v = j - (1u << nb);
  unsigned pattern = 0, npos = 0, nbits = 0;
  while (nbits < nb && state != 2) {
    const unsigned bit = (v >> nbits) & 1u;
    nbits++;
    if (state == 0) state = bit ? 1u : 2u;
    else {
      if (bit) { pattern |= 1u << npos; state = 0; }
      npos++;
    }
  }
  return nbits | (state == 2 ? FLAT_END : 0u) | (state == 1 ? FLAT_B : 0u) | (pattern << 16) | (npos << 24);
}
constexpr unsigned ENC_WORDS = 512 + NEE_SIZE, DEC_WORDS = 512 + NE_SIZE / 2;  // table sizes in 32-bit words (NE: 16-bit entries)
ZFB_HD uint32_t enc_table_word(unsigned i) { return i < 512u ? enc_entry(i) : nee_entry(i - 512u); }
ZFB_HD uint32_t dec_entry(unsigned state, unsigned v);
ZFB_HD uint32_t dec_table_word(unsigned i)
{
  return i < 512u ? dec_entry(i >> 8, i & 255u) : ne_entry(2u * (i - 512u)) | (ne_entry(2u * (i - 512u) + 1u) << 16);
}

struct tables {
  const uint32_t* enc;  // ENC_WORDS entries: ENC, then NEE
  const uint32_t* dec;  // DEC_WORDS entries: DEC, then NE
  uint32_t enc_s, dec_s;  // device: 32-bit shared-memory addresses of the two tables
  const uint32_t* flat = nullptr;  // FLAT_WORDS entries (flat parser only)
  uint32_t flat_s = 0;
};
// FLAT entry j of the state whose byte offset is st (0 or FLAT_B)
ZFB_HD uint32_t flat_at(const tables& tb, uint32_t st, uint32_t j)
{
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tb.flat_s + st + 4u * j));
  return v;
#else
  return tb.flat[(st >> 2) + j];
#endif
}
ZFB_HD uint32_t step_index(uint32_t c, unsigned b)  // (1 << b) | the b lowest bits of c, b <= 8
{
#if defined(__CUDA_ARCH__)
  uint32_t lo, one;
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(lo) : "r"(b));
  asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(one) : "r"(b));
  return (c & lo) | one;
#else
  return (1u << b) | (c & ((1u << b) - 1u));
#endif
}
ZFB_HD uint32_t lowmask(unsigned m)  // the m lowest bits, all of them for m >= 32
{
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(v) : "r"(m));
  return v;
#else
  return m >= 32u ? 0xffffffffu : (1u << m) - 1u;
#endif
}
// Table lookups.  On the device the tables live in shared memory and are read with ld.shared through a 32-bit
// address: going through the generic pointer makes the compiler rebuild the shared window base
// (S2UR SR_CgaCtaId -> ULEA) inside the coder loops, on their critical path.
ZFB_HD uint32_t enc_at(const tables& tb, uint32_t i)
{
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tb.enc_s + 4u * i));
  return v;
#else
  return tb.enc[i];
#endif
}
// DEC entry at a byte address inside the table (device: shared-memory address; host: offset from tb.dec)
ZFB_HD uint32_t dec_load(const tables& tb, uint32_t addr)
{
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
#else
  return tb.dec[(addr - tb.dec_s) >> 2];
#endif
}
ZFB_HD uint32_t ne_at(const tables& tb, uint32_t i)  // NE entry i (16-bit entries behind the 512 DEC words)
{
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(tb.dec_s + 2048u + 2u * i));
  return v;
#else
  return (tb.dec[512u + (i >> 1)] >> (16u * (i & 1u))) & 0xffffu;
#endif
}
ZFB_HD uint32_t dec_at(const tables& tb, uint32_t i)
{
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tb.dec_s + 4u * i));
  return v;
#else
  return tb.dec[i];
#endif
}

// ---------------------------------------------------------------------------------------------
// bit planes of 64 x uint32 coefficients, finished lazily in two halves
// ---------------------------------------------------------------------------------------------
// a[] / b[] hold coefficients 0..31 / 32..63.  After split16(): words 16..31 of each array carry the
// upper 16 bits of all 32 coefficients, words 0..15 the lower 16 bits.  finish(16) / finish(0) complete
// the transpose of that half; afterwards word k of a/b is plane k of coefficients 0..31 / 32..63.
struct planes64 {
  uint32_t a[32], b[32];

  ZFB_HD static void stage16(uint32_t (&w)[32])
  {
    ZFB_UNROLL for (int k = 0; k < 16; k++) {
      const uint32_t lo = bperm(w[k], w[k + 16], 0x5410), hi = bperm(w[k], w[k + 16], 0x7632);
      w[k] = lo; w[k + 16] = hi;
    }
  }
  // stages 8, 4, 2, 1 on words [base, base+16).
  // NEG folds the negabinary XOR of float coefficients (mask 0xaaaaaaaa = the odd bits) into stage 1, whose mask it
  // shares: flipping the odd bits of every coefficient is complementing the odd PLANES, and stage 1 is where words of
  // odd and even index meet.  NEG = 1 (encoder: coefficients -> planes) complements the odd words it produces;
  // NEG = 2 (decoder: planes -> coefficients) complements the odd words it consumes.  One LOP3 either way.
  template <int NEG = 0> ZFB_HD static void stages_low(uint32_t (&w)[32], int base)
  {
    ZFB_UNROLL for (int k = 0; k < 8; k++) {
      const uint32_t p = w[base + k], q = w[base + k + 8];
      w[base + k] = bperm(p, q, 0x6240); w[base + k + 8] = bperm(p, q, 0x7351);
    }
    ZFB_UNROLL for (int g = 0; g < 16; g += 8)
      ZFB_UNROLL for (int k = 0; k < 4; k++) {
        const uint32_t p = w[base + g + k], q = w[base + g + k + 4];
        w[base + g + k] = bitsel(p, q << 4, 0xf0f0f0f0u);
        w[base + g + k + 4] = bitsel(q, p >> 4, 0x0f0f0f0fu);
      }
    ZFB_UNROLL for (int g = 0; g < 16; g += 4)
      ZFB_UNROLL for (int k = 0; k < 2; k++) {
        const uint32_t p = w[base + g + k], q = w[base + g + k + 2];
        w[base + g + k] = bitsel(p, q << 2, 0xccccccccu);
        w[base + g + k + 2] = bitsel(q, p >> 2, 0x33333333u);
      }
    ZFB_UNROLL for (int g = 0; g < 16; g += 2) {
      const uint32_t p = w[base + g], q = w[base + g + 1];
      if (NEG == 0) {
        w[base + g] = bitsel(p, q << 1, 0xaaaaaaaau);
        w[base + g + 1] = bitsel(q, p >> 1, 0x55555555u);
      } else if (NEG == 1) {
        w[base + g] = bitsel(p, q << 1, 0xaaaaaaaau);
        w[base + g + 1] = lop3<0x27>(q, p >> 1, 0x55555555u);  // ~((q & ~m) | ((p >> 1) & m))
      } else {
        w[base + g] = lop3<0x72>(p, q << 1, 0xaaaaaaaau);      // (p & ~m) | (~(q << 1) & m)
        w[base + g + 1] = lop3<0x8D>(q, p >> 1, 0x55555555u);  // (~q & ~m) | ((p >> 1) & m)
      }
    }
  }
  ZFB_HD void split16() { stage16(a); stage16(b); }
  template <int NEG = 0> ZFB_HD void finish(int base) { stages_low<NEG>(a, base); stages_low<NEG>(b, base); }
};

// Finish (or, equally, un-finish: the stages are involutions) the transpose of the 16 planes parked in rows
// r0 .. r0+7: both 16-word groups (coefficients 0..31 and 32..63) go through stages 8, 4, 2, 1.
template <int NEG = 0, class PR> ZFB_HD void finish_rows(const PR& ps, int r0)
{
  planes64 Q;
  ZFB_UNROLL for (int r = 0; r < 8; r++) {
    const u32x4 v = ps.ld(r0 + r);
    Q.a[2 * r] = v.x; Q.b[2 * r] = v.y; Q.a[2 * r + 1] = v.z; Q.b[2 * r + 1] = v.w;
  }
  Q.template finish<NEG>(0);
  ZFB_UNROLL for (int r = 0; r < 8; r++) {
    u32x4 v;
    v.x = Q.a[2 * r]; v.y = Q.b[2 * r]; v.z = Q.a[2 * r + 1]; v.w = Q.b[2 * r + 1];
    ps.st(r0 + r, v);
  }
}

// ---------------------------------------------------------------------------------------------
// 32-bit word sinks / writer
// ---------------------------------------------------------------------------------------------
struct sink32 {  // slot on a word boundary, whole number of 32-bit words (maxbits % 64 == 0)
  uint32_t* dst;
  unsigned widx, nwords;
  ZFB_HD sink32(uint32_t* d, unsigned n) : dst(d), widx(0), nwords(n) {}
  ZFB_HD void word(uint32_t w)
  {
    if (widx < nwords) dst[widx] = w;
    widx++;
  }
  ZFB_HD void word_if(bool on, uint32_t w)
  {
    if (on & (widx < nwords)) dst[widx] = w;
    widx += on ? 1u : 0u;
  }
  ZFB_HD void word2(uint32_t w0, uint32_t w1)
  {
    if (widx + 2 <= nwords) { dst[widx] = w0; dst[widx + 1] = w1; }
    else if (widx < nwords) dst[widx] = w0;
    widx += 2;
  }
  ZFB_HD void finish()
  {
    for (; widx < nwords; widx++) dst[widx] = 0;
  }
  ZFB_HD unsigned words() const { return widx; }
  ZFB_HD void pad_to(unsigned bits) { while (32u * widx < bits && widx < nwords) dst[widx++] = 0; }
};

#if defined(__CUDACC__)
struct smem_sink32 {  // sink32 for a slot in shared memory, addressed through its 32-bit shared address (tile kernels)
  uint32_t base;
  unsigned widx, nwords;
  __device__ __forceinline__ smem_sink32(uint32_t b, unsigned n) : base(b), widx(0), nwords(n) {}
  __device__ __forceinline__ void put_word(uint32_t w) const
  {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4u * widx), "r"(w) : "memory");
  }
  __device__ __forceinline__ void word(uint32_t w)
  {
    if (widx < nwords) put_word(w);
    widx++;
  }
  __device__ __forceinline__ void word_if(bool on, uint32_t w)
  {
    if (on & (widx < nwords)) put_word(w);
    widx += on ? 1u : 0u;
  }
  __device__ __forceinline__ void word2(uint32_t w0, uint32_t w1)
  {
    word(w0);
    word(w1);
  }
  __device__ __forceinline__ void finish()
  {
    for (; widx < nwords; widx++) put_word(0u);
  }
  __device__ __forceinline__ unsigned words() const { return widx; }
  __device__ __forceinline__ void pad_to(unsigned bits)
  {
    while (32u * widx < bits && widx < nwords) { put_word(0u); widx++; }
  }
};
#endif

struct or_sink32 {  // arbitrary bit offset in a pre-zeroed stream (maxbits % 64 != 0)
  uint32_t* base;
  size_t startbit;
  unsigned maxbits, j;
  ZFB_HD or_sink32(uint32_t* b, size_t sb, unsigned mb) : base(b), startbit(sb), maxbits(mb), j(0) {}
  ZFB_HD static void or32(uint32_t* p, uint32_t v)
  {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
  }
  ZFB_HD void word(uint32_t w)
  {
    const unsigned off = 32u * j++;
    if (off >= maxbits) return;
    const unsigned valid = maxbits - off;
    if (valid < 32) w &= mask32(valid);
    if (!w) return;
    const size_t pos = startbit + off;
    const unsigned s = (unsigned)(pos & 31);
    or32(base + (pos >> 5), w << s);
    const uint32_t c = carry_out(w, s);
    if (c) or32(base + (pos >> 5) + 1, c);
  }
  ZFB_HD void word_if(bool on, uint32_t w) { if (on) word(w); }
  ZFB_HD void word2(uint32_t w0, uint32_t w1) { word(w0); word(w1); }
  ZFB_HD void finish() {}
  ZFB_HD unsigned words() const { return j; }
  ZFB_HD void pad_to(unsigned) {}  // the stream is pre-zeroed
};

template <class Sink> struct writer32 {
  Sink& sink;
  uint32_t acc;
  unsigned cnt;  // bits held in acc, < 32
  ZFB_HD explicit writer32(Sink& s) : sink(s), acc(0), cnt(0) {}
  ZFB_HD unsigned bits() const { return 32u * sink.words() + cnt; }  // bits emitted so far
  ZFB_HD void put(uint32_t v, unsigned len)  // len in [0, 32]; bits of v above len are zero
  {
    acc |= v << cnt;
    const uint32_t c = carry_out(v, cnt);
    cnt += len;
    // no branch: the lanes of a warp fill their words at different times, so a flush branch would be taken by a
    // few lanes on nearly every call
    const bool full = cnt >= 32;
    sink.word_if(full, acc);
    acc = full ? c : acc;
    cnt &= 31u;
  }
  ZFB_HD void put64(uint32_t xl, uint32_t xh)  // exactly 64 bits
  {
    const uint32_t w0 = acc | (xl << cnt);
    const uint32_t w1 = shl_pair(xl, xh, cnt);
    acc = carry_out(xh, cnt);
    sink.word2(w0, w1);
  }
  ZFB_HD void putv(uint32_t xl, uint32_t xh, unsigned len)  // len in [0, 64]; bits of (xh:xl) above len are zero
  {
    const uint32_t w0 = acc | (xl << cnt);
    const uint32_t w1 = shl_pair(xl, xh, cnt);
    const uint32_t w2 = carry_out(xh, cnt);
    const unsigned t = cnt + len;  // < 96
    const bool f1 = t >= 32u, f2 = t >= 64u;
    sink.word_if(f1, w0);
    sink.word_if(f2, w1);
    acc = f2 ? w2 : (f1 ? w1 : w0);
    cnt = t & 31u;
  }
  ZFB_HD void flush()
  {
    if (cnt) sink.word(acc);
    acc = 0; cnt = 0;
  }
};

#if defined(__CUDACC__)
// The bit writer of the tile kernels (slot in shared memory), arranged for the ALU pipe that bounds the encode kernel:
// the pending bits are one 32-bit word and a put is ONE wide multiply-add on the FMA pipe -- v * 2^(pos & 31) + pending
// as a 64-bit product: its low word is the word being filled, its high word what spills into the next -- instead of two
// shifts, an OR and a carry extraction on the ALU pipe.  The word index comes from the bit position (carried between
// puts), and the only ALU work left is the power of two, the new word index and the "did we cross a word" compare that
// predicates the store.  Words at or behind the end of the slot are dropped (fixed-rate truncation).
template <> struct writer32<smem_sink32> {
  smem_sink32& sink;
  uint32_t acc;   // pending bits (below bit pos & 31)
  unsigned pos;   // bits emitted so far
  unsigned wi;    // pos >> 5
  __device__ __forceinline__ explicit writer32(smem_sink32& s) : sink(s), acc(0u), pos(0u), wi(0u) {}
  __device__ __forceinline__ unsigned bits() const { return pos; }
  __device__ __forceinline__ void put(uint32_t v, unsigned len)  // len in [0, 32]; bits of v above len are zero
  {
    const uint32_t p = __funnelshift_l(0u, 1u, pos);  // 1 << (pos & 31)
    const uint64_t a = (uint64_t)v * p + acc;
    pos += len;
    const unsigned wn = pos >> 5;
    // crossed a word boundary: the finished word goes to the slot (unless it lies behind the slot's end) and the spill
    // becomes the pending word; all predicated, no branch (the lanes of a warp cross at different puts)
    asm volatile(
        "{\n.reg .pred q, r;\n"
        "setp.ne.u32 q, %1, %2;\n"
        "setp.lt.and.u32 r, %2, %3, q;\n"
        "@r st.shared.u32 [%4], %5;\n"
        "selp.u32 %0, %6, %5, q;\n}\n"
        : "=r"(acc)
        : "r"(wn), "r"(wi), "r"(sink.nwords), "r"(sink.base + 4u * wi), "r"((uint32_t)a), "r"((uint32_t)(a >> 32))
        : "memory");
    wi = wn;
  }
  __device__ __forceinline__ void put64(uint32_t xl, uint32_t xh) { put(xl, 32u); put(xh, 32u); }
  __device__ __forceinline__ void putv(uint32_t xl, uint32_t xh, unsigned len)  // len in [0, 64]
  {
    const unsigned l0 = len < 32u ? len : 32u;
    put(xl, l0);
    put(xh, len - l0);
  }
  __device__ __forceinline__ void flush()
  {
    if ((pos & 31u) && wi < sink.nwords) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sink.base + 4u * wi), "r"(acc) : "memory");
    const unsigned w = (pos + 31u) >> 5;
    sink.widx = w < sink.nwords ? w : sink.nwords;  // for sink.finish(): the rest of the slot is zero padding
    acc = 0u;
  }
};
#endif

ZFB_HD int msb64(uint32_t lo, uint32_t hi)  // (hi:lo) != 0
{
  return hi ? 32 + msb32(hi) : msb32(lo);
}

// ---------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------
// The embedded coder proper, shared by the float (NP = 32 planes) and double (NP = 64) front ends: planes
// NP-1 .. kmin of 64 coefficients, read from the plane rows.  On entry the top 16 planes (rows NP/2-8 ..
// NP/2-1) are finished; every lower group of 16 planes is still in its half-transposed form and is finished
// only when the budget reaches it.
//
// Per plane (n = coefficients already significant): ONE put of n + 1 bits (the verbatim bits and the first
// group test), then -- only if the rest of the plane holds a one -- table steps of two bytes each.  No case
// split on n: the plane is handled as a 64-bit value.
template <int NP, class Sink, class PR>
ZFB_HD unsigned encode_planes(const PR& ps, unsigned maxbits, uint32_t header, unsigned hbits, int kmin, Sink& sink,
                              const tables& tb)
{
  writer32<Sink> bw(sink);
  bw.put(header, hbits);

  // Three loops, one per regime of n (the regimes follow each other, n never decreases).  Each lane walks its own
  // planes (k is per lane), so the lanes of a warp go through a regime together even when they reach it at
  // different planes: the expensive middle regime costs the warp the slowest lane's planes, not one pass per
  // distinct plane index.
  unsigned n = 0;
  int k = NP - 1;
  uint32_t xl = 0u, xh = 0u;
  // next plane into (xl, xh); false when the block is finished (budget covered or no planes left)
  auto next = [&]() -> bool {
    if (k < kmin || bw.bits() >= maxbits) return false;
    if (NP == 32 ? k == 15 : (k < NP - 16 && (k & 15) == 15)) finish_rows<1>(ps, (k - 15) >> 1);  // next 16 planes: finish their transpose now
    ps.ld64(k, xl, xh);
    return true;
  };
  bool more = next();
  // 1. few coefficients significant and every one-bit of the rest within its first byte (the top planes of a block):
  //    verbatim bits, group test and the byte's code are one put of at most 32 bits
  ZFB_NOUNROLL while (more && n <= 15u && !xh && !(xl >> (n + 8u))) {
    const uint32_t b0 = xl >> n;
    const uint32_t e = enc_at(tb, 256u + b0);  // b0 == 0: nothing behind the (zero) group test
    bw.put((xl & mask32(n)) | ((b0 ? 1u : 0u) << n) | ((e & 0xffffu) << (n + 1u)), n + 1u + ((e >> 16) & 0x1fu));
    n += e >> 24;
    k--;
    more = next();
  }
  // 2. the general plane: one put of n + 1 bits (the verbatim bits and the first group test), then -- only if the
  //    rest of the plane holds a one -- table steps of two bytes each
  ZFB_NOUNROLL while (more && n < 60u) {
    // y = the not yet significant coefficients; t = first group test (y != 0)
    const uint64_t x = ((uint64_t)xh << 32) | xl;
    uint64_t y = x >> n;
    const uint32_t t = y ? 1u : 0u;
    {
      // the n verbatim bits with the test bit behind them: bits >= n of x are y, so x ^ ((y ^ t) << n) clears them
      // and leaves t at position n
      const uint64_t v = x ^ ((y ^ (uint64_t)t) << n);
      bw.putv((uint32_t)v, (uint32_t)(v >> 32), n + 1u);
    }
    if (t) {
      const unsigned plast = (unsigned)msb64((uint32_t)y, (uint32_t)(y >> 32));
      // the one of coefficient 63 is implicit: neither it nor a group test behind it is coded
      const bool implicit = n + plast == 63u;
      unsigned nb = (plast >> 3) + 1u;  // bytes of y up to the one holding the last one-bit
      n += plast + 1u;
      for (;;) {
        const uint32_t yl = (uint32_t)y;
        const uint32_t e0 = enc_at(tb, (yl & 0xffu) + (nb == 1u ? 256u : 0u));
        const uint32_t e1 = enc_at(tb, ((yl >> 8) & 0xffu) + (nb <= 2u ? 256u : 0u));  // nb == 1: byte 1 is 0 -> ENC[256] = nothing
        const unsigned l0 = (e0 >> 16) & 0x1fu;
        uint32_t code = (e0 & 0xffffu) | ((e1 & 0xffffu) << l0);
        unsigned len = l0 + ((e1 >> 16) & 0x1fu);
        if (nb <= 2u) {
          if (implicit) { len -= 2u; code &= ~(1u << len); }
          bw.put(code, len);
          break;
        }
        bw.put(code, len);
        nb -= 2u;
        y >>= 16;
      }
    }
    k--;
    more = next();
  }
  // 3. at most four coefficients left (or none: row 0 of the table is empty): 32 verbatim bits, then the other
  //    n - 32 with the row's whole group-test code (NEE) behind them
  ZFB_NOUNROLL while (more) {
    const unsigned row = 64u - n, s = n - 32u;  // s in 28..32
    const uint32_t e = enc_at(tb, 512u + 16u * row + fsr(xh, 0u, s));
    const uint32_t code = e & 0xffu;
    bw.put(xl, 32u);
    bw.putv((xh & mask32(s)) | (s < 32u ? code << s : 0u), carry_out(code, s & 31u), s + ((e >> 8) & 15u));
    n += e >> 12;
    k--;
    more = next();
  }
  const unsigned total = bw.bits();  // bits emitted (the fixed-rate sinks truncate / pad to the slot)
  bw.flush();
  sink.finish();
  return total;
}

// General form (zfp_stream's minbits / maxbits / maxprec / minexp): returns the bits the block occupies.
template <class Sink, class PR>
ZFB_HD unsigned encode_block(const float (&f)[64], const stream_params& sp, Sink& sink, const tables& tb,
                             const PR& ps)
{
  typedef traits<float> T;
  const unsigned maxbits = sp.maxbits;
  float mx = 0.f;
  ZFB_UNROLL for (int i = 0; i < 64; i++) {
#if defined(__CUDA_ARCH__)
    mx = fmaxf(mx, fabsf(f[i]));  // NaN operands are ignored, exactly like `if (max < f) max = f`
#else
    const float a = f[i] < 0 ? -f[i] : f[i];
    mx = (mx < a) ? a : mx;
#endif
  }
  const int E = T::expfield(mx);
  const int emax = E - 126;
  const int prec = block_precision<3>(emax, sp);
  if (!(mx > 0.f) || prec == 0) {  // all zero, or nothing above the accuracy floor: a single 0 bit
    sink.word(0u);
    sink.finish();
    sink.pad_to(sp.minbits);
    return sp.minbits > 1u ? sp.minbits : 1u;
  }
  const int kmin = prec < 32 ? 32 - prec : 0;

  planes64 P;
  {
    int32_t ib[64];
    const int se = 30 - emax;
    const float s = T::pow2(se);
    ZFB_UNROLL for (int i = 0; i < 64; i++) ib[i] = T::trunc_cast(s * f[i]);
    if (se > 127) { ZFB_UNROLL for (int i = 0; i < 64; i++) ib[i] = T::int_min(); }  // see zfb_codec.cuh
    fwd_xform<int32_t, 3>(ib);
    ZFB_UNROLL for (int i = 0; i < 32; i++) {
      // negabinary (i + M) ^ M: the XOR is folded into the last transpose stage (stages_low<1>)
      P.a[i] = (uint32_t)ib[order<3>::at(i)] + T::NBMASK;
      P.b[i] = (uint32_t)ib[order<3>::at(i + 32)] + T::NBMASK;
    }
  }
  P.split16();
  P.finish<1>(16);
  // rows 8..15: planes 16..31, finished.  rows 0..7: the lower halves, still un-transposed; they are
  // finished only if the budget reaches plane 15.
  ZFB_UNROLL for (int r = 0; r < 16; r++) {
    u32x4 v;
    v.x = P.a[2 * r]; v.y = P.b[2 * r]; v.z = P.a[2 * r + 1]; v.w = P.b[2 * r + 1];
    ps.st(r, v);
  }

  unsigned used = encode_planes<32>(ps, maxbits, 2u * (uint32_t)(E + 1) + 1u, 9u, kmin, sink, tb);
  used = used < maxbits ? used : maxbits;
  sink.pad_to(sp.minbits);
  return used < sp.minbits ? sp.minbits : used;
}

// Fixed-rate form.
template <class Sink, class PR>
ZFB_HD void encode_block(const float (&f)[64], unsigned maxbits, Sink& sink, const tables& tb, const PR& ps)
{
  encode_block(f, fixed_rate_params(maxbits), sink, tb, ps);
}


// ---------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------
// Register window over a slot: (hi:lo) holds the next `avail` >= 32 stream bits (fewer only at the end).
struct reader32 {
  const uint32_t* w;
  unsigned nwords, widx;
  uint32_t lo, hi;
  unsigned avail;
  unsigned bit0;  // bit offset of the slot inside word 0
  // p = first word of the slot region, n = words that may be read from p, startbit < 32
  ZFB_HD reader32(const uint32_t* p, unsigned n, unsigned startbit) : w(p), nwords(n), widx(2), avail(64), bit0(startbit)
  {
    lo = n > 0 ? p[0] : 0u;
    hi = n > 1 ? p[1] : 0u;
    if (startbit) consume(startbit);
  }
  ZFB_HD void consume(unsigned k)  // k in [0, 32]
  {
    lo = fsr(lo, hi, k);
    hi = k < 32 ? hi >> k : 0u;
    avail -= k;
    if (avail < 32) {  // hi == 0 here
      const uint32_t nw = widx < nwords ? w[widx] : 0u;
      widx++;
      lo |= nw << avail;
      hi = carry_out(nw, avail);
      avail += 32;
    }
  }
  ZFB_HD void consume_small(unsigned k)  // k in [0, 31]
  {
    lo = fsr(lo, hi, k);
    hi >>= k;
    avail -= k;
    if (avail < 32) {
      const uint32_t nw = widx < nwords ? w[widx] : 0u;
      widx++;
      lo |= nw << avail;
      hi = carry_out(nw, avail);
      avail += 32;
    }
  }
  ZFB_HD void consume_entry(uint32_t e)  // consume dec_nbits(e) <= 8 bits; e's upper bits are ignored by the shifts
  {
#if defined(__CUDA_ARCH__)
    lo = __funnelshift_r(lo, hi, e);
    hi = __funnelshift_r(hi, 0u, e);
#else
    const unsigned k = e & 31u;
    lo = fsr(lo, hi, k);
    hi >>= k;
#endif
    avail -= e & 31u;
    if (avail < 32) {
      const uint32_t nw = widx < nwords ? w[widx] : 0u;
      widx++;
      lo |= nw << avail;
      hi = carry_out(nw, avail);
      avail += 32;
    }
  }
  ZFB_HD uint32_t take(unsigned k)  // k in [0, 32]
  {
    const uint32_t v = lo & mask32(k);
    consume(k);
    return v;
  }
  ZFB_HD uint32_t word_at(unsigned i) const { return i < nwords ? w[i] : 0u; }
  // 64 stream bits starting `pos` bits into the slot, straight from memory (independent of the window)
  ZFB_HD void read64_at(unsigned pos, uint32_t& xl, uint32_t& xh) const
  {
    const unsigned p = pos + bit0, i = p >> 5, sh = p & 31u;
    const uint32_t w0 = word_at(i), w1 = word_at(i + 1), w2 = word_at(i + 2);
    xl = fsr(w0, w1, sh);
    xh = fsr(w1, w2, sh);
  }
};

// Bits of a slot by position (the parser keeps no streaming window: every access is a fresh, position-indexed fetch,
// so the only loop-carried reader state is the bit position).  R needs word_at(i) and bit0.
template <class R> ZFB_HD uint32_t fetch32(const R& r, unsigned pos)
{
  const unsigned p = pos + r.bit0, i = p >> 5;
  return fsr(r.word_at(i), r.word_at(i + 1), p & 31u);
}
template <class R> ZFB_HD void fetch64(const R& r, unsigned pos, uint32_t& lo, uint32_t& hi)
{
  const unsigned p = pos + r.bit0, i = p >> 5, sh = p & 31u;
  const uint32_t w0 = r.word_at(i), w1 = r.word_at(i + 1), w2 = r.word_at(i + 2);
  lo = fsr(w0, w1, sh);
  hi = fsr(w1, w2, sh);
}

// Unchecked reader over words in shared memory (tile kernels: the slot buffer is followed by other shared data, so
// the look-ahead words of the last slot are readable; what they hold is never used).
struct smem_words {
  uint32_t base;  // shared-memory address of the slot's first word
  static constexpr unsigned bit0 = 0;
  ZFB_HD uint32_t word_at(unsigned i) const
  {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + 4u * i));
    return v;
#else
    (void)i;
    return 0u;
#endif
  }
};

// Shared slot parser (float: NP = 32 planes, 9 header bits; double: NP = 64, 12): fills the plane rows with
// the decoded planes (finished form) and returns the first plane index that was NOT decoded (kmin - 1 .. NP-1),
// so the caller knows which groups of 16 planes are non-zero.  The header has been consumed by the caller.
//
// Per plane (n = coefficients already significant): the min(n, budget) verbatim bits come out of one 64-bit
// fetch; the group-test code behind them is parsed by
//   * table steps (<= 8 stream bits each, DEC) while neither coefficient 63 nor the end of the budget can be
//     reached within a step (n <= 54 and >= 8 bits left), and
//   * token steps otherwise: one group test plus one zero run ending in a one, by count-trailing-zeros, with zfp's
//     end rules done exactly (the one of coefficient 63 is implicit; a run cut off by the budget still deposits
//     its one).
template <int NP, class Reader, class PR>
ZFB_HD int parse_planes(const PR& ps, unsigned maxbits, unsigned hbits, int kmin, Reader& r, const tables& tb)
{
  {
    u32x4 z;
    z.x = z.y = z.z = z.w = 0u;
    ZFB_UNROLL for (int q = 0; q < NP / 2; q++) ps.st(q, z);
  }
  unsigned pos = hbits;  // bit position inside the slot
  unsigned n = 0;
  int k = NP - 1;
  // planes leave row-wise: the odd plane of a row is met first and waits in `cur` for the even one (one conflict-free
  // 128-bit store per two planes; the rows were zeroed, so a row whose even plane is never decoded is flushed as is)
  u32x4 cur;
  cur.x = cur.y = cur.z = cur.w = 0u;
  auto emit = [&](int kk, uint32_t lo, uint32_t hi) {
    if (kk & 1) { cur.z = lo; cur.w = hi; }
    else { cur.x = lo; cur.y = hi; ps.st(kk >> 1, cur); }
  };
  // One loop per regime of n, each lane walking its own planes (k is per lane): the lanes of a warp go through a
  // regime together even when they reach it at different planes (see encode_planes).
  // 1. few coefficients significant (the top planes of a block): the verbatim bits and the whole group-test code of
  //    the plane sit in one 32-bit fetch, one table step finishes the plane
  ZFB_NOUNROLL for (; k >= kmin && n <= 23u && pos + 31u <= maxbits; k--) {
    const uint32_t c = fetch32(r, pos);
    const uint32_t e = dec_at(tb, (c >> n) & 0xffu);
    if ((int32_t)e >= 0) break;  // the step did not end on a zero group test: a plane for the general loop
    emit(k, (c & mask32(n)) | (dec_pat(e) << n), 0u);
    pos += n + dec_nbits(e);
    n += dec_npos(e);
  }
  for (;;) {
  // 2. the general plane; leaves for loop 3 once at most four coefficients are left and the budget is not near its end
  ZFB_NOUNROLL for (; k >= kmin; k--) {
    if (pos >= maxbits || (n >= 60u && pos + 71u <= maxbits)) break;
    unsigned rem = maxbits - pos;
    uint32_t rl, rh;
    fetch64(r, pos, rl, rh);
    const unsigned m = n < rem ? n : rem;  // verbatim bits
    uint64_t x = ((uint64_t)rh << 32) | rl;
    x = m >= 64u ? x : x & ~(~(uint64_t)0 << m);
    pos += m;
    rem -= m;
    if (n < 64u && rem) {
      // group tests; c = the next cb valid stream bits (refilled when a step might need more)
      uint32_t c = m <= 32u ? fsr(rl, rh, m) : fetch32(r, pos);
      unsigned cb = 32u;
      bool inrun = false;  // inside a zero run (the group test that opened it has been consumed)
      for (;;) {
        bool stepped = false;
        if (n >= 60u && !inrun && rem >= 7u) {
          // at most four coefficients left, at a group test: one NE lookup finishes the plane
          const uint32_t e = ne_at(tb, 128u * (64u - n) + (c & 127u));
          x |= (uint64_t)(e >> 8) << n;
          pos += e & 15u;
          n += (e >> 4) & 15u;
          break;
        }
        if (rem >= 8u) {
          // table step: <= 8 stream bits; taken unless it would read the bit of coefficient 63 (whose one is implicit)
          const uint32_t e = dec_at(tb, (c & 0xffu) | (inrun ? 256u : 0u));
          const unsigned np = dec_npos(e);
          if (n + np <= 63u) {
            const unsigned nb = dec_nbits(e);
            x |= (uint64_t)dec_pat(e) << n;
            n += np;
            pos += nb; rem -= nb;
            c >>= nb; cb -= nb;
            inrun = (e >> 30) == 1u;
            if ((int32_t)e < 0) break;  // a zero group test: the plane is done
            if (inrun && n == 63u) { x |= (uint64_t)1 << 63; n = 64u; inrun = false; }
            stepped = true;
          }
        }
        if (!stepped) {
          // token step: one group test and one zero run, zfp's end rules done exactly
          if (cb < 32u) c = fetch32(r, pos);
          cb = 0u;
          const unsigned hb = inrun ? 0u : 1u;
          if (!inrun && !(c & 1u)) { pos += 1u; break; }
          const uint32_t z = inrun ? c : c >> 1;
          const unsigned wb = 32u - hb;  // run bits visible in the window
          const unsigned L = 63u - n, left = rem - hb;
          const unsigned lb = L < left ? L : left;  // the run ends here at the latest (block end / budget end)
          const unsigned r0 = z ? (unsigned)(msb32(z & (0u - z))) : 32u;
          unsigned used;
          if (r0 < (lb < wb ? lb : wb)) {  // the run's one is in the window
            x |= (uint64_t)1 << (n + r0);
            n += r0 + 1u;
            used = hb + r0 + 1u;
            inrun = false;
          } else if (lb <= wb) {  // cut off: the one is implicit (coefficient 63) or deposited by the budget rule
            x |= (uint64_t)1 << (n + lb);
            n += lb + 1u;
            used = hb + lb;
            inrun = false;
          } else {  // a run longer than the window
            n += wb;
            used = 32u;
            inrun = true;
          }
          pos += used;
          rem -= used;
        }
        if (n >= 64u || !rem) break;
        if (cb < 8u) { c = fetch32(r, pos); cb = 32u; }
      }
      if (inrun && !rem) {  // the budget ran out inside a zero run (after a table step): the one is deposited all the same
        x |= (uint64_t)1 << n;
        n++;
      }
    }
    emit(k, (uint32_t)x, (uint32_t)(x >> 32));
  }
  if (k < kmin || pos >= maxbits) break;
  // 3. at most four coefficients left (none: row 0 of NE is empty): n verbatim bits and one lookup for the rest.
  //    (A separate loop for n == 64 was slower: the lanes of a warp reach 64 at different planes and the two loops
  //    then run one after the other.)
  ZFB_NOUNROLL for (; k >= kmin && pos + 71u <= maxbits; k--) {
    uint32_t rl, rh;
    fetch64(r, pos, rl, rh);
    const unsigned row = 64u - n, s = n - 32u;  // s in 28..32
    const uint32_t e = ne_at(tb, 128u * row + (fetch32(r, pos + n) & 127u));
    emit(k, rl, (rh & mask32(s)) | (s < 32u ? (e >> 8) << s : 0u));
    pos += n + (e & 15u);
    n += (e >> 4) & 15u;
  }
  if (k < kmin || pos >= maxbits) break;
  }  // the budget is near its end: back to the general loop for the last planes
  if (k >= 0 && !(k & 1)) {  // stopped with the odd plane of a row pending (k = first plane NOT decoded)
    cur.x = 0u; cur.y = 0u;
    ps.st(k >> 1, cur);
  }
  return k;
}

// ---------------------------------------------------------------------------------------------
// Flat slot parser: ONE loop whose iteration is "the verbatim bits of a plane start (if the lane is at one) plus one
// table step", the same instruction sequence for every lane whatever plane it is in and however far it is.
// parse_planes above nests a table-step loop inside the plane loop and has a loop per regime, so a warp pays for its
// slowest lane in every plane and for every regime in turn (16 of 32 lanes active on the benchmark field); in
// fixed-rate mode all lanes of a warp consume the same number of bits, so their TOTAL step counts are close (22.8 per
// lane, 25.6 for the slowest lane of a warp at 8 bits per value) and the flat loop's trip count is their maximum.
//   * a step parses b = min(8, bits of budget left, distance to coefficient 63) stream bits through the FLAT table,
//     indexed by (state, b, the b bits): short steps at the end of the budget and near the last coefficient are the
//     same code with another table address; b = 0 (nothing left to parse in this plane) is a lookup without effect;
//   * zfp's two end rules are predicated instructions behind the step: in a zero run at coefficient 63 the one is
//     implicit, and a zero run cut off by the budget still deposits its one.
// There is no other path.  Same contract as parse_planes: fills the plane rows (zeroed here first), returns the first
// plane NOT decoded.  Checked against parse_planes on 800 k random slots and on encoded fields (tools/flat_check.cpp).
#ifdef ZFB_FLAT_STATS
struct flat_stats { unsigned iters, slow; };
#endif
template <int NP, class Reader, class PR>
ZFB_HD int parse_planes_flat(const PR& ps, unsigned maxbits, unsigned hbits, int kmin, Reader& r, const tables& tb
#ifdef ZFB_FLAT_STATS
                             , flat_stats* stt = nullptr
#endif
)
{
  {
    u32x4 z;
    z.x = z.y = z.z = z.w = 0u;
    ZFB_UNROLL for (int q = 0; q < NP / 2; q++) ps.st(q, z);
  }
  unsigned pos = hbits, n = 0;
  uint32_t st = 0u;  // 0: at a test bit, FLAT_B: inside a zero run
  int k = NP - 1;
  uint32_t xl = 0u, xh = 0u;
  bool mid = false;  // inside a plane (its verbatim bits are done)
  ZFB_NOUNROLL while (k >= kmin && pos < maxbits) {
#ifdef ZFB_FLAT_STATS
    if (stt) stt->iters++;
#endif
    const unsigned rem = maxbits - pos;
    const unsigned nr = n < rem ? n : rem;
    const unsigned m = mid ? 0u : nr;  // verbatim bits of this iteration
    uint32_t rl, rh;
    fetch64(r, pos, rl, rh);
    const unsigned p2 = pos + m;
    const uint32_t c = fetch32(r, p2);
    const uint32_t vl = rl & lowmask(m), vh = rh & lowmask(m > 32u ? m - 32u : 0u);
    xl = mid ? xl : vl;
    xh = mid ? xh : vh;
    // stream bits this step may parse: at most 8, not behind the budget, and not so many that the coefficient index
    // could pass 63 (state A: a test bit and b - 1 run bits advance at most b - 1 coefficients; state B: b)
    const unsigned rem2 = rem - m, lim = 64u - n - (st ? 1u : 0u);
    unsigned b = rem2 < lim ? rem2 : lim;
    b = b < 8u ? b : 8u;
    const uint32_t e = flat_at(tb, st, step_index(c, b));
    const uint64_t pat = (uint64_t)byte_of(e, 2) << n;
    xl |= (uint32_t)pat;
    xh |= (uint32_t)(pat >> 32);
    n += byte_of(e, 3);
    pos = p2 + (e & 31u);
    st = e & FLAT_B;
    if (st && (n == 63u || pos == maxbits)) {
      // the one of coefficient 63 is implicit; a zero run cut off by the budget still deposits its one
      const uint64_t o = (uint64_t)1 << n;
      xl |= (uint32_t)o;
      xh |= (uint32_t)(o >> 32);
      n++;
      st = 0u;
    }
    const bool done = b == 0u || (e & FLAT_END) || n >= 64u;
    if (done) {
      ps.st64(k, xl, xh);
      k--;
      st = 0u;
    }
    mid = !done;
  }
  if (mid) {  // the budget ended inside a plane
    ps.st64(k, xl, xh);
    k--;
  }
  return k;
}

// Phase 1: parse the slot into bit planes (left in P).  Returns false for an all-zero block.
template <class Reader, class PR>
ZFB_HD bool decode_planes(planes64& P, int& emax, bool& low_used_out, const stream_params& sp, Reader& r,
                          const tables& tb, const PR& ps)
{
  const uint32_t head = fetch32(r, 0);
  emax = 0;
  low_used_out = false;
  if (!(head & 1u)) return false;
  emax = (int)((head >> 1) & 0xffu) - 127;
  const int prec = block_precision<3>(emax, sp);
  const int kend = parse_planes<32>(ps, sp.maxbits, 9u, prec < 32 ? 32 - prec : 0, r, tb);
  ZFB_UNROLL for (int q = 0; q < 16; q++) {
    const u32x4 v = ps.ld(q);
    P.a[2 * q] = v.x; P.b[2 * q] = v.y; P.a[2 * q + 1] = v.z; P.b[2 * q + 1] = v.w;
  }
  low_used_out = kend < 15;
  return true;
}

template <class Reader, class PR>
ZFB_HD bool decode_planes(planes64& P, int& emax, bool& low_used_out, unsigned maxbits, Reader& r, const tables& tb,
                          const PR& ps)
{
  return decode_planes(P, emax, low_used_out, fixed_rate_params(maxbits), r, tb, ps);
}

// Phase 2: planes -> coefficients -> inverse transform -> floats.
ZFB_HD void decode_finish(float (&f)[64], planes64& P, int emax, bool nonzero, bool low_used)
{
  typedef traits<float> T;
  if (!nonzero) {
    ZFB_UNROLL for (int i = 0; i < 64; i++) f[i] = 0.f;
    return;
  }
  // negabinary (u ^ M) - M: the XOR is folded into the last transpose stage (stages_low<2>); a lower half that was
  // never decoded (all-zero planes) comes out of that stage as the constant M
  P.finish<2>(16);
  if (low_used) P.finish<2>(0);
  else { ZFB_UNROLL for (int i = 0; i < 16; i++) { P.a[i] = T::NBMASK; P.b[i] = T::NBMASK; } }
  P.split16();
  int32_t ib[64];
  ZFB_UNROLL for (int i = 0; i < 32; i++) {
    ib[order<3>::at(i)] = (int32_t)(P.a[i] - T::NBMASK);
    ib[order<3>::at(i + 32)] = (int32_t)(P.b[i] - T::NBMASK);
  }
  inv_xform<int32_t, 3>(ib);
  const float s = T::pow2(emax - 30);
  ZFB_UNROLL for (int i = 0; i < 64; i++) f[i] = s * T::to_scalar(ib[i]);
}

template <class Reader, class PR>
ZFB_HD void decode_block(float (&f)[64], const stream_params& sp, Reader& r, const tables& tb, const PR& ps)
{
  planes64 P;
  int emax;
  bool low_used;
  const bool nonzero = decode_planes(P, emax, low_used, sp, r, tb, ps);
  decode_finish(f, P, emax, nonzero, low_used);
}

template <class Reader, class PR>
ZFB_HD void decode_block(float (&f)[64], unsigned maxbits, Reader& r, const tables& tb, const PR& ps)
{
  decode_block(f, fixed_rate_params(maxbits), r, tb, ps);
}

}  // namespace f3
}  // namespace zfb
