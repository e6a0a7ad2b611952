// zfb_codec_f1.cuh -- float32, 1D blocks (4 coefficients): the HACC particle-array instantiation.
//
// Same bit stream as encode_block<float,1> / decode_block<float,1> (zfp 0.5.5 fixed-rate, SURVEY
// Appendix A; reference call sites zfpCompressorGpu.hpp:145,249 with zfp_field_1d), organised so that a
// block costs ~200 instructions instead of a 32-iteration plane loop:
//
//   * planes above the highest set bit of any coefficient are single zero test bits: skipped with one clz
//   * "head" planes (while fewer than 4 coefficients are significant): one table step per plane, keyed by
//     (n, 4-bit plane) for the encoder and (n, next 7 stream bits) for the decoder
//   * once all 4 coefficients are significant every further plane is 4 verbatim bits, so the rest of the
//     slot is the 4-way bit interleave of the coefficients' remaining bits, MSB first: done without a loop
//     over planes through byte spread / gather steps.
//
// Works for any maxbits that is a multiple of 32 (CBench's wra rounding gives 64); other budgets take the
// generic path.  Host-device so tests/host_emul.cpp can bit-compare it with the oracle on the CPU.
#pragma once
#include <math.h>
#include "zfb_codec.cuh"
#include "zfb_codec_f3.cuh"

namespace zfb {
namespace f1 {

using f3::carry_out;
using f3::fsr;
using f3::mask32;
using f3::popc32;

ZFB_HD uint32_t brev32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
ZFB_HD uint32_t byte_perm(uint32_t x, uint32_t y, unsigned sel)  // PRMT, plain byte selectors (0..7) only
{
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, y, sel);
#else
  const uint64_t xy = (uint64_t)x | ((uint64_t)y << 32);
  uint32_t r = 0;
  for (int j = 0; j < 4; j++) r |= (uint32_t)((xy >> (8u * ((sel >> (4 * j)) & 7u))) & 0xffu) << (8 * j);
  return r;
#endif
}
ZFB_HD int clz32(uint32_t v)  // v != 0
{
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return __builtin_clz(v);
#endif
}
ZFB_HD int ctz32(uint32_t v)  // v != 0
{
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

// ---------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------
// zfp's coder for one plane of a 4-coefficient block: n coefficients already significant, plane value x
// (bit i = coefficient i).  Returns the emitted bits (LSB first), their count, and the new n.
ZFB_HD void plane_code(unsigned n, unsigned x, unsigned& code, unsigned& len, unsigned& newn)
{
  code = x & ((1u << n) - 1u);
  len = n;
  x >>= n;
  while (n < 4u) {
    const unsigned t = x != 0u;
    code |= t << len; len++;
    if (!t) break;
    while (n < 3u) {
      const unsigned b = x & 1u;
      code |= b << len; len++;
      if (b) break;
      x >>= 1; n++;
    }
    x >>= 1; n++;
  }
  newn = n;
}

// ENC[n*16 + x] = code | len << 8 | newn << 12      (n = 0..3)
ZFB_HD uint16_t enc_entry(unsigned idx)
{
  unsigned code, len, newn;
  plane_code(idx >> 4, idx & 15u, code, len, newn);
  return (uint16_t)(code | (len << 8) | (newn << 12));
}

// DEC[n*128 + next 7 stream bits] = x | len << 4 | newn << 8   (n = 0..3): the unique plane whose code is a
// prefix of the 7 bits (codes are at most 7 bits long and prefix free for a given n)
ZFB_HD uint16_t dec_entry(unsigned idx)
{
  const unsigned n = idx >> 7, bits = idx & 127u;
  for (unsigned x = 0; x < 16u; x++) {
    unsigned code, len, newn;
    plane_code(n, x, code, len, newn);
    if ((bits & ((1u << len) - 1u)) == code) return (uint16_t)(x | (len << 4) | (newn << 8));
  }
  return 0;
}

// SPREAD[b]: bit j of byte b -> bit 4j
ZFB_HD uint32_t spread_entry(unsigned b)
{
  uint32_t r = 0;
  for (int j = 0; j < 8; j++) r |= ((b >> j) & 1u) << (4 * j);
  return r;
}

// UNSPREAD[b]: the inverse for two planes at a time.  Byte b = the raw nibbles of planes j (low) and j + 1 (high);
// coefficient i's two bits go to bits 8 i and 8 i + 1.
ZFB_HD uint32_t unspread_entry(unsigned b)
{
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (((b >> i) & 1u) | (((b >> (4 + i)) & 1u) << 1)) << (8 * i);
  return r;
}

struct tables {
  const uint16_t* enc;       // 64
  const uint16_t* dec;       // 512
  const uint32_t* spread;    // 256
  const uint32_t* unspread;  // 256
};

ZFB_HD uint32_t gather4(uint32_t w)  // bits 0,4,8,...,28 of w -> bits 0..7
{
  w &= 0x11111111u;
  w = (w | (w >> 3)) & 0x03030303u;
  w = (w | (w >> 6)) & 0x000f000fu;
  w = (w | (w >> 12)) & 0xffu;
  return w;
}

ZFB_HD uint32_t spread2(uint32_t x)  // bit j (j < 16) -> bit 2j
{
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
ZFB_HD uint32_t gather2(uint32_t x)  // bit 2j -> bit j
{
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0f0f0f0fu;
  x = (x | (x >> 4)) & 0x00ff00ffu;
  x = (x | (x >> 8)) & 0x0000ffffu;
  return x;
}
ZFB_HD uint32_t spread3(uint32_t x)  // bit j (j < 10) -> bit 3j
{
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
ZFB_HD uint32_t gather3(uint32_t x)  // bit 3j (j < 10) -> bit j
{
  x &= 0x09249249u;
  x = (x | (x >> 2)) & 0x030c30c3u;
  x = (x | (x >> 4)) & 0x0300f00fu;
  x = (x | (x >> 8)) & 0x030000ffu;
  x = (x | (x >> 16)) & 0x3ffu;
  return x;
}

// The head of a slot, i.e. the planes coded while n <= 2 coefficients are significant, has a simple
// shape: RUNS of planes in which no further coefficient becomes significant -- each such plane is its n
// verbatim bits followed by a 0 group test, n + 1 bits -- separated by EVENT planes, whose code comes from
// the (n, nibble) table.  Runs are emitted / parsed many planes at a time with the spread / gather helpers
// above, so the loops below make at most a handful of iterations whatever the data (each coefficient
// causes at most one event).  From n >= 3 on the code of a plane IS its raw nibble: the bit of coefficient 3
// doubles as the group test (its one is implicit), so the rest of the slot is the 4-way bit interleave.

// ---------------------------------------------------------------------------------------------
// encode: writes maxbits/32 words to out[]
// ---------------------------------------------------------------------------------------------
template <int MAXW>
ZFB_HD void encode_block(const float (&f)[4], unsigned maxbits, uint32_t (&out)[MAXW], const tables& tb)
{
  typedef traits<float> T;
  const unsigned nw = maxbits >> 5;
  ZFB_UNROLL for (int i = 0; i < MAXW; i++) out[i] = 0u;
  float mx = 0.f;
  ZFB_UNROLL for (int i = 0; i < 4; i++) mx = fmaxf(mx, fabsf(f[i]));  // NaN operands are ignored, exactly like `if (max < f) max = f`
  if (!(mx > 0.f)) return;
  const int E = T::expfield(mx);
  const int emax = E - 126;
  int32_t ib[4];
  {
    const int se = 30 - emax;
    const float s = T::pow2(se);
    ZFB_UNROLL for (int i = 0; i < 4; i++) ib[i] = T::trunc_cast(s * f[i]);
    if (se > 127) { ZFB_UNROLL for (int i = 0; i < 4; i++) ib[i] = T::int_min(); }
  }
  fwd_lift(ib[0], ib[1], ib[2], ib[3]);
  uint32_t u[4];
  ZFB_UNROLL for (int i = 0; i < 4; i++) u[i] = ((uint32_t)ib[i] + T::NBMASK) ^ T::NBMASK;

  // bit position bookkeeping: bits are OR-ed into out[] at `pos`; everything past maxbits is dropped
  unsigned pos = 0;
  uint64_t acc0 = 0, acc1 = 0;  // stream bits 0..63 and 64..127
  auto put = [&](uint32_t v, unsigned len) {  // len <= 32, v < 2^len
    if (pos < 64u) acc0 |= (uint64_t)v << pos;
    if (MAXW > 2) {
      if (pos >= 64u) { if (pos < 128u) acc1 |= (uint64_t)v << (pos - 64u); }
      else if (pos > 32u) acc1 |= (uint64_t)v >> (64u - pos);
    }
    pos += len;
  };
  auto store = [&]() {
    out[0] = (uint32_t)acc0;
    if (MAXW > 1) out[1 % MAXW] = (uint32_t)(acc0 >> 32);
    if (MAXW > 2) { out[2 % MAXW] = (uint32_t)acc1; out[3 % MAXW] = (uint32_t)(acc1 >> 32); }
  };
  put(2u * (uint32_t)(E + 1) + 1u, 9);
  const uint32_t any = u[0] | u[1] | u[2] | u[3];
  if (!any) { store(); return; }  // cannot happen for finite input (the block maximum maps to >= 2^29); kept for safety
  int k = 31 - clz32(any);
  pos += (unsigned)(31 - k);  // planes above k: one zero group-test bit each
  unsigned n = 0;
  // Is there budget left?  Only an early exit for 32- and 64-bit slots, where the writer drops whatever lands beyond bit
  // 63 and the caller stores the low word(s) alone: there the head steps simply run on (a block whose budget ends inside
  // its head is the rare case) and the test costs nothing.
  auto room = [&]() -> bool { return MAXW <= 2 || pos < maxbits; };
  // The coefficients are carried left-aligned at the current plane: bit 31 of t[i] is plane k of u[i] (no per-use shift
  // or mask by k; planes below plane 0 are zero bits).
  uint32_t t[4];
  ZFB_UNROLL for (int i = 0; i < 4; i++) t[i] = u[i] << (31 - k);
  auto nibble = [&]() -> unsigned {  // plane k as a nibble (bit i = coefficient i): four funnel shifts
    unsigned x = t[3] >> 31;
    x = (x << 1) | (t[2] >> 31);
    x = (x << 1) | (t[1] >> 31);
    return (x << 1) | (t[0] >> 31);
  };
  auto advance = [&](unsigned s) {  // s < 32 planes coded
    ZFB_UNROLL for (int i = 0; i < 4; i++) t[i] <<= s;
    k -= (int)s;
  };
  // One head step = a run (possibly empty) of planes in which coefficients nn..3 are all zero AND the event plane behind
  // it: the lanes of a warp then walk the same code whatever mix of runs and events their blocks hold.  `nn` (1 or 2) is
  // the number of significant coefficients on entry (n == nn); with a literal argument the step is specialised at
  // compile time (one spread form, fixed cap and table row).
  auto head_step = [&](const unsigned nn) {
    uint32_t rest = t[3] | t[2];
    if (nn < 2u) rest |= t[1];
    unsigned L = rest ? (unsigned)clz32(rest) : 32u;  // planes from k down without a bit of coefficients nn..3
    if (L > (unsigned)k + 1u) L = (unsigned)k + 1u;   // (plane 0 is the last one)
    const unsigned cap = nn == 1u ? 16u : 10u;        // planes whose codes fit one 32-bit put
    const bool capped = L > cap;  // (a capped run is simply continued by the next step)
    if (capped) L = cap;
    if (L) {
      uint32_t code;
      if (nn == 1u) code = spread2(brev32(t[0]) & mask32(L));
      else code = spread3(brev32(t[0]) & mask32(L)) | (spread3(brev32(t[1]) & mask32(L)) << 1);
      put(code, L * (nn + 1u));
      advance(L);
    }
    if (!capped && k >= 0 && room()) {
      // event plane k
      const unsigned e = tb.enc[nn * 16u + nibble()];
      put(e & 0xffu, (e >> 8) & 15u);
      n = e >> 12;
      advance(1u);
    }
  };
  // Every event raises n, so a block passes through each of n = 0, 1, 2 at most once: three straight-line steps, each
  // specialised for its n, instead of a loop whose trip count is the warp's maximum and whose body serves every n.
  // Plane k is the top set plane, so the n == 0 step is its event alone.
  if (room()) {
    const unsigned e = tb.enc[nibble()];
    put(e & 0xffu, (e >> 8) & 15u);
    n = e >> 12;
    advance(1u);
  }
  if (n == 1u && k >= 0 && room()) head_step(1u);
  if (n == 2u && k >= 0 && room()) head_step(2u);
  while (n < 3u && k >= 0 && room()) head_step(n);  // only behind a capped run (> 16 / 10 planes without an event)
  if (n >= 3u && k >= 0 && pos < (MAXW <= 2 ? 64u : maxbits)) {
    // planes k, k-1, ... : the raw nibble of every plane = interleave of the coefficients' bits below plane k+1
    uint32_t v[4];
    ZFB_UNROLL for (int i = 0; i < 4; i++) v[i] = brev32(t[i]);  // bit j = bit (k - j) of u[i]
    if (MAXW <= 2) {
      // at most 54 bits of the slot are left: the top 8 (16) planes, interleaved, are OR-ed in with one shift; planes
      // that do not exist are zero bits of v[], what falls beyond the slot is dropped by the shift.  (A table indexed by
      // the un-reversed byte saves the four BREV and is slower: 0.78 -> 0.80 ms.)
      auto chunk = [&](unsigned sh) -> uint32_t {
        return tb.spread[(v[0] >> sh) & 0xffu] | (tb.spread[(v[1] >> sh) & 0xffu] << 1) |
               (tb.spread[(v[2] >> sh) & 0xffu] << 2) | (tb.spread[(v[3] >> sh) & 0xffu] << 3);
      };
      uint64_t W = chunk(0u);
      if (MAXW == 2) W |= (uint64_t)chunk(8u) << 32;
      acc0 |= W << pos;
    } else {
      const unsigned planes = (unsigned)k + 1u;
      // 32 stream bits (8 planes) per step; the writer drops what falls beyond maxbits
      for (unsigned c = 0; c * 8u < planes && pos < maxbits; c++) {
        const unsigned sh = 8u * c;
        uint32_t w = tb.spread[(v[0] >> sh) & 0xffu] | (tb.spread[(v[1] >> sh) & 0xffu] << 1) |
                     (tb.spread[(v[2] >> sh) & 0xffu] << 2) | (tb.spread[(v[3] >> sh) & 0xffu] << 3);
        const unsigned left = planes - 8u * c;  // planes available in this chunk
        unsigned len = left >= 8u ? 32u : 4u * left;
        put(w & mask32(len), len);
      }
    }
  }
  // bits past maxbits (a multiple of 32) live in words the caller does not store
  store();
  (void)nw;
}

// One plane whose code is cut short by the end of the bit budget (`bits` < code length): zfp's decoder loop
// conditions, bit by bit.  Returns the plane nibble | new n << 4.  Kept out of line: it runs at most once per
// block and must not be if-converted into the table path.
#if defined(__CUDACC__)
__device__ __host__ __noinline__
#else
inline
#endif
unsigned truncated_plane(uint32_t w, unsigned bits, unsigned n)
{
  unsigned m = n < bits ? n : bits;
  bits -= m;
  unsigned x = w & ((1u << m) - 1u);
  uint32_t ww = w >> m;
  while (n < 4u && bits) {
    bits--;
    const unsigned tt = ww & 1u; ww >>= 1;
    if (!tt) break;
    while (n < 3u && bits) {
      bits--;
      const unsigned b = ww & 1u; ww >>= 1;
      if (b) break;
      n++;
    }
    x |= 1u << n;
    n++;
  }
  return x | (n << 4);
}

// ---------------------------------------------------------------------------------------------
// decode: reads maxbits/32 words from in[]
// ---------------------------------------------------------------------------------------------
template <int MAXW>
ZFB_HD void decode_block(float (&f)[4], unsigned maxbits, const uint32_t (&in)[MAXW], const tables& tb)
{
  typedef traits<float> T;
  if (!(in[0] & 1u)) {
    ZFB_UNROLL for (int i = 0; i < 4; i++) f[i] = 0.f;
    return;
  }
  const int emax = (int)((in[0] >> 1) & 0xffu) - 127;
  unsigned pos = 9;
  const uint64_t in0 = (uint64_t)in[0] | ((uint64_t)(MAXW > 1 ? in[1 % MAXW] : 0u) << 32);
  const uint64_t in1 = MAXW > 2 ? ((uint64_t)in[2 % MAXW] | ((uint64_t)in[3 % MAXW] << 32)) : 0;
  auto peek = [&](unsigned p) -> uint32_t {  // 32 stream bits starting at p (zeros beyond the slot words)
    if (MAXW <= 2) return p < 64u ? (uint32_t)(in0 >> p) : 0u;
    if (p >= 128u) return 0u;
    if (p >= 64u) return (uint32_t)(in1 >> (p - 64u));
    uint64_t v = in0 >> p;
    if (p > 32u) v |= in1 << (64u - p);
    return (uint32_t)v;
  };
  // 32- and 64-bit slots: the unread rest of the slot is carried in one register pair, shifted down as bits are
  // consumed (no position arithmetic, zeros shift in behind the slot's end)
  uint64_t cur = in0 >> 9;
  auto window = [&]() -> uint32_t { return MAXW <= 2 ? (uint32_t)cur : peek(pos); };
  uint32_t u[4] = {0u, 0u, 0u, 0u};
  unsigned remaining = maxbits - 9u;
  auto consume = [&](unsigned len) {  // len <= 32
    pos += len; remaining -= len;
    if (MAXW <= 2) cur >>= len;
  };
  // leading all-zero planes: zero group tests while n == 0
  int k = 31;
  {
    uint32_t w = window();
    if (MAXW > 2 && remaining < 32u) w &= mask32(remaining);  // (32/64-bit slots: zeros shift in behind the slot's end)
    const unsigned z = w ? (unsigned)ctz32(w) : 32u;   // zero test bits before the first one
    const unsigned skip = z < remaining ? z : remaining;
    // at most 32 planes exist; a slot that is all zero after the header decodes to zero coefficients
    const unsigned sk = skip > 32u ? 32u : skip;
    consume(sk); k -= (int)sk;
  }
  unsigned n = 0;
  // the event plane k, its code in the low bits of `we`, with nn coefficients significant on entry
  auto event_plane = [&](const uint32_t we, const unsigned nn) {
    const unsigned e = tb.dec[nn * 128u + (we & 127u)];
    const unsigned len = (e >> 4) & 15u;
    unsigned x = e & 15u;
    n = e >> 8;
    if (len <= remaining) consume(len);
    else {
      // the budget ends inside this plane's code (at most once per block): exact bit-serial steps, out of line
      const unsigned xn = truncated_plane(we, remaining, nn);
      x = xn & 15u;
      n = xn >> 4;
      remaining = 0;
    }
    ZFB_UNROLL for (int i = 0; i < 4; i++) u[i] |= ((x >> i) & 1u) << k;
    k--;
  };
  // One head step = a run (possibly empty) AND the event plane behind it, as in the encoder; nn = n on entry (a literal
  // argument specialises the step).  A run that fills the whole window (L == R > 0) has not met its event yet: the next
  // step continues it.
  auto head_step = [&](const unsigned nn) {
    const uint32_t w = window();
    const unsigned cw = nn + 1u;                                   // bits per run-plane code
    unsigned R = (remaining < 32u ? remaining : 32u);
    if (nn == 2u && R > 30u) R = 30u;
    R = nn == 0u ? R : (nn == 1u ? R >> 1 : (R * 11u) >> 5);       // whole codes in the window (nn == 2: floor(R / 3) for R <= 30)
    if (R > (unsigned)k + 1u) R = (unsigned)k + 1u;
    // group-test bits of the run-plane codes: the last bit of every code
    const uint32_t tests = (nn == 0u ? 0xffffffffu : (nn == 1u ? 0xaaaaaaaau : 0x24924924u)) & mask32(R * cw);
    const uint32_t t = w & tests;
    unsigned L = R;
    if (t) { const unsigned z = (unsigned)ctz32(t); L = nn == 0u ? z : (nn == 1u ? z >> 1 : (z * 11u) >> 5); }
    const bool event = L == 0u || L < R;
    uint32_t we = w;
    if (L) {
      if (nn == 1u) u[0] |= brev32(gather2(w) & mask32(L)) >> (31 - k);
      else if (nn == 2u) {
        u[0] |= brev32(gather3(w) & mask32(L)) >> (31 - k);
        u[1] |= brev32(gather3(w >> 1) & mask32(L)) >> (31 - k);
      }
      consume(L * cw); k -= (int)L;
      we = window();
    }
    // (R == 0: fewer than nn + 1 bits are left, the budget ends inside this plane's code -- handled exactly in event_plane)
    if (event && k >= 0 && remaining) event_plane(we, nn);
  };
  // Every event raises n: three straight-line steps, each specialised for its n (see the encoder).  Behind the skipped
  // zero tests the next bit is a one, so the n == 0 step is its event alone.
  if (k >= 0 && remaining) event_plane(window(), 0u);
  if (n == 1u && k >= 0) head_step(1u);  // (with nothing left a step parses zeros and changes nothing)
  if (n == 2u && k >= 0) head_step(2u);
  while (n < 3u && k >= 0 && remaining) head_step(n);  // only behind a run longer than one window
  if (n >= 3u && k >= 0 && remaining) {
    if (MAXW <= 2) {
      // the rest of the slot (zeros shift in behind its end) is the interleave of the top planes: two planes per table
      // lookup, coefficient i's bits collected in byte i; planes below plane 0 fall off in the final shift
      const uint64_t r = cur;
      auto chunk = [&](uint32_t w) -> uint32_t {
        return tb.unspread[w & 0xffu] | (tb.unspread[(w >> 8) & 0xffu] << 2) | (tb.unspread[(w >> 16) & 0xffu] << 4) |
               (tb.unspread[w >> 24] << 6);
      };
      const uint32_t o0 = chunk((uint32_t)r);
      const uint32_t o1 = MAXW == 2 ? chunk((uint32_t)(r >> 32)) : 0u;
      // coefficient i's 16 plane bits, most significant plane first, are byte 3 - i of the two bit-reversed words: one
      // PRMT puts them in the top half, the shift drops them at plane k
      const uint32_t b0 = brev32(o0), b1 = brev32(o1);
      const uint32_t keep = 0xffff0000u >> (31 - k);
      ZFB_UNROLL for (int i = 0; i < 4; i++)
        u[i] |= (byte_perm(b0, b1, (unsigned)(((3 - i) << 12) | ((7 - i) << 8))) >> (31 - k)) & keep;
    } else {
      const unsigned planes = (unsigned)k + 1u;
      uint32_t v[4] = {0u, 0u, 0u, 0u};
      for (unsigned c = 0; c * 8u < planes && remaining; c++) {
        uint32_t w = peek(pos);
        const unsigned left = planes - 8u * c;
        unsigned len = left >= 8u ? 32u : 4u * left;
        if (len > remaining) len = remaining;
        w &= mask32(len);
        ZFB_UNROLL for (int i = 0; i < 4; i++) v[i] |= gather4(w >> i) << (8u * c);
        pos += len; remaining -= len;
      }
      ZFB_UNROLL for (int i = 0; i < 4; i++) u[i] |= brev32(v[i]) >> (31 - k);
    }
  }
  int32_t ib[4];
  ZFB_UNROLL for (int i = 0; i < 4; i++) ib[i] = (int32_t)((u[i] ^ T::NBMASK) - T::NBMASK);
  inv_lift(ib[0], ib[1], ib[2], ib[3]);
  const float s = T::pow2(emax - 30);
  ZFB_UNROLL for (int i = 0; i < 4; i++) f[i] = s * T::to_scalar(ib[i]);
}

// ---------------------------------------------------------------------------------------------
// Variable rate (zfp fixed accuracy / precision, no bit budget in the way): one table step per bit plane.
// Valid when sp.maxbits covers the longest codable block (1 + 8 + 3 + 4 * 32 = 140 bits) and sp.minbits <= 1,
// which is what zfp_stream_set_accuracy / set_precision give; other parameter sets take the generic coder.
// ---------------------------------------------------------------------------------------------
ZFB_HD bool var_params_ok(const stream_params& sp) { return sp.maxbits >= 140u && sp.minbits <= 1u; }

// Appends the block to the caller's bit writer (put(v, len) with len <= 64); returns its length in bits.
template <class Writer>
ZFB_HD unsigned encode_block_var(const float (&f)[4], const stream_params& sp, Writer& bw, const tables& tb)
{
  typedef traits<float> T;
  float mx = 0.f;
  ZFB_UNROLL for (int i = 0; i < 4; i++) mx = fmaxf(mx, fabsf(f[i]));  // NaN operands are ignored, exactly like `if (max < f) max = f`
  const int E = T::expfield(mx);
  const int emax = E - 126;
  const int prec = block_precision<1>(emax, sp);
  if (!(mx > 0.f) || prec == 0) { bw.put(0, 1); return 1u; }
  const int kmin = prec < 32 ? 32 - prec : 0;
  int32_t ib[4];
  {
    const int se = 30 - emax;
    const float s = T::pow2(se);
    ZFB_UNROLL for (int i = 0; i < 4; i++) ib[i] = T::trunc_cast(s * f[i]);
    if (se > 127) { ZFB_UNROLL for (int i = 0; i < 4; i++) ib[i] = T::int_min(); }  // see zfb_codec.cuh
  }
  fwd_xform<int32_t, 1>(ib);
  uint32_t u[4];
  ZFB_UNROLL for (int i = 0; i < 4; i++) u[i] = ((uint32_t)ib[i] + T::NBMASK) ^ T::NBMASK;
  bw.put(2u * (uint32_t)(E + 1) + 1u, 9);
  unsigned used = 9;
  int k = 31;
  // planes above every coefficient's top bit: one zero test bit each
  const uint32_t any = u[0] | u[1] | u[2] | u[3];
  const int ktop = any ? 31 - clz32(any) : -1;
  {
    const int stop = ktop + 1 > kmin ? ktop + 1 : kmin;  // first plane that is not skipped
    const unsigned z = (unsigned)(32 - stop);
    if (z) { bw.put(0, z); used += z; k = stop - 1; }
  }
  // head: one table step per plane until all four coefficients are significant, eight planes per batch: their
  // nibbles come from one byte of each coefficient through the spread table (plane k in the top nibble of W), and
  // the batch's codes (<= 7 bits each) are collected in a 64-bit word before they go to the writer
  unsigned n = 0;
  while (k >= kmin && n < 4u) {
    const unsigned sh = 31u - (unsigned)k;
    const uint32_t W = tb.spread[(u[0] << sh) >> 24] | (tb.spread[(u[1] << sh) >> 24] << 1) |
                       (tb.spread[(u[2] << sh) >> 24] << 2) | (tb.spread[(u[3] << sh) >> 24] << 3);
    uint64_t acc = 0;
    unsigned cnt = 0;
    ZFB_NOUNROLL for (unsigned i = 0; i < 8u && k >= kmin && n < 4u; i++, k--) {
      const unsigned x = (W >> (28u - 4u * i)) & 15u;
      const unsigned e = tb.enc[n * 16u + x];
      acc |= (uint64_t)(e & 0xffu) << cnt;
      cnt += (e >> 8) & 15u;
      n = e >> 12;
    }
    bw.put(acc, cnt);
    used += cnt;
  }
  // tail: every further plane is its four bits verbatim, i.e. the rest of the block is the 4-way bit interleave of
  // the coefficients' bits k .. kmin, most significant first: 32 stream bits per step through the byte spread table
  if (k >= kmin) {
    const unsigned R = (unsigned)(k - kmin + 1);
    uint32_t v[4];
    ZFB_UNROLL for (int c = 0; c < 4; c++) v[c] = (brev32(u[c]) >> (31 - k)) & mask32(R);  // bit i = bit k - i of u[c]
    ZFB_UNROLL for (unsigned w = 0; w < 4; w++) {
      if (8u * w < R) {
        const uint32_t t = tb.spread[(v[0] >> (8u * w)) & 0xffu] | (tb.spread[(v[1] >> (8u * w)) & 0xffu] << 1) |
                           (tb.spread[(v[2] >> (8u * w)) & 0xffu] << 2) | (tb.spread[(v[3] >> (8u * w)) & 0xffu] << 3);
        const unsigned left = 4u * R - 32u * w;
        bw.put(t, left < 32u ? left : 32u);
      }
    }
    used += 4u * R;
  }
  return used;
}

// Decodes the block that starts at the reader's position (an f3::reader32-like window: lo, consume(k), take(k));
// returns its length in bits.
template <class Reader>
ZFB_HD unsigned decode_block_var(float (&f)[4], const stream_params& sp, Reader& r, const tables& tb)
{
  typedef traits<float> T;
  const uint32_t head = r.lo;
  if (!(head & 1u)) {
    ZFB_UNROLL for (int i = 0; i < 4; i++) f[i] = 0.f;
    r.consume(1);
    return 1u;
  }
  const int emax = (int)((head >> 1) & 0xffu) - 127;
  r.consume(9);
  unsigned used = 9;
  const int prec = block_precision<1>(emax, sp);
  const int kmin = prec < 32 ? 32 - prec : 0;
  uint32_t u[4] = {0u, 0u, 0u, 0u};
  unsigned n = 0;
  int k = 31;
  while (k >= kmin && n < 4u) {  // head: one table step per plane, the nibbles of eight planes collected in W
    const int kb = k;
    uint32_t W = 0u;
    ZFB_NOUNROLL for (unsigned i = 0; i < 8u && k >= kmin && n < 4u; i++, k--) {
      const unsigned e = tb.dec[n * 128u + (r.lo & 127u)];
      const unsigned len = (e >> 4) & 15u;
      W |= (e & 15u) << (28u - 4u * i);  // plane kb - i in nibble 7 - i
      n = e >> 8;
      r.consume(len);
      used += len;
    }
    // de-interleave: gather4(W >> c) has plane kb - 7 + j of coefficient c in bit j
    ZFB_UNROLL for (int c = 0; c < 4; c++) u[c] |= (gather4(W >> c) << 24) >> (31 - kb);
  }
  if (k >= kmin) {  // tail: de-interleave the remaining 4 R bits, 32 at a time
    const unsigned R = (unsigned)(k - kmin + 1);
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    ZFB_UNROLL for (unsigned w = 0; w < 4; w++) {
      if (8u * w < R) {
        const unsigned left = 4u * R - 32u * w;
        const uint32_t t = r.take(left < 32u ? left : 32u);
        ZFB_UNROLL for (int c = 0; c < 4; c++) v[c] |= gather4(t >> c) << (8u * w);
      }
    }
    ZFB_UNROLL for (int c = 0; c < 4; c++) u[c] |= brev32(v[c]) >> (31 - k);
    used += 4u * R;
  }
  int32_t ib[4];
  ZFB_UNROLL for (int i = 0; i < 4; i++) ib[i] = (int32_t)((u[i] ^ T::NBMASK) - T::NBMASK);
  inv_xform<int32_t, 1>(ib);
  const float s = T::pow2(emax - 30);
  ZFB_UNROLL for (int i = 0; i < 4; i++) f[i] = s * T::to_scalar(ib[i]);
  return used;
}

}  // namespace f1
}  // namespace zfb
