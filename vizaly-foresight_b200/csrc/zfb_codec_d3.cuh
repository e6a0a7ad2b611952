// zfb_codec_d3.cuh -- float64, 3D blocks: 64 coefficients of 64 bits, 64 bit planes, 12-bit header.
//
// Same bit stream as encode_block<double,3> / decode_block<double,3> in zfb_codec.cuh (zfp 0.5.5 fixed
// rate; reference call sites zfpCompressorGpu.hpp:145,249 with zfp_type_double).  A bit
// plane of a 3D block is 64 bits wide whatever the scalar type, so the embedded coder is the shared
// f3::encode_planes / f3::parse_planes with NP = 64; what is specific here is the front end: int64
// block-floating-point cast and lifting, and the 64x64 bit transpose done as two 32-plane groups (the
// coefficients' upper words give planes 63..32, the lower words planes 31..0), each finished lazily in
// halves of 16 planes like the float codec.  Plane rows: 32 rows of 16 bytes per block.
#pragma once
#include "zfb_codec_f3.cuh"

namespace zfb {
namespace d3 {

using f3::plane_rows;
using f3::planes64;
using f3::tables;
using f3::u32x4;

// General form (zfp_stream's minbits / maxbits / maxprec / minexp): returns the bits the block occupies.
template <class Sink, class PR>
ZFB_HD unsigned encode_block(const double (&f)[64], const stream_params& sp, Sink& sink, const tables& tb, const PR& ps)
{
  typedef traits<double> T;
  const unsigned maxbits = sp.maxbits;
  double mx = 0.0;
  ZFB_UNROLL for (int i = 0; i < 64; i++) {
#if defined(__CUDA_ARCH__)
    mx = fmax(mx, fabs(f[i]));  // NaN operands are ignored, exactly like `if (max < f) max = f`
#else
    const double a = f[i] < 0 ? -f[i] : f[i];
    mx = (mx < a) ? a : mx;
#endif
  }
  const int E = T::expfield(mx);
  const int emax = E - 1022;
  const int prec = block_precision<3>(emax, sp);
  if (!(mx > 0.0) || prec == 0) {  // all zero, or nothing above the accuracy floor: a single 0 bit
    sink.word(0u);
    sink.finish();
    sink.pad_to(sp.minbits);
    return sp.minbits > 1u ? sp.minbits : 1u;
  }
  const int kmin = 64 - prec;

  planes64 PH, PL;  // upper / lower words of the negabinary coefficients in sequency order
  {
    int64_t ib[64];
    const int se = 62 - emax;
    const double s = T::pow2(se);
    ZFB_UNROLL for (int i = 0; i < 64; i++) ib[i] = T::trunc_cast(s * f[i]);
    if (se > 1023) { ZFB_UNROLL for (int i = 0; i < 64; i++) ib[i] = T::int_min(); }  // see zfb_codec.cuh
    fwd_xform<int64_t, 3>(ib);
    ZFB_UNROLL for (int i = 0; i < 32; i++) {
      // negabinary (i + M) ^ M: the XOR is folded into the last transpose stage (planes64::stages_low<1>)
      const uint64_t ua = (uint64_t)ib[order<3>::at(i)] + T::NBMASK;
      const uint64_t ub = (uint64_t)ib[order<3>::at(i + 32)] + T::NBMASK;
      PH.a[i] = (uint32_t)(ua >> 32); PL.a[i] = (uint32_t)ua;
      PH.b[i] = (uint32_t)(ub >> 32); PL.b[i] = (uint32_t)ub;
    }
  }
  PH.split16();
  PH.finish<1>(16);  // planes 63..48 ready; the other three groups of 16 planes are finished on demand
  PL.split16();
  ZFB_UNROLL for (int r = 0; r < 16; r++) {
    u32x4 v;
    v.x = PL.a[2 * r]; v.y = PL.b[2 * r]; v.z = PL.a[2 * r + 1]; v.w = PL.b[2 * r + 1];
    ps.st(r, v);
    v.x = PH.a[2 * r]; v.y = PH.b[2 * r]; v.z = PH.a[2 * r + 1]; v.w = PH.b[2 * r + 1];
    ps.st(16 + r, v);
  }
  unsigned used = f3::encode_planes<64>(ps, maxbits, 2u * (uint32_t)(E + 1) + 1u, 12u, kmin, sink, tb);
  used = used < maxbits ? used : maxbits;
  sink.pad_to(sp.minbits);
  return used < sp.minbits ? sp.minbits : used;
}

// Fixed-rate form.
template <class Sink, class PR>
ZFB_HD void encode_block(const double (&f)[64], unsigned maxbits, Sink& sink, const tables& tb, const PR& ps)
{
  encode_block(f, fixed_rate_params(maxbits), sink, tb, ps);
}

// Phase 1: parse the slot; planes come back in PH (63..32) and PL (31..0).  kend = first plane not decoded.
template <class Reader, class PR>
ZFB_HD bool decode_planes(planes64& PH, planes64& PL, int& emax, int& kend, const stream_params& sp, Reader& r,
                          const tables& tb, const PR& ps)
{
  const uint32_t head = f3::fetch32(r, 0);
  emax = 0;
  kend = 63;
  if (!(head & 1u)) return false;
  emax = (int)((head >> 1) & 0x7ffu) - 1023;
  kend = f3::parse_planes<64>(ps, sp.maxbits, 12u, 64 - block_precision<3>(emax, sp), r, tb);
  ZFB_UNROLL for (int q = 0; q < 16; q++) {
    u32x4 v = ps.ld(q);
    PL.a[2 * q] = v.x; PL.b[2 * q] = v.y; PL.a[2 * q + 1] = v.z; PL.b[2 * q + 1] = v.w;
    v = ps.ld(16 + q);
    PH.a[2 * q] = v.x; PH.b[2 * q] = v.y; PH.a[2 * q + 1] = v.z; PH.b[2 * q + 1] = v.w;
  }
  return true;
}

template <class Reader, class PR>
ZFB_HD bool decode_planes(planes64& PH, planes64& PL, int& emax, int& kend, unsigned maxbits, Reader& r, const tables& tb,
                          const PR& ps)
{
  return decode_planes(PH, PL, emax, kend, fixed_rate_params(maxbits), r, tb, ps);
}

// Phase 2: planes -> coefficients -> inverse transform -> doubles.
ZFB_HD void decode_finish(double (&f)[64], planes64& PH, planes64& PL, int emax, bool nonzero, int kend)
{
  typedef traits<double> T;
  if (!nonzero) {
    ZFB_UNROLL for (int i = 0; i < 64; i++) f[i] = 0.0;
    return;
  }
  // a group of 16 planes that was never reached is all zero: its transpose stages are skipped
  // negabinary (u ^ M) - M: the XOR is folded into the last transpose stage (stages_low<2>); a group that was never
  // reached comes out of that stage as the constant 0xaaaaaaaa
  auto fill = [](planes64& P, int base) { ZFB_UNROLL for (int i = 0; i < 16; i++) { P.a[base + i] = 0xaaaaaaaau; P.b[base + i] = 0xaaaaaaaau; } };
  PH.finish<2>(16);
  if (kend < 47) PH.finish<2>(0); else fill(PH, 0);
  if (kend < 31) PL.finish<2>(16); else fill(PL, 16);
  if (kend < 15) PL.finish<2>(0); else fill(PL, 0);
  PH.split16();
  PL.split16();
  int64_t ib[64];
  ZFB_UNROLL for (int i = 0; i < 32; i++) {
    const uint64_t ua = ((uint64_t)PH.a[i] << 32) | PL.a[i], ub = ((uint64_t)PH.b[i] << 32) | PL.b[i];
    ib[order<3>::at(i)] = (int64_t)(ua - T::NBMASK);
    ib[order<3>::at(i + 32)] = (int64_t)(ub - T::NBMASK);
  }
  inv_xform<int64_t, 3>(ib);
  const double s = T::pow2(emax - 62);
  ZFB_UNROLL for (int i = 0; i < 64; i++) f[i] = s * T::to_scalar(ib[i]);
}

template <class Reader, class PR>
ZFB_HD void decode_block(double (&f)[64], const stream_params& sp, Reader& r, const tables& tb, const PR& ps)
{
  planes64 PH, PL;
  int emax, kend;
  const bool nonzero = decode_planes(PH, PL, emax, kend, sp, r, tb, ps);
  decode_finish(f, PH, PL, emax, nonzero, kend);
}

template <class Reader, class PR>
ZFB_HD void decode_block(double (&f)[64], unsigned maxbits, Reader& r, const tables& tb, const PR& ps)
{
  decode_block(f, fixed_rate_params(maxbits), r, tb, ps);
}

}  // namespace d3
}  // namespace zfb
