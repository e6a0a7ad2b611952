/*
 * oracle/zfp_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Serial CPU restatement of the zfp codec that CBench's zfp plugins call (LLNL zfp 0.5.5,
 * pinned by /root/reference/scripts/thirdparty/build_zfp.sh:6-8): fixed rate (zfp_gpu plugin)
 * and, through zfo_compress_ex / zfo_decompress_ex, fixed accuracy and fixed precision (the
 * CPU zfp plugin's "abs" / "rel", CBench/compressors/zfp/zfpCompressor.hpp:83-93,117-120).
 * zfp itself is NOT vendored in the reference tree and cannot be fetched here, so
 * this file restates the published algorithm and anchors on
 * the reference's own call sites:
 *
 *   CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp:110-118  field dims (x = LAST size, fastest)
 *   CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp:129      zfp_stream_set_rate(zfp, bits, type, dims, 8)
 *   CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp:134      zfp_stream_maximum_size
 *   CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp:145      zfp_compress   (cbytes = its return)
 *   CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp:249      zfp_decompress
 *
 * PARITY UNPINNED: the reference holds no golden vectors / known-answer tests for
 * this path and no libzfp exists in this image, so the
 * restatement is checked only through codec invariants and the documented transform
 * matrices (tests/test_oracle_codec.py).  A runtime dlopen("libzfp.so") cross-check
 * is attempted by tests/test_oracle_codec.py::test_against_real_libzfp_if_present.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this file.  The product path (vizaly-foresight_b200/csrc) never does.
 *
 * Style: one scalar loop per step, no SIMD, no threads inside the codec.  The only
 * parallelism offered is zfo_roundtrip_slabs(), which mimics CBench's MPI ranks (one
 * independent sub-field per rank, main.cpp:262-280) with OpenMP threads for the CPU
 * baseline timing.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ZFO_FLOAT 1
#define ZFO_DOUBLE 2

/* ------------------------------------------------------------------------- */
/* bit stream: 64-bit little-endian words, bits appended LSB first (A.9)      */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint64_t *words; /* output / input words                         */
  size_t pos;      /* absolute bit position of the next bit         */
} bitio;

static void put_bit(bitio *s, unsigned bit)
{
  if (bit)
    s->words[s->pos >> 6] |= (uint64_t)1 << (s->pos & 63);
  s->pos++;
}

/* write the n (0..64) low bits of v, LSB first; return v >> n.  The buffer is
 * pre-zeroed, so bits are OR-ed in a word at a time (zfp's stream_write_bits
 * also works a word at a time; the result is the same bit string).            */
static uint64_t put_bits(bitio *s, uint64_t v, unsigned n)
{
  if (n) {
    size_t w = s->pos >> 6;
    unsigned sh = (unsigned)(s->pos & 63);
    uint64_t val = n < 64 ? v & (((uint64_t)1 << n) - 1) : v;
    s->words[w] |= val << sh;
    if (sh + n > 64)
      s->words[w + 1] |= val >> (64 - sh);
    s->pos += n;
  }
  return n < 64 ? v >> n : 0;
}

static unsigned get_bit(bitio *s)
{
  unsigned b = (unsigned)((s->words[s->pos >> 6] >> (s->pos & 63)) & 1u);
  s->pos++;
  return b;
}

static uint64_t get_bits(bitio *s, unsigned n)
{
  uint64_t v = 0;
  if (n) {
    size_t w = s->pos >> 6;
    unsigned sh = (unsigned)(s->pos & 63);
    v = s->words[w] >> sh;
    if (sh + n > 64)
      v |= s->words[w + 1] << (64 - sh);
    if (n < 64)
      v &= ((uint64_t)1 << n) - 1;
    s->pos += n;
  }
  return v;
}

/* ------------------------------------------------------------------------- */
/* A.1  rate -> per-block bit budget (zfp_stream_set_rate)                     */
/* ------------------------------------------------------------------------- */
unsigned zfo_maxbits(int type, int dims, double rate, int wra)
{
  unsigned n = 1u << (2 * dims);
  unsigned bits = (unsigned)floor(n * rate + 0.5);
  unsigned floor_bits = (type == ZFO_FLOAT) ? 1 + 8u : 1 + 11u;
  if (bits < floor_bits)
    bits = floor_bits;
  if (wra) {
    bits += 63u;
    bits &= ~63u;
  }
  return bits;
}

static size_t nblocks(int dims, size_t nx, size_t ny, size_t nz)
{
  size_t b = (nx + 3) / 4;
  if (dims >= 2) b *= (ny + 3) / 4;
  if (dims >= 3) b *= (nz + 3) / 4;
  return b;
}

/* bytes zfp_compress returns: ceil(B*maxbits/64) words (stream flush pads)   */
size_t zfo_stream_bytes(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits)
{
  size_t bits = nblocks(dims, nx, ny, nz) * (size_t)maxbits;
  return ((bits + 63) / 64) * 8;
}

/* bytes zfp_stream_maximum_size returns for a fixed-rate stream (header room of
 * 148 bits included, zfpCompressorGpu.hpp:134 mallocs this much)              */
size_t zfo_stream_capacity(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits)
{
  size_t bits = 148 + nblocks(dims, nx, ny, nz) * (size_t)maxbits;
  return ((bits + 63) & ~(size_t)63) / 8;
}

/* ------------------------------------------------------------------------- */
/* stream parameters (zfp_stream: minbits, maxbits, maxprec, minexp) and the   */
/* three mode setters of zfp 0.5.5 [zfp-recall: src/zfp.c]                      */
/* ------------------------------------------------------------------------- */
#define ZFO_MIN_BITS 1u
#define ZFO_MAX_BITS 16657u
#define ZFO_MAX_PREC 64u
#define ZFO_MIN_EXP (-1074)

typedef struct {
  unsigned minbits, maxbits, maxprec;
  int minexp;
} zfo_params;

/* zfp_stream_set_rate(zfp, rate, type, dims, wra): zfpCompressorGpu.hpp:129 */
void zfo_set_rate(zfo_params *zp, int type, int dims, double rate, int wra)
{
  zp->minbits = zp->maxbits = zfo_maxbits(type, dims, rate, wra);
  zp->maxprec = ZFO_MAX_PREC;
  zp->minexp = ZFO_MIN_EXP;
}

/* zfp_stream_set_accuracy(zfp, tolerance): zfp/zfpCompressor.hpp:118 ("abs").
 * tolerance = x * 2^emin with 0.5 <= x < 1, then emin--, so 2^minexp <= tolerance */
void zfo_set_accuracy(zfo_params *zp, double tolerance)
{
  int emin = ZFO_MIN_EXP;
  if (tolerance > 0) {
    frexp(tolerance, &emin);
    emin--;
  }
  zp->minbits = ZFO_MIN_BITS;
  zp->maxbits = ZFO_MAX_BITS;
  zp->maxprec = ZFO_MAX_PREC;
  zp->minexp = emin;
}

/* zfp_stream_set_precision(zfp, precision): zfp/zfpCompressor.hpp:120 ("rel") */
void zfo_set_precision(zfo_params *zp, unsigned precision)
{
  zp->minbits = ZFO_MIN_BITS;
  zp->maxbits = ZFO_MAX_BITS;
  zp->maxprec = precision ? (precision < ZFO_MAX_PREC ? precision : ZFO_MAX_PREC) : ZFO_MAX_PREC;
  zp->minexp = ZFO_MIN_EXP;
}

/* zfp_stream_maximum_size(zfp, field): zfp/zfpCompressor.hpp:123 mallocs this much */
size_t zfo_maximum_size(int type, int dims, size_t nx, size_t ny, size_t nz, const zfo_params *zp)
{
  unsigned values = 1u << (2 * dims);
  unsigned tprec = (type == ZFO_FLOAT) ? 32u : 64u;
  unsigned maxbits = 1 + ((type == ZFO_FLOAT) ? 8u : 11u);
  maxbits += values - 1 + values * (zp->maxprec < tprec ? zp->maxprec : tprec);
  if (maxbits > zp->maxbits) maxbits = zp->maxbits;
  if (maxbits < zp->minbits) maxbits = zp->minbits;
  return ((148 + nblocks(dims, nx, ny, nz) * (size_t)maxbits + 63) & ~(size_t)63) / 8;
}

/* ------------------------------------------------------------------------- */
/* A.7  sequency orderings                                                     */
/* ------------------------------------------------------------------------- */
static const unsigned char PERM1[4] = {0, 1, 2, 3};
static const unsigned char PERM2[16] = {0, 1, 4, 5, 2, 8, 6, 9, 3, 12, 10, 7, 13, 11, 14, 15};
static const unsigned char PERM3[64] = {
    0,  1,  4,  16, 20, 17, 5,  2,  8,  32, 21, 6,  18, 24, 9,  33, 36, 3,  12, 48, 22, 25,
    37, 40, 34, 10, 7,  19, 28, 13, 49, 52, 41, 38, 26, 23, 29, 53, 11, 35, 44, 14, 50, 56,
    42, 27, 39, 45, 30, 54, 57, 60, 51, 15, 43, 46, 58, 61, 55, 31, 62, 59, 47, 63};

const unsigned char *zfo_perm(int dims)
{
  return dims == 1 ? PERM1 : dims == 2 ? PERM2 : PERM3;
}

/* A.3  pad a partially filled run of 4 (n valid values, stride s) */
#define DEFINE_PAD(NAME, T)                              \
  static void NAME(T *p, unsigned n, unsigned s)         \
  {                                                      \
    switch (n) {                                         \
    case 0: p[0 * s] = 0;        /* fall through */      \
    case 1: p[1 * s] = p[0 * s]; /* fall through */      \
    case 2: p[2 * s] = p[1 * s]; /* fall through */      \
    case 3: p[3 * s] = p[0 * s]; /* fall through */      \
    default: break;                                      \
    }                                                    \
  }
DEFINE_PAD(pad_f, float)
DEFINE_PAD(pad_d, double)

/* ------------------------------------------------------------------------- */
/* The codec is written once per scalar type through a macro "template".       */
/* Scalar/Int/UInt, EBITS, EBIAS, NBMASK, precision p = bits of Int.           */
/* ------------------------------------------------------------------------- */
#define DEFINE_CODEC(SFX, Scalar, Int, UInt, EBITS, EBIAS, NBMASK, FREXP, LDEXP, FABS, PAD)          \
                                                                                                      \
  /* A.6 forward lifting of 4 values at stride s */                                                   \
  static void fwd_lift_##SFX(Int *p, unsigned s)                                                      \
  {                                                                                                   \
    Int x = p[0], y = p[s], z = p[2 * s], w = p[3 * s];                                               \
    x += w; x >>= 1; w -= x;                                                                          \
    z += y; z >>= 1; y -= z;                                                                          \
    x += z; x >>= 1; z -= x;                                                                          \
    w += y; w >>= 1; y -= w;                                                                          \
    w += y >> 1; y -= w >> 1;                                                                         \
    p[0] = x; p[s] = y; p[2 * s] = z; p[3 * s] = w;                                                   \
  }                                                                                                   \
                                                                                                      \
  static void inv_lift_##SFX(Int *p, unsigned s)                                                      \
  {                                                                                                   \
    Int x = p[0], y = p[s], z = p[2 * s], w = p[3 * s];                                               \
    y += w >> 1; w -= y >> 1;                                                                         \
    y += w; w = (Int)((UInt)w << 1); w -= y;                                                          \
    z += x; x = (Int)((UInt)x << 1); x -= z;                                                          \
    y += z; z = (Int)((UInt)z << 1); z -= y;                                                          \
    w += x; x = (Int)((UInt)x << 1); x -= w;                                                          \
    p[0] = x; p[s] = y; p[2 * s] = z; p[3 * s] = w;                                                   \
  }                                                                                                   \
                                                                                                      \
  static void fwd_xform_##SFX(Int *p, int dims)                                                       \
  {                                                                                                   \
    unsigned x, y, z;                                                                                 \
    if (dims == 1) {                                                                                  \
      fwd_lift_##SFX(p, 1);                                                                           \
    } else if (dims == 2) {                                                                           \
      for (y = 0; y < 4; y++) fwd_lift_##SFX(p + 4 * y, 1);                                           \
      for (x = 0; x < 4; x++) fwd_lift_##SFX(p + x, 4);                                               \
    } else {                                                                                          \
      for (z = 0; z < 4; z++) for (y = 0; y < 4; y++) fwd_lift_##SFX(p + 4 * y + 16 * z, 1);          \
      for (x = 0; x < 4; x++) for (z = 0; z < 4; z++) fwd_lift_##SFX(p + 16 * z + x, 4);              \
      for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) fwd_lift_##SFX(p + x + 4 * y, 16);              \
    }                                                                                                 \
  }                                                                                                   \
                                                                                                      \
  static void inv_xform_##SFX(Int *p, int dims)                                                       \
  {                                                                                                   \
    unsigned x, y, z;                                                                                 \
    if (dims == 1) {                                                                                  \
      inv_lift_##SFX(p, 1);                                                                           \
    } else if (dims == 2) {                                                                           \
      for (x = 0; x < 4; x++) inv_lift_##SFX(p + x, 4);                                               \
      for (y = 0; y < 4; y++) inv_lift_##SFX(p + 4 * y, 1);                                           \
    } else {                                                                                          \
      for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) inv_lift_##SFX(p + x + 4 * y, 16);              \
      for (x = 0; x < 4; x++) for (z = 0; z < 4; z++) inv_lift_##SFX(p + 16 * z + x, 4);              \
      for (z = 0; z < 4; z++) for (y = 0; y < 4; y++) inv_lift_##SFX(p + 4 * y + 16 * z, 1);          \
    }                                                                                                 \
  }                                                                                                   \
                                                                                                      \
  /* A.8 embedded coder: encode `size` negabinary coefficients within `maxbits` */                    \
  static unsigned encode_ints_##SFX(bitio *s, unsigned maxbits, unsigned maxprec, const UInt *data,   \
                                    unsigned size)                                                    \
  {                                                                                                   \
    unsigned intprec = 8 * (unsigned)sizeof(UInt);                                                    \
    unsigned kmin = intprec > maxprec ? intprec - maxprec : 0;                                        \
    unsigned bits = maxbits;                                                                          \
    unsigned i, k, m, n;                                                                              \
    uint64_t x;                                                                                       \
    for (k = intprec, n = 0; bits && k-- > kmin;) {                                                   \
      x = 0;                                                                                          \
      for (i = 0; i < size; i++)                                                                      \
        x += (uint64_t)((data[i] >> k) & 1u) << i;                                                    \
      m = n < bits ? n : bits;                                                                        \
      bits -= m;                                                                                      \
      x = put_bits(s, x, m);                                                                          \
      for (; n < size && bits && (bits--, put_bit(s, !!x), !!x); x >>= 1, n++)                        \
        for (; n < size - 1 && bits && (bits--, put_bit(s, (unsigned)(x & 1u)), !(x & 1u));           \
             x >>= 1, n++)                                                                            \
          ;                                                                                           \
    }                                                                                                 \
    return maxbits - bits;                                                                            \
  }                                                                                                   \
                                                                                                      \
  static unsigned decode_ints_##SFX(bitio *s, unsigned maxbits, unsigned maxprec, UInt *data,         \
                                    unsigned size)                                                    \
  {                                                                                                   \
    unsigned intprec = 8 * (unsigned)sizeof(UInt);                                                    \
    unsigned kmin = intprec > maxprec ? intprec - maxprec : 0;                                        \
    unsigned bits = maxbits;                                                                          \
    unsigned i, k, m, n;                                                                              \
    uint64_t x;                                                                                       \
    for (i = 0; i < size; i++)                                                                        \
      data[i] = 0;                                                                                    \
    for (k = intprec, n = 0; bits && k-- > kmin;) {                                                   \
      m = n < bits ? n : bits;                                                                        \
      bits -= m;                                                                                      \
      x = get_bits(s, m);                                                                             \
      for (; n < size && bits && (bits--, get_bit(s)); x += (uint64_t)1 << n++)                       \
        for (; n < size - 1 && bits && (bits--, !get_bit(s)); n++)                                    \
          ;                                                                                           \
      for (i = 0; x; i++, x >>= 1)                                                                    \
        data[i] += (UInt)(x & 1u) << k;                                                               \
    }                                                                                                 \
    return maxbits - bits;                                                                            \
  }                                                                                                   \
                                                                                                      \
  /* A.4 block exponent */                                                                            \
  static int exponent_block_##SFX(const Scalar *p, unsigned n)                                        \
  {                                                                                                   \
    Scalar max = 0;                                                                                   \
    do {                                                                                              \
      Scalar f = FABS(*p++);                                                                          \
      if (max < f)                                                                                    \
        max = f;                                                                                      \
    } while (--n);                                                                                    \
    if (max > 0) {                                                                                    \
      int e;                                                                                          \
      FREXP(max, &e);                                                                                 \
      return e > 1 - EBIAS ? e : 1 - EBIAS;                                                           \
    }                                                                                                 \
    return -EBIAS;                                                                                    \
  }                                                                                                   \
                                                                                                      \
  /* precision(maxexp, maxprec, minexp, dims) = MIN(maxprec, MAX(0, maxexp - minexp + 2*(dims+1))) */  \
  static unsigned precision_##SFX(int maxexp, const zfo_params *zp, int dims)                         \
  {                                                                                                   \
    int p = maxexp - zp->minexp + 2 * (dims + 1);                                                     \
    if (p < 0) p = 0;                                                                                 \
    return (unsigned)p > zp->maxprec ? zp->maxprec : (unsigned)p;                                     \
  }                                                                                                   \
                                                                                                      \
  /* A.5 encode one full block of 4^dims scalars: at least minbits, at most maxbits bits */            \
  static unsigned encode_block_##SFX(bitio *s, const Scalar *fblock, int dims, const zfo_params *zp)  \
  {                                                                                                   \
    unsigned size = 1u << (2 * dims);                                                                 \
    size_t start = s->pos;                                                                            \
    unsigned bits = 1;                                                                                \
    int emax = exponent_block_##SFX(fblock, size);                                                    \
    unsigned maxprec = precision_##SFX(emax, zp, dims);                                               \
    unsigned e = maxprec ? (unsigned)(emax + EBIAS) : 0;                                              \
    if (e) {                                                                                          \
      Int iblock[64];                                                                                 \
      UInt ublock[64];                                                                                \
      const unsigned char *perm = zfo_perm(dims);                                                     \
      unsigned i;                                                                                     \
      Scalar sc = LDEXP((Scalar)1, (8 * (int)sizeof(Scalar) - 2) - emax);                             \
      bits += EBITS;                                                                                  \
      put_bits(s, 2 * (uint64_t)e + 1, bits);                                                         \
      for (i = 0; i < size; i++) {                                                                    \
        const Scalar prod = sc * fblock[i]; /* product rounded to Scalar (SSE: no excess precision), then C truncation */ \
        iblock[i] = (Int)prod;                                                                        \
      }                                                                                               \
      fwd_xform_##SFX(iblock, dims);                                                                  \
      for (i = 0; i < size; i++)                                                                      \
        ublock[i] = ((UInt)iblock[perm[i]] + NBMASK) ^ NBMASK;                                        \
      bits += encode_ints_##SFX(s, zp->maxbits - bits, maxprec, ublock, size);                        \
    } else {                                                                                          \
      put_bit(s, 0);                                                                                  \
    }                                                                                                 \
    if (bits < zp->minbits)                                                                           \
      bits = zp->minbits; /* stream_pad: the buffer is pre-zeroed */                                  \
    s->pos = start + bits;                                                                            \
    return bits;                                                                                      \
  }                                                                                                   \
                                                                                                      \
  static unsigned decode_block_##SFX(bitio *s, Scalar *fblock, int dims, const zfo_params *zp)        \
  {                                                                                                   \
    unsigned size = 1u << (2 * dims);                                                                 \
    size_t start = s->pos;                                                                            \
    unsigned bits = 1;                                                                                \
    unsigned i;                                                                                       \
    if (get_bit(s)) {                                                                                 \
      Int iblock[64];                                                                                 \
      UInt ublock[64];                                                                                \
      const unsigned char *perm = zfo_perm(dims);                                                     \
      int emax;                                                                                       \
      unsigned maxprec;                                                                               \
      Scalar sc;                                                                                      \
      bits += EBITS;                                                                                  \
      emax = (int)get_bits(s, EBITS) - EBIAS;                                                         \
      maxprec = precision_##SFX(emax, zp, dims);                                                      \
      bits += decode_ints_##SFX(s, zp->maxbits - bits, maxprec, ublock, size);                        \
      for (i = 0; i < size; i++)                                                                      \
        iblock[perm[i]] = (Int)((ublock[i] ^ NBMASK) - NBMASK);                                       \
      inv_xform_##SFX(iblock, dims);                                                                  \
      sc = LDEXP((Scalar)1, emax - (8 * (int)sizeof(Scalar) - 2));                                    \
      for (i = 0; i < size; i++) {                                                                    \
        const Scalar conv = (Scalar)iblock[i];                                                        \
        const Scalar prod = sc * conv;                                                                \
        fblock[i] = prod;                                                                             \
      }                                                                                               \
    } else {                                                                                          \
      for (i = 0; i < size; i++)                                                                      \
        fblock[i] = 0;                                                                                \
    }                                                                                                 \
    if (bits < zp->minbits)                                                                           \
      bits = zp->minbits; /* stream_skip */                                                           \
    s->pos = start + bits;                                                                            \
    return bits;                                                                                      \
  }                                                                                                   \
                                                                                                      \
  /* A.2/A.3 gather one (possibly partial) block at (x,y,z) */                                        \
  static void gather_##SFX(Scalar *q, const Scalar *data, int dims, size_t nx, size_t ny, size_t nz,  \
                           size_t x, size_t y, size_t z)                                              \
  {                                                                                                   \
    unsigned mx = (unsigned)(nx - x < 4 ? nx - x : 4);                                                \
    unsigned my = dims >= 2 ? (unsigned)(ny - y < 4 ? ny - y : 4) : 1;                                \
    unsigned mz = dims >= 3 ? (unsigned)(nz - z < 4 ? nz - z : 4) : 1;                                \
    unsigned i, j, k;                                                                                 \
    for (k = 0; k < mz; k++) {                                                                        \
      for (j = 0; j < my; j++) {                                                                      \
        for (i = 0; i < mx; i++)                                                                      \
          q[16 * k + 4 * j + i] = data[(x + i) + nx * ((y + j) + ny * (z + k))];                      \
        PAD(q + 16 * k + 4 * j, mx, 1);                                                               \
      }                                                                                               \
      if (dims >= 2)                                                                                  \
        for (i = 0; i < 4; i++)                                                                       \
          PAD(q + 16 * k + i, my, 4);                                                                 \
    }                                                                                                 \
    if (dims >= 3)                                                                                    \
      for (j = 0; j < 4; j++)                                                                         \
        for (i = 0; i < 4; i++)                                                                       \
          PAD(q + 4 * j + i, mz, 16);                                                                 \
  }                                                                                                   \
                                                                                                      \
  static void scatter_##SFX(const Scalar *q, Scalar *data, int dims, size_t nx, size_t ny, size_t nz, \
                            size_t x, size_t y, size_t z)                                             \
  {                                                                                                   \
    unsigned mx = (unsigned)(nx - x < 4 ? nx - x : 4);                                                \
    unsigned my = dims >= 2 ? (unsigned)(ny - y < 4 ? ny - y : 4) : 1;                                \
    unsigned mz = dims >= 3 ? (unsigned)(nz - z < 4 ? nz - z : 4) : 1;                                \
    unsigned i, j, k;                                                                                 \
    for (k = 0; k < mz; k++)                                                                          \
      for (j = 0; j < my; j++)                                                                        \
        for (i = 0; i < mx; i++)                                                                      \
          data[(x + i) + nx * ((y + j) + ny * (z + k))] = q[16 * k + 4 * j + i];                      \
  }                                                                                                   \
                                                                                                      \
  /* zfp_compress: blocks in z, y, x order, then stream_flush (pad to a 64-bit word); returns the  */  \
  /* stream size in bytes.  `out` must be pre-sized to zfo_maximum_size() and is zeroed here.      */  \
  static size_t compress_##SFX(const Scalar *data, int dims, size_t nx, size_t ny, size_t nz,         \
                               const zfo_params *zp, uint64_t *out, size_t cap_bytes,                 \
                               uint64_t *block_offsets)                                               \
  {                                                                                                   \
    bitio s;                                                                                          \
    size_t x, y, z, b = 0;                                                                            \
    Scalar q[64];                                                                                     \
    if (dims < 2) ny = 1;                                                                             \
    if (dims < 3) nz = 1;                                                                             \
    memset(out, 0, cap_bytes);                                                                        \
    s.words = out;                                                                                    \
    s.pos = 0;                                                                                        \
    for (z = 0; z < nz; z += 4)                                                                       \
      for (y = 0; y < ny; y += 4)                                                                     \
        for (x = 0; x < nx; x += 4) {                                                                 \
          if (block_offsets) block_offsets[b] = s.pos;                                                \
          b++;                                                                                        \
          gather_##SFX(q, data, dims, nx, ny, nz, x, y, z);                                           \
          encode_block_##SFX(&s, q, dims, zp);                                                        \
        }                                                                                             \
    if (block_offsets) block_offsets[b] = s.pos;                                                      \
    return ((s.pos + 63) / 64) * 8;                                                                   \
  }                                                                                                   \
                                                                                                      \
  static size_t decompress_##SFX(Scalar *data, int dims, size_t nx, size_t ny, size_t nz,             \
                                 const zfo_params *zp, const uint64_t *in)                            \
  {                                                                                                   \
    bitio s;                                                                                          \
    size_t x, y, z;                                                                                   \
    Scalar q[64];                                                                                     \
    if (dims < 2) ny = 1;                                                                             \
    if (dims < 3) nz = 1;                                                                             \
    s.words = (uint64_t *)in;                                                                         \
    s.pos = 0;                                                                                        \
    for (z = 0; z < nz; z += 4)                                                                       \
      for (y = 0; y < ny; y += 4)                                                                     \
        for (x = 0; x < nx; x += 4) {                                                                 \
          decode_block_##SFX(&s, q, dims, zp);                                                        \
          scatter_##SFX(q, data, dims, nx, ny, nz, x, y, z);                                          \
        }                                                                                             \
    return ((s.pos + 63) / 64) * 8;                                                                   \
  }

DEFINE_CODEC(f, float, int32_t, uint32_t, 8, 127, 0xaaaaaaaau, frexpf, ldexpf, fabsf, pad_f)
DEFINE_CODEC(d, double, int64_t, uint64_t, 11, 1023, 0xaaaaaaaaaaaaaaaaull, frexp, ldexp, fabs, pad_d)

/* ------------------------------------------------------------------------- */
/* public entry points (ctypes-friendly)                                       */
/* ------------------------------------------------------------------------- */

/* General form: any (minbits, maxbits, maxprec, minexp).  `out` must hold cap_bytes >= zfo_maximum_size();
 * returns the bytes zfp_compress returns (stream flushed to a 64-bit word), 0 on bad arguments.
 * block_offsets (nullable): nblocks+1 absolute bit positions, [b] = first bit of block b.            */
size_t zfo_compress_ex(int type, int dims, size_t nx, size_t ny, size_t nz, const zfo_params *zp,
                       const void *in, void *out, size_t cap_bytes, uint64_t *block_offsets)
{
  if (dims < 1 || dims > 3 || !nx || !zp)
    return 0;
  if (type == ZFO_FLOAT) {
    if (zp->maxbits < 9) return 0;
    return compress_f((const float *)in, dims, nx, ny, nz, zp, (uint64_t *)out, cap_bytes, block_offsets);
  }
  if (type == ZFO_DOUBLE) {
    if (zp->maxbits < 12) return 0;
    return compress_d((const double *)in, dims, nx, ny, nz, zp, (uint64_t *)out, cap_bytes, block_offsets);
  }
  return 0;
}

/* returns the bytes consumed (zfp_decompress's return) */
size_t zfo_decompress_ex(int type, int dims, size_t nx, size_t ny, size_t nz, const zfo_params *zp,
                         const void *in, void *out)
{
  if (dims < 1 || dims > 3 || !nx || !zp)
    return 0;
  if (type == ZFO_FLOAT)
    return decompress_f((float *)out, dims, nx, ny, nz, zp, (const uint64_t *)in);
  if (type == ZFO_DOUBLE)
    return decompress_d((double *)out, dims, nx, ny, nz, zp, (const uint64_t *)in);
  return 0;
}

/* Fixed-rate form (zfp_gpu plugin): returns compressed bytes (what zfp_compress returns), 0 on bad
 * arguments.  `out` must hold zfo_stream_bytes() bytes.                                             */
size_t zfo_compress(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                    const void *in, void *out)
{
  zfo_params zp;
  zp.minbits = zp.maxbits = maxbits;
  zp.maxprec = ZFO_MAX_PREC;
  zp.minexp = ZFO_MIN_EXP;
  if (dims < 1 || dims > 3 || !nx)
    return 0;
  return zfo_compress_ex(type, dims, nx, ny, nz, &zp, in, out, zfo_stream_bytes(dims, nx, ny, nz, maxbits), NULL);
}

size_t zfo_decompress(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                      const void *in, void *out)
{
  zfo_params zp;
  zp.minbits = zp.maxbits = maxbits;
  zp.maxprec = ZFO_MAX_PREC;
  zp.minexp = ZFO_MIN_EXP;
  if (dims < 1 || dims > 3 || !nx || (type != ZFO_FLOAT && type != ZFO_DOUBLE))
    return 0;
  zfo_decompress_ex(type, dims, nx, ny, nz, &zp, in, out);
  return zfo_stream_bytes(dims, nx, ny, nz, maxbits);
}

void zfo_fwd_xform_i32(int32_t *p, int dims) { fwd_xform_f(p, dims); }
void zfo_inv_xform_i32(int32_t *p, int dims) { inv_xform_f(p, dims); }
void zfo_fwd_xform_i64(int64_t *p, int dims) { fwd_xform_d(p, dims); }
void zfo_inv_xform_i64(int64_t *p, int dims) { inv_xform_d(p, dims); }

unsigned zfo_encode_ints_u32(const uint32_t *data, unsigned size, unsigned maxbits, unsigned maxprec,
                             uint64_t *out /* >= maxbits/64+2 words, zeroed by callee */)
{
  bitio s;
  memset(out, 0, ((size_t)maxbits / 64 + 2) * 8);
  s.words = out;
  s.pos = 0;
  return encode_ints_f(&s, maxbits, maxprec, data, size);
}

unsigned zfo_decode_ints_u32(uint32_t *data, unsigned size, unsigned maxbits, unsigned maxprec,
                             const uint64_t *in)
{
  bitio s;
  s.words = (uint64_t *)in;
  s.pos = 0;
  return decode_ints_f(&s, maxbits, maxprec, data, size);
}

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* CPU-baseline helper: `nslabs` independent sub-fields (CBench ranks, main.cpp:262-280),
 * each slab compressed then decompressed by one thread.  Slab i starts at
 * in + i*slab_elems (same for out), streams at stream + i*slab_stream_bytes.
 * Returns wall seconds for [compress all, decompress all]; t_enc/t_dec get the
 * max-over-"ranks" of each phase (CBench reports min throughput = max time).      */
double zfo_roundtrip_slabs(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits,
                           const void *in, void *stream, void *out, int nslabs, int nthreads,
                           double *t_enc, double *t_dec)
{
  size_t esz = type == ZFO_FLOAT ? 4 : 8;
  size_t slab_elems = nx * (dims >= 2 ? ny : 1) * (dims >= 3 ? nz : 1);
  size_t sbytes = zfo_stream_bytes(dims, nx, ny, nz, maxbits);
  double enc_max = 0, dec_max = 0;
  double t0 = now_s();
  int i;
#ifdef _OPENMP
  if (nthreads > 0)
    omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(max : enc_max, dec_max)
  for (i = 0; i < nslabs; i++) {
    const char *src = (const char *)in + (size_t)i * slab_elems * esz;
    char *dst = (char *)out + (size_t)i * slab_elems * esz;
    char *st = (char *)stream + (size_t)i * sbytes;
    double a = now_s(), b, c;
    zfo_compress(type, dims, nx, ny, nz, maxbits, src, st);
    b = now_s();
    zfo_decompress(type, dims, nx, ny, nz, maxbits, st, dst);
    c = now_s();
    if (b - a > enc_max) enc_max = b - a;
    if (c - b > dec_max) dec_max = c - b;
  }
  if (t_enc) *t_enc = enc_max;
  if (t_dec) *t_dec = dec_max;
  return now_s() - t0;
}
