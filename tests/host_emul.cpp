// tests/host_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the device block codec
// (vizaly-foresight_b200/csrc/zfb_codec.cuh, which is __host__ __device__) for the CPU so that the
// exact kernel arithmetic can be bit-compared with the oracle in this GPU-less container.
// Never linked into the product library.
#include <cstring>
#include <cstdint>
#include <cstddef>
#include "../vizaly-foresight_b200/csrc/zfb_codec.cuh"
#include "../vizaly-foresight_b200/csrc/zfb_codec_f3.cuh"
#include "../vizaly-foresight_b200/csrc/zfb_codec_f1.cuh"
#include "../vizaly-foresight_b200/csrc/zfb_codec_d3.cuh"
#include <type_traits>

using namespace zfb;

static int g_use_f3 = 1;  // float/3D goes through the fast coder (zfb_codec_f3.cuh) unless switched off
static uint32_t g_enc[f3::ENC_WORDS];
static uint32_t g_dec[f3::DEC_WORDS];
static f3::tables host_tables()
{
  static bool built = false;
  if (!built) {
    for (unsigned i = 0; i < f3::ENC_WORDS; i++) g_enc[i] = f3::enc_table_word(i);
    for (unsigned i = 0; i < f3::DEC_WORDS; i++) g_dec[i] = f3::dec_table_word(i);
    built = true;
  }
  f3::tables t;
  t.enc = g_enc; t.dec = g_dec; t.enc_s = t.dec_s = 0;
  return t;
}

static uint16_t g1_enc[64], g1_dec[512];
static uint32_t g1_spread[256], g1_unspread[256];
static f1::tables host_tables1()
{
  static bool built = false;
  if (!built) {
    for (unsigned i = 0; i < 64; i++) g1_enc[i] = f1::enc_entry(i);
    for (unsigned i = 0; i < 512; i++) g1_dec[i] = f1::dec_entry(i);
    for (unsigned i = 0; i < 256; i++) { g1_spread[i] = f1::spread_entry(i); g1_unspread[i] = f1::unspread_entry(i); }
    built = true;
  }
  f1::tables t;
  t.enc = g1_enc; t.dec = g1_dec; t.spread = g1_spread; t.unspread = g1_unspread;
  return t;
}

template <typename Scalar, int DIMS>
static void run(bool enc, size_t nx, size_t ny, size_t nz, unsigned maxbits, Scalar* data, uint64_t* stream,
                size_t stream_words)
{
  constexpr int BS = 1 << (2 * DIMS);
  if (DIMS < 2) ny = 1;
  if (DIMS < 3) nz = 1;
  const size_t bx = (nx + 3) / 4, by = (ny + 3) / 4, bz = (nz + 3) / 4;
  const bool aligned = (maxbits % 64) == 0;
  if (enc) std::memset(stream, 0, stream_words * 8);
  for (size_t b = 0; b < bx * by * bz; b++) {
    const size_t ix = b % bx, iy = (b / bx) % by, iz = b / (bx * by);
    const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
    const unsigned mx = (unsigned)(nx - x0 < 4 ? nx - x0 : 4);
    const unsigned my = DIMS >= 2 ? (unsigned)(ny - y0 < 4 ? ny - y0 : 4) : 1;
    const unsigned mz = DIMS >= 3 ? (unsigned)(nz - z0 < 4 ? nz - z0 : 4) : 1;
    Scalar f[BS];
    const size_t startbit = b * (size_t)maxbits;
    if (enc) {
      for (int i = 0; i < BS; i++) f[i] = 0;
      for (unsigned k = 0; k < mz; k++)
        for (unsigned j = 0; j < my; j++)
          for (unsigned i = 0; i < mx; i++) f[16 * k + 4 * j + i] = data[(x0 + i) + nx * ((y0 + j) + ny * (z0 + k))];
      if (mx < 4 || my < (DIMS >= 2 ? 4u : 1u) || mz < (DIMS >= 3 ? 4u : 1u)) pad_block<Scalar, DIMS>(f, mx, my, mz);
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 1) {
        if (g_use_f3 && mx == 4 && (maxbits % 32) == 0 && maxbits <= 128) {
          const f1::tables tb = host_tables1();
          uint32_t* s32 = reinterpret_cast<uint32_t*>(stream) + startbit / 32;
          // the same instantiations as the kernels: <1> for 32-bit slots, <2> for 64, <4> for 96 / 128
          if (maxbits == 32) { uint32_t o[1]; f1::encode_block<1>(f, maxbits, o, tb); s32[0] = o[0]; }
          else if (maxbits == 64) { uint32_t o[2]; f1::encode_block<2>(f, maxbits, o, tb); s32[0] = o[0]; s32[1] = o[1]; }
          else {
            uint32_t o[4];
            f1::encode_block<4>(f, maxbits, o, tb);
            for (unsigned i = 0; i < maxbits / 32; i++) s32[i] = o[i];
          }
          continue;
        }
      }
      if constexpr (std::is_same<Scalar, double>::value && DIMS == 3) {
        if (g_use_f3) {
          const f3::tables tb = host_tables();
          uint32_t* s32 = reinterpret_cast<uint32_t*>(stream);
          f3::u32x4 rows[32];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          if (aligned) {
            f3::sink32 sink(s32 + startbit / 32, maxbits / 32);
            d3::encode_block(f, maxbits, sink, tb, ps);
          } else {
            f3::or_sink32 sink(s32, startbit, maxbits);
            d3::encode_block(f, maxbits, sink, tb, ps);
          }
          continue;
        }
      }
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 3) {
        if (g_use_f3) {
          const f3::tables tb = host_tables();
          uint32_t* s32 = reinterpret_cast<uint32_t*>(stream);
          f3::u32x4 rows[16];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          if (aligned) {
            f3::sink32 sink(s32 + startbit / 32, maxbits / 32);
            f3::encode_block(f, maxbits, sink, tb, ps);
          } else {
            f3::or_sink32 sink(s32, startbit, maxbits);
            f3::encode_block(f, maxbits, sink, tb, ps);
          }
          continue;
        }
      }
      if (aligned) {
        aligned_sink sink(stream + startbit / 64, maxbits / 64);
        encode_block<Scalar, DIMS>(f, maxbits, sink);
      } else {
        or_sink sink(stream, startbit, maxbits);
        encode_block<Scalar, DIMS>(f, maxbits, sink);
      }
    } else {
      bool done = false;
      if constexpr (std::is_same<Scalar, double>::value && DIMS == 3) {
        if (g_use_f3) {
          const f3::tables tb = host_tables();
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
          f3::reader32 r(s32 + (startbit >> 5), (unsigned)(2 * stream_words - (startbit >> 5)), (unsigned)(startbit & 31));
          f3::u32x4 rows[32];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          d3::decode_block(f, maxbits, r, tb, ps);
          done = true;
        }
      }
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 1) {
        if (g_use_f3 && mx == 4 && (maxbits % 32) == 0 && maxbits <= 128) {
          const f1::tables tb = host_tables1();
          uint32_t in4[4] = {0, 0, 0, 0};
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream) + startbit / 32;
          for (unsigned i = 0; i < maxbits / 32; i++) in4[i] = s32[i];
          if (maxbits == 32) { const uint32_t in1[1] = {in4[0]}; f1::decode_block<1>(f, maxbits, in1, tb); }
          else if (maxbits == 64) { const uint32_t in2[2] = {in4[0], in4[1]}; f1::decode_block<2>(f, maxbits, in2, tb); }
          else f1::decode_block<4>(f, maxbits, in4, tb);
          done = true;
        }
      }
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 3) {
        if (g_use_f3) {
          const f3::tables tb = host_tables();
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
          f3::reader32 r(s32 + (startbit >> 5), (unsigned)(2 * stream_words - (startbit >> 5)), (unsigned)(startbit & 31));
          f3::u32x4 rows[16];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          f3::decode_block(f, maxbits, r, tb, ps);
          done = true;
        }
      }
      if (!done) {
        slot_reader r(stream + (startbit >> 6), (unsigned)(stream_words - (startbit >> 6)), (unsigned)(startbit & 63));
        decode_block<Scalar, DIMS>(f, maxbits, r);
      }
      for (unsigned k = 0; k < mz; k++)
        for (unsigned j = 0; j < my; j++)
          for (unsigned i = 0; i < mx; i++) data[(x0 + i) + nx * ((y0 + j) + ny * (z0 + k))] = f[16 * k + 4 * j + i];
    }
  }
}

template <typename Scalar>
static int dispatch(bool enc, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, void* data, void* stream,
                    size_t stream_words)
{
  Scalar* d = (Scalar*)data;
  uint64_t* s = (uint64_t*)stream;
  if (dims == 1) run<Scalar, 1>(enc, nx, ny, nz, maxbits, d, s, stream_words);
  else if (dims == 2) run<Scalar, 2>(enc, nx, ny, nz, maxbits, d, s, stream_words);
  else if (dims == 3) run<Scalar, 3>(enc, nx, ny, nz, maxbits, d, s, stream_words);
  else return -1;
  return 0;
}

// Variable-rate emulation: the device codec's general encode_block / decode_block (stream_params) per block,
// staged words appended bit-tight on the host (the device does that in compact_var_kernel).
template <typename Scalar, int DIMS>
static size_t run_var(bool enc, size_t nx, size_t ny, size_t nz, const stream_params& sp, Scalar* data, uint64_t* stream,
                      size_t stream_words, uint64_t* offsets)
{
  constexpr int BS = 1 << (2 * DIMS);
  if (DIMS < 2) ny = 1;
  if (DIMS < 3) nz = 1;
  const size_t bx = (nx + 3) / 4, by = (ny + 3) / 4, bz = (nz + 3) / 4;
  if (enc) std::memset(stream, 0, stream_words * 8);
  uint64_t pos = 0;
  for (size_t b = 0; b < bx * by * bz; b++) {
    const size_t ix = b % bx, iy = (b / bx) % by, iz = b / (bx * by);
    const size_t x0 = 4 * ix, y0 = 4 * iy, z0 = 4 * iz;
    const unsigned mx = (unsigned)(nx - x0 < 4 ? nx - x0 : 4);
    const unsigned my = DIMS >= 2 ? (unsigned)(ny - y0 < 4 ? ny - y0 : 4) : 1;
    const unsigned mz = DIMS >= 3 ? (unsigned)(nz - z0 < 4 ? nz - z0 : 4) : 1;
    Scalar f[BS];
    if (offsets) offsets[b] = pos;
    if (enc) {
      for (int i = 0; i < BS; i++) f[i] = 0;
      for (unsigned k = 0; k < mz; k++)
        for (unsigned j = 0; j < my; j++)
          for (unsigned i = 0; i < mx; i++) f[16 * k + 4 * j + i] = data[(x0 + i) + nx * ((y0 + j) + ny * (z0 + k))];
      if (mx < 4 || my < (DIMS >= 2 ? 4u : 1u) || mz < (DIMS >= 3 ? 4u : 1u)) pad_block<Scalar, DIMS>(f, mx, my, mz);
      uint64_t stage[96];
      std::memset(stage, 0xff, sizeof stage);  // garbage behind the written words must not leak into the stream
      unsigned bits = 0;
      bool done = false;
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 3) {
        if (g_use_f3) {  // the fast float/3D coder, as enc_var_kernel uses it
          const f3::tables tb = host_tables();
          f3::u32x4 rows[16];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          f3::sink32 sink(reinterpret_cast<uint32_t*>(stage), 160);
          bits = f3::encode_block(f, sp, sink, tb, ps);
          done = true;
        }
      }
      if constexpr (std::is_same<Scalar, double>::value && DIMS == 3) {
        if (g_use_f3) {  // the double/3D table coder, as enc_var_f64_kernel uses it
          const f3::tables tb = host_tables();
          f3::u32x4 rows[32];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          f3::sink32 sink(reinterpret_cast<uint32_t*>(stage), 160);
          bits = d3::encode_block(f, sp, sink, tb, ps);
          done = true;
        }
      }
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 1) {
        if (g_use_f3 && f1::var_params_ok(sp)) {  // per-plane table coder of the float/1D variable-rate kernels
          const f1::tables tb = host_tables1();
          stage_sink sink(stage);
          bit_writer<stage_sink> bw(sink);
          bits = f1::encode_block_var(f, sp, bw, tb);
          bw.flush();
          done = true;
        }
      }
      if (!done && DIMS == 1) {  // append form, as enc_var1_kernel strings the blocks of a group together
        stage_sink sink(stage);
        bit_writer<stage_sink> bw(sink);
        bw.put(0x2au, 7);  // something already in the writer: the block must land behind it untouched
        bits = encode_block_append<Scalar, DIMS>(f, sp, bw);
        bw.flush();
        uint64_t carry = 0;  // drop the 7 leading bits again
        const unsigned nwords = (bits + 7 + 63) / 64;
        for (unsigned j = nwords; j-- > 0;) {
          const uint64_t w = stage[j];
          stage[j] = (w >> 7) | (carry << 57);
          carry = w & 0x7fu;
        }
        if ((stage[0], carry) != 0x2au) return 0;
        done = true;
      }
      if (!done) {
        stage_sink sink(stage);
        bits = encode_block<Scalar, DIMS>(f, sp, sink);
      }
      for (unsigned j = 0; 64 * j < bits; j++) {
        uint64_t w = stage[j];
        const unsigned valid = bits - 64 * j;
        if (valid < 64) w &= lowmask64(valid);
        const uint64_t at = pos + 64 * j;
        const unsigned sh = (unsigned)(at & 63);
        if ((at >> 6) < stream_words) stream[at >> 6] |= w << sh;
        if (sh && (at >> 6) + 1 < stream_words) stream[(at >> 6) + 1] |= w >> (64 - sh);
      }
      pos += bits;
    } else {
      slot_reader r(stream + (pos >> 6), (unsigned)(stream_words - (pos >> 6)), (unsigned)(pos & 63));
      slot_reader r2 = r;
      Scalar dummy[BS], f2[BS];
      const unsigned bits = decode_block<Scalar, DIMS, false>(f, sp, r);
      const unsigned bits2 = decode_block<Scalar, DIMS, true>(dummy, sp, r2);  // index walk must agree
      if (bits != bits2) return 0;
      if constexpr (std::is_same<Scalar, double>::value && DIMS == 3) {
        if (g_use_f3) {  // the double/3D table decoder must reproduce the generic one bit for bit
          const f3::tables tb = host_tables();
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
          f3::reader32 r3(s32 + (pos >> 5), (unsigned)(2 * stream_words - (pos >> 5)), (unsigned)(pos & 31));
          f3::u32x4 rows[32];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          d3::decode_block(f2, sp, r3, tb, ps);
          if (std::memcmp(f, f2, sizeof f2)) return 0;
        }
      }
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 1) {
        if (g_use_f3 && f1::var_params_ok(sp)) {  // the float/1D table decoder must reproduce the generic one
          const f1::tables tb = host_tables1();
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
          f3::reader32 r3(s32 + (pos >> 5), (unsigned)(2 * stream_words - (pos >> 5)), (unsigned)(pos & 31));
          const unsigned bits3 = f1::decode_block_var(f2, sp, r3, tb);
          if (bits3 != bits || std::memcmp(f, f2, 4 * sizeof(float))) return 0;
        }
      }
      if constexpr (std::is_same<Scalar, float>::value && DIMS == 3) {
        if (g_use_f3) {  // the fast float/3D decoder must reproduce the generic one bit for bit
          const f3::tables tb = host_tables();
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stream);
          f3::reader32 r3(s32 + (pos >> 5), (unsigned)(2 * stream_words - (pos >> 5)), (unsigned)(pos & 31));
          f3::u32x4 rows[16];
          f3::plane_rows ps;
          ps.p = rows; ps.stride = 1;
          f3::decode_block(f2, sp, r3, tb, ps);
          if (std::memcmp(f, f2, sizeof f2)) return 0;
        }
      }
      pos += bits;
      for (unsigned k = 0; k < mz; k++)
        for (unsigned j = 0; j < my; j++)
          for (unsigned i = 0; i < mx; i++) data[(x0 + i) + nx * ((y0 + j) + ny * (z0 + k))] = f[16 * k + 4 * j + i];
    }
  }
  if (offsets) offsets[bx * by * bz] = pos;
  return (size_t)((pos + 63) / 64) * 8;
}

template <typename Scalar>
static size_t dispatch_var(bool enc, int dims, size_t nx, size_t ny, size_t nz, const stream_params& sp, void* data,
                           void* stream, size_t stream_words, uint64_t* offsets)
{
  Scalar* d = (Scalar*)data;
  uint64_t* s = (uint64_t*)stream;
  if (dims == 1) return run_var<Scalar, 1>(enc, nx, ny, nz, sp, d, s, stream_words, offsets);
  if (dims == 2) return run_var<Scalar, 2>(enc, nx, ny, nz, sp, d, s, stream_words, offsets);
  if (dims == 3) return run_var<Scalar, 3>(enc, nx, ny, nz, sp, d, s, stream_words, offsets);
  return 0;
}

extern "C" {
// params: {minbits, maxbits, maxprec, minexp}; returns the stream size in bytes (0 = failure)
size_t emul_compress_var(int type, int dims, size_t nx, size_t ny, size_t nz, const unsigned* params, const void* in,
                         void* stream, size_t stream_words, uint64_t* offsets)
{
  stream_params sp;
  sp.minbits = params[0]; sp.maxbits = params[1]; sp.maxprec = params[2]; sp.minexp = (int)params[3];
  return type == 1 ? dispatch_var<float>(true, dims, nx, ny, nz, sp, (void*)in, stream, stream_words, offsets)
                   : dispatch_var<double>(true, dims, nx, ny, nz, sp, (void*)in, stream, stream_words, offsets);
}
size_t emul_decompress_var(int type, int dims, size_t nx, size_t ny, size_t nz, const unsigned* params, const void* stream,
                           void* out, size_t stream_words)
{
  stream_params sp;
  sp.minbits = params[0]; sp.maxbits = params[1]; sp.maxprec = params[2]; sp.minexp = (int)params[3];
  return type == 1 ? dispatch_var<float>(false, dims, nx, ny, nz, sp, out, (void*)stream, stream_words, nullptr)
                   : dispatch_var<double>(false, dims, nx, ny, nz, sp, out, (void*)stream, stream_words, nullptr);
}
void emul_use_f3(int on) { g_use_f3 = on; }
// type: 1 float, 2 double (same codes as the oracle and include/zfb.h)
int emul_compress(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, const void* in, void* stream,
                  size_t stream_words)
{
  return type == 1 ? dispatch<float>(true, dims, nx, ny, nz, maxbits, (void*)in, stream, stream_words)
                   : dispatch<double>(true, dims, nx, ny, nz, maxbits, (void*)in, stream, stream_words);
}
int emul_decompress(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, const void* stream,
                    void* out, size_t stream_words)
{
  return type == 1 ? dispatch<float>(false, dims, nx, ny, nz, maxbits, out, (void*)stream, stream_words)
                   : dispatch<double>(false, dims, nx, ny, nz, maxbits, out, (void*)stream, stream_words);
}
void emul_transpose32(uint32_t* a)
{
  uint32_t t[32];
  for (int i = 0; i < 32; i++) t[i] = a[i];
  transpose32(t);
  for (int i = 0; i < 32; i++) a[i] = t[i];
}
void emul_transpose64(uint64_t* a)
{
  uint64_t t[64];
  for (int i = 0; i < 64; i++) t[i] = a[i];
  transpose64(t);
  for (int i = 0; i < 64; i++) a[i] = t[i];
}
}
