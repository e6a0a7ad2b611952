// Development aid (CPU): the flat slot parser against parse_planes on random and on realistic (encoded) slots,
// plus a SIMT cost model (32 consecutive blocks = one warp).   g++ -O2 -std=c++17 -DZFB_FLAT_STATS tools/flat_check.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <random>
#include <vector>
#include <algorithm>
#include "../vizaly-foresight_b200/csrc/zfb_codec.cuh"
#include "../vizaly-foresight_b200/csrc/zfb_codec_f3.cuh"
using namespace zfb;

static uint32_t g_enc[f3::ENC_WORDS], g_dec[f3::DEC_WORDS], g_flat[f3::FLAT_WORDS];
struct words { const uint32_t* w; unsigned nwords; static constexpr unsigned bit0 = 0;
  uint32_t word_at(unsigned i) const { return i < nwords ? w[i] : 0xdeadbeefu; } };  // look-ahead words hold garbage

template <int NP> static bool compare(const uint32_t* slot, unsigned maxbits, unsigned hbits, int kmin, const f3::tables& tb,
                                      f3::flat_stats* st)
{
  f3::u32x4 ra[NP / 2], rb[NP / 2];
  f3::plane_rows pa, pb; pa.p = ra; pa.stride = 1; pb.p = rb; pb.stride = 1;
  words r; r.w = slot; r.nwords = (maxbits + 31) / 32;
  const int ka = f3::parse_planes<NP>(pa, maxbits, hbits, kmin, r, tb);
  const int kb = f3::parse_planes_flat<NP>(pb, maxbits, hbits, kmin, r, tb, st);
  if (ka != kb || memcmp(ra, rb, sizeof ra)) {
    printf("MISMATCH maxbits %u kmin %d: kend %d vs %d\n", maxbits, kmin, ka, kb);
    for (int q = 0; q < NP / 2; q++)
      if (memcmp(&ra[q], &rb[q], 16)) printf("  row %d: %08x %08x %08x %08x | %08x %08x %08x %08x\n", q, ra[q].x, ra[q].y, ra[q].z, ra[q].w, rb[q].x, rb[q].y, rb[q].z, rb[q].w);
    return false;
  }
  return true;
}

int main(int argc, char** argv)
{
  for (unsigned i = 0; i < f3::ENC_WORDS; i++) g_enc[i] = f3::enc_table_word(i);
  for (unsigned i = 0; i < f3::DEC_WORDS; i++) g_dec[i] = f3::dec_table_word(i);
  for (unsigned i = 0; i < f3::FLAT_WORDS; i++) g_flat[i] = f3::flat_entry(i);
  f3::tables tb; tb.enc = g_enc; tb.dec = g_dec; tb.enc_s = tb.dec_s = 0; tb.flat = g_flat;
  std::mt19937_64 rng(12345);
  long bad = 0, cases = 0;
  // 1. random slots with various densities of ones, budgets and kmin
  const int ncases = argc > 1 ? atoi(argv[1]) : 400000;
  for (int t = 0; t < ncases && !bad; t++) {
    const unsigned maxbits = 16 + rng() % 2100;
    uint32_t slot[80];
    const int dens = rng() % 6;  // probability of a one: 1/2, 1/4, ... or mixed
    for (auto& w : slot) {
      uint32_t v = (uint32_t)rng();
      for (int d = 0; d < dens; d++) v &= (uint32_t)rng();
      if (dens == 5 && (rng() & 3) == 0) v = (uint32_t)rng() | (uint32_t)rng();
      w = v;
    }
    const int kmin = (rng() & 3) ? 0 : (int)(rng() % 32);
    cases++;
    if (!compare<32>(slot, maxbits, 9, kmin, tb, nullptr)) bad++;
    const int kmin2 = (rng() & 3) ? 0 : (int)(rng() % 64);
    if (!compare<64>(slot, maxbits, 12, kmin2, tb, nullptr)) bad++;
  }
  printf("random slots: %ld cases, %ld mismatches\n", cases, bad);
  if (bad) return 1;
  // 2. realistic slots: a lognormal-ish field, 32 x-consecutive blocks per warp; SIMT model
  const int NX = 256, NY = 16, NZ = 16;
  std::vector<float> fld((size_t)NX * NY * NZ);
  std::normal_distribution<float> nd(0.f, 1.f);
  for (int z = 0; z < NZ; z++) for (int y = 0; y < NY; y++) for (int x = 0; x < NX; x++) {
    const float s = sinf(0.05f * x) * cosf(0.07f * y) + 0.5f * sinf(0.11f * z + 0.013f * x * y * 0.1f);
    fld[((size_t)z * NY + y) * NX + x] = expf(2.0f * s + 0.05f * nd(rng));
  }
  for (unsigned rate : {4u, 8u, 16u}) {
    const unsigned maxbits = rate * 64;
    long warps = 0, it_sum = 0, it_max_sum = 0, slow_any = 0, slow_sum = 0, old_bad = 0;
    for (int bz = 0; bz < NZ / 4; bz++) for (int by = 0; by < NY / 4; by++) for (int bw = 0; bw < NX / 128; bw++) {
      unsigned itmax = 0; std::vector<unsigned> slows;
      for (int l = 0; l < 32; l++) {
        const int bx = bw * 32 + l;
        float f[64];
        for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++)
          f[16 * k + 4 * j + i] = fld[((size_t)(4 * bz + k) * NY + 4 * by + j) * NX + 4 * bx + i];
        uint32_t slot[80]; memset(slot, 0, sizeof slot);
        f3::u32x4 rows[16]; f3::plane_rows ps; ps.p = rows; ps.stride = 1;
        f3::sink32 sink(slot, maxbits / 32);
        f3::encode_block(f, maxbits, sink, tb, ps);
        f3::flat_stats st = {0, 0};
        if (slot[0] & 1u) {
          if (!compare<32>(slot, maxbits, 9, 0, tb, &st)) old_bad++;
        }
        it_sum += st.iters; itmax = std::max(itmax, st.iters); slow_sum += st.slow; slows.push_back(st.slow);
      }
      warps++; it_max_sum += itmax; slow_any += *std::max_element(slows.begin(), slows.end());
    }
    printf("rate %2u: warps %ld  iterations per lane %.1f  per warp (max) %.1f  slow steps per lane %.2f  (lane max per warp %.2f) mismatches %ld\n",
           rate, warps, (double)it_sum / (32.0 * warps), (double)it_max_sum / warps, (double)slow_sum / (32.0 * warps), (double)slow_any / warps, old_bad);
    bad += old_bad;
  }
  return bad ? 1 : 0;
}
