// zfb_ws.cuh -- warp-specialised decode (+ fused metrics) kernel for float 3D fields, sm_100a.
//
// Same path as dec3_f32_tile_kernel (zfp_decompress behind ZFPCompressorGpu::decompress, zfpCompressorGpu.hpp:249, and
// the metric passes of main.cpp:306-352), same codec functions (f3::parse_planes / f3::decode_finish), other schedule.
// The tile kernel runs parse and inverse transform one after the other in every thread: the parse is a chain of
// dependent shared-memory fetches and table look-ups with divergent trip counts (latency bound, ~50 live registers), the
// transform is 2 400 straight-line instructions on 64 + 64 live values (issue bound, 128+ registers).  At 128 registers
// only 16 warps fit an SM, so the parse phases leave the schedulers idle (issue slots 53 % busy, profiles/r02_notes.md).
// Here the two halves run in different warps of one 640-thread CTA per SM and trade registers with setmaxnreg:
//   * 16 PARSER warps (56 registers each): a warp owns a sequence of warp-tiles (32 x-consecutive blocks); its slot
//     span arrives by cp.async.bulk one tile ahead, is re-laid into a bank-conflict-free padded area and parsed into
//     the warp's plane-row buffer (32 blocks x 256 B);
//   * 4 TRANSFORM warps (232 registers, one per scheduler): take the plane rows of "their" four parser warps in turn,
//     hand the buffer back at once, and run inverse transposes / lifting / conversion, the metric partials against the
//     original (16 coalesced 128-bit global loads per lane, issued before the rows are awaited) and the 16 coalesced
//     128-bit stores.
// Hand-over by mbarriers (rows_full / rows_empty per parser warp); parser and transform warp walk the same static tile
// sequence, so no descriptors are passed.  A scheduler then always holds four latency-bound warps and one issue-bound
// warp.
#pragma once
#include "zfb_kernels.cuh"

namespace zfb {
namespace ws {

constexpr int T_WARPS = 4;                          // transform warps = warpgroup 0
#ifndef ZFB_WS_P_WARPS
#define ZFB_WS_P_WARPS 16
#endif
constexpr int P_WARPS = ZFB_WS_P_WARPS;             // parser warps = warpgroups 1..
constexpr int THREADS = 32 * (T_WARPS + P_WARPS);
constexpr int ROWS_BYTES = 32 * 256;                // plane rows of one warp-tile
#ifndef ZFB_WS_REGS_T
#define ZFB_WS_REGS_T 232
#endif
#ifndef ZFB_WS_REGS_P
#define ZFB_WS_REGS_P 56
#endif

__host__ __device__ inline unsigned land_bytes(unsigned W) { return 32u * W * 8u; }
__host__ __device__ inline unsigned pad_bytes(unsigned W) { return (32u * (2u * W + 1u) * 4u + 16u + 15u) & ~15u; }
// dynamic shared memory: [rows: P x 8 KiB][landing buffers: P x 256 W][padded slots: P x pad_bytes][meta: P x 128 B]
//                        [mbarriers: 3 P]
__host__ __device__ inline size_t smem_bytes(unsigned W)
{
  return (size_t)P_WARPS * (ROWS_BYTES + land_bytes(W) + pad_bytes(W) + 128u) + 3u * P_WARPS * 8u;
}

struct wtile {
  unsigned x0, y0, z0;  // element coordinates of the warp-tile's first value
  size_t blk0;          // first block (stream order)
  unsigned nvalid;      // blocks that exist (<= 32)
};
__device__ __forceinline__ wtile wtile_of(const tile_geom& tg, unsigned wpr, unsigned w)
{
  wtile t;
  const unsigned row = w / wpr, xw = w - row * wpr;
  const unsigned iz = row / tg.nby, iy = row - iz * tg.nby;
  t.x0 = xw * 128u; t.y0 = iy * 4u; t.z0 = iz * 4u;
  t.blk0 = ((size_t)iz * tg.nby_t + iy) * tg.nbx + (size_t)xw * 32u;
  const unsigned rem = tg.nbx - xw * 32u;
  t.nvalid = rem < 32u ? rem : 32u;
  return t;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar_s)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar_s),
      "r"(parity)
      : "memory");
}

template <bool METRICS>
__global__ void __launch_bounds__(THREADS, 1)
dec3_f32_ws_kernel(const float* __restrict__ orig, const uint64_t* __restrict__ stream, float* __restrict__ out,
                   tile_geom tg, unsigned maxbits, double* __restrict__ part)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ dec_coder_smem cs;
  __shared__ double red[7 * T_WARPS];

  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const unsigned W = tg.W, W2 = 2u * W, pstride = W2 + 1u;
  const uint32_t smem_s = ptx::smem_u32(smem_raw);
  const uint32_t rows_s = smem_s;
  const uint32_t land_s = rows_s + P_WARPS * ROWS_BYTES;
  const uint32_t pad_s = land_s + P_WARPS * land_bytes(W);
  const uint32_t meta_s = pad_s + P_WARPS * pad_bytes(W);
  const uint32_t bars_s = meta_s + P_WARPS * 128u;  // [p]: slots landed, [P + p]: rows full, [2P + p]: rows empty
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (bars_s - smem_s));

  if (tid == 0) {
    for (int p = 0; p < P_WARPS; p++) {
      ptx::mbar_init(&bars[p], 1);
      ptx::mbar_init(&bars[P_WARPS + p], 32);
      ptx::mbar_init(&bars[2 * P_WARPS + p], 32);
    }
    ptx::fence_barrier_init();
  }
  const f3::tables tb = build_coder_tables(cs);  // includes the __syncthreads that publishes the barriers

  const unsigned wpr = (tg.nbx + 31u) >> 5;            // warp-tiles per block row
  const unsigned nwt = wpr * tg.nby * tg.nbz;          // warp-tiles in the field
  const unsigned stride = gridDim.x * P_WARPS;

  if (warp >= (unsigned)T_WARPS) {
    // ------------------------------------------------------------------ parser warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ZFB_WS_REGS_P));
    const unsigned p = warp - T_WARPS;
    const uint32_t my_land = land_s + p * land_bytes(W), my_pad = pad_s + p * pad_bytes(W);
    const uint32_t b_slot = bars_s + 8u * p, b_full = bars_s + 8u * (P_WARPS + p), b_empty = bars_s + 8u * (2 * P_WARPS + p);
    f3::smem_rows_t<0> ps;
    ps.base = rows_s + p * ROWS_BYTES + 16u * lane;
    ps.stride = 512u;
    f3::smem_words r;
    r.base = my_pad + 4u * lane * pstride;

    auto issue_slots = [&](unsigned w) {  // lane 0
      const wtile t = wtile_of(tg, wpr, w);
      const uint32_t bytes = t.nvalid * W * 8u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_slot), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(my_land),
                   "l"(stream + t.blk0 * W), "r"(bytes), "r"(b_slot)
                   : "memory");
    };

    unsigned w = blockIdx.x * P_WARPS + p;
    if (lane == 0 && w < nwt) issue_slots(w);
    for (unsigned it = 0; w < nwt; w += stride, it++) {
      const unsigned rem = tg.nbx - (w % wpr) * 32u;
      const unsigned nvalid = rem < 32u ? rem : 32u;
      mbar_wait_s(b_slot, it & 1u);
      // re-lay the 32 slots into the padded area: slot i at 32-bit word i * (2W + 1) (odd stride: conflict free)
      if (W2 == 16u) {
        uint32_t v[16];
#pragma unroll
        for (int m = 0; m < 16; m++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[m]) : "r"(my_land + 4u * (lane + 32u * m)));
#pragma unroll
        for (int m = 0; m < 16; m++) {
          const unsigned i = lane + 32u * m;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_pad + 4u * ((i >> 4) * 17u + (i & 15u))), "r"(v[m]) : "memory");
        }
      } else {
        for (unsigned i = lane; i < 32u * W2; i += 32u) {
          uint32_t v;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(my_land + 4u * i));
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_pad + 4u * ((i / W2) * pstride + i % W2)), "r"(v) : "memory");
        }
      }
      __syncwarp();
      if (lane == 0 && w + stride < nwt) {
        ptx::fence_proxy_async();  // the landing buffer has been read (generic proxy) before the bulk copy refills it
        issue_slots(w + stride);
      }
      mbar_wait_s(b_empty, (it & 1u) ^ 1u);  // the transform warp has taken the previous tile's rows
      uint32_t meta = 0u;
      if (lane < nvalid) {
        const uint32_t head = f3::fetch32(r, 0);
        if (head & 1u) {
          const int emax = (int)((head >> 1) & 0xffu) - 127;
          const stream_params sp = fixed_rate_params(maxbits);
          const int prec = block_precision<3>(emax, sp);
          const int kend = f3::parse_planes<32>(ps, sp.maxbits, 9u, prec < 32 ? 32 - prec : 0, r, tb);
          meta = 1u | (((head >> 1) & 0xffu) << 1) | (kend < 15 ? 512u : 0u);
        }
      }
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(meta_s + 128u * p + 4u * lane), "r"(meta) : "memory");
      mbar_arrive(b_full);
    }
    return;
  }

  // -------------------------------------------------------------------- transform warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ZFB_WS_REGS_T));
  maccf m;
  m.init();
  const size_t rs = tg.nx, ps2 = (size_t)tg.ny * tg.nx;
  for (unsigned it = 0;; it++) {
    if (blockIdx.x * P_WARPS + warp + it * stride >= nwt) break;
#pragma unroll 1
    for (unsigned q = 0; q < (unsigned)(P_WARPS / T_WARPS); q++) {
      const unsigned p = warp + T_WARPS * q;
      const unsigned w = blockIdx.x * P_WARPS + p + it * stride;
      if (w >= nwt) break;
      const wtile t = wtile_of(tg, wpr, w);
      const bool active = lane < t.nvalid;
      const size_t off = ((size_t)t.z0 * tg.ny + t.y0) * tg.nx + t.x0 + 4u * lane;
      float4 o[16];
      if (METRICS && active) {
        const float* og = orig + off;
#pragma unroll
        for (int r16 = 0; r16 < 16; r16++) {
          const float* a = og + (size_t)(r16 >> 2) * ps2 + (size_t)(r16 & 3) * rs;
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(o[r16].x), "=f"(o[r16].y), "=f"(o[r16].z), "=f"(o[r16].w)
                       : "l"(a));
        }
      }
      const uint32_t b_full = bars_s + 8u * (P_WARPS + p), b_empty = bars_s + 8u * (2 * P_WARPS + p);
      mbar_wait_s(b_full, it & 1u);
      f3::planes64 P;
      f3::smem_rows_t<0> ps;
      ps.base = rows_s + p * ROWS_BYTES + 16u * lane;
      ps.stride = 512u;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const f3::u32x4 v = ps.ld(k);
        P.a[2 * k] = v.x; P.b[2 * k] = v.y; P.a[2 * k + 1] = v.z; P.b[2 * k + 1] = v.w;
      }
      uint32_t meta;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(meta) : "r"(meta_s + 128u * p + 4u * lane));
      mbar_arrive(b_empty);  // release: the loads above are ordered before the parser's next writes
      if (active) {
        float f[64];
        f3::decode_finish(f, P, (int)((meta >> 1) & 0xffu) - 127, (meta & 1u) != 0u, (meta & 512u) != 0u);
        if (METRICS) {
          frow fr;
          fr.init();
#pragma unroll
          for (int r16 = 0; r16 < 16; r16++) {
            fr.add4(o[r16], f[4 * r16], f[4 * r16 + 1], f[4 * r16 + 2], f[4 * r16 + 3]);
            if ((r16 & 3) == 3) fr.flush(m);
          }
        }
        float* dst = out + off;
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
      }
    }
  }
  if (METRICS) {
    // reduction over the four transform warps only (the parser warps are gone): named barrier 1, 128 threads
    macc mm;
    mm.s_sq = m.s_sq; mm.s_abs = m.s_abs; mm.s_rel = m.s_rel;
    mm.m_abs = (double)m.m_abs; mm.m_rel = (double)m.m_rel; mm.m_o = (double)m.m_o; mm.m_no = (double)m.m_no;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mm.s_sq += shfl_xor_d(mm.s_sq, o);
      mm.s_abs += shfl_xor_d(mm.s_abs, o);
      mm.s_rel += shfl_xor_d(mm.s_rel, o);
      mm.m_abs = fmax(mm.m_abs, shfl_xor_d(mm.m_abs, o));
      mm.m_rel = fmax(mm.m_rel, shfl_xor_d(mm.m_rel, o));
      mm.m_o = fmax(mm.m_o, shfl_xor_d(mm.m_o, o));
      mm.m_no = fmax(mm.m_no, shfl_xor_d(mm.m_no, o));
    }
    if (lane == 0) {
      double* rr = red + 7 * warp;
      rr[0] = mm.s_sq; rr[1] = mm.s_abs; rr[2] = mm.s_rel; rr[3] = mm.m_abs; rr[4] = mm.m_rel; rr[5] = mm.m_o; rr[6] = mm.m_no;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * T_WARPS) : "memory");
    if (tid == 0) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = -1.7976931348623157e308, a6 = -1.7976931348623157e308;
      for (int wi = 0; wi < T_WARPS; wi++) {  // fixed order
        const double* rr = red + 7 * wi;
        a0 += rr[0]; a1 += rr[1]; a2 += rr[2];
        a3 = fmax(a3, rr[3]); a4 = fmax(a4, rr[4]); a5 = fmax(a5, rr[5]); a6 = fmax(a6, rr[6]);
      }
      double* pr = part + (size_t)blockIdx.x * PART_STRIDE;
      pr[0] = a0; pr[1] = a1; pr[2] = a2; pr[3] = a3; pr[4] = a4; pr[5] = a5; pr[6] = a6; pr[7] = 0.0;
    }
  }
}

}  // namespace ws
}  // namespace zfb
