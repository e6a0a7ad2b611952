// zfb_spectrum.cuh -- shell binning of a 3D power spectrum (what PAT's NYX analysis plots).
//
// Foresight's NYX workflow runs gimlet's sim_stats on the original and on every decompressed field and PAT plots the
// RATIO of the two spectra (Analysis/pat/nyx/cinema.py:84-130 reads columns 2 and 3 -- k and P(k) -- of the
// *_ps3d.txt files, README.md:19-25).  gimlet is not part of the reference tree, so there is no normalisation to match;
// the ratio does not depend on one.  This is the device half of that analysis for a field that is already resident:
// the caller transforms it (torch.fft.rfftn -> nz x ny x (nx/2+1) complex values, x fastest) and this kernel adds
// |F|^2 into shells of integer radius round(|k|), with the Hermitian weight 2 for the modes whose mirror image the
// half-spectrum leaves out.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace zfb {

constexpr int PK_MAX_BINS = 2048;  // shells 0 .. min(n)/2 of fields up to 4094^3

// per CTA: shared-memory bins (count, sum of |k|, sum of |F|^2), flushed to row blockIdx.x of `rows`
// rows layout: [gridDim.x][3][nbins] doubles
__global__ void __launch_bounds__(256) pk_bin_kernel(const float2* __restrict__ fhat, unsigned nx, unsigned ny, unsigned nz,
                                                     unsigned nbins, double* __restrict__ rows)
{
  extern __shared__ double sbins[];  // [3][nbins]
  for (unsigned i = threadIdx.x; i < 3u * nbins; i += blockDim.x) sbins[i] = 0.0;
  __syncthreads();
  const unsigned nxh = nx / 2u + 1u;
  const size_t nlines = (size_t)ny * nz;
  for (size_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const unsigned iy = (unsigned)(line % ny), iz = (unsigned)(line / ny);
    const int ky = iy <= ny / 2u ? (int)iy : (int)iy - (int)ny;
    const int kz = iz <= nz / 2u ? (int)iz : (int)iz - (int)nz;
    const float k2yz = (float)ky * (float)ky + (float)kz * (float)kz;
    const float2* row = fhat + line * nxh;
    for (unsigned ix = threadIdx.x; ix < nxh; ix += blockDim.x) {
      const float2 v = row[ix];
      const float kk = sqrtf(k2yz + (float)ix * (float)ix);  // exact for the integer radicands that occur (< 2^24)
      const unsigned b = (unsigned)(kk + 0.5f);
      if (b < nbins) {
        const double w = (ix == 0u || (!(nx & 1u) && ix == nx / 2u)) ? 1.0 : 2.0;
        const double p = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
        atomicAdd(&sbins[b], w);
        atomicAdd(&sbins[nbins + b], w * (double)kk);
        atomicAdd(&sbins[2u * nbins + b], w * p);
      }
    }
  }
  __syncthreads();
  double* out = rows + (size_t)blockIdx.x * 3u * nbins;
  for (unsigned i = threadIdx.x; i < 3u * nbins; i += blockDim.x) out[i] = sbins[i];
}

// fixed-order sum of the CTA rows -> out[3][nbins]
__global__ void __launch_bounds__(256) pk_finalize_kernel(const double* __restrict__ rows, unsigned nrows, unsigned nbins,
                                                          double* __restrict__ out)
{
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < 3u * nbins; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (unsigned r = 0; r < nrows; r++) s += rows[(size_t)r * 3u * nbins + i];
    out[i] = s;
  }
}

}  // namespace zfb
