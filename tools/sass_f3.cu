// Development aid: compiles only the float/3D tile kernels so that their SASS can be inspected quickly
//   nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -o build/f3.cubin tools/sass_f3.cu
//   cuobjdump -sass build/f3.cubin
#include "../vizaly-foresight_b200/csrc/zfb_kernels.cuh"
namespace zfb {
template __global__ void dec3_f32_tile_kernel<true, false>(const __grid_constant__ CUtensorMap, const uint64_t* __restrict__,
                                                           float* __restrict__, tile_geom, unsigned, double* __restrict__, double,
                                                           double, unsigned long long* __restrict__);
template __global__ void enc3_f32_tile_kernel<false>(const __grid_constant__ CUtensorMap, uint64_t* __restrict__, tile_geom,
                                                     unsigned, double* __restrict__);
}
namespace zfb {
template __global__ void dec3_f32_tile_kernel<true, true>(const __grid_constant__ CUtensorMap, const uint64_t* __restrict__,
                                                          float* __restrict__, tile_geom, unsigned, double* __restrict__, double,
                                                          double, unsigned long long* __restrict__);
}
