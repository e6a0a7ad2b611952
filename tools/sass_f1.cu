// Development aid: compiles only the float/1D line kernels (64-bit slots) so that their SASS can be inspected quickly
//   nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -o build/f1.cubin tools/sass_f1.cu
#include "../vizaly-foresight_b200/csrc/zfb_kernels.cuh"
namespace zfb {
template __global__ void enc1_f32_kernel<2>(const float* __restrict__, uint32_t* __restrict__, size_t, unsigned);
template __global__ void dec1_f32_kernel<2, true>(const uint32_t* __restrict__, float* __restrict__, const float* __restrict__, size_t,
                                                  unsigned, double* __restrict__);
}
