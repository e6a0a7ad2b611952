// Development aid: compiles only the warp-specialised float/3D kernels (registers, spills, SASS)
#include "../vizaly-foresight_b200/csrc/zfb_ws.cuh"
namespace zfb { namespace ws {
template __global__ void dec3_f32_ws_kernel<true>(const float* __restrict__, const uint64_t* __restrict__, float* __restrict__, tile_geom, unsigned, double* __restrict__);
template __global__ void dec3_f32_ws_kernel<false>(const float* __restrict__, const uint64_t* __restrict__, float* __restrict__, tile_geom, unsigned, double* __restrict__);
}}
// (the float64 kernel is not a template: it is compiled with the header)
namespace zfb { namespace ws {
template __global__ void enc3_f32_warp_kernel<false>(const __grid_constant__ CUtensorMap, uint64_t* __restrict__, tile_geom, unsigned, double* __restrict__);
}}
