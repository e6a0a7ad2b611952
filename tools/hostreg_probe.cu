// Probe (GPU box): what does pinning a caller's pageable buffer cost?  cudaHostRegister / Unregister of a malloc'd,
// touched buffer, whole and in 64 MiB chunks, with and without transparent huge pages, against the staged copy path.
//   nvcc -O2 -o build/hostreg_probe tools/hostreg_probe.cu && build/hostreg_probe [GiB]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv)
{
  const size_t gib = argc > 1 ? atoi(argv[1]) : 2;
  const size_t n = gib << 30;
  void* d; cudaMalloc(&d, n);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int thp = 0; thp < 2; thp++) {
    char* h = (char*)aligned_alloc(2 << 20, n);
    if (thp) madvise(h, n, MADV_HUGEPAGE);
    double t0 = now();
    for (size_t i = 0; i < n; i += 4096) h[i] = (char)i;
    double t1 = now();
    printf("thp=%d first touch %.1f ms (%.1f GB/s)\n", thp, 1e3 * (t1 - t0), n / (t1 - t0) / 1e9);
    // pageable memcpy as is
    t0 = now(); cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); t1 = now();
    printf("  pageable cudaMemcpy H2D %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), n / (t1 - t0) / 1e9);
    for (int rep = 0; rep < 2; rep++) {
      t0 = now(); cudaError_t e = cudaHostRegister(h, n, cudaHostRegisterDefault); t1 = now();
      if (e) { printf("  register failed: %s\n", cudaGetErrorString(e)); break; }
      double t2 = now(); cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); double t3 = now();
      double t4 = now(); cudaHostUnregister(h); double t5 = now();
      printf("  whole: register %.1f ms (%.1f GB/s), H2D %.1f ms (%.1f GB/s), unregister %.1f ms\n", 1e3 * (t1 - t0), n / (t1 - t0) / 1e9,
             1e3 * (t3 - t2), n / (t3 - t2) / 1e9, 1e3 * (t5 - t4));
    }
    // chunked: register chunk i+1 on a helper thread while chunk i is copied
    const size_t ch = 64 << 20, nch = n / ch;
    for (int nthreads = 1; nthreads <= 4; nthreads *= 2) {
      t0 = now();
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t] { for (size_t c = t; c < nch; c += nthreads) cudaHostRegister(h + c * ch, ch, cudaHostRegisterDefault); });
      for (auto& x : th) x.join();
      t1 = now();
      double t2 = now();
      for (size_t c = 0; c < nch; c++) cudaMemcpyAsync((char*)d + c * ch, h + c * ch, ch, cudaMemcpyHostToDevice, s);
      cudaStreamSynchronize(s);
      double t3 = now();
      for (size_t c = 0; c < nch; c++) cudaHostUnregister(h + c * ch);
      double t4 = now();
      printf("  64 MiB chunks, %d thread(s): register %.1f ms (%.1f GB/s), H2D %.1f ms, unregister %.1f ms\n", nthreads, 1e3 * (t1 - t0),
             n / (t1 - t0) / 1e9, 1e3 * (t3 - t2), 1e3 * (t4 - t3));
    }
    // D2H into the pageable buffer (registered)
    cudaHostRegister(h, n, cudaHostRegisterDefault);
    t0 = now(); cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); t1 = now();
    printf("  registered D2H %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), n / (t1 - t0) / 1e9);
    cudaHostUnregister(h);
    free(h);
  }
  // fresh malloc (untouched) + register: does registering fault the pages in faster than touching?
  {
    char* h = (char*)aligned_alloc(2 << 20, n);
    madvise(h, n, MADV_HUGEPAGE);
    double t0 = now(); cudaError_t e = cudaHostRegister(h, n, cudaHostRegisterDefault); double t1 = now();
    printf("untouched thp buffer: register %.1f ms (%s)\n", 1e3 * (t1 - t0), cudaGetErrorString(e));
    t0 = now(); cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); t1 = now();
    printf("  D2H %.1f ms (%.1f GB/s)\n", 1e3 * (t1 - t0), n / (t1 - t0) / 1e9);
    t0 = now(); cudaHostUnregister(h); t1 = now();
    printf("  unregister %.1f ms\n", 1e3 * (t1 - t0));
    free(h);
  }
  return 0;
}
