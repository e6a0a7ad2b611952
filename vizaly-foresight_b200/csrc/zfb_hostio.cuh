// zfb_hostio.cuh -- host side of the host-pointer entry points: moving PAGEABLE caller buffers at PCIe speed.
//
// CBench hands a plugin the loader's `new T[]` array and expects malloc'd outputs (main.cpp:262-280, 395), i.e.
// pageable memory.  cudaMemcpyAsync on pageable memory is staged by the driver through one small pinned buffer and is
// synchronous for the host, so a slab pipeline over three streams collapses into a serial chain.  Here the staging is
// explicit: a slab is copied by a few threads between the caller's buffer and a pinned double buffer, and the DMA of
// slab i runs while the threads copy slab i+1 (or i-1 on the way out).  Pinned caller buffers take the direct path.
#pragma once
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#if defined(__linux__)
#include <sched.h>
#endif

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace zfb {

// One thread's share of a staging copy.  The destination is never read again by this CPU (it is either the pinned buffer
// the DMA engine reads next or the caller's output array), so it is written with non-temporal stores: no read-for-
// ownership of the destination lines, i.e. two memory transfers per byte instead of three on a host whose DRAM
// bandwidth is what bounds the pageable path (glibc's memcpy only switches to such stores far above the 4 MiB a thread
// gets here).  ZFB_COPY_NT=0 goes back to memcpy.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__CUDA_ARCH__)
__attribute__((target("avx2"))) inline void stream_copy_avx2(char* d, const char* s, size_t blocks128)
{
  for (size_t i = 0; i < blocks128; i++) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(s + 128 * i)), b = _mm256_loadu_si256((const __m256i*)(s + 128 * i + 32)),
                  c = _mm256_loadu_si256((const __m256i*)(s + 128 * i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + 128 * i + 96));
    _mm256_stream_si256((__m256i*)(d + 128 * i), a);
    _mm256_stream_si256((__m256i*)(d + 128 * i + 32), b);
    _mm256_stream_si256((__m256i*)(d + 128 * i + 64), c);
    _mm256_stream_si256((__m256i*)(d + 128 * i + 96), e);
  }
}
#define ZFB_HAVE_AVX2_COPY 1
#endif

inline void stream_copy(char* d, const char* s, size_t n)
{
#if defined(__x86_64__)
  static const bool nt = [] { const char* e = std::getenv("ZFB_COPY_NT"); return !(e && e[0] == '0'); }();
  if (nt && n >= 4096) {
#if defined(ZFB_HAVE_AVX2_COPY)
    static const bool avx2 = [] { const char* e = std::getenv("ZFB_COPY_NT"); return __builtin_cpu_supports("avx2") && !(e && e[0] == '1' && e[1] == '6'); }();
    if (avx2) {  // 32-byte stores (ZFB_COPY_NT=16 keeps the 16-byte form)
      size_t head = (size_t)((32 - ((uintptr_t)d & 31)) & 31);
      std::memcpy(d, s, head);
      d += head; s += head; n -= head;
      const size_t blocks = n / 128;
      stream_copy_avx2(d, s, blocks);
      _mm_sfence();
      std::memcpy(d + 128 * blocks, s + 128 * blocks, n - 128 * blocks);
      return;
    }
#endif
    size_t head = (size_t)((16 - ((uintptr_t)d & 15)) & 15);
    std::memcpy(d, s, head);
    d += head; s += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; i++) {
      const __m128i a = _mm_loadu_si128((const __m128i*)(s + 64 * i)), b = _mm_loadu_si128((const __m128i*)(s + 64 * i + 16)),
                    c = _mm_loadu_si128((const __m128i*)(s + 64 * i + 32)), e = _mm_loadu_si128((const __m128i*)(s + 64 * i + 48));
      _mm_stream_si128((__m128i*)(d + 64 * i), a);
      _mm_stream_si128((__m128i*)(d + 64 * i + 16), b);
      _mm_stream_si128((__m128i*)(d + 64 * i + 32), c);
      _mm_stream_si128((__m128i*)(d + 64 * i + 48), e);
    }
    _mm_sfence();
    std::memcpy(d + 64 * blocks, s + 64 * blocks, n - 64 * blocks);
    return;
  }
#endif
  std::memcpy(d, s, n);
}

// A few persistent threads that split one memcpy between them; the caller takes a share and returns when all are done.
class copy_pool {
 public:
  explicit copy_pool(unsigned nthreads) : n_(nthreads < 1 ? 1 : nthreads)
  {
    for (unsigned i = 1; i < n_; i++) th_.emplace_back([this, i] { worker(i); });
  }
  ~copy_pool()
  {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  unsigned threads() const { return n_; }
  void copy(void* dst, const void* src, size_t bytes)
  {
    if (bytes < (size_t)(4u << 20) || n_ == 1) { stream_copy((char*)dst, (const char*)src, bytes); return; }
    {
      std::lock_guard<std::mutex> l(mu_);
      dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes;
      pending_ = n_ - 1;
      gen_++;
    }
    cv_.notify_all();
    share(0);
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void share(unsigned i)
  {
    // 2 MiB granularity keeps every thread on whole (huge) pages
    const size_t unit = (size_t)2 << 20, units = (bytes_ + unit - 1) / unit;
    const size_t u0 = units * i / n_, u1 = units * (i + 1) / n_;
    const size_t b0 = u0 * unit, b1 = u1 * unit < bytes_ ? u1 * unit : bytes_;
    if (b1 > b0) stream_copy(dst_ + b0, src_ + b0, b1 - b0);
  }
  void worker(unsigned i)
  {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      share(i);
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  unsigned n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  unsigned pending_ = 0;
  bool stop_ = false;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0;
};

inline unsigned default_copy_threads()
{
  if (const char* e = std::getenv("ZFB_COPY_THREADS")) {
    const long v = std::atol(e);
    if (v >= 1 && v <= 64) return (unsigned)v;
  }
  unsigned cpus = std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = (unsigned)CPU_COUNT(&set);
#endif
  unsigned ranks = 1;  // ranks sharing this host (one process per GPU)
  for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE"})
    if (const char* e = std::getenv(name)) {
      const long v = std::atol(e);
      if (v >= 1) { ranks = (unsigned)v; break; }
    }
  unsigned t = cpus / ranks;
  return t < 2 ? 2 : t > 16 ? 16 : t;
}

// pinned double buffer, grown on demand
struct pinned_pair {
  void* p[2] = {nullptr, nullptr};
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (cap >= bytes) return 0;
    release();
    for (int i = 0; i < 2; i++)
      if (cudaHostAlloc(&p[i], bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        release();
        return -1;
      }
    cap = bytes;
    return 0;
  }
  void release()
  {
    for (int i = 0; i < 2; i++) {
      if (p[i]) cudaFreeHost(p[i]);
      p[i] = nullptr;
    }
    cap = 0;
  }
};

// true if the DMA engine can read / write the buffer directly (cudaHostAlloc / cudaHostRegister memory)
inline bool is_pinned(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace zfb
