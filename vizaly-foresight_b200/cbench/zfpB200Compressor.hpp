// zfpB200Compressor.hpp -- CBench compressor plugin backed by the B200 codec (include/zfb.h).
//
// Drop-in for CBench/compressors/zfp_gpu/zfpCompressorGpu.hpp (class ZFPCompressorGpu, factory name
// "zfp_gpu"): same parameter ("bits", int-truncated exactly like zfpCompressorGpu.hpp:83,99), same
// dimension convention (n[0] slowest, zfp_field_3d(input, type, n[2], n[1], n[0]), :110-118), same
// rate semantics (zfp_stream_set_rate(zfp, bits, type, dims, 8), :129), malloc'd outputs that main.cpp
// std::free()s (main.cpp:395), return 1 / 0, messages on std::cout.  Differences, all documented in
// "rate" (double) and "wra" are accepted as extensions; the compressed input buffer is
// freed in decompress() (ZFPCompressorGpu leaks it); abs / rel modes are rejected loudly (libzfp's CUDA
// backend is fixed-rate only as well: zfp_compress returns 0 there and the plugin prints
// "compression failed", zfpCompressorGpu.hpp:145-150).
#ifndef ZFB_ZFP_B200_COMPRESSOR_HPP
#define ZFB_ZFP_B200_COMPRESSOR_HPP

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "cbench_api.hpp"
#include "zfb.h"

namespace zfb_cbench {

// One codec context per process (= per MPI rank = per GPU), shared by the compressor and the GPU
// metrics so that the metric pass can pick up the partials fused into decompress().
struct Shared {
  zfb_ctx* ctx = nullptr;
  int rc = ZFB_ENODEV;
  const void* last_orig = nullptr;  // host pointer given to the last compress()
  static Shared& get()
  {
    static Shared s;
    return s;
  }
  int ensure()
  {
    if (ctx) return ZFB_OK;
    int dev = 0;
    if (const char* e = std::getenv("ZFB_DEVICE")) dev = std::atoi(e);
    else if (const char* r = std::getenv("OMPI_COMM_WORLD_LOCAL_RANK")) dev = std::atoi(r);
    else if (const char* r2 = std::getenv("LOCAL_RANK")) dev = std::atoi(r2);
    rc = zfb_create(&ctx, dev);
    return rc;
  }
  void shutdown()
  {
    if (ctx) zfb_destroy(ctx);
    ctx = nullptr;
  }
};

inline double to_double(const std::string& s)
{
  std::stringstream ss(s);
  double v = 0;
  ss >> v;
  return v;
}

}  // namespace zfb_cbench

class ZFPCompressorB200 : public CompressorInterface {
  size_t compressedSize;

  struct Shape {
    int dims;
    size_t nx, ny, nz, numel;
  };
  static bool shape_of(const size_t* n, Shape& s)
  {
    s.dims = 1;
    s.numel = n[0];
    for (int i = 1; i < 5; i++)
      if (n[i] != 0) { s.numel *= n[i]; s.dims++; }
    s.nx = s.ny = s.nz = 1;
    if (s.dims == 1) s.nx = s.numel;
    else if (s.dims == 2) { s.nx = n[1]; s.ny = n[0]; }
    else if (s.dims == 3) { s.nx = n[2]; s.ny = n[1]; s.nz = n[0]; }
    else return false;  // the CUDA path of zfp stops at 3D as well
    return true;
  }
  static int dtype_of(const std::string& t)
  {
    if (t == "float") return ZFB_FLOAT;
    if (t == "double") return ZFB_DOUBLE;
    return 0;
  }
  // bits-per-value -> bits per block, CBench semantics
  bool maxbits_of(int dtype, int dims, unsigned& mb)
  {
    if (compressorParameters.find("abs") != compressorParameters.end() ||
        compressorParameters.find("rel") != compressorParameters.end()) {
      std::cout << "zfp_gpu (B200): only fixed-rate (\"bits\") is supported on the GPU path\n";
      return false;
    }
    double rate = 4;  // default of the reference plugin (int bits = 4)
    int wra = 8;
    auto it = compressorParameters.find("bits");
    if (it != compressorParameters.end()) rate = (double)(int)zfb_cbench::to_double(it->second);
    it = compressorParameters.find("rate");
    if (it != compressorParameters.end()) rate = zfb_cbench::to_double(it->second);
    it = compressorParameters.find("wra");
    if (it != compressorParameters.end()) wra = (int)zfb_cbench::to_double(it->second);
    mb = zfb_maxbits(dtype, dims, rate, wra);
    return mb != 0;
  }

 public:
  ZFPCompressorB200() : compressedSize(0) { compressorName = "zfp_gpu"; }
  ~ZFPCompressorB200() {}

  void init()
  {
    if (zfb_cbench::Shared::get().ensure() != ZFB_OK)
      std::cout << "zfp_gpu (B200): no usable CUDA device, compress/decompress will fail (no CPU fallback)\n";
  }

  int compress(void* input, void*& output, std::string dataType, size_t dataTypeSize, size_t* n)
  {
    zfb_cbench::Shared& sh = zfb_cbench::Shared::get();
    Shape s;
    const int dtype = dtype_of(dataType);
    unsigned mb = 0;
    if (sh.ensure() != ZFB_OK || !shape_of(n, s) || !dtype || !maxbits_of(dtype, s.dims, mb)) {
      std::cout << "compression failed\n";
      return 0;
    }
    Timer cTime;
    cTime.start();
    output = std::malloc(zfb_stream_capacity(s.dims, s.nx, s.ny, s.nz, mb));
    size_t got = 0;
    const int rc = output ? zfb_compress_host(sh.ctx, dtype, s.dims, s.nx, s.ny, s.nz, mb, input, output, &got) : ZFB_ENOMEM;
    if (rc != ZFB_OK || !got) {
      std::cout << "compression failed\n";
      return 0;
    }
    compressedSize = got;
    cbytes = got;
    sh.last_orig = input;
    cTime.stop();
    log << "\n" << compressorName << " ~ InputBytes: " << dataTypeSize * s.numel << ", OutputBytes: " << cbytes
        << ", cRatio: " << (dataTypeSize * s.numel / (float)cbytes) << ", #elements: " << s.numel << std::endl;
    log << compressorName << " ~ CompressTime: " << cTime.getDuration() << " s " << std::endl;
    return 1;
  }

  int decompress(void*& input, void*& output, std::string dataType, size_t dataTypeSize, size_t* n)
  {
    zfb_cbench::Shared& sh = zfb_cbench::Shared::get();
    Shape s;
    const int dtype = dtype_of(dataType);
    unsigned mb = 0;
    if (sh.ensure() != ZFB_OK || !shape_of(n, s) || !dtype || !maxbits_of(dtype, s.dims, mb) || !input) {
      std::cout << "decompression failed\n";
      return 0;
    }
    Timer dTime;
    dTime.start();
    output = std::malloc(s.numel * dataTypeSize);
    // the field uploaded by compress() is still resident: the metric partials are fused into this pass
    const int rc = output ? zfb_decompress_host(sh.ctx, dtype, s.dims, s.nx, s.ny, s.nz, mb, input, output, sh.last_orig)
                          : ZFB_ENOMEM;
    if (rc != ZFB_OK) {
      std::cout << "decompression failed\n";
      return 0;
    }
    std::free(input);
    input = NULL;
    dTime.stop();
    log << compressorName << " ~ DecompressTime: " << dTime.getDuration() << " s " << std::endl;
    return 1;
  }

  void close()
  {
    zfb_cbench::Shared& sh = zfb_cbench::Shared::get();
    if (sh.ctx) zfb_release_resident(sh.ctx);
    sh.last_orig = nullptr;
  }
};

#endif
