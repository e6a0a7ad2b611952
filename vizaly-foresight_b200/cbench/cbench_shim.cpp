// cbench_shim.cpp -- libzfb_cbench.so: the plugin CLASSES behind a C interface, so that a harness without a C++
// toolchain at hand (bench.py's end-to-end leg, tests) can drive them exactly as CBench/main.cpp:262-352 does:
//   compress(data, cdata) -> decompress(cdata, decomp) -> one new metric object per metric name, execute(data, decomp)
//   -> std::free(decomp).
// The caller's array plays the loader's `new T[]` (pageable memory); the outputs are the plugin's own malloc's.
#include <cstdlib>
#include <cstring>
#include <string>

#include "gpuMetrics.hpp"
#include "zfpB200Compressor.hpp"

namespace {
struct handle {
  CompressorInterface* comp;
};
const char* const kMetrics[5] = {"absolute_error", "relative_error", "mse", "psnr", "minmax"};
}  // namespace

extern "C" void* zfbc_open(const char* name)
{
  handle* h = new handle;
  h->comp = new ZFPCompressorB200(name && *name ? name : "zfp_gpu");
  h->comp->init();
  return h;
}

extern "C" int zfbc_set(void* hv, const char* key, const char* value)
{
  if (!hv || !key || !value) return -1;
  static_cast<handle*>(hv)->comp->compressorParameters[key] = value;
  return 0;
}

// One scalar through the plugin.  n = CBench's extents (n0 slowest, unused = 0), dtype "float" | "double".
// metrics5 = global values of absolute_error, relative_error, mse, psnr, minmax; seconds3 = wall time of compress,
// decompress and the five metric executes; returns 0, or -1 if a plugin call failed.
extern "C" int zfbc_step(void* hv, void* data, const char* dtype, size_t n0, size_t n1, size_t n2, double* metrics5,
                         double* seconds3, size_t* cbytes)
{
  if (!hv || !data) return -1;
  CompressorInterface* comp = static_cast<handle*>(hv)->comp;
  size_t n[5] = {n0, n1, n2, 0, 0};
  size_t numel = n0;
  for (int i = 1; i < 3; i++)
    if (n[i]) numel *= n[i];
  const std::string dt = dtype && *dtype ? dtype : "float";
  const size_t esz = dt == "double" ? 8 : 4;
  Timer t;
  void* cdata = NULL;
  void* decomp = NULL;
  t.start();
  const int okc = comp->compress(data, cdata, dt, esz, n);
  t.stop();
  if (seconds3) seconds3[0] = t.getDuration();
  if (!okc) return -1;
  t.start();
  const int okd = comp->decompress(cdata, decomp, dt, esz, n);
  t.stop();
  if (seconds3) seconds3[1] = t.getDuration();
  if (!okd) return -1;
  if (cbytes) *cbytes = comp->getCompressedSize();
  t.start();
  if (dt == "float") {  // the reference's metric classes read both arrays as float
    for (int i = 0; i < 5; i++) {
      MetricInterface* m = zfb_create_gpu_metric(kMetrics[i]);
      m->init(MPI_COMM_WORLD);
      m->execute(data, decomp, numel);
      if (metrics5) metrics5[i] = m->getGlobalValue();
      m->close();
      delete m;
    }
  }
  t.stop();
  if (seconds3) seconds3[2] = t.getDuration();
  std::free(decomp);  // main.cpp:395
  comp->clearLog();
  return 0;
}

extern "C" void zfbc_close(void* hv)
{
  if (!hv) return;
  handle* h = static_cast<handle*>(hv);
  h->comp->close();
  delete h->comp;
  delete h;
  zfb_cbench::Shared::get().shutdown();
}
