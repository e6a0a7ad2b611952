// tests/dropin_check.cpp -- TEST INFRASTRUCTURE.  Compiles the new plugins against the REFERENCE's own
// CompressorInterface / MetricInterface headers (found under /root/reference at test time, never copied),
// the way CBench/main.cpp would see them, and drives them through base-class pointers.
#include <algorithm>
#include <iostream>
#include <limits>
#include <list>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <mpi.h>
#include "utils.hpp"  // as main.cpp: python_histogram & friends come from the includer

#define ZFB_IN_CBENCH_TREE
#include "zfpB200Compressor.hpp"
#include "gpuMetrics.hpp"

int main()
{
  CompressorInterface* c = new ZFPCompressorB200();
  c->compressorParameters["bits"] = "8";
  c->init();
  MetricInterface* m = zfb_create_gpu_metric("psnr");
  m->init(MPI_COMM_WORLD);
  std::cout << c->getCompressorName() << " " << c->getParamsInfo() << " " << m->getMetricName() << std::endl;
  // the histogram text must be byte-identical to the reference's formatter
  std::vector<float> h(1024);
  for (int i = 0; i < 1024; i++) h[i] = 0.001f * (float)i;
  const bool same = zfb_cbench::histogram_script(1024, -1.5f, 2.25f, h) == python_histogram(1024, -1.5f, 2.25f, h);
  std::cout << "histogram_script " << (same ? "identical" : "DIFFERENT") << std::endl;
  float data[64] = {0};
  size_t n[5] = {4, 4, 4, 0, 0};
  void* out = NULL;
  const int rc = c->compress(data, out, "float", 4, n);  // no GPU here: must fail loudly, return 0
  std::cout << "compress rc " << rc << std::endl;
  return same ? 0 : 1;
}
