// cbench_mini.cpp -- the scalar loop of CBench/main.cpp:217-432 for ONE in-memory field, driving the new
// plugins exactly as main.cpp drives any plugin: parameters as strings, compress, decompress, metric
// loop, then the CSV row and the human-readable log.  File I/O, JSON and MPI launch stay with the real
// CBench (out of scope); this driver exists so the plugin classes can be exercised and tested here.
//
//   cbench_mini <raw float32 file | synth> <n0> <n1> <n2> <bits | abs=TOL | rel=PREC> [--hist] [--prefix P]
//               [--field NAME] [--name zfp|zfp_gpu]
//     dims with extent 0 are unused (CBench's size_t n[5] convention), n0 slowest.
//   prints:  CSV header, CSV row, then the metric log (same text main.cpp appends to its metrics file)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "gpuMetrics.hpp"
#include "zfpB200Compressor.hpp"

int main(int argc, char** argv)
{
  if (argc < 6) {
    std::cerr << "usage: cbench_mini <file.f32|synth> n0 n1 n2 <bits|abs=TOL|rel=PREC> [--hist] [--prefix P] [--field NAME] [--name zfp|zfp_gpu]\n";
    return 2;
  }
  const std::string src = argv[1];
  size_t n[5] = {(size_t)atoll(argv[2]), (size_t)atoll(argv[3]), (size_t)atoll(argv[4]), 0, 0};
  const std::string bits = argv[5];
  bool hist = false;
  std::string prefix = "zfp_gpu_" + bits + "_", field = "field", cname = "zfp_gpu";
  for (int i = 6; i < argc; i++) {
    if (!strcmp(argv[i], "--hist")) hist = true;
    else if (!strcmp(argv[i], "--prefix") && i + 1 < argc) prefix = argv[++i];
    else if (!strcmp(argv[i], "--field") && i + 1 < argc) field = argv[++i];
    else if (!strcmp(argv[i], "--name") && i + 1 < argc) cname = argv[++i];
  }
  size_t numel = n[0];
  for (int i = 1; i < 3; i++)
    if (n[i]) numel *= n[i];
  std::vector<float> data(numel);
  if (src == "synth") {
    for (size_t i = 0; i < numel; i++) data[i] = 100.0f * sinf(0.001f * (float)(i % 100003)) + 0.01f * (float)(i % 17);
  } else {
    std::ifstream f(src, std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(data.data()), numel * sizeof(float))) {
      std::cerr << "cannot read " << numel << " floats from " << src << "\n";
      return 2;
    }
  }

  const std::vector<std::string> metrics = {"absolute_error", "relative_error", "mse", "psnr", "minmax"};
  std::stringstream csv, info;
  csv << "Compressor_field__params, name, ";
  for (const auto& m : metrics) csv << m << ", ";
  csv << "Compression Throughput(MB/s), DeCompression Throughput(MB/s), Compression Ratio\n";

  CompressorInterface* comp = new ZFPCompressorB200(cname.c_str());
  comp->init();
  // strConvert::toStr of the JSON value, main.cpp:196-204; "abs=0.003" / "rel=16" set that key instead of "bits"
  const size_t eq = bits.find('=');
  if (eq == std::string::npos) comp->compressorParameters["bits"] = bits;
  else comp->compressorParameters[bits.substr(0, eq)] = bits.substr(eq + 1);
  csv << comp->getCompressorName() << "_" << field << "__" << comp->getParamsInfo() << ", " << prefix << ", ";

  Timer cclock, dclock;
  void* cdata = NULL;
  cclock.start();
  const int okc = comp->compress(data.data(), cdata, "float", sizeof(float), n);
  cclock.stop();
  void* decomp = NULL;
  dclock.start();
  const int okd = okc ? comp->decompress(cdata, decomp, "float", sizeof(float), n) : 0;
  dclock.stop();
  if (!okc || !okd) {
    std::cerr << "cbench_mini: the GPU path failed (there is no CPU fallback)\n";
    return 3;
  }
  info << comp->getParamsInfo() << "\n\nField: " << field << "\n";
  for (const auto& name : metrics) {
    MetricInterface* m = zfb_create_gpu_metric(name);
    if (hist) m->parameters["histogram"] = field;
    m->init(MPI_COMM_WORLD);
    m->execute(data.data(), decomp, numel);
    info << m->getLog();
    csv << m->getGlobalValue() << ", ";
    if (!m->additionalOutput.empty()) info << "[histogram script: " << m->additionalOutput.size() << " bytes]\n";
    m->close();
    delete m;
  }
  const double mb = (double)(numel * sizeof(float)) / (1024.0 * 1024.0);
  csv << mb / cclock.getDuration() << ", " << mb / dclock.getDuration() << ", "
      << (numel * sizeof(float)) / (float)comp->getCompressedSize() << "\n";
  std::cout << csv.str() << info.str();
  std::free(decomp);
  comp->close();
  delete comp;
  zfb_cbench::Shared::get().shutdown();
  return 0;
}
