// cbench_mini.cpp -- the scalar loop of CBench/main.cpp:217-432 for ONE in-memory field, driving the new
// plugins exactly as main.cpp drives any plugin: parameters as strings, compress, decompress, metric
// loop, then the CSV row and the human-readable log.  File I/O, JSON and MPI launch stay with the real
// CBench (out of scope); this driver exists so the plugin classes can be exercised and tested here.
//
//   cbench_mini <raw float32 file | synth> <n0> <n1> <n2> <bits | abs=TOL | rel=PREC | default> [--hist] [--prefix P]
//               [--field NAME] [--name zfp|zfp_gpu|quant] [--scalars K]
//     dims with extent 0 are unused (CBench's size_t n[5] convention), n0 slowest.
//     --scalars K: K scalars one after the other, each loaded into a fresh `new float[]` that is released before
//       the next one is loaded (ioMgr->close(), main.cpp:398) -- the allocator hands out the same addresses again,
//       with different contents (scalar j = the input scaled by j + 1).
//     --name quant: a stand-in for the compressors that are not ours (SZ, BLOSC, ...): uniform quantisation on the
//       CPU with step = the <bits> argument; only the GPU metric classes touch the device then.
//   prints:  CSV header, one CSV row per scalar, then the metric log (same text main.cpp appends to its metrics file)
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

// test stand-in for a CPU compressor of another family: q = round(x / step) as int32; malloc-family buffers
class QuantCompressor : public CompressorInterface {
  float step = 1.0f;
  size_t count = 0;

 public:
  QuantCompressor() { compressorName = "quant"; }
  void init() {}
  int compress(void* input, void*& output, std::string, size_t, size_t* n)
  {
    step = (float)zfb_cbench::to_double(compressorParameters["bits"]);
    count = n[0];
    for (int i = 1; i < 5; i++)
      if (n[i]) count *= n[i];
    int32_t* q = (int32_t*)std::malloc(count * sizeof(int32_t));
    const float* x = (const float*)input;
    for (size_t i = 0; i < count; i++) q[i] = (int32_t)lrintf(x[i] / step);
    output = q;
    cbytes = count * sizeof(int32_t);
    return 1;
  }
  int decompress(void*& input, void*& output, std::string, size_t, size_t*)
  {
    float* y = (float*)std::malloc(count * sizeof(float));
    const int32_t* q = (const int32_t*)input;
    for (size_t i = 0; i < count; i++) y[i] = (float)q[i] * step;
    std::free(input);
    input = NULL;
    output = y;
    return 1;
  }
  void close() {}
};

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
  int nscalars = 1;
  std::string prefix = "zfp_gpu_" + bits + "_", field = "field", cname = "zfp_gpu";
  for (int i = 6; i < argc; i++) {
    if (!strcmp(argv[i], "--hist")) hist = true;
    else if (!strcmp(argv[i], "--prefix") && i + 1 < argc) prefix = argv[++i];
    else if (!strcmp(argv[i], "--field") && i + 1 < argc) field = argv[++i];
    else if (!strcmp(argv[i], "--name") && i + 1 < argc) cname = argv[++i];
    else if (!strcmp(argv[i], "--scalars") && i + 1 < argc) nscalars = atoi(argv[++i]);
  }
  size_t numel = n[0];
  for (int i = 1; i < 3; i++)
    if (n[i]) numel *= n[i];
  std::vector<float> base(numel);
  if (src == "synth") {
    for (size_t i = 0; i < numel; i++) base[i] = 100.0f * sinf(0.001f * (float)(i % 100003)) + 0.01f * (float)(i % 17);
  } else {
    std::ifstream f(src, std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(base.data()), numel * sizeof(float))) {
      std::cerr << "cannot read " << numel << " floats from " << src << "\n";
      return 2;
    }
  }

  const std::vector<std::string> metrics = {"absolute_error", "relative_error", "mse", "psnr", "minmax"};
  std::stringstream csv, info;
  csv << "Compressor_field__params, name, ";
  for (const auto& m : metrics) csv << m << ", ";
  csv << "Compression Throughput(MB/s), DeCompression Throughput(MB/s), Compression Ratio\n";

  CompressorInterface* comp = cname == "quant" ? (CompressorInterface*)new QuantCompressor()
                                               : (CompressorInterface*)new ZFPCompressorB200(cname.c_str());
  comp->init();
  // strConvert::toStr of the JSON value, main.cpp:196-204; "abs=0.003" / "rel=16" set that key instead of "bits"
  // "default" = no compressor parameter at all
  const size_t eq = bits.find('=');
  if (bits == "default") {}
  else if (eq == std::string::npos) comp->compressorParameters["bits"] = bits;
  else comp->compressorParameters[bits.substr(0, eq)] = bits.substr(eq + 1);

  for (int sc = 0; sc < nscalars; sc++) {
    // the loader's array (utils.hpp:137-165: new T[]); released at the end of the scalar like ioMgr->close()
    float* data = new float[numel];
    for (size_t i = 0; i < numel; i++) data[i] = base[i] * (float)(sc + 1);
    const std::string fname = nscalars > 1 ? field + std::to_string(sc) : field;
    csv << comp->getCompressorName() << "_" << fname << "__" << comp->getParamsInfo() << ", " << prefix << ", ";

    Timer cclock, dclock;
    void* cdata = NULL;
    cclock.start();
    const int okc = comp->compress(data, cdata, "float", sizeof(float), n);
    cclock.stop();
    void* decomp = NULL;
    dclock.start();
    const int okd = okc ? comp->decompress(cdata, decomp, "float", sizeof(float), n) : 0;
    dclock.stop();
    if (!okc || !okd) {
      std::cerr << "cbench_mini: the GPU path failed (there is no CPU fallback)\n";
      return 3;
    }
    info << comp->getParamsInfo() << "\n\nField: " << fname << "\n";
    for (const auto& name : metrics) {
      MetricInterface* m = zfb_create_gpu_metric(name);
      if (hist) m->parameters["histogram"] = fname;
      m->init(MPI_COMM_WORLD);
      m->execute(data, decomp, numel);
      info << m->getLog();
      csv << m->getGlobalValue() << ", ";
      if (!m->additionalOutput.empty()) info << "[histogram script: " << m->additionalOutput.size() << " bytes]\n";
      m->close();
      delete m;
    }
    const double mb = (double)(numel * sizeof(float)) / (1024.0 * 1024.0);
    csv << mb / cclock.getDuration() << ", " << mb / dclock.getDuration() << ", "
        << (numel * sizeof(float)) / (float)comp->getCompressedSize() << "\n";
    std::free(decomp);   // main.cpp:395
    delete[] data;       // ioMgr->close(), main.cpp:398
  }
  std::cout << csv.str() << info.str();
  comp->close();
  delete comp;
  zfb_cbench::Shared::get().shutdown();
  return 0;
}
