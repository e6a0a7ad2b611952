// cbench_api.hpp -- the two CBench plugin base classes, as the new plugins need to see them.
//
// Inside the Foresight tree the plugins include the reference's own headers
//   CBench/compressors/compressorInterface.hpp:24-69  (CompressorInterface)
//   CBench/metrics/metricInterface.hpp:40-66          (MetricInterface)
// (define ZFB_IN_CBENCH_TREE).  Outside the tree -- this repository's tests and the
// mini driver -- this header supplies base classes with the same public surface, member names and
// semantics, so that the plugin sources compile unchanged in both places.
#ifndef ZFB_CBENCH_API_HPP
#define ZFB_CBENCH_API_HPP

#ifdef ZFB_IN_CBENCH_TREE
#include "compressorInterface.hpp"
#include "metricInterface.hpp"
#else

#include <cstddef>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef ZFB_WITH_MPI
#include <mpi.h>
#else
#include "mpi_single.hpp"
#endif

#include <chrono>
// Wall-clock stopwatch with the interface the plugins use from CBench/utils/timer.hpp:19-45.
class Timer {
  std::chrono::time_point<std::chrono::system_clock> t0, t1;

 public:
  void start() { t0 = std::chrono::system_clock::now(); }
  void stop() { t1 = std::chrono::system_clock::now(); }
  double getDuration() { return std::chrono::duration<double>(t1 - t0).count(); }
};

// Base of every compressor plugin: string-keyed parameters filled from the JSON by the driver, a log
// stream, and the compressed size of the last compress() call.
class CompressorInterface {
 public:
  std::unordered_map<std::string, std::string> compressorParameters;

  virtual ~CompressorInterface() {}
  virtual void init() = 0;
  // n: size_t[5], n[0] slowest ... unused entries 0.  Returns 1 on success, 0 on failure.
  virtual int compress(void* input, void*& output, std::string dataType, size_t dataTypeSize, size_t* n) = 0;
  virtual int decompress(void*& input, void*& output, std::string dataType, size_t dataTypeSize, size_t* n) = 0;
  virtual void close() = 0;

  std::string getCompressorName() { return compressorName; }
  std::string getLog() { return log.str(); }
  void clearLog() { log.str(""); }
  size_t getCompressedSize() { return cbytes; }
  std::string getCompressorInfo() { return "\nCompressor: " + compressorName + "\n"; }
  // key+value of every parameter joined by '_' (lands in the CSV's first column and histogram file names)
  std::string getParamsInfo()
  {
    std::string out;
    for (const auto& kv : compressorParameters) {
      if (!out.empty()) out += "_";
      out += kv.first + kv.second;
    }
    return out;
  }

 protected:
  std::string compressorName;
  std::stringstream log;
  size_t cbytes = 0;
};

// Base of every metric plugin: local / global value, log stream, free-form parameters ("histogram"),
// and an extra output string (the matplotlib script for histograms).
class MetricInterface {
 public:
  std::unordered_map<std::string, std::string> parameters;
  std::string additionalOutput;

  virtual ~MetricInterface() {}
  virtual void init(MPI_Comm _comm) = 0;
  virtual void execute(void* original, void* approx, size_t n) = 0;
  virtual void close() = 0;

  double getLocalValue() { return val; }
  double getGlobalValue() { return total_val; }
  std::string getMetricName() { return metricName; }
  std::string getMetricInfo() { return "\nMetric: " + metricName + "\n"; }
  std::string getLog() { return log.str(); }
  void clearLog() { log.str(""); }

 protected:
  double val = 0;
  double total_val = 0;
  std::string metricName;
  std::stringstream log;
  MPI_Comm comm;
};

#endif  // ZFB_IN_CBENCH_TREE
#endif
