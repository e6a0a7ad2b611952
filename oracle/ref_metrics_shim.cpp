/*
 * oracle/ref_metrics_shim.cpp -- TEST INFRASTRUCTURE.
 * Thin extern "C" door onto the reference's OWN metric classes, compiled from the headers where
 * they lie under /root/reference/CBench/metrics (never copied into this repo) against
 * oracle/mpi_stub/mpi.h.  Built by oracle/Makefile into oracle/_ref/libcbench_ref.so.
 * Used to (1) validate oracle/metrics_ref.c and (2) serve as the "reference" CPU baseline
 * for the metric pass.
 *
 * The headers rely on their includer for <unordered_map>, <algorithm>, <limits> and
 * python_histogram (utils.hpp), exactly as CBench/main.cpp provides them.
 */
#include <unordered_map>
#include <algorithm>
#include <limits>
#include <vector>
#include <string>
#include <sstream>
#include <iostream>
#include <cstring>
#include <mpi.h>
#include "utils.hpp"
#include "metricFactory.hpp"

extern "C" {

/* Run reference metric `name` on float arrays; returns getGlobalValue().
 * If want_hist != 0 parameters["histogram"] is set the way main.cpp:320-332 does.
 * log_out / extra_out (nullable) receive getLog() / additionalOutput, truncated to cap bytes. */
double cbref_metric(const char *name, const float *orig, const float *approx, size_t n, int want_hist,
                    char *log_out, size_t log_cap, char *extra_out, size_t extra_cap, double *local_val)
{
  MetricInterface *m = MetricsFactory::createMetric(name);
  if (!m)
    return std::numeric_limits<double>::quiet_NaN();
  if (want_hist)
    m->parameters["histogram"] = "field";
  m->init(MPI_COMM_WORLD);
  m->execute((void *)orig, (void *)approx, n);
  double v = m->getGlobalValue();
  if (local_val)
    *local_val = m->getLocalValue();
  if (log_out && log_cap) {
    std::string s = m->getLog();
    std::strncpy(log_out, s.c_str(), log_cap - 1);
    log_out[log_cap - 1] = 0;
  }
  if (extra_out && extra_cap) {
    std::strncpy(extra_out, m->additionalOutput.c_str(), extra_cap - 1);
    extra_out[extra_cap - 1] = 0;
  }
  m->close();
  delete m;
  return v;
}

/* the reference's histogram-script formatter (utils.hpp:321-352) */
size_t cbref_python_histogram(size_t bins, float lo, float hi, const float *hist, char *out, size_t cap)
{
  std::vector<float> h(hist, hist + bins);
  std::string s = python_histogram(bins, lo, hi, h);
  if (out && cap) {
    std::strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
  }
  return s.size();
}
}
