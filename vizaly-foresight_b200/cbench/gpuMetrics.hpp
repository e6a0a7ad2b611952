// gpuMetrics.hpp -- CBench metric plugins computed on the B200 (include/zfb.h).
//
// Drop-ins for the five reference metrics (factory names unchanged):
//   absolute_error  CBench/metrics/absoluteError.hpp:65-143
//   relative_error  CBench/metrics/relativeError.hpp:66-155
//   mse             CBench/metrics/meansquareError.hpp:55-78
//   psnr            CBench/metrics/psnrError.hpp:56-96
//   minmax          CBench/metrics/minmaxMetric.hpp:60-135
// Each execute(original, approx, n) asks the codec context for the metric partials of (original, approx):
// when the pair is the one the zfp_gpu plugin just produced, they were already accumulated inside the
// decompress kernel and nothing is uploaded or recomputed; otherwise both arrays are uploaded once and
// one streaming kernel produces all partials, which the remaining four metrics of the loop then reuse.
// The cross-rank combination keeps the reference's collectives (MPI_Allreduce MAX / SUM on the same
// quantities) so CSV values and log lines are the same; histograms take a second device pass because
// their bin width depends on the global maximum, exactly as in the reference.
#ifndef ZFB_GPU_METRICS_HPP
#define ZFB_GPU_METRICS_HPP

#include <cmath>
#include <cstdio>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "cbench_api.hpp"
#include "zfb.h"
#include "zfpB200Compressor.hpp"

namespace zfb_cbench {

// Text of the matplotlib script CBench writes for a histogram (utils.hpp:321-352): values through
// std::to_string, i.e. fixed notation with six decimals.
inline std::string histogram_script(size_t bins, float lo, float hi, const std::vector<float>& h)
{
  std::ostringstream o;
  o << "import sys\nimport numpy as np\nimport matplotlib\nmatplotlib.use('agg')\nimport matplotlib.pyplot as plt\n";
  o << "y=[";
  for (size_t i = 0; i < bins; i++) o << std::to_string(h[i]) << (i + 1 < bins ? ", " : "]");
  o << "\nminVal=" << std::to_string(lo) << "\nmaxVal=" << std::to_string(hi) << "\n";
  o << "plotName=sys.argv[0]\nplotName = plotName.replace('.py','.png')\nnumVals = len(y)\n"
       "x = np.linspace(minVal, maxVal, numVals+1)[1:]\nplt.plot(x,y, linewidth=0.5)\nplt.title(plotName)\n"
       "plt.yscale(\"linear\") #log,linear,symlog,logit\nplt.ylabel(\"Frequency\")\nplt.xticks(rotation=90)\n"
       "plt.tight_layout()\nplt.savefig(plotName, dpi=300)\n";
  return o.str();
}

// Shared body of the five metrics.
class GpuMetricBase : public MetricInterface {
 protected:
  int numRanks = 0, myRank = 0;
  zfb_partials local;   // this rank
  zfb_partials global;  // after the reference's all-reduces
  unsigned long long global_n = 0;

  bool partials(void* original, void* approx, size_t n)
  {
    Shared& sh = Shared::get();
    if (sh.ensure() != ZFB_OK) {
      std::cout << metricName << ": no usable CUDA device (no CPU fallback)\n";
      return false;
    }
    sh.metric_begins(metricName);  // a metric name coming round again = the next scalar: stale device copies are dropped
    // like the reference, both arrays are read as float regardless of the loader's type
    if (zfb_metrics_host(sh.ctx, ZFB_FLOAT, original, approx, n, &local) != ZFB_OK) {
      std::cout << metricName << ": " << zfb_last_error(sh.ctx) << "\n";
      return false;
    }
    global = local;
    double mx[4] = {local.max_abs, local.max_rel, local.max_orig, local.neg_min_orig}, gmx[4];
    double sm[3] = {local.sum_sq, local.sum_abs, local.sum_rel}, gsm[3];
    unsigned long long ln = n;
    MPI_Allreduce(mx, gmx, 4, MPI_DOUBLE, MPI_MAX, comm);
    MPI_Allreduce(sm, gsm, 3, MPI_DOUBLE, MPI_SUM, comm);
    MPI_Allreduce(&ln, &global_n, 1, MPI_UNSIGNED_LONG_LONG, MPI_SUM, comm);
    global.max_abs = gmx[0]; global.max_rel = gmx[1]; global.max_orig = gmx[2]; global.neg_min_orig = gmx[3];
    global.sum_sq = gsm[0]; global.sum_abs = gsm[1]; global.sum_rel = gsm[2];
    global.n = global_n;
    return true;
  }

  // second pass: 1024-bin histogram over [lo, hi]; which = 0 abs error, 1 rel error, 2 approx values
  void histogram(int which, double lo, double hi, void* original, void* approx, size_t n, bool normalise)
  {
    Shared& sh = Shared::get();
    std::vector<unsigned long long> cnt(ZFB_NBINS, 0), gcnt(ZFB_NBINS, 0);
    uint64_t oob = 0;
    if (zfb_histogram_host(sh.ctx, ZFB_FLOAT, which, lo, hi, original, approx, n, (uint64_t*)cnt.data(), &oob) != ZFB_OK) {
      std::cout << metricName << ": histogram failed: " << zfb_last_error(sh.ctx) << "\n";
      return;
    }
    MPI_Allreduce(cnt.data(), gcnt.data(), ZFB_NBINS, MPI_UNSIGNED_LONG_LONG, MPI_SUM, comm);
    std::vector<float> h(ZFB_NBINS);
    for (size_t i = 0; i < (size_t)ZFB_NBINS; i++)
      h[i] = normalise ? ((float)gcnt[i]) / (float)global_n : (float)gcnt[i];
    if (myRank == 0) additionalOutput = histogram_script(ZFB_NBINS, (float)lo, (float)hi, h);
  }

 public:
  void init(MPI_Comm _comm)
  {
    comm = _comm;
    MPI_Comm_size(comm, &numRanks);
    MPI_Comm_rank(comm, &myRank);
  }
  void close() {}
};

}  // namespace zfb_cbench

class gpuAbsoluteError : public zfb_cbench::GpuMetricBase {
 public:
  gpuAbsoluteError() { metricName = "absolute_error"; }
  void execute(void* original, void* approx, size_t n)
  {
    if (!partials(original, approx, n)) return;
    val = local.max_abs;
    total_val = global.max_abs;
    log << "-Max Abs Error: " << total_val << std::endl;
    log << " Total Abs Error: " << global.sum_abs << std::endl;
    log << " Mean Abs Error: " << global.sum_abs / global_n << std::endl;
    MPI_Barrier(comm);
    if (parameters.find("histogram") != parameters.end() && total_val != 0)
      histogram(0, 0.0, total_val, original, approx, n, true);
  }
};

class gpuRelativeError : public zfb_cbench::GpuMetricBase {
 public:
  gpuRelativeError() { metricName = "relative_error"; }
  void execute(void* original, void* approx, size_t n)
  {
    if (!partials(original, approx, n)) return;
    val = local.max_rel;
    total_val = global.max_rel;
    log << "-Max Rel Error: " << total_val << std::endl;
    log << " Total Rel Error: " << global.sum_rel << std::endl;
    log << " Mean Rel Error: " << global.sum_rel / global_n << std::endl;
    MPI_Barrier(comm);
    if (parameters.find("histogram") != parameters.end() && total_val != 0)
      histogram(1, 0.0, total_val, original, approx, n, true);
  }
};

class gpuMeanSquareError : public zfb_cbench::GpuMetricBase {
 public:
  gpuMeanSquareError() { metricName = "mse"; }
  void execute(void* original, void* approx, size_t n)
  {
    if (!partials(original, approx, n)) return;
    val = local.sum_sq / n;
    total_val = global.sum_sq / (double)global_n;
    log << "-mse: " << total_val << std::endl;
    MPI_Barrier(comm);
  }
};

class gpuPsnr : public zfb_cbench::GpuMetricBase {
 public:
  gpuPsnr() { metricName = "psnr"; }
  void execute(void* original, void* approx, size_t n)
  {
    if (!partials(original, approx, n)) return;
    // the reference starts its running maximum at -999999999 (psnrError.hpp:58)
    const double lmax = std::max(local.max_orig, -999999999.0), gmax = std::max(global.max_orig, -999999999.0);
    val = 10 * log10(pow(lmax, 2.0) / (local.sum_sq / n));
    total_val = 10 * log10(pow(gmax, 2.0) / (global.sum_sq / global_n));
    log << " local_psnr: " << val << std::endl;
    log << "-psnr: " << total_val << std::endl;
    MPI_Barrier(comm);
  }
};

class gpuMinMax : public zfb_cbench::GpuMetricBase {
 public:
  gpuMinMax() { metricName = "minmax"; }
  void execute(void* original, void* approx, size_t n)
  {
    if (!partials(original, approx, n)) return;
    const double lmin = -local.neg_min_orig, gmin = -global.neg_min_orig;
    log << " local_minmax: " << lmin << " " << local.max_orig << std::endl;
    log << "-minmax: " << gmin << " " << global.max_orig << std::endl;
    MPI_Barrier(comm);
    val = local.max_orig;
    total_val = global.max_orig;
    if (parameters.find("histogram") != parameters.end() && global.max_orig != 0)
      histogram(2, gmin, global.max_orig, NULL, approx, n, false);  // raw counts of the approx values
  }
};

// What MetricsFactory::createMetric (metricFactory.hpp:21-35) returns once the GPU metrics are enabled.
inline MetricInterface* zfb_create_gpu_metric(const std::string& name)
{
  if (name == "absolute_error") return new gpuAbsoluteError();
  if (name == "relative_error") return new gpuRelativeError();
  if (name == "mse") return new gpuMeanSquareError();
  if (name == "psnr") return new gpuPsnr();
  if (name == "minmax") return new gpuMinMax();
  return NULL;
}

#endif
