/*
 * oracle/metrics_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of CBench's five error metrics, one scalar loop each, following the
 * reference headers literally (float-typed subtraction, float-rounded relative error,
 * int(err/binSize) binning, un-normalised minmax histogram):
 *
 *   CBench/metrics/absoluteError.hpp:65-143   zfo_abs_error
 *   CBench/metrics/relativeError.hpp:66-155   zfo_rel_error
 *   CBench/metrics/meansquareError.hpp:55-78  zfo_mse
 *   CBench/metrics/psnrError.hpp:56-96        zfo_psnr
 *   CBench/metrics/minmaxMetric.hpp:60-135    zfo_minmax
 *
 * PINNED: tests/test_oracle_metrics.py checks every function here against the reference's own
 * headers compiled into oracle/_ref/libcbench_ref.so (when /root/reference is present) and against
 * tests/golden/metrics_*.json generated from that build.
 *
 * Single-rank semantics: every MPI_Allreduce is the identity, so "global" == "local".  Multi-rank
 * combination is exercised by tests/test_multirank_gloo.py on the partials.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#define NBINS 1024

/* out[0]=max abs err, out[1]=sum abs err, out[2]=mean abs err; hist (nullable) gets NBINS fractions of n */
void zfo_abs_error(const float *o, const float *a, size_t n, double out[3], float *hist)
{
  double sum = 0, mx = 0; /* the reference's vector starts with n zeros, so max >= 0 */
  size_t i;
  for (i = 0; i < n; i++) {
    double err = (double)fabsf(o[i] - a[i]);
    sum += err;
    if (err > mx) mx = err;
  }
  out[0] = mx;
  out[1] = sum;
  out[2] = sum / (double)n;
  if (hist) {
    uint64_t cnt[NBINS];
    memset(cnt, 0, sizeof cnt);
    memset(hist, 0, NBINS * sizeof(float));
    if (mx != 0) {
      double bin = mx / NBINS;
      for (i = 0; i < n; i++) {
        double err = (double)fabsf(o[i] - a[i]);
        int pos = (int)(err / bin);
        if (pos >= NBINS) pos = pos - 1;
        cnt[pos]++;
      }
      for (i = 0; i < NBINS; i++)
        hist[i] = (float)cnt[i] / (float)n;
    }
  }
}

/* relError<float>(o, a, 1): the T=float return type rounds the quotient to float */
static float rel_one(float o, float a)
{
  double abs_err = (double)fabsf(o - a);
  if ((double)fabsf(o) < 1.0)
    return (float)abs_err;
  return (float)(abs_err / (double)fabsf(o));
}

void zfo_rel_error(const float *o, const float *a, size_t n, double out[3], float *hist)
{
  double sum = 0, mx = 0;
  size_t i;
  for (i = 0; i < n; i++) {
    double err = (double)rel_one(o[i], a[i]);
    sum += err;
    if (err > mx) mx = err;
  }
  out[0] = mx;
  out[1] = sum;
  out[2] = sum / (double)n;
  if (hist) {
    uint64_t cnt[NBINS];
    memset(cnt, 0, sizeof cnt);
    memset(hist, 0, NBINS * sizeof(float));
    if (mx != 0) {
      double bin = mx / NBINS;
      for (i = 0; i < n; i++) {
        double err = (double)rel_one(o[i], a[i]);
        int pos = (int)(err / bin);
        if (pos >= NBINS) pos = pos - 1;
        cnt[pos]++;
      }
      for (i = 0; i < NBINS; i++)
        hist[i] = (float)cnt[i] / (float)n; /* float / size_t promotes to float */
    }
  }
}

/* returns sum of squared error / n; *sum_sq gets the un-normalised sum */
double zfo_mse(const float *o, const float *a, size_t n, double *sum_sq)
{
  double s = 0;
  size_t i;
  for (i = 0; i < n; i++) {
    double d = (double)(o[i] - a[i]);
    s += d * d; /* pow(d, 2.0) is exact d*d */
  }
  if (sum_sq) *sum_sq = s;
  return s / (double)n;
}

double zfo_psnr(const float *o, const float *a, size_t n)
{
  double mx = -999999999, s = 0;
  size_t i;
  for (i = 0; i < n; i++) {
    if ((double)o[i] > mx) mx = o[i];
    double d = (double)(o[i] - a[i]);
    s += d * d;
  }
  return 10 * log10((mx * mx) / (s / (double)n));
}

/* out[0]=min(orig), out[1]=max(orig); hist (nullable) gets NBINS raw counts of approx values.
 * The reference indexes out of bounds when an approx value leaves [min,max] (minmaxMetric.hpp:109-116);
 * here such values are clamped to the end bins and *oob (nullable) counts them.                     */
void zfo_minmax(const float *o, const float *a, size_t n, double out[2], float *hist, uint64_t *oob)
{
  double mx = -DBL_MAX, mn = DBL_MAX;
  size_t i;
  uint64_t bad = 0;
  for (i = 0; i < n; i++) {
    if ((double)o[i] > mx) mx = o[i];
    if ((double)o[i] < mn) mn = o[i];
  }
  out[0] = mn;
  out[1] = mx;
  if (hist) {
    uint64_t cnt[NBINS];
    memset(cnt, 0, sizeof cnt);
    memset(hist, 0, NBINS * sizeof(float));
    if (mx != 0) {
      double bin = (mx - mn) / NBINS;
      for (i = 0; i < n; i++) {
        double v = (double)a[i];
        double q = ((mx - mn) * ((v - mn) / (mx - mn))) / bin;
        int pos;
        if (!(q >= 0)) { pos = 0; bad++; }
        else if (q >= (double)NBINS + 1) { pos = NBINS - 1; bad++; }
        else {
          pos = (int)q;
          if (pos >= NBINS) pos = pos - 1;
          if (pos >= NBINS) { pos = NBINS - 1; bad++; }
        }
        cnt[pos]++;
      }
      for (i = 0; i < NBINS; i++)
        hist[i] = (float)cnt[i];
    }
  }
  if (oob) *oob = bad;
}
