/*
 * oracle/cbench_step_ref.c -- TEST INFRASTRUCTURE (CPU baseline timing only).
 *
 * One CBench "scalar iteration" per rank as main.cpp does it (main.cpp:262-352): compress, decompress,
 * then the five metric passes, each rank on its own independent sub-field.  Ranks are OpenMP threads here
 * (the image has no MPI); every rank runs the serial restatements in zfp_ref.c / metrics_ref.c.
 * Used by bench.py's cpu_baseline leg and `bench.py --impl reference`.
 */
#include <stddef.h>
#include <stdint.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

size_t zfo_compress(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, const void *in, void *out);
size_t zfo_decompress(int type, int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, const void *in, void *out);
size_t zfo_stream_bytes(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits);
void zfo_abs_error(const float *o, const float *a, size_t n, double out[3], float *hist);
void zfo_rel_error(const float *o, const float *a, size_t n, double out[3], float *hist);
double zfo_mse(const float *o, const float *a, size_t n, double *sum_sq);
double zfo_psnr(const float *o, const float *a, size_t n);
void zfo_minmax(const float *o, const float *a, size_t n, double out[2], float *hist, uint64_t *oob);

static double wall(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* float32 only (the reference metrics cast to float*).  times[3] receives the max over ranks of
 * {compress, decompress, metrics} seconds; vals[5] the rank-0 metric values (abs, rel, mse, psnr, max).
 * Returns the wall time of the whole step. */
double zfo_cbench_step(int dims, size_t nx, size_t ny, size_t nz, unsigned maxbits, const float *in, void *stream,
                       float *out, int nranks, int nthreads, double times[3], double vals[5])
{
  size_t n = nx * (dims >= 2 ? ny : 1) * (dims >= 3 ? nz : 1);
  size_t sbytes = zfo_stream_bytes(dims, nx, ny, nz, maxbits);
  double tc = 0, td = 0, tm = 0;
  double t0 = wall();
  int r;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(static, 1) reduction(max : tc, td, tm)
  for (r = 0; r < nranks; r++) {
    const float *src = in + (size_t)r * n;
    float *dst = out + (size_t)r * n;
    char *st = (char *)stream + (size_t)r * sbytes;
    double a = wall(), b, c, d;
    double v3[3], mm[2], ss;
    zfo_compress(1, dims, nx, ny, nz, maxbits, src, st);
    b = wall();
    zfo_decompress(1, dims, nx, ny, nz, maxbits, st, dst);
    c = wall();
    zfo_abs_error(src, dst, n, v3, NULL);
    if (r == 0 && vals) vals[0] = v3[0];
    zfo_rel_error(src, dst, n, v3, NULL);
    if (r == 0 && vals) vals[1] = v3[0];
    ss = zfo_mse(src, dst, n, NULL);
    if (r == 0 && vals) vals[2] = ss;
    ss = zfo_psnr(src, dst, n);
    if (r == 0 && vals) vals[3] = ss;
    zfo_minmax(src, dst, n, mm, NULL, NULL);
    if (r == 0 && vals) vals[4] = mm[1];
    d = wall();
    if (b - a > tc) tc = b - a;
    if (c - b > td) td = c - b;
    if (d - c > tm) tm = d - c;
  }
  if (times) { times[0] = tc; times[1] = td; times[2] = tm; }
  return wall() - t0;
}
