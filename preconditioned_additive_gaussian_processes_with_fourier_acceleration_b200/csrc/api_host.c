/* Part 2 of the C ABI (B200 extensions): device selection, memory helpers, events, counters. */
#include <stdio.h>
#include <stdlib.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

int Nfft4GPB200Init(int device) { return b200_init(device); }
int Nfft4GPB200DeviceCount(void) { return b200_device_count(); }
const char *Nfft4GPB200Version(void) { return "nfft4gp_b200 0.1 (sm_100a; N=32, n_os=64, m=4 Kaiser-Bessel)"; }
long long Nfft4GPB200LaunchCount(void) { return b200_launch_count(); }
void Nfft4GPB200LaunchCountReset(void) { b200_launch_count_reset(); }

void *Nfft4GPB200HostAlloc(size_t bytes) { return b200_host_alloc(bytes); }
void Nfft4GPB200HostFree(void *p) { b200_host_free(p); }
void *Nfft4GPB200DeviceAlloc(size_t bytes) { return b200_malloc(bytes); }
void Nfft4GPB200DeviceFree(void *p) { b200_free(p); }
int Nfft4GPB200MemcpyH2D(void *dst_dev, const void *src_host, size_t bytes) { return b200_h2d(dst_dev, src_host, bytes) || b200_sync(); }
int Nfft4GPB200MemcpyD2H(void *dst_host, const void *src_dev, size_t bytes) { return b200_d2h(dst_host, src_dev, bytes); }
int Nfft4GPB200DeviceSynchronize(void) { return b200_sync(); }

void *Nfft4GPB200EventCreate(void) { return b200_event_create(); }
void Nfft4GPB200EventDestroy(void *ev) { b200_event_destroy(ev); }
int Nfft4GPB200EventRecord(void *ev) { return b200_event_record(ev); }
float Nfft4GPB200EventElapsedMs(void *start, void *stop) { return b200_event_elapsed_ms(start, stop); }

/* per-kernel-class CUDA-event timing on the library stream (bench.py roofline) */
void Nfft4GPB200ProfileStart(void) { b200_prof_start(); }
int Nfft4GPB200ProfileStop(double *sums_ms, long long *counts, int ncls) { return b200_prof_stop(sums_ms, counts, ncls); }

/* fp64 tensor-core (DMMA.8x8x4) throughput of this GPU measured now, TFLOP/s: the roof of the 3-D spread / gather kernels */
double Nfft4GPB200MeasureFp64TensorPeak(void) { return b200_init(-1) ? -1.0 : b200_fp64_tensor_peak_tflops(); }

/* ------------------------------------------------------------------ setup building blocks for the parity tests */
int Nfft4GPB200HostPotrf(int n, double *A) { return b200_potrf_lower(n, A, n); }
int Nfft4GPB200HostSyev(int n, double *A, double *w, int want_vectors) { return want_vectors ? b200_syev(n, A, w) : b200_syev_values(n, A, w); }

/* host-only: symmetric tridiagonal eigen-decomposition used by the stochastic Lanczos quadrature (dense_small.c:b200_stev) */
int Nfft4GPB200HostStev(int n, double *d, double *e, double *Z) { return b200_stev(n, d, e, Z); }

int Nfft4GPB200Dgemm(char transa, char transb, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                     double beta, double *C, int ldc)
{
   size_t ca = (size_t)lda * (size_t)(transa == 'N' ? K : M), cb = (size_t)ldb * (size_t)(transb == 'N' ? N : K), cc = (size_t)ldc * (size_t)N;
   double *dA, *dB, *dC;
   int rc;
   if (b200_init(-1))
      return -1;
   dA = (double *)b200_malloc(sizeof(double) * ca);
   dB = (double *)b200_malloc(sizeof(double) * cb);
   dC = (double *)b200_malloc(sizeof(double) * cc);
   rc = !dA || !dB || !dC || b200_h2d(dA, A, sizeof(double) * ca) || b200_h2d(dB, B, sizeof(double) * cb) || b200_h2d(dC, C, sizeof(double) * cc) ||
        b200k_dgemm(transa, transb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc) || b200_d2h(C, dC, sizeof(double) * cc);
   b200_free(dA);
   b200_free(dB);
   b200_free(dC);
   return rc ? -1 : 0;
}

static double *upload_compact(const double *data, int n, int ldim, int d)
{
   double *d_X = (double *)b200_malloc(sizeof(double) * (size_t)n * d);
   int t;
   for (t = 0; d_X && t < d; t++)
      if (b200_h2d(d_X + (size_t)t * n, data + (size_t)t * ldim, sizeof(double) * (size_t)n))
      {
         b200_free(d_X);
         return NULL;
      }
   return d_X;
}

int Nfft4GPB200Fps(const double *data, int n, int ldim, int d, int k, int *perm_out)
{
   double c[64], *d_X, *d_c;
   int t, i, rc, *d_p;
   if (b200_init(-1) || d > 64 || k > n)
      return -1;
   for (t = 0; t < d; t++)
   {
      double s = 0.0;
      for (i = 0; i < n; i++)
         s += data[(size_t)t * ldim + i] / n;
      c[t] = s;
   }
   d_X = upload_compact(data, n, ldim, d);
   d_c = (double *)b200_malloc(sizeof(double) * 64);
   d_p = (int *)b200_malloc(sizeof(int) * (size_t)k);
   rc = !d_X || !d_c || !d_p || b200_h2d(d_c, c, sizeof(double) * (size_t)d) || b200k_fps(d_X, n, n, d, d_c, k, d_p) ||
        b200_d2h(perm_out, d_p, sizeof(int) * (size_t)k);
   b200_free(d_X);
   b200_free(d_c);
   b200_free(d_p);
   return rc ? -1 : 0;
}

int Nfft4GPB200KnnPattern(const double *data, int n, int ldim, int d, int lfil, int *ia, int *ja)
{
   long long nnz = b200k_knn_nnz(n, lfil);
   double *d_X;
   int rc, *d_ia, *d_ja;
   if (b200_init(-1))
      return -1;
   d_X = upload_compact(data, n, ldim, d);
   d_ia = (int *)b200_malloc(sizeof(int) * ((size_t)n + 1));
   d_ja = (int *)b200_malloc(sizeof(int) * (size_t)nnz);
   rc = !d_X || !d_ia || !d_ja || b200k_knn_pattern(d_X, n, n, d, lfil, d_ia, d_ja) || b200_d2h(ia, d_ia, sizeof(int) * ((size_t)n + 1)) ||
        b200_d2h(ja, d_ja, sizeof(int) * (size_t)nnz);
   b200_free(d_X);
   b200_free(d_ia);
   b200_free(d_ja);
   return rc ? -1 : 0;
}

/* host-only: the range partition of the fused spread / gather launches (nfft_host.c:b200_partition_chunks) */
int Nfft4GPB200DebugPartitionChunks(const int *chunks, int nchunk, int wcnt, int rounds, const double *weight, double flush, int *first)
{
   if (!chunks || nchunk < 0 || wcnt < 1 || rounds < 1 || !weight || !first)
      return -1;
   b200_partition_chunks(chunks, nchunk, wcnt, rounds, weight, flush, first);
   return 0;
}
