/* A driver in the reference's host language (C) that uses libnfft4gp_b200.so exactly the way the reference's own TESTS
 * drivers use nfft4gp (TESTS/TEST1/foo.cpp:209-293: kernel build + matvec timing; TESTS/TEST3/foo3.cpp:251-291: dense
 * additive kernel handle + preconditioner setup + Krylov solve) -- no Python, no torch, only include/nfft4gp_b200.h.
 *
 *   test3_solve [n] [d] [dwindows] [kernel 0|1] [f] [l_ratio] [mu] [afn_rank] [lfil]
 *
 * Synthetic data as the reference generates it (Nfft4GPDatasetUniformRandom, utils.c:266-286: n^{1/d} U(0,1) after
 * srand(906)); windows {0..dw-1}, {dw..2dw-1}, ...; right-hand side U(0,1) - 0.5 (Nfft4GPVecRand, vecops.c:15-25).
 * Prints the quantities the reference drivers print plus one machine-readable line `RESULT key=value ...`. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "nfft4gp_b200.h"

static double wtime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
   int n = argc > 1 ? atoi(argv[1]) : 20000, d = argc > 2 ? atoi(argv[2]) : 6, dw = argc > 3 ? atoi(argv[3]) : 3;
   int kernel = argc > 4 ? atoi(argv[4]) : 0;
   double f = argc > 5 ? atof(argv[5]) : 1.0, l_ratio = argc > 6 ? atof(argv[6]) : 0.25, mu = argc > 7 ? atof(argv[7]) : 0.01;
   int rank = argc > 8 ? atoi(argv[8]) : 300, lfil = argc > 9 ? atoi(argv[9]) : 100;
   int nwindows = (d + dw - 1) / dw, i, j, iters = 0, iters0 = 0, maxits = 1000;
   double l = l_ratio * pow((double)n, 1.0 / (double)d), scale = pow((double)n, 1.0 / (double)d);
   double *X = (double *)malloc(sizeof(double) * (size_t)n * d), *x = (double *)malloc(sizeof(double) * (size_t)n);
   double *y = (double *)calloc((size_t)n, sizeof(double)), *sol = (double *)calloc((size_t)n, sizeof(double));
   double *res = (double *)malloc(sizeof(double) * (size_t)n), *K = NULL, *dK = NULL, *hist = NULL, *hist0 = NULL;
   int *windows = (int *)malloc(sizeof(int) * (size_t)nwindows * dw);
   double t0, t_build, t_mv, t_setup, t_solve, t_solve0, rel = 0.0, rel0 = 0.0, nb = 0.0, nr = 0.0;
   void *h, *dense, *afn;
   pnfft4gp_kernel pk;
   func_kernel build = kernel == 0 ? Nfft4GPNFFTAdditiveKernelGaussianKernel : Nfft4GPNFFTAdditiveKernelMatern12Kernel;

   if (Nfft4GPB200Init(-1))
   {
      printf("no CUDA device: libnfft4gp_b200 has no CPU fallback\n");
      return 2;
   }
   srand(906);
   for (i = 0; i < n * d; i++)
      X[i] = scale * ((double)rand() / (double)RAND_MAX);
   for (i = 0; i < n; i++)
      x[i] = (double)rand() / (double)RAND_MAX - 0.5;
   for (i = 0; i < nwindows; i++)
      for (j = 0; j < dw; j++)
         windows[i * dw + j] = i * dw + j < d ? i * dw + j : -1; /* the last window may be shorter (nfft_interface.c:630-636) */

   /* TEST1/foo.cpp:209-233: handle, hyper-parameters written directly, kernel build */
   h = Nfft4GPNFFTAdditiveKernelParamCreate(X, n, n, d, windows, nwindows, dw);
   pk = (pnfft4gp_kernel)h;
   pk->_params[0] = f;
   pk->_params[1] = l;
   pk->_noise_level = mu;
   t0 = wtime();
   if (build(h, X, n, n, d, NULL, 0, NULL, 0, &K, &dK))
      return 1;
   t_build = wtime() - t0;
   /* TEST1/foo.cpp:243-251: matvec timing */
   Nfft4GPAdditiveNFFTMatSymv(K, n, 1.0, x, 0.0, y);
   t0 = wtime();
   for (i = 0; i < 10; i++)
      Nfft4GPAdditiveNFFTMatSymv(K, n, 1.0, x, 0.0, y);
   t_mv = (wtime() - t0) / 10.0;

   /* TEST3/foo3.cpp:251-265: dense additive kernel handle for the preconditioner setup */
   dense = Nfft4GPKernelAdditiveKernelParamCreate(X, n, n, d, windows, nwindows, dw,
                                                  kernel == 0 ? Nfft4GPKernelGaussianKernel : Nfft4GPKernelMatern12Kernel);
   pk = (pnfft4gp_kernel)dense;
   pk->_params[0] = f;
   pk->_params[1] = l;
   pk->_noise_level = mu;
   afn = Nfft4GPPrecondAFNCreate();
   Nfft4GPPrecondAFNSetRank(afn, rank);
   Nfft4GPPrecondAFNSetLfil(afn, lfil);
   t0 = wtime();
   if (Nfft4GPPrecondAFNSetupWithKernel(X, n, n, d, Nfft4GPKernelAdditiveKernel, dense, 0, afn))
      return 1;
   t_setup = wtime() - t0;

   /* the solve: K sol = x to 1e-6, with and without the preconditioner (INC/_solvers.h:102) */
   t0 = wtime();
   Nfft4GPSolverPcg(K, n, Nfft4GPAdditiveNFFTMatSymv, afn, Nfft4GPPrecondAFNSolve, sol, x, maxits, 0, 1e-6, &rel, &hist, &iters, 0);
   t_solve = wtime() - t0;
   for (i = 0; i < n; i++)
      res[i] = x[i];
   Nfft4GPAdditiveNFFTMatSymv(K, n, -1.0, sol, 1.0, res); /* res = b - K sol */
   for (i = 0; i < n; i++)
   {
      nb += x[i] * x[i];
      nr += res[i] * res[i];
   }
   for (i = 0; i < n; i++)
      y[i] = 0.0;
   t0 = wtime();
   Nfft4GPSolverPcg(K, n, Nfft4GPAdditiveNFFTMatSymv, NULL, NULL, y, x, maxits, 0, 1e-6, &rel0, &hist0, &iters0, 0);
   t_solve0 = wtime() - t0;

   printf("n = %d, d = %d, %d windows of %d, kernel = %d, f = %g, l = %g, mu = %g\n", n, d, nwindows, dw, kernel, f, l, mu);
   printf("kernel build %.4f s, matvec %.6f s, AFN(rank %d, lfil %d) setup %.4f s\n", t_build, t_mv, rank, lfil, t_setup);
   printf("PCG + AFN: %d iterations, %.4f s, relative residual %.3e (true %.3e); plain CG: %d iterations, %.4f s, %.3e\n", iters, t_solve,
          rel, sqrt(nr / nb), iters0, t_solve0, rel0);
   printf("RESULT n=%d iters_afn=%d iters_plain=%d rel_afn=%.6e true_rel_afn=%.6e rel_plain=%.6e matvec_s=%.6e launches=%lld\n", n, iters,
          iters0, rel, sqrt(nr / nb), rel0, t_mv, Nfft4GPB200LaunchCount());

   Nfft4GPPrecondAFNFree(afn);
   Nfft4GPKernelParamFree(dense);
   Nfft4GPAdditiveNFFTKernelFree(h);
   free(hist);
   free(hist0);
   free(X);
   free(x);
   free(y);
   free(sol);
   free(res);
   free(windows);
   return 0;
}
