/* TEST1-style driver in C (TESTS/TEST1/foo.cpp, the reference's NFFT-vs-dense experiment behind its Fig. 2) on
 * libnfft4gp_b200.so: reads the reference's text formats -- feature file `n d` + d rows of n values (feature-major), window
 * file `nwindows dwindows` + nwindows * dwindows integers, -1 padding a short last window (TEST4/foo.cpp:9-120) -- samples
 * n_sample points, builds the dense additive kernel and the NFFT additive kernel side by side (foo.cpp:209-233), applies
 * both and their gradient forms to x = U(0,1) - 0.5 and prints the error lines of foo.cpp:250-293.
 *
 *   test1_matvec <data prefix> <window file> <n_sample (<=0: all)> <kernel 0|1> <f> <l> <mu> <seed>
 *
 * The dense products use the lower triangles the func_kernel returns (what the reference's Nfft4GPDenseMatSymv does). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"

static double *read_features(const char *path, int *n, int *d)
{
   FILE *fp = fopen(path, "r");
   double *data;
   size_t i, cnt;
   if (!fp || fscanf(fp, "%d %d", n, d) != 2 || *n <= 0 || *d <= 0)
   {
      printf("Error reading data file %s\n", path);
      exit(1);
   }
   cnt = (size_t)(*n) * (size_t)(*d);
   data = (double *)malloc(sizeof(double) * cnt);
   for (i = 0; i < cnt; i++)
      if (fscanf(fp, "%lf", data + i) != 1)
      {
         printf("Error reading data file at entry %zu\n", i);
         exit(1);
      }
   fclose(fp);
   return data;
}

static int *read_windows(const char *path, int *nw, int *dw)
{
   FILE *fp = fopen(path, "r");
   int *w, i;
   if (!fp || fscanf(fp, "%d %d", nw, dw) != 2 || *nw <= 0 || *dw <= 0)
   {
      printf("Error reading window file %s\n", path);
      exit(1);
   }
   w = (int *)malloc(sizeof(int) * (size_t)(*nw) * (size_t)(*dw));
   for (i = 0; i < (*nw) * (*dw); i++)
      if (fscanf(fp, "%d", w + i) != 1)
      {
         printf("Error reading window file at entry %d\n", i);
         exit(1);
      }
   fclose(fp);
   return w;
}

/* y = A x with A symmetric, lower triangle stored column-major (upper part ignored) */
static void lower_symv(const double *A, int n, const double *x, double *y)
{
   int i, j;
   for (i = 0; i < n; i++)
      y[i] = 0.0;
   for (j = 0; j < n; j++)
   {
      const double *col = A + (size_t)j * n;
      double xj = x[j], s = 0.0;
      y[j] += col[j] * xj;
      for (i = j + 1; i < n; i++)
      {
         y[i] += col[i] * xj;
         s += col[i] * x[i];
      }
      y[j] += s;
   }
}

int main(int argc, char **argv)
{
   int n1, d, nw, dw, n, kernel, i, j;
   long seed;
   double f, l, mu, *Xall, *X, *x, *y_nfft, *dy_nfft, *y_exact, *dy_exact, *K = NULL, *dK = NULL, *Kn = NULL, *dKn = NULL;
   double err = 0.0, err1 = 0.0, derr[3] = {0, 0, 0}, derr1[3] = {0, 0, 0}, nrmy = 0.0, nrmdy[3] = {0, 0, 0};
   int *windows, *sample;
   void *dense, *h;
   pnfft4gp_kernel pk;
   char path[4096];
   if (argc != 9)
   {
      printf("Usage: %s <data prefix> <window file> <n_sample> <kernel> <f> <l> <mu> <seed>\n", argv[0]);
      return 1;
   }
   snprintf(path, sizeof(path), "%s.train.feature", argv[1]);
   Xall = read_features(path, &n1, &d);
   windows = read_windows(argv[2], &nw, &dw);
   n = atoi(argv[3]);
   kernel = atoi(argv[4]);
   f = atof(argv[5]);
   l = atof(argv[6]);
   mu = atof(argv[7]);
   seed = atol(argv[8]);
   if (n <= 0 || n > n1)
      n = n1;
   printf("n_sample = %d of %d, d = %d, kernel = %d, f = %10.10f, l = %10.10f, mu = %10.10f\n", n, n1, d, kernel, f, l, mu);
   printf("nwindows = %d, dwindows = %d\n", nw, dw);
   printf("Random seed: %ld\n", seed);
   srand((unsigned)seed);
   sample = (int *)malloc(sizeof(int) * (size_t)n1);
   for (i = 0; i < n1; i++)
      sample[i] = i;
   if (n < n1)
      for (i = 0; i < n; i++) /* partial Fisher-Yates: the first n entries are a random sample without replacement */
      {
         int r = i + (int)((double)rand() / ((double)RAND_MAX + 1.0) * (double)(n1 - i)), t = sample[i];
         sample[i] = sample[r];
         sample[r] = t;
      }
   X = (double *)malloc(sizeof(double) * (size_t)n * d);
   for (j = 0; j < d; j++)
      for (i = 0; i < n; i++)
         X[(size_t)j * n + i] = Xall[(size_t)j * n1 + sample[i]];
   free(Xall);
   free(sample);
   if (Nfft4GPB200Init(-1))
   {
      printf("no CUDA device: libnfft4gp_b200 has no CPU fallback\n");
      return 2;
   }
   dense = Nfft4GPKernelAdditiveKernelParamCreate(X, n, n, d, windows, nw, dw, kernel == 0 ? Nfft4GPKernelGaussianKernel : Nfft4GPKernelMatern12Kernel);
   h = Nfft4GPNFFTAdditiveKernelParamCreate(X, n, n, d, windows, nw, dw);
   pk = (pnfft4gp_kernel)dense;
   pk->_params[0] = f;
   pk->_params[1] = l;
   pk->_noise_level = mu;
   pk = (pnfft4gp_kernel)h;
   pk->_params[0] = f;
   pk->_params[1] = l;
   pk->_noise_level = mu;
   if (Nfft4GPKernelAdditiveKernel(dense, X, n, n, d, NULL, 0, NULL, 0, &K, &dK) ||
       (kernel == 0 ? Nfft4GPNFFTAdditiveKernelGaussianKernel : Nfft4GPNFFTAdditiveKernelMatern12Kernel)(h, X, n, n, d, NULL, 0, NULL, 0, &Kn, &dKn))
      return 1;
   x = (double *)malloc(sizeof(double) * (size_t)n);
   y_nfft = (double *)calloc((size_t)n, sizeof(double));
   dy_nfft = (double *)calloc((size_t)3 * n, sizeof(double));
   y_exact = (double *)calloc((size_t)n, sizeof(double));
   dy_exact = (double *)calloc((size_t)3 * n, sizeof(double));
   for (i = 0; i < n; i++)
      x[i] = (double)rand() / (double)RAND_MAX - 0.5;
   Nfft4GPAdditiveNFFTMatSymv(Kn, n, 1.0, x, 0.0, y_nfft);
   Nfft4GPAdditiveNFFTGradMatSymv(dKn, n, 1.0, x, 0.0, dy_nfft);
   lower_symv(K, n, x, y_exact);
   for (j = 0; j < 3; j++)
      lower_symv(dK + (size_t)j * n * n, n, x, dy_exact + (size_t)j * n);
   for (i = 0; i < n; i++)
   {
      double diff = y_nfft[i] - y_exact[i];
      nrmy += y_exact[i] * y_exact[i];
      err += diff * diff;
      if (fabs(diff) > err1)
         err1 = fabs(diff);
      for (j = 0; j < 3; j++)
      {
         double dd = dy_nfft[(size_t)j * n + i] - dy_exact[(size_t)j * n + i];
         nrmdy[j] += dy_exact[(size_t)j * n + i] * dy_exact[(size_t)j * n + i];
         derr[j] += dd * dd;
         if (fabs(dd) > derr1[j])
            derr1[j] = fabs(dd);
      }
   }
   nrmy = sqrt(nrmy);
   for (j = 0; j < 3; j++)
      nrmdy[j] = sqrt(nrmdy[j]);
   printf("Linf Error. Abs: %24.20e ; Rel: %24.20e\n", err1, err1 / nrmy);
   for (j = 0; j < 3; j++)
      printf("Linf Gradient Error %d. Abs: %24.20e ; Rel: %24.20e\n", j + 1, derr1[j], derr1[j] / nrmdy[j]);
   printf("L2 Error. Abs: %24.20e ; Rel: %24.20e\n", sqrt(err), sqrt(err) / nrmy);
   for (j = 0; j < 3; j++)
      printf("L2 Gradient Error %d. Abs: %24.20e ; Rel: %24.20e\n", j + 1, sqrt(derr[j]), sqrt(derr[j]) / nrmdy[j]);
   printf("RESULT n=%d rel_l2=%.6e rel_l2_df=%.6e rel_l2_dl=%.6e rel_l2_dmu=%.6e norm_exact=%.10e\n", n, sqrt(err) / nrmy, sqrt(derr[0]) / nrmdy[0],
          sqrt(derr[1]) / nrmdy[1], sqrt(derr[2]) / nrmdy[2], nrmy);
   Nfft4GPKernelFree(K);
   Nfft4GPKernelFree(dK);
   Nfft4GPKernelParamFree(dense);
   Nfft4GPAdditiveNFFTKernelFree(h);
   free(X);
   free(windows);
   free(x);
   free(y_nfft);
   free(dy_nfft);
   free(y_exact);
   free(dy_exact);
   return 0;
}
