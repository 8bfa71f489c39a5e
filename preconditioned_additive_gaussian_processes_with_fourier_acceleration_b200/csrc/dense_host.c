/* Dense kernel handles (row K1 of SURVEY.md 8(a)): the func_kernel implementations that build dense blocks of
 *    K = f^2 ( exp(-r^2 / 2 l^2) + mu I )          SRC/linearalg/kernels.c:680  (Gaussian;    :495,523,549)
 *    K = f^2 ( exp(-r / l)       + mu I )          SRC/linearalg/kernels.c:2390 (Matern-1/2;  :2218-2222)
 *    additive = mean over the feature windows       SRC/linearalg/kernels.c:3046-3494
 * for the preconditioner setups.  The blocks are evaluated on the GPU (k_dense_additive); through the func_kernel ABI
 * they are returned in host memory exactly as the reference does (column-major; full lower when permr == NULL,
 * K(permr, permr) lower when permc == NULL, K(permr, permc) full and without the noise term otherwise).  The setups
 * of this library recognise these functions and keep the blocks on the device instead (b200_dense_view_*).
 *
 * Deliberate deviations from the reference (both are reference bugs):
 *  - the additive kernel honours `permr` / `n` against the handle's own copy of the features; the reference indexes that
 *    copy with the caller's n, which breaks every call made with a sub-sampled `data` (nys.c:549, fsai.c:374);
 *  - a short LAST window is detected from its trailing -1 entries (the reference tests windows[w * dwindows - 1] of the
 *    wrong windows, kernels.c:3054-3060).
 * Gradient blocks: dKp != NULL returns the three stacked blocks [d/df; d/dl; d/dmu] of the same call shape (kernels.c:495-549,
 * 2218-2222), which the preconditioners' gradient sides and the GP optimizer use. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

typedef struct
{
   int magic;
   int n, P, dw;
   int kernel; /* 0 Gauss, 1 Matern-1/2 */
   int *dims;  /* host [P] */
   double *d_Xw;
   int *d_dims;
} b200_dense;

/* reference: kernels.c:404-441 */
void *Nfft4GPKernelParamCreate(int max_n, int omp)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)b200_xcalloc(1, sizeof(nfft4gp_kernel));
   k->_max_n = max_n;
   k->_omp = omp;
   k->_ldwork = 0;
   return k;
}

static void dense_release(b200_dense *D)
{
   if (!D)
      return;
   b200_free(D->d_Xw);
   b200_free(D->d_dims);
   free(D->dims);
   free(D);
}

/* reference: kernels.c:443-470 */
void Nfft4GPKernelParamFree(void *str)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)str;
   if (!k)
      return;
   if (k->_external && ((b200_dense *)k->_external)->magic == B200_MAGIC_DENSE)
      dense_release((b200_dense *)k->_external);
   if (k->_own_buffer)
      free(k->_buffer);
   if (k->_own_fkernel_buffer_params)
      Nfft4GPKernelParamFree(k->_fkernel_buffer_params);
   free(k->_ibufferp);
   free(k->_dwork);
   free(k);
}

/* reference: kernels.c (func_free for matrices returned through *Kp) */
void Nfft4GPKernelFree(void *str) { free(str); }

static int dense_kernel_id(func_kernel f)
{
   if (f == Nfft4GPKernelGaussianKernel)
      return 0;
   if (f == Nfft4GPKernelMatern12Kernel)
      return 1;
   return -1;
}

/* reference: kernels.c:3046-3096 */
void *Nfft4GPKernelAdditiveKernelParamCreate(double *data, int n, int ldim, int d, int *windows, int nwindows, int dwindows, func_kernel fkernel)
{
   nfft4gp_kernel *k;
   b200_dense *D;
   int kid = dense_kernel_id(fkernel), i, j, skip = 1;
   int *fw = windows;
   double *dst;
   (void)d;
   if (kid < 0)
   {
      printf("Error: Nfft4GPKernelAdditiveKernelParamCreate: the window kernel must be Nfft4GPKernelGaussianKernel or Nfft4GPKernelMatern12Kernel.\n");
      return NULL;
   }
   k = (nfft4gp_kernel *)Nfft4GPKernelParamCreate(n, 1);
   D = (b200_dense *)b200_xcalloc(1, sizeof(b200_dense));
   D->magic = B200_MAGIC_DENSE;
   D->n = n;
   D->P = nwindows;
   D->dw = dwindows;
   D->kernel = kid;
   D->dims = (int *)b200_xcalloc((size_t)nwindows, sizeof(int));
   k->_iparams[0] = nwindows;
   k->_iparams[1] = dwindows;
   k->_iparams[2] = 0;
   while (skip < dwindows && windows[nwindows * dwindows - skip] < 0)
   {
      skip++;
      k->_iparams[2]++;
   }
   k->_ibufferp = (int **)b200_xmalloc(sizeof(int *));
   k->_ibufferp[0] = windows;
   k->_fkernel_buffer = fkernel;
   k->_fkernel_buffer_params = Nfft4GPKernelParamCreate(n, 1);
   k->_own_fkernel_buffer_params = 1;
   k->_buffer = (double *)b200_xcalloc((size_t)n * nwindows * dwindows, sizeof(double));
   k->_own_buffer = 1;
   dst = k->_buffer;
   for (i = 0; i < nwindows; i++)
   {
      int actual = 0;
      for (j = 0; j < dwindows; j++)
         if (fw[0] >= 0)
         {
            memcpy(dst, data + (size_t)fw[0] * ldim, sizeof(double) * (size_t)n);
            fw++;
            dst += n;
            actual++;
         }
      D->dims[i] = actual;
   }
   k->_external = D;
   return k;
}

static int dense_ensure_device(nfft4gp_kernel *k, b200_dense *D)
{
   size_t sz = sizeof(double) * (size_t)D->n * D->P * D->dw;
   if (D->d_Xw)
      return 0;
   if (b200_init(-1))
      return -1;
   D->d_Xw = (double *)b200_malloc(sz);
   D->d_dims = (int *)b200_malloc(sizeof(int) * (size_t)D->P);
   if (!D->d_Xw || !D->d_dims)
      return -1;
   return b200_h2d(D->d_Xw, k->_buffer, sz) || b200_h2d(D->d_dims, D->dims, sizeof(int) * (size_t)D->P) || b200_sync();
}

/* ------------------------------------------------------------------ device view used by the setups */
int b200_dense_view_acquire(func_kernel fkernel, void *fkernel_params, const double *data, int n, int ldim, int d, b200_dense_view *v)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)fkernel_params;
   memset(v, 0, sizeof(*v));
   if (!k)
      return 1;
   v->f = k->_params[0];
   v->l = k->_params[1];
   v->mu = k->_noise_level;
   v->d_all = d;
   if (fkernel == Nfft4GPKernelAdditiveKernel)
   {
      b200_dense *D = (b200_dense *)k->_external;
      if (!D || D->magic != B200_MAGIC_DENSE)
         return 1;
      if (D->n != n)
      {
         printf("Error: the dense additive kernel handle was created for n = %d points, called with n = %d.\n", D->n, n);
         return -1;
      }
      if (dense_ensure_device(k, D))
         return -1;
      v->d_Xw = D->d_Xw;
      v->d_dims = D->d_dims;
      v->n = n;
      v->P = D->P;
      v->dw = D->dw;
      v->kernel = D->kernel;
      return 0;
   }
   if (dense_kernel_id(fkernel) >= 0)
   {
      /* plain kernel over all d features of `data`: one window of dimension d, uploaded for the duration of the setup */
      double *d_X, *tmp = NULL;
      int *d_dims;
      if (b200_init(-1))
         return -1;
      d_X = (double *)b200_malloc(sizeof(double) * (size_t)n * d);
      d_dims = (int *)b200_malloc(sizeof(int));
      if (!d_X || !d_dims)
         return -1;
      if (ldim == n)
      {
         if (b200_h2d(d_X, data, sizeof(double) * (size_t)n * d))
            return -1;
      }
      else
      {
         int t;
         tmp = (double *)b200_xmalloc(sizeof(double) * (size_t)n * d);
         for (t = 0; t < d; t++)
            memcpy(tmp + (size_t)t * n, data + (size_t)t * ldim, sizeof(double) * (size_t)n);
         if (b200_h2d(d_X, tmp, sizeof(double) * (size_t)n * d))
            return -1;
      }
      if (b200_h2d(d_dims, &d, sizeof(int)) || b200_sync())
         return -1;
      free(tmp);
      v->d_Xw = d_X;
      v->d_dims = d_dims;
      v->n = n;
      v->P = 1;
      v->dw = d;
      v->kernel = dense_kernel_id(fkernel);
      v->owns = 1;
      return 0;
   }
   return 1;
}

void b200_dense_view_release(b200_dense_view *v)
{
   if (v->owns)
   {
      b200_free((void *)v->d_Xw);
      b200_free((void *)v->d_dims);
   }
   memset(v, 0, sizeof(*v));
}

int b200_dense_view_block(const b200_dense_view *v, double noise_diag, const int *d_rows, int nr, const int *d_cols, int nc,
                          int lower_only, double *d_K, int ldk)
{
   if (lower_only && b200_memset(d_K, 0, sizeof(double) * (size_t)ldk * (size_t)nc))
      return -1;
   return b200k_dense_additive(v->d_Xw, v->n, v->P, v->dw, v->d_dims, v->kernel, v->f * v->f, v->l, noise_diag, d_rows, nr, d_cols, nc,
                               lower_only, d_K, ldk);
}

/* dK/dl of the same block (no noise term; zero on the diagonal) */
int b200_dense_view_block_dl(const b200_dense_view *v, const int *d_rows, int nr, const int *d_cols, int nc, int lower_only, double *d_K,
                             int ldk)
{
   if (lower_only && b200_memset(d_K, 0, sizeof(double) * (size_t)ldk * (size_t)nc))
      return -1;
   return b200k_dense_additive(v->d_Xw, v->n, v->P, v->dw, v->d_dims, v->kernel + 2, v->f * v->f, v->l, 0.0, d_rows, nr, d_cols, nc,
                               lower_only, d_K, ldk);
}

/* ------------------------------------------------------------------ func_kernel entry points (host in, host out) */
static int dense_call(func_kernel self, void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc,
                      double **Kp, double **dKp)
{
   b200_dense_view v;
   int rc, nr, nc, lower, *d_rows = NULL, *d_cols = NULL;
   double *d_K, *K, *dK = NULL, noise;
   size_t cnt, i;
   if (!Kp && !dKp)
      return 0;
   rc = b200_dense_view_acquire(self, str, data, n, ldim, d, &v);
   if (rc)
   {
      printf("Error: dense kernel handle not recognised.\n");
      return -1;
   }
   if (permr && permc)
   {
      nr = kr;
      nc = kc;
      lower = 0;
      noise = 0.0;
   }
   else if (permr)
   {
      nr = nc = kr;
      lower = 1;
      noise = v.mu * v.f * v.f;
   }
   else
   {
      nr = nc = n;
      lower = 1;
      noise = v.mu * v.f * v.f;
   }
   d_K = (double *)b200_malloc(sizeof(double) * (size_t)nr * nc);
   if (permr)
   {
      d_rows = (int *)b200_malloc(sizeof(int) * (size_t)nr);
      b200_h2d(d_rows, permr, sizeof(int) * (size_t)nr);
   }
   if (permr && permc)
   {
      d_cols = (int *)b200_malloc(sizeof(int) * (size_t)nc);
      b200_h2d(d_cols, permc, sizeof(int) * (size_t)nc);
   }
   rc = b200_dense_view_block(&v, noise, d_rows, nr, permc ? d_cols : d_rows, nc, lower, d_K, nr);
   cnt = (size_t)nr * nc;
   K = (Kp && *Kp) ? *Kp : (double *)b200_xmalloc(sizeof(double) * cnt);
   rc = rc || b200_d2h(K, d_K, sizeof(double) * cnt);
   if (!rc && dKp)
   {
      /* gradient blocks stacked [d/df; d/dl; d/dmu] (kernels.c:495-549, 2218-2222): d/df = (2/f) K including the noise
       * diagonal of the symmetric call shapes, d/dmu = f^2 on that diagonal and zero in the K(permr, permc) shape */
      dK = *dKp ? *dKp : (double *)b200_xmalloc(sizeof(double) * 3 * cnt);
      rc = b200_dense_view_block_dl(&v, d_rows, nr, permc ? d_cols : d_rows, nc, lower, d_K, nr) ||
           b200_d2h(dK + cnt, d_K, sizeof(double) * cnt);
      for (i = 0; i < cnt; i++)
      {
         dK[i] = 2.0 / v.f * K[i];
         dK[2 * cnt + i] = 0.0;
      }
      if (lower)
         for (i = 0; i < (size_t)nr; i++)
            dK[2 * cnt + i * nr + i] = v.f * v.f;
   }
   b200_free(d_K);
   b200_free(d_rows);
   b200_free(d_cols);
   b200_dense_view_release(&v);
   if (rc)
   {
      if (!(Kp && *Kp))
         free(K);
      if (dKp && !*dKp)
         free(dK);
      return -1;
   }
   if (Kp)
      *Kp = K;
   else
      free(K);
   if (dKp)
      *dKp = dK;
   return 0;
}

int Nfft4GPKernelAdditiveKernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc, double **Kp, double **dKp)
{
   return dense_call(Nfft4GPKernelAdditiveKernel, str, data, n, ldim, d, permr, kr, permc, kc, Kp, dKp);
}

int Nfft4GPKernelGaussianKernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc, double **Kp, double **dKp)
{
   return dense_call(Nfft4GPKernelGaussianKernel, str, data, n, ldim, d, permr, kr, permc, kc, Kp, dKp);
}

int Nfft4GPKernelMatern12Kernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc, double **Kp, double **dKp)
{
   return dense_call(Nfft4GPKernelMatern12Kernel, str, data, n, ldim, d, permr, kr, permc, kc, Kp, dKp);
}
