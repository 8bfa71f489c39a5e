/* Dense Cholesky "preconditioner" (the exact solver of the reference's small-n experiments), SRC/preconds/chol.c.
 *   setup   chol.c:428-640   K (+ noise) from the dense func_kernel, optional shift sqrt(n) ulp(||K||_F) (`stable`), K = L L'
 *   solve   chol.c:111-136   x = L^{-T} L^{-1} rhs      (here: two GEMVs with the explicit L^{-1} built next to L on the device)
 *   logdet  chol.c:293-311   2 sum log L_ii
 *   trace   chol.c:231-291   tr(K^{-1} dK_i), i = f, l, mu   (reference: 2 sum dL_ii / L_ii with dL = L Phi(L^{-1} dK_i L^{-T}),
 *                            whose diagonal is half the diagonal of L^{-1} dK_i L^{-T}: the same number)
 *   dvp     chol.c:138-229   y_i = K^{-1} (dK_i x)
 * The factor, its inverse and (with require_grad) the full K and dK/dl stay in HBM: n is limited by 2..4 n^2 doubles.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

#define B200_MAGIC_CHOL 0x42323043 /* "B20C" */
#define CHOL_MAX_N 46340           /* n^2 fits an int (vector kernels take int lengths) */

typedef struct
{
   int magic;
   int stable;
   int n, has_grad, own_kernel;
   double f;           /* hyper-parameter of the setup (own kernels: dK/df = (2/f) K, dK/dmu = f^2 I) */
   double *d_L;        /* n x n, lower triangle = L */
   double *d_Linv;     /* n x n explicit inverse, upper triangle zero */
   double *d_K;        /* require_grad: full symmetric K with the noise diagonal, without the shift */
   double *d_dK[3];    /* require_grad: full dK/df, dK/dl, dK/dmu; [0] and [2] NULL for this library's own kernels */
   double *d_work;     /* 2n + 4 */
   double *d_x, *d_rhs;
   int cap;
   double logdet;
   double traces[3];
   int traces_ready;
   double tset;
} b200_chol;

static int stage(double **d_x, double **d_rhs, int *cap, int n)
{
   if (*cap < n)
   {
      b200_free(*d_x);
      b200_free(*d_rhs);
      *d_x = (double *)b200_malloc(sizeof(double) * (size_t)n);
      *d_rhs = (double *)b200_malloc(sizeof(double) * (size_t)3 * n);
      *cap = n;
   }
   return (*d_x && *d_rhs) ? 0 : -1;
}

void *Nfft4GPPrecondCholCreate(void)
{
   b200_chol *s = (b200_chol *)b200_xcalloc(1, sizeof(b200_chol));
   s->magic = B200_MAGIC_CHOL;
   return s;
}

static void chol_release(b200_chol *s)
{
   int i;
   b200_free(s->d_L);
   b200_free(s->d_Linv);
   b200_free(s->d_K);
   for (i = 0; i < 3; i++)
   {
      b200_free(s->d_dK[i]);
      s->d_dK[i] = NULL;
   }
   b200_free(s->d_work);
   b200_free(s->d_x);
   b200_free(s->d_rhs);
   s->d_L = s->d_Linv = s->d_K = s->d_work = s->d_x = s->d_rhs = NULL;
   s->n = s->cap = s->has_grad = s->traces_ready = 0;
}

void Nfft4GPPrecondCholReset(void *str) /* chol.c:66-99: keeps the handle and the `stable` switch */
{
   b200_chol *s = (b200_chol *)str;
   if (s && s->magic == B200_MAGIC_CHOL)
      chol_release(s);
}

void Nfft4GPPrecondCholFree(void *str)
{
   b200_chol *s = (b200_chol *)str;
   if (!s || s->magic != B200_MAGIC_CHOL)
      return;
   chol_release(s);
   s->magic = 0;
   free(s);
}

void Nfft4GPPrecondCholSetStable(void *str, int stable)
{
   b200_chol *s = (b200_chol *)str;
   if (s && s->magic == B200_MAGIC_CHOL)
      s->stable = stable;
}

/* host lower triangle (column-major, ld n) -> full symmetric, in place */
static void symmetrize_lower(int n, double *A)
{
   int i, j;
   for (j = 0; j < n; j++)
      for (i = j + 1; i < n; i++)
         A[(size_t)i * n + j] = A[(size_t)j * n + i];
}

int Nfft4GPPrecondCholSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad,
                                      void *vchol_mat)
{
   b200_chol *s = (b200_chol *)vchol_mat;
   nfft4gp_kernel *pk = (nfft4gp_kernel *)fkernel_params;
   b200_dense_view v;
   size_t nn = (size_t)n * (size_t)n;
   int own, info = 0, rc = -1, i, *idx = NULL, *d_idx = NULL;
   double *h = NULL, *hd = NULL, *diag = NULL;
   if (!s || s->magic != B200_MAGIC_CHOL || !fkernel || !fkernel_params || n <= 0)
      return -1;
   if (n > CHOL_MAX_N)
   {
      printf("Error: the dense Cholesky preconditioner is limited to n <= %d.\n", CHOL_MAX_N);
      return -1;
   }
   if (b200_init(-1))
      return -1;
   chol_release(s);
   own = b200_dense_view_acquire(fkernel, fkernel_params, data, n, ldim, d, &v);
   if (own < 0)
      return -1;
   s->n = n;
   s->own_kernel = own == 0;
   s->f = pk->_params[0];
   s->d_L = (double *)b200_malloc(sizeof(double) * nn);
   s->d_Linv = (double *)b200_malloc(sizeof(double) * nn);
   s->d_work = (double *)b200_malloc(sizeof(double) * ((size_t)2 * n + 4));
   if (!s->d_L || !s->d_Linv || !s->d_work)
      goto done;
   if (require_grad)
   {
      s->d_K = (double *)b200_malloc(sizeof(double) * nn);
      s->d_dK[1] = (double *)b200_malloc(sizeof(double) * nn);
      if (!s->d_K || !s->d_dK[1])
         goto done;
   }
   if (own == 0)
   {
      double noise = v.mu * v.f * v.f;
      if (b200_dense_view_block(&v, noise, NULL, n, NULL, n, 1, s->d_L, n))
         goto done;
      if (require_grad && (b200_dense_view_block(&v, noise, NULL, n, NULL, n, 0, s->d_K, n) ||
                           b200_dense_view_block_dl(&v, NULL, n, NULL, n, 0, s->d_dK[1], n)))
         goto done;
   }
   else
   {
      /* foreign kernel: the reference's own host call (chol.c:447 / 478), then upload */
      i = fkernel(fkernel_params, data, n, ldim, d, NULL, 0, NULL, 0, &h, require_grad ? &hd : NULL);
      if (i || !h || (require_grad && !hd))
      {
         printf("Error: the kernel function failed in the Cholesky setup.\n");
         goto done;
      }
      if (b200_h2d(s->d_L, h, sizeof(double) * nn))
         goto done;
      if (require_grad)
      {
         symmetrize_lower(n, h);
         if (b200_h2d(s->d_K, h, sizeof(double) * nn))
            goto done;
         for (i = 0; i < 3; i++)
         {
            symmetrize_lower(n, hd + (size_t)i * nn);
            if (i != 1)
               s->d_dK[i] = (double *)b200_malloc(sizeof(double) * nn);
            if (!s->d_dK[i] || b200_h2d(s->d_dK[i], hd + (size_t)i * nn, sizeof(double) * nn))
               goto done;
         }
      }
      if (b200_sync())
         goto done;
   }
   if (s->stable)
   {
      /* chol.c:449-463: shift by sqrt(n) ulp(||K||_F) */
      double fro, nu;
      if (!h)
      {
         h = (double *)b200_xmalloc(sizeof(double) * nn);
         if (b200_d2h(h, s->d_L, sizeof(double) * nn))
            goto done;
      }
      fro = b200_sym_fro_lower(n, h, n);
      nu = b200_chol_shift(n, fro);
      if (b200k_add_diag(n, s->d_L, n, nu))
         goto done;
   }
   if (b200k_potrf_inverse(n, s->d_L, s->d_Linv, &info))
      goto done;
   if (info)
   {
      printf("Error: Cholesky of the %d x %d kernel matrix failed at pivot %d.\n", n, n, info);
      goto done;
   }
   /* log det = 2 sum log L_ii */
   idx = (int *)b200_xmalloc(sizeof(int) * (size_t)n);
   diag = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   for (i = 0; i < n; i++)
      idx[i] = i * (n + 1);
   d_idx = (int *)b200_malloc(sizeof(int) * (size_t)n);
   if (!d_idx || b200_h2d(d_idx, idx, sizeof(int) * (size_t)n) || b200k_gather_idx(n, d_idx, s->d_L, s->d_work) ||
       b200_d2h(diag, s->d_work, sizeof(double) * (size_t)n))
      goto done;
   s->logdet = 0.0;
   for (i = 0; i < n; i++)
      s->logdet += log(diag[i]);
   s->logdet *= 2.0;
   s->has_grad = require_grad != 0;
   rc = 0;
done:
   free(h);
   free(hd);
   free(idx);
   free(diag);
   b200_free(d_idx);
   if (own == 0)
      b200_dense_view_release(&v);
   if (rc)
      chol_release(s);
   return rc;
}

/* x = L^{-T} (L^{-1} rhs); d_x and d_rhs may alias */
static int chol_solve_dev_n(b200_chol *s, double *d_x, const double *d_rhs)
{
   int n = s->n;
   return b200k_gemv('N', n, n, 1.0, s->d_Linv, d_rhs, 0.0, s->d_work) || b200k_gemv('T', n, n, 1.0, s->d_Linv, s->d_work, 0.0, d_x);
}

int b200_chol_solve_dev(void *vc, int n, double *d_x, const double *d_rhs)
{
   b200_chol *s = (b200_chol *)vc;
   if (!s || s->magic != B200_MAGIC_CHOL || !s->d_Linv || n != s->n)
   {
      printf("Error: Cholesky preconditioner not set up for n = %d.\n", n);
      return -1;
   }
   return chol_solve_dev_n(s, d_x, d_rhs);
}

int Nfft4GPPrecondCholSolve(void *vchol_mat, int n, double *x, double *rhs) /* chol.c:111-136 */
{
   b200_chol *s = (b200_chol *)vchol_mat;
   if (!s || s->magic != B200_MAGIC_CHOL || stage(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   if (b200_h2d(s->d_rhs, rhs, sizeof(double) * (size_t)n) || b200_chol_solve_dev(s, n, s->d_x, s->d_rhs))
      return -1;
   return b200_d2h(x, s->d_x, sizeof(double) * (size_t)n);
}

double Nfft4GPPrecondCholLogdet(void *vchol_mat) /* chol.c:293-311 */
{
   b200_chol *s = (b200_chol *)vchol_mat;
   if (!s || s->magic != B200_MAGIC_CHOL || !s->d_L)
      return 0.0;
   return s->logdet;
}

static int dot_to_host(int n, const double *a, const double *b, double *d_out, double *out)
{
   return b200k_dot(n, a, b, d_out) || b200_d2h(out, d_out, sizeof(double));
}

int Nfft4GPPrecondCholTrace(void *vchol_mat, double **tracesp) /* chol.c:231-291 */
{
   b200_chol *s = (b200_chol *)vchol_mat;
   double *tr;
   int i;
   if (!s || s->magic != B200_MAGIC_CHOL)
      return -1;
   if (!s->has_grad)
   {
      printf("Setup chol without gradient, trace not supported.\n");
      return -1;
   }
   if (!tracesp)
   {
      printf("Trace pointer cannot be NULL\n");
      return -1;
   }
   if (!s->traces_ready)
   {
      /* tr(K^{-1} dK_i) = <K^{-1}, dK_i>_F with K^{-1} = Linv' Linv */
      int n = s->n, nn = n * n;
      double *d_Kinv = (double *)b200_malloc(sizeof(double) * (size_t)nn), *d_out = s->d_work + 2 * (size_t)n;
      int rc;
      if (!d_Kinv)
         return -1;
      rc = b200k_dgemm('T', 'N', n, n, n, 1.0, s->d_Linv, n, s->d_Linv, n, 0.0, d_Kinv, n);
      if (s->own_kernel)
      {
         double kk = 0.0, ll = 0.0;
         rc = rc || dot_to_host(nn, d_Kinv, s->d_K, d_out, &kk) || dot_to_host(nn, s->d_Linv, s->d_Linv, d_out, &ll);
         s->traces[0] = 2.0 / s->f * kk;
         s->traces[2] = s->f * s->f * ll;
      }
      else
         rc = rc || dot_to_host(nn, d_Kinv, s->d_dK[0], d_out, &s->traces[0]) || dot_to_host(nn, d_Kinv, s->d_dK[2], d_out, &s->traces[2]);
      rc = rc || dot_to_host(nn, d_Kinv, s->d_dK[1], d_out, &s->traces[1]);
      b200_free(d_Kinv);
      if (rc)
         return -1;
      s->traces_ready = 1;
   }
   tr = *tracesp ? *tracesp : (double *)b200_xcalloc(3, sizeof(double));
   for (i = 0; i < 3; i++)
      tr[i] = 2.0 * tr[i] + s->traces[i]; /* the reference accumulates into the caller's array and then doubles it (chol.c:270-280) */
   *tracesp = tr;
   return 0;
}

int Nfft4GPPrecondCholDvp(void *vchol_mat, int n, int *mask, double *x, double **yp) /* chol.c:138-229 */
{
   b200_chol *s = (b200_chol *)vchol_mat;
   double *y, *d_y;
   int i, rc = 0;
   if (!s || s->magic != B200_MAGIC_CHOL)
      return -1;
   if (!s->has_grad)
   {
      printf("Setup chol without gradient, dvp not supported.\n");
      return -1;
   }
   if (!yp)
   {
      printf("output pointer cannot be NULL\n");
      return -1;
   }
   if (n != s->n || stage(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   y = *yp ? *yp : (double *)b200_xcalloc((size_t)3 * n, sizeof(double));
   d_y = s->d_rhs; /* 3n */
   rc = b200_h2d(s->d_x, x, sizeof(double) * (size_t)n);
   for (i = 0; i < 3 && !rc; i++)
   {
      double *yi = d_y + (size_t)i * n;
      if (mask && !mask[i])
         continue;
      if (s->d_dK[i])
         rc = b200k_gemv('N', n, n, 1.0, s->d_dK[i], s->d_x, 0.0, yi);
      else if (i == 0)
         rc = b200k_gemv('N', n, n, 2.0 / s->f, s->d_K, s->d_x, 0.0, yi);
      else
         rc = b200k_copy(n, s->d_x, yi) || b200k_scale(n, s->f * s->f, yi);
      rc = rc || chol_solve_dev_n(s, yi, yi);
      rc = rc || b200_d2h(y + (size_t)i * n, yi, sizeof(double) * (size_t)n);
   }
   if (rc)
   {
      if (!*yp)
         free(y);
      return -1;
   }
   *yp = y;
   return 0;
}
