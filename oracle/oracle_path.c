/* ORACLE (test infrastructure, not product code) -- see oracle_path.h.
 * The NFFT3 fastsum boundary (nfft3_shim/) is pinned on the reference's shipped real-NFFT3 output (TESTS/TEST2, see
 * nfft3_shim/fastsum.h); everything in this
 * file is validated against the reference's own sources (oracle/_ref) by
 * tests/test_oracle_vs_ref.py where /root/reference is available and against the
 * golden fixtures of tests/golden elsewhere. */
#include "oracle_path.h"
#include "fastsum.h"
#include "kernels.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void)
{
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU baseline sets its thread count explicitly.  libgomp is one
 * runtime per process, so this also governs oracle/_ref (the reference's own sources) loaded next to this library. */
void oracle_set_num_threads(int nthreads)
{
#ifdef _OPENMP
   if (nthreads > 0)
      omp_set_num_threads(nthreads);
#else
   (void)nthreads;
#endif
}

/* ------------------------------------------------------------------ */
/* vector ops (SRC/linearalg/vecops.c:3-155)                            */
/* ------------------------------------------------------------------ */
static double vec_dot(const double *x, const double *y, int n)
{
   double s = 0.0;
   long long i;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : s) schedule(static) if (n > 65536)
#endif
   for (i = 0; i < n; i++)
      s += x[i] * y[i];
   return s;
}

static double vec_norm2(const double *x, int n) { return sqrt(vec_dot(x, x, n)); } /* vecops.c:3-7 */

static void vec_scale(double *x, size_t n, double a) /* vecops.c:71-100: scale == 0 -> fill */
{
   size_t i;
   if (a == 0.0)
      memset(x, 0, n * sizeof(double));
   else
      for (i = 0; i < n; i++)
         x[i] *= a;
}

static void vec_axpy(double a, const double *x, size_t n, double *y) /* vecops.c:102-155 */
{
   size_t i;
   if (a == 0.0)
      return;
   for (i = 0; i < n; i++)
      y[i] += a * x[i];
}

/* ------------------------------------------------------------------ */
/* one feature window (str_adj, nfft_interface.h:35-57)                 */
/* ------------------------------------------------------------------ */
typedef struct
{
   int d, n;
   int N, p, m, NN; /* 32, 1, 4, 64: nfft_interface.c:18-27 */
   int kernel;      /* 0 Gauss, 1 Matern-1/2 */
   double sigma[2];
   double mu, scale, kernel_scale;
   double *x; /* scaled, centred, point-major n x d */
   fastsum_plan *orig, *deriv;
} o_window;

typedef struct
{
   int n, nwindows, dwindows, skip_last;
   double f, l, mu;
   o_window *win;
   double *dwork; /* 3n, nfft_interface.c:624 */
} o_additive;

/* first-call scaling of a window: centre per feature, max radius, scale to 0.25,
 * transpose to point-major (nfft_interface.c:150-213).  `cols` = the window's
 * feature columns, contiguous, n values each. */
static void window_scale_points(o_window *w, const double *cols)
{
   int n = w->n, d = w->d, i, j;
   double radius = 0.0;
   double *tmp = (double *)malloc(sizeof(double) * (size_t)n * d);
   memcpy(tmp, cols, sizeof(double) * (size_t)n * d);
   for (i = 0; i < d; i++)
   {
      double c = 0.0;
      for (j = 0; j < n; j++)
         c += tmp[(size_t)i * n + j];
      c /= (double)n;
      for (j = 0; j < n; j++)
         tmp[(size_t)i * n + j] -= c;
   }
   for (j = 0; j < n; j++)
   {
      double r = 0.0;
      for (i = 0; i < d; i++)
         r += tmp[(size_t)i * n + j] * tmp[(size_t)i * n + j];
      r = sqrt(r);
      if (r > radius)
         radius = r;
   }
   if (radius > 0.25 || radius < 0.125)
   {
      w->scale = 0.25 / radius;
      for (j = 0; j < n * d; j++)
         tmp[j] *= w->scale;
   }
   else
      w->scale = 1.0;
   for (i = 0; i < d; i++)
      for (j = 0; j < n; j++)
         w->x[(size_t)j * d + i] = tmp[(size_t)i * n + j];
   free(tmp);
}

static void window_free_plans(o_window *w)
{
   if (w->orig)
   {
      fastsum_plan *pl[2];
      int q;
      pl[0] = w->orig;
      pl[1] = w->deriv;
      for (q = 0; q < 2; q++)
      {
         fastsum_finalize_source_nodes(pl[q]);
         fastsum_finalize_target_nodes(pl[q]);
         fastsum_finalize_kernel(pl[q]);
         nfft_free(pl[q]);
      }
      w->orig = w->deriv = NULL;
   }
}

/* every-call part of nfft_interface.c:215-259 (Gauss) / 351-395 (Matern-1/2) */
static void window_build(o_window *w, int kernel, double f, double l, double mu)
{
   size_t bytes = sizeof(double) * (size_t)w->n * w->d;
   window_free_plans(w);
   w->kernel = kernel;
   w->sigma[0] = (kernel == 0) ? l * w->scale * sqrt(2.0) : l * w->scale;
   w->sigma[1] = w->sigma[0];
   w->mu = mu;
   w->orig = (fastsum_plan *)nfft_malloc(sizeof(fastsum_plan));
   w->deriv = (fastsum_plan *)nfft_malloc(sizeof(fastsum_plan));
   fastsum_init_guru_kernel(w->orig, w->d, kernel == 0 ? gaussian : laplacian_rbf, w->sigma,
                            STORE_PERMUTATION_X_ALPHA, w->N, w->p, 0.0, 0.0);
   fastsum_init_guru_kernel(w->deriv, w->d, kernel == 0 ? xx_gaussian : der_laplacian_rbf, w->sigma + 1,
                            STORE_PERMUTATION_X_ALPHA, w->N, w->p, 0.0, 0.0);
   fastsum_init_guru_source_nodes(w->orig, w->n, w->NN, w->m);
   fastsum_init_guru_target_nodes(w->orig, w->n, w->NN, w->m);
   fastsum_init_guru_source_nodes(w->deriv, w->n, w->NN, w->m);
   fastsum_init_guru_target_nodes(w->deriv, w->n, w->NN, w->m);
   memcpy(w->orig->x, w->x, bytes);
   memcpy(w->orig->y, w->x, bytes);
   memcpy(w->deriv->x, w->x, bytes);
   memcpy(w->deriv->y, w->x, bytes);
   fastsum_precompute(w->orig);
   fastsum_precompute(w->deriv);
   w->kernel_scale = f;
}

/* Nfft4GPNFFTMatSymv, nfft_interface.c:400-497 */
static void window_matvec(o_window *w, int n, double alpha, const double *x, double beta, double *y)
{
   double ff = w->kernel_scale * w->kernel_scale;
   int i;
   for (i = 0; i < n; i++)
      w->orig->alpha[i] = alpha * x[i];
   fastsum_trafo(w->orig);
   if (beta == 0.0)
      for (i = 0; i < n; i++)
         y[i] = ff * (creal(w->orig->f[i]) + w->mu * creal(w->orig->alpha[i]));
   else
   {
      if (beta != 1.0)
         vec_scale(y, (size_t)n, beta);
      for (i = 0; i < n; i++)
         y[i] += ff * (creal(w->orig->f[i]) + w->mu * creal(w->orig->alpha[i]));
   }
}

/* Nfft4GPNFFTGradMatSymv, nfft_interface.c:499-620 */
static void window_grad_matvec(o_window *w, int n, double alpha, const double *x, double beta, double *y)
{
   double ff = w->kernel_scale * w->kernel_scale;
   double f2 = w->kernel_scale * 2.0;
   double *y0 = y, *y1 = y + n, *y2 = y + 2 * (size_t)n;
   double scale;
   int i;
   for (i = 0; i < n; i++)
   {
      w->orig->alpha[i] = alpha * x[i];
      w->deriv->alpha[i] = alpha * x[i];
   }
   fastsum_trafo(w->orig);
   fastsum_trafo(w->deriv);
   scale = (w->kernel == 0) ? 2.0 * w->scale * sqrt(2.0) / w->deriv->kernel_param[0]
                            : w->scale / w->deriv->kernel_param[0];
   scale *= ff;
   if (beta == 0.0)
   {
      for (i = 0; i < n; i++)
      {
         y0[i] = f2 * (creal(w->orig->f[i]) + w->mu * creal(w->orig->alpha[i]));
         y1[i] = scale * creal(w->deriv->f[i]);
         y2[i] = ff * creal(w->orig->alpha[i]);
      }
   }
   else
   {
      if (beta != 1.0)
         vec_scale(y, 3 * (size_t)n, beta);
      for (i = 0; i < n; i++)
      {
         y0[i] += f2 * (creal(w->orig->f[i]) + w->mu * creal(w->orig->alpha[i]));
         y1[i] += scale * creal(w->deriv->f[i]);
         y2[i] += ff * creal(w->orig->alpha[i]);
      }
   }
}

/* Nfft4GPNFFTAdditiveKernelParamCreate, nfft_interface.c:622-674 */
void *oracle_additive_create(const double *data, int n, int ldim, int d, const int *windows, int nwindows, int dwindows)
{
   o_additive *a = (o_additive *)calloc(1, sizeof(o_additive));
   int skip = 1, i, j;
   const int *fw = windows;
   double *cols = (double *)malloc(sizeof(double) * (size_t)n * dwindows);
   (void)d;
   a->n = n;
   a->nwindows = nwindows;
   a->dwindows = dwindows;
   a->skip_last = 0;
   while (skip < dwindows && windows[nwindows * dwindows - skip] < 0)
   {
      skip++;
      a->skip_last++;
   }
   a->win = (o_window *)calloc((size_t)nwindows, sizeof(o_window));
   a->dwork = (double *)malloc(sizeof(double) * 3 * (size_t)n);
   for (i = 0; i < nwindows; i++)
   {
      o_window *w = &a->win[i];
      int actual = 0;
      for (j = 0; j < dwindows; j++)
      {
         if (fw[0] >= 0)
         {
            memcpy(cols + (size_t)actual * n, data + (size_t)fw[0] * ldim, sizeof(double) * (size_t)n);
            fw++;
            actual++;
         }
      }
      w->d = actual;
      w->n = n;
      w->N = 32;
      w->p = 1;
      w->m = 4;
      w->NN = 2;
      while (2 * w->N > w->NN)
         w->NN *= 2;
      w->x = (double *)malloc(sizeof(double) * (size_t)n * (actual > 0 ? actual : 1));
      w->scale = -1.0;
      w->kernel_scale = 1.0;
      if (actual > 0)
         window_scale_points(w, cols);
   }
   free(cols);
   return a;
}

/* Nfft4GPNFFTAdditiveKernel{Gaussian,Matern12}Kernel, nfft_interface.c:676-794 */
int oracle_additive_build(void *h, int kernel, double f, double l, double mu)
{
   o_additive *a = (o_additive *)h;
   int i;
   if (kernel != 0 && kernel != 1)
      return -1;
   a->f = f;
   a->l = l;
   a->mu = mu;
   for (i = 0; i < a->nwindows; i++)
      window_build(&a->win[i], kernel, f, l, mu);
   return 0;
}

/* Nfft4GPAdditiveNFFTMatSymv, nfft_interface.c:796-817 */
int oracle_additive_matvec(void *h, int n, double alpha, double *x, double beta, double *y)
{
   o_additive *a = (o_additive *)h;
   double scale = 1.0 / (double)a->nwindows * alpha;
   int i;
   memset(a->dwork, 0, sizeof(double) * (size_t)n);
   for (i = 0; i < a->nwindows; i++)
      window_matvec(&a->win[i], n, scale, x, 1.0, a->dwork);
   vec_scale(y, (size_t)n, beta);
   vec_axpy(1.0, a->dwork, (size_t)n, y);
   return 0;
}

/* Nfft4GPAdditiveNFFTGradMatSymv, nfft_interface.c:819-840 */
int oracle_additive_grad_matvec(void *h, int n, double alpha, double *x, double beta, double *y)
{
   o_additive *a = (o_additive *)h;
   double scale = 1.0 / (double)a->nwindows * alpha;
   int i;
   memset(a->dwork, 0, sizeof(double) * 3 * (size_t)n);
   for (i = 0; i < a->nwindows; i++)
      window_grad_matvec(&a->win[i], n, scale, x, 1.0, a->dwork);
   vec_scale(y, 3 * (size_t)n, beta);
   vec_axpy(1.0, a->dwork, 3 * (size_t)n, y);
   return 0;
}

void oracle_additive_free(void *h)
{
   o_additive *a = (o_additive *)h;
   int i;
   if (!a)
      return;
   for (i = 0; i < a->nwindows; i++)
   {
      window_free_plans(&a->win[i]);
      free(a->win[i].x);
   }
   free(a->win);
   free(a->dwork);
   free(a);
}

int oracle_additive_nwindows(void *h) { return ((o_additive *)h)->nwindows; }
int oracle_additive_window_dim(void *h, int w) { return ((o_additive *)h)->win[w].d; }
double oracle_additive_window_scale(void *h, int w) { return ((o_additive *)h)->win[w].scale; }
const double *oracle_additive_window_points(void *h, int w) { return ((o_additive *)h)->win[w].x; }
const double *oracle_additive_window_bhat(void *h, int w, int which)
{
   o_window *ww = &((o_additive *)h)->win[w];
   fastsum_plan *p = which ? ww->deriv : ww->orig;
   return p ? (const double *)p->b : NULL;
}

/* ------------------------------------------------------------------ */
/* PCG, SRC/solvers/pcg.c:3-206 (same control flow, same quirks:        */
/* *piter stays 0 unless convergence is confirmed by the true-residual  */
/* recheck; rel_res_v[ii] is overwritten with the ABSOLUTE true         */
/* residual at the recheck).                                            */
/* ------------------------------------------------------------------ */
int oracle_pcg(void *mat, int n, oracle_symmatvec matvec, void *prec, oracle_solve precond,
               double *x, double *rhs, int maxits, int atol, double tol,
               double *prel_res, double *rel_res_v, int *piter)
{
   int iter = 0, ii;
   double rho = 1.0, alpha, beta;
   double normb, normr, normr2, tolb;
   double *r, *z, *p, *q;

   normb = vec_norm2(rhs, n);
   if (normb < DBL_EPSILON)
   {
      memset(x, 0, sizeof(double) * (size_t)n);
      *prel_res = 0.0;
      *piter = 0;
      rel_res_v[0] = 0.0;
      return 0;
   }
   tolb = atol ? tol : tol * normb;
   if (maxits > n)
      maxits = n;
   r = (double *)malloc(sizeof(double) * (size_t)n);
   z = (double *)malloc(sizeof(double) * (size_t)n);
   p = (double *)malloc(sizeof(double) * (size_t)n);
   q = (double *)malloc(sizeof(double) * (size_t)n);
   memcpy(r, rhs, sizeof(double) * (size_t)n);
   matvec(mat, n, -1.0, x, 1.0, r);
   normr = vec_norm2(r, n);
   if (normr < tolb)
   {
      *prel_res = normr / normb;
      *piter = 0;
      rel_res_v[0] = *prel_res;
      free(r); free(z); free(p); free(q);
      return 0;
   }
   normr2 = normr;
   memset(rel_res_v, 0, sizeof(double) * ((size_t)maxits + 1));
   rel_res_v[0] = normr / normb;
   for (ii = 1; ii <= maxits; ii++)
   {
      double rho1, pq;
      if (prec)
         precond(prec, n, z, r);
      else
         memcpy(z, r, sizeof(double) * (size_t)n);
      rho1 = rho;
      rho = vec_dot(z, r, n);
      if (rho == 0.0)
         break;
      if (ii == 1)
         memcpy(p, z, sizeof(double) * (size_t)n);
      else
      {
         beta = rho / rho1;
         if (beta == 0.0)
            break;
         vec_scale(p, (size_t)n, beta);
         vec_axpy(1.0, z, (size_t)n, p);
      }
      matvec(mat, n, 1.0, p, 0.0, q);
      pq = vec_dot(q, p, n);
      if (pq <= 0)
         break;
      alpha = rho / pq;
      vec_axpy(alpha, p, (size_t)n, x);
      vec_axpy(-alpha, q, (size_t)n, r);
      normr = vec_norm2(r, n);
      normr2 = normr;
      rel_res_v[ii] = normr / normb;
      if (normr <= tolb)
      {
         memcpy(r, rhs, sizeof(double) * (size_t)n);
         matvec(mat, n, -1.0, x, 1.0, r);
         normr2 = vec_norm2(r, n);
         rel_res_v[ii] = normr2;
         if (normr2 <= tolb)
         {
            iter = ii;
            break;
         }
      }
   }
   *prel_res = normr2 / normb;
   *piter = iter;
   free(r); free(z); free(p); free(q);
   return 0;
}

/* ------------------------------------------------------------------ */
/* Nystrom apply, SRC/preconds/nys.c:115-173:                           */
/*   y = rhs(perm); z = U'y; y = (y - U z)/eta; z .*= s; y += U z; x(perm) = y */
/* U is n x k column-major; work holds k + n doubles.                   */
/* ------------------------------------------------------------------ */
int oracle_nys_solve(void *vn, int n, double *x, double *rhs)
{
   oracle_nys *N = (oracle_nys *)vn;
   double *z = N->work, *y = N->work + N->k;
   int i, c, k = N->k;
   (void)n;
   n = N->n;
   if (N->perm)
      for (i = 0; i < n; i++)
         y[i] = rhs[N->perm[i]];
   else
   {
      y = x;
      memcpy(y, rhs, sizeof(double) * (size_t)n);
   }
   for (c = 0; c < k; c++)
      z[c] = vec_dot(N->U + (size_t)c * n, y, n);
   for (i = 0; i < n; i++)
      y[i] *= 1.0 / N->eta;
   for (c = 0; c < k; c++)
      vec_axpy(-1.0 / N->eta * z[c], N->U + (size_t)c * n, (size_t)n, y);
   for (c = 0; c < k; c++)
      z[c] *= N->s[c];
   for (c = 0; c < k; c++)
      vec_axpy(z[c], N->U + (size_t)c * n, (size_t)n, y);
   if (N->perm)
      for (i = 0; i < n; i++)
         x[N->perm[i]] = y[i];
   return 0;
}

/* FSAI apply, SRC/preconds/fsai.c:106-123 + CSR products matops.c:139-272:
 *   work = G rhs ; x = G' work */
int oracle_fsai_solve(void *vf, int n, double *x, double *rhs)
{
   oracle_fsai *F = (oracle_fsai *)vf;
   int i, j;
   (void)n;
   n = F->n;
   for (i = 0; i < n; i++)
   {
      double r = 0.0;
      for (j = F->ia[i]; j < F->ia[i + 1]; j++)
         r += F->a[j] * rhs[F->ja[j]];
      F->work[i] = r;
   }
   memset(x, 0, sizeof(double) * (size_t)n);
   for (i = 0; i < n; i++)
      for (j = F->ia[i]; j < F->ia[i + 1]; j++)
         x[F->ja[j]] += F->a[j] * F->work[i];
   return 0;
}

/* k x k lower-triangular solves with the Cholesky factor (column-major), chol.c:111-136 */
static void chol_solve(const double *L, int k, double *v)
{
   int i, j;
   for (i = 0; i < k; i++)
   {
      double s = v[i];
      for (j = 0; j < i; j++)
         s -= L[(size_t)j * k + i] * v[j];
      v[i] = s / L[(size_t)i * k + i];
   }
   for (i = k - 1; i >= 0; i--)
   {
      double s = v[i];
      for (j = i + 1; j < k; j++)
         s -= L[(size_t)i * k + j] * v[j];
      v[i] = s / L[(size_t)i * k + i];
   }
}

/* AFN apply, SRC/preconds/afn.c:109-136 (uncompiled design) = MATLAB afn_solve.m:39-69.
 * work holds 2n doubles + the Schur FSAI's own work (n-k). */
int oracle_afn_solve(void *va, int n, double *x, double *rhs)
{
   oracle_afn *A = (oracle_afn *)va;
   int k = A->k, n2, i, j;
   double *rp, *rp2, *y, *y2;
   (void)n;
   n = A->n;
   n2 = n - k;
   rp = A->work;
   rp2 = rp + k;
   y = A->work + n;
   y2 = y + k;
   for (i = 0; i < n; i++)
      rp[i] = rhs[A->perm[i]];
   memcpy(y, rp, sizeof(double) * (size_t)k);
   chol_solve(A->L, k, y);
   for (j = 0; j < n2; j++) /* rp2 -= K12' y */
      rp2[j] -= vec_dot(A->K12 + (size_t)j * k, y, k);
   oracle_fsai_solve(&A->schur, n2, y2, rp2);
   for (j = 0; j < n2; j++) /* rp -= K12 y2 */
      vec_axpy(-y2[j], A->K12 + (size_t)j * k, (size_t)k, rp);
   memcpy(y, rp, sizeof(double) * (size_t)k);
   chol_solve(A->L, k, y);
   for (i = 0; i < n; i++)
      x[A->perm[i]] = y[i];
   return 0;
}

/* ------------------------------------------------------------------ */
/* dense additive kernel, full-matrix call shape of                      */
/* SRC/linearalg/kernels.c:3099-3494: mean over windows of               */
/* f^2 exp(-r^2/(2 l^2)) (kernels.c:495) or f^2 exp(-r/l) (:2218-2222),  */
/* plus f^2 mu on the diagonal.                                          */
/* ------------------------------------------------------------------ */
int oracle_dense_additive(const double *data, int n, int ldim, int d, const int *windows, int nwindows, int dwindows,
                          int kernel, double f, double l, double mu, double *K)
{
   long long i;
   double ff = f * f;
   (void)d;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
   for (i = 0; i < n; i++)
   {
      int j, w, t;
      for (j = 0; j <= i; j++)
      {
         double acc = 0.0;
         for (w = 0; w < nwindows; w++)
         {
            double r2 = 0.0;
            for (t = 0; t < dwindows; t++)
            {
               int ft = windows[w * dwindows + t];
               if (ft >= 0)
               {
                  double df = data[(size_t)ft * ldim + i] - data[(size_t)ft * ldim + j];
                  r2 += df * df;
               }
            }
            acc += (kernel == 0) ? exp(-r2 / (2.0 * l * l)) : exp(-sqrt(r2) / l);
         }
         acc = ff * acc / (double)nwindows;
         if (i == j)
            acc += ff * mu;
         K[(size_t)j * n + i] = acc;
         K[(size_t)i * n + j] = acc;
      }
   }
   return 0;
}

int oracle_dense_symv(const double *K, int n, double alpha, const double *x, double beta, double *y)
{
   long long i;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
   for (i = 0; i < n; i++)
   {
      double s = 0.0;
      int j;
      for (j = 0; j < n; j++)
         s += K[(size_t)j * n + i] * x[j];
      y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
   }
   return 0;
}
