/* Preconditioner setups (C host side, rows N2 / G2 / A2 of SURVEY.md 8(a)).
 *   Nystrom  SRC/preconds/nys.c:518-660       K11 Cholesky + shift (chol.c:439-466), G = L^{-1}, U = K1 G', SVD through eig(U'U)
 *   FSAI     SRC/preconds/fsai.c:302-673      KNN pattern (kernels.c:121-277) + one k x k Cholesky solve per row
 *   AFN      MATLAB afn_setup.m:30-109 (SRC/preconds/afn.c is not compiled by the reference): rank estimate, farthest
 *            point sampling, Cholesky of K11, K12, L^{-1} K12, FSAI of the Schur complement; Nystrom when the rank is low
 * Everything O(n) and above runs on the GPU (setup_kernels.cu, precond_kernels.cu); the O(k^3) factorizations of the
 * landmark block run here (dense_small.c).  A dense kernel of this library (Nfft4GPKernel{Additive,Gaussian,Matern12}Kernel)
 * is evaluated on the device; a foreign func_kernel is honoured by the Nystrom setup only (two host calls, as the
 * reference makes them).  require_grad is accepted and ignored (the reference's own prediction passes 1, nfft_interface.c:936):
 * the derivative factors only serve Dvp / Trace / Logdet of the GP loss gradient (SURVEY.md 8(f) "next"), which are not exported. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

static double wtime(void)
{
   struct timeval t;
   gettimeofday(&t, NULL);
   return (double)t.tv_sec + 1e-6 * (double)t.tv_usec;
}

static int *identity_perm(int n)
{
   int *p = (int *)b200_xmalloc(sizeof(int) * (size_t)(n ? n : 1)), i;
   for (i = 0; i < n; i++)
      p[i] = i;
   return p;
}

/* feature-major n x d copy of the caller's data on the device, leading dimension n */
static double *upload_features(const double *data, int n, int ldim, int d)
{
   double *d_X = (double *)b200_malloc(sizeof(double) * (size_t)n * d);
   int t;
   if (!d_X)
      return NULL;
   for (t = 0; t < d; t++)
      if (b200_h2d(d_X + (size_t)t * n, data + (size_t)t * ldim, sizeof(double) * (size_t)n))
      {
         b200_free(d_X);
         return NULL;
      }
   if (b200_sync())
   {
      b200_free(d_X);
      return NULL;
   }
   return d_X;
}

static int *upload_ints(const int *h, size_t count)
{
   int *d = (int *)b200_malloc(sizeof(int) * (count ? count : 1));
   if (d && count && (b200_h2d(d, h, sizeof(int) * count) || b200_sync()))
   {
      b200_free(d);
      return NULL;
   }
   return d;
}

/* Shifted Cholesky of the k x k landmark block (device, lower triangle, ld k), chol.c:439-466: the stabilising shift
 * sqrt(k) ulp(||K11||_F) is derived on the host from a copy of the block, the blocked factorisation and the explicit
 * inverse of the factor (dtrtri, matops.c:40-48) run on the device.  d_K11 <- L, d_Linv <- L^{-1}. */
static int landmark_cholesky(int k, double *d_K11, double *d_Linv)
{
   double *h = (double *)b200_xmalloc(sizeof(double) * (size_t)k * k), fro, nu;
   int info = 0;
   if (b200_d2h(h, d_K11, sizeof(double) * (size_t)k * k))
   {
      free(h);
      return -1;
   }
   fro = b200_sym_fro_lower(k, h, k);
   nu = b200_chol_shift(k, fro);
   free(h);
   if (b200k_add_diag(k, d_K11, k, nu) || b200k_potrf_inverse(k, d_K11, d_Linv, &info))
      return -1;
   if (info)
      printf("Error: Cholesky of the %d x %d landmark block failed at pivot %d.\n", k, k, info);
   return info;
}

/* ================================================================== Nystrom (nys.c:518-660) */
static int nys_setup_perm(b200_nys *s, double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params,
                          const int *perm_in, int k, int require_grad)
{
   nfft4gp_kernel *pk = (nfft4gp_kernel *)fkernel_params;
   b200_dense_view v, v11;
   int own, rc = -1, i, j, *perm = NULL, *d_perm = NULL, quirk = 0;
   double *K11 = NULL, *AA = NULL, *w = NULL, *sv = NULL, *Vs = NULL;
   double *d_K1 = NULL, *d_U = NULL, *d_small = NULL, *d_small2 = NULL;
   double noise_level, f2, eta;
   double *grad_keep[5] = {NULL, NULL, NULL, NULL, NULL};
   size_t kk;
   if (b200_init(-1))
      return -1;
   if (k > n)
      k = n;
   if (k <= 0)
   {
      printf("Error: Nystrom setup needs a positive rank.\n");
      return -1;
   }
   kk = (size_t)k * (size_t)k;
   own = b200_dense_view_acquire(fkernel, fkernel_params, data, n, ldim, d, &v);
   if (own < 0)
      return -1;
   noise_level = pk->_noise_level;
   f2 = pk->_params[0] * pk->_params[0];
   eta = noise_level * f2; /* nys.c:641 */
   perm = (int *)b200_xmalloc(sizeof(int) * (size_t)n);
   if (perm_in)
      memcpy(perm, perm_in, sizeof(int) * (size_t)n);
   else
      for (i = 0; i < n; i++)
         perm[i] = i;
   K11 = (double *)b200_xcalloc(kk, sizeof(double));
   d_K1 = (double *)b200_malloc(sizeof(double) * (size_t)n * k);
   d_U = (double *)b200_malloc(sizeof(double) * (size_t)n * k);
   d_small = (double *)b200_malloc(sizeof(double) * kk);
   d_small2 = (double *)b200_malloc(sizeof(double) * kk);
   if (!d_K1 || !d_U || !d_small || !d_small2)
      goto done;
   if (own == 0)
   {
      /* K1 = K(perm, perm[0:k]) (n x k, no noise: nys.c:597) and K11 = its leading block (noise 0: nys.c:562) */
      d_perm = upload_ints(perm, (size_t)n);
      /* Reference quirk, reproduced on request (NFFT4GP_B200_REFERENCE_QUIRKS=1) so that the reference's own drivers print what they
       * print on the reference's code: nys.c:549-566 forms K11 by calling the kernel on the k sub-sampled points with n = k, and the
       * reference's dense ADDITIVE kernel ignores its data argument and strides its own feature copy with that n (kernels.c:3160),
       * i.e. it reads the FIRST P dw k values of the copy as a [P dw] x k feature table.  The same view here: stride k, rows 0..k-1. */
      if (fkernel == Nfft4GPKernelAdditiveKernel && getenv("NFFT4GP_B200_REFERENCE_QUIRKS") && atoi(getenv("NFFT4GP_B200_REFERENCE_QUIRKS")) > 0)
      {
         v11 = v;
         v11.n = k;
         v11.owns = 0;
         quirk = 1;
      }
      if (!d_perm || b200_dense_view_block(&v, 0.0, d_perm, n, d_perm, k, 0, d_K1, n) ||
          b200_dense_view_block(quirk ? &v11 : &v, 0.0, quirk ? NULL : d_perm, k, quirk ? NULL : d_perm, k, 1, d_small2, k))
         goto done;
   }
   else
   {
      /* foreign kernel: the reference's own two host calls, then upload */
      double *K1h = NULL;
      pk->_noise_level = 0.0;
      i = fkernel(fkernel_params, data, n, ldim, d, perm, n, perm, k, &K1h, NULL);
      pk->_noise_level = noise_level;
      if (i || !K1h)
      {
         printf("Error: the kernel function failed in the Nystrom setup.\n");
         free(K1h);
         goto done;
      }
      for (j = 0; j < k; j++)
         for (i = j; i < k; i++)
            K11[(size_t)j * k + i] = K1h[(size_t)j * n + i];
      i = b200_h2d(d_K1, K1h, sizeof(double) * (size_t)n * k) || b200_h2d(d_small2, K11, sizeof(double) * kk) || b200_sync();
      free(K1h);
      if (i)
         goto done;
   }
   /* K11 + shift = L L' (chol.c:439-466), G = L^{-1} (nys.c:576), U = K1 G' (dtrmm R,L,T: nys.c:602) */
   if (landmark_cholesky(k, d_small2, d_small) || b200k_dgemm('N', 'T', n, k, k, 1.0, d_K1, n, d_small, k, 0.0, d_U, n))
      goto done;
   if (require_grad)
   {
      /* nys.c:555-610 with require_grad: keep K1, dK1/dl, L^{-1} and G_g = L^{-1} dK11_g L^{-T} (g = f, l) for Dvp / Trace */
      double *g_K1, *g_dK1l, *g_Linv, *g_G0, *g_G1, *t1, *t2;
      if (own != 0)
      {
         printf("Error: the Nystrom setup with require_grad needs one of this library's dense kernels.\n");
         goto done;
      }
      g_K1 = (double *)b200_malloc(sizeof(double) * (size_t)n * k);
      g_dK1l = (double *)b200_malloc(sizeof(double) * (size_t)n * k);
      g_Linv = (double *)b200_malloc(sizeof(double) * kk);
      g_G0 = (double *)b200_malloc(sizeof(double) * kk);
      g_G1 = (double *)b200_malloc(sizeof(double) * kk);
      t1 = (double *)b200_malloc(sizeof(double) * kk);
      t2 = (double *)b200_malloc(sizeof(double) * kk);
      i = !g_K1 || !g_dK1l || !g_Linv || !g_G0 || !g_G1 || !t1 || !t2 ||
          b200_d2d(g_K1, d_K1, sizeof(double) * (size_t)n * k) || b200_d2d(g_Linv, d_small, sizeof(double) * kk) ||
          b200_dense_view_block_dl(&v, d_perm, n, d_perm, k, 0, g_dK1l, n) ||
          /* dK11/df = (2/f) K11 (noise 0, no shift): G_0 = (2/f) L^{-1} K11 L^{-T} */
          b200_dense_view_block(quirk ? &v11 : &v, 0.0, quirk ? NULL : d_perm, k, quirk ? NULL : d_perm, k, 0, t1, k) ||
          b200k_dgemm('N', 'N', k, k, k, 2.0 / pk->_params[0], d_small, k, t1, k, 0.0, t2, k) ||
          b200k_dgemm('N', 'T', k, k, k, 1.0, t2, k, d_small, k, 0.0, g_G0, k) ||
          b200_dense_view_block_dl(quirk ? &v11 : &v, quirk ? NULL : d_perm, k, quirk ? NULL : d_perm, k, 0, t1, k) ||
          b200k_dgemm('N', 'N', k, k, k, 1.0, d_small, k, t1, k, 0.0, t2, k) ||
          b200k_dgemm('N', 'T', k, k, k, 1.0, t2, k, d_small, k, 0.0, g_G1, k) || b200_sync();
      b200_free(t1);
      b200_free(t2);
      if (i)
      {
         b200_free(g_K1);
         b200_free(g_dK1l);
         b200_free(g_Linv);
         b200_free(g_G0);
         b200_free(g_G1);
         goto done;
      }
      grad_keep[0] = g_K1;
      grad_keep[1] = g_dK1l;
      grad_keep[2] = g_Linv;
      grad_keep[3] = g_G0;
      grad_keep[4] = g_G1;
   }
   /* SVD of U through the eigen-decomposition of U'U (matops.c:65-137): U = W S V'  =>  W = U V S^{-1} */
   AA = (double *)b200_xmalloc(sizeof(double) * kk);
   w = (double *)b200_xmalloc(sizeof(double) * (size_t)k);
   sv = (double *)b200_xmalloc(sizeof(double) * (size_t)k);
   Vs = (double *)b200_xmalloc(sizeof(double) * kk);
   if (b200k_dgemm('T', 'N', k, k, n, 1.0, d_U, n, d_U, n, 0.0, d_small2, k) || b200_d2h(AA, d_small2, sizeof(double) * kk))
      goto done;
   for (j = 0; j < k; j++) /* exact symmetry for the eigensolver */
      for (i = j + 1; i < k; i++)
      {
         double m = 0.5 * (AA[(size_t)j * k + i] + AA[(size_t)i * k + j]);
         AA[(size_t)j * k + i] = AA[(size_t)i * k + j] = m;
      }
   if (b200_syev(k, AA, w))
   {
      printf("Error: the eigensolver did not converge in the Nystrom setup.\n");
      goto done;
   }
   /* descending singular values; column c of W = U * V(:, k-1-c) / sigma_c.  The reference takes sqrt of a possibly
    * negative eigenvalue (NaN) and mis-steps for sigma < 1e-12 (matops.c:114-129); here: clamp at 0, scale by 1e12. */
   for (j = 0; j < k; j++)
   {
      double ev = w[k - 1 - j], sig = sqrt(ev > 0.0 ? ev : 0.0), sc = sig < 1e-12 ? 1e12 : 1.0 / sig;
      double val;
      for (i = 0; i < k; i++)
         Vs[(size_t)j * k + i] = AA[(size_t)(k - 1 - j) * k + i] * sc;
      val = 1.0 / (sig * sig + eta); /* nys.c:643-647 */
      sv[j] = val > 0.0 ? val : 0.0;
   }
   if (b200_h2d(d_small, Vs, sizeof(double) * kk) || b200k_dgemm('N', 'N', n, k, k, 1.0, d_U, n, d_small, k, 0.0, d_K1, n) || b200_sync())
      goto done;
   /* the handle takes ownership of d_K1 (= W) */
   rc = b200_nys_set_device_factors(s, n, k, perm, d_K1, sv, eta);
   d_K1 = NULL;
   if (!rc && require_grad)
   {
      s->d_gK1 = grad_keep[0];
      s->d_gdK1l = grad_keep[1];
      s->d_gLinv = grad_keep[2];
      s->d_gG[0] = grad_keep[3];
      s->d_gG[1] = grad_keep[4];
      s->d_gvec = (double *)b200_malloc(sizeof(double) * (4 * (size_t)n + 8 * (size_t)k));
      s->f = pk->_params[0];
      s->has_grad = 1;
      memset(grad_keep, 0, sizeof(grad_keep));
      rc = s->d_gvec ? 0 : -1;
   }
done:
   for (i = 0; i < 5; i++)
      b200_free(grad_keep[i]);
   b200_free(d_K1);
   b200_free(d_U);
   b200_free(d_small);
   b200_free(d_small2);
   b200_free(d_perm);
   free(perm);
   free(K11);
   free(AA);
   free(w);
   free(sv);
   free(Vs);
   if (own == 0)
      b200_dense_view_release(&v);
   return rc;
}

int Nfft4GPPrecondNysSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad,
                                     void *vnys_mat)
{
   b200_nys *s = (b200_nys *)vnys_mat;
   double t0 = wtime();
   int rc;
   if (!s || s->magic != B200_MAGIC_NYS || !fkernel || !fkernel_params)
      return -1;
   rc = nys_setup_perm(s, data, n, ldim, d, fkernel, fkernel_params, s->perm, s->k_setup, require_grad);
   s->tset = wtime() - t0;
   return rc;
}

/* ================================================================== FSAI (fsai.c:302-673) */
/* fsai.c:302-673.  pattern_i / pattern_j: the caller's CSR pattern (host, lower triangular, diagonal last in every row; NULL = the
 * KNN pattern of Nfft4GPDistanceEuclidKnn built on the GPU).  require_grad also builds the derivative factors (fsai.c:495-560). */
static int fsai_setup_impl(const int *pattern_i, const int *pattern_j, double *data, int n, int ldim, int d, func_kernel fkernel,
                           void *fkernel_params, int require_grad, b200_fsai *s)
{
   b200_dense_view v;
   double t0 = wtime(), *d_X = NULL, *d_a = NULL, *d_da = NULL;
   int own, rc = -1, lfil, kmax, bad = 0, *d_ia = NULL, *d_ja = NULL, i;
   long long nnz;
   if (!s || s->magic != B200_MAGIC_FSAI || !fkernel || !fkernel_params || b200_init(-1))
      return -1;
   own = b200_dense_view_acquire(fkernel, fkernel_params, data, n, ldim, d, &v);
   if (own)
   {
      if (own > 0)
         printf("Error: the FSAI setup needs one of this library's dense kernels (Nfft4GPKernel{Additive,Gaussian,Matern12}Kernel).\n");
      return -1;
   }
   if (pattern_i)
   {
      nnz = pattern_i[n];
      kmax = 1;
      for (i = 0; i < n; i++)
         if (pattern_i[i + 1] - pattern_i[i] > kmax)
            kmax = pattern_i[i + 1] - pattern_i[i];
   }
   else
   {
      lfil = s->lfil < 1 ? 1 : s->lfil;
      kmax = lfil < n ? lfil : n;
      nnz = b200k_knn_nnz(n, lfil);
   }
   if (nnz > 2147483647LL)
   {
      printf("Error: FSAI pattern with %lld entries exceeds the 32-bit index range of the reference's CSR format.\n", nnz);
      goto done;
   }
   d_ia = (int *)b200_malloc(sizeof(int) * ((size_t)n + 1));
   d_ja = (int *)b200_malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
   d_a = (double *)b200_malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
   if (require_grad)
      d_da = (double *)b200_malloc(sizeof(double) * 3 * (size_t)(nnz ? nnz : 1));
   if (!d_ia || !d_ja || !d_a || (require_grad && !d_da))
      goto done;
   if (pattern_i)
   {
      if (b200_h2d(d_ia, pattern_i, sizeof(int) * ((size_t)n + 1)) || b200_h2d(d_ja, pattern_j, sizeof(int) * (size_t)nnz) || b200_sync())
         goto done;
   }
   else
   {
      d_X = upload_features(data, n, ldim, d);
      if (!d_X || b200k_knn_pattern(d_X, n, n, d, s->lfil < 1 ? 1 : s->lfil, d_ia, d_ja))
         goto done;
   }
   if (require_grad ? b200k_fsai_rows_grad(v.d_Xw, v.d_dims, v.n, v.P, v.dw, v.kernel, v.f * v.f, v.l, v.mu * v.f * v.f, n, d_ia, d_ja, kmax, d_a,
                                           d_da, (size_t)nnz, &bad)
                    : b200k_fsai_rows(v.d_Xw, v.d_dims, v.n, v.P, v.dw, v.kernel, v.f * v.f, v.l, v.mu * v.f * v.f, n, d_ia, d_ja, NULL, NULL, 0,
                                      kmax, d_a, &bad))
      goto done;
   if (bad)
   {
      printf("Error: FSAI breakdown: %d rows have a kernel block that is not positive definite.\n", bad);
      goto done;
   }
   rc = b200_fsai_set_device_factors(s, n, (int)nnz, d_ia, d_ja, d_a);
   d_ia = d_ja = NULL;
   d_a = NULL;
   if (!rc && require_grad)
   {
      rc = b200_fsai_set_device_grad(s, d_da);
      d_da = NULL;
   }
done:
   b200_free(d_X);
   b200_free(d_ia);
   b200_free(d_ja);
   b200_free(d_a);
   b200_free(d_da);
   b200_dense_view_release(&v);
   s->tset = wtime() - t0;
   return rc;
}

int Nfft4GPPrecondFsaiSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad,
                                      void *vfsai_mat)
{
   return fsai_setup_impl(NULL, NULL, data, n, ldim, d, fkernel, fkernel_params, require_grad, (b200_fsai *)vfsai_mat);
}

/* fsai.c:314-673: the caller's pattern; as in the reference the preconditioner takes ownership of A_i / A_j (malloc'd by the caller) */
int Nfft4GPPrecondFsaiSetupWithKernelPattern(int *A_i, int *A_j, double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params,
                                             int require_grad, void *vfsai_mat)
{
   int rc;
   if (!A_i || !A_j)
      return -1;
   rc = fsai_setup_impl(A_i, A_j, data, n, ldim, d, fkernel, fkernel_params, require_grad, (b200_fsai *)vfsai_mat);
   free(A_i);
   free(A_j);
   return rc;
}

/* ================================================================== AFN (afn_setup.m:30-109) */

/* counter-based generator for the sub-samples of the rank estimation (the reference uses rand(); MATLAB randsample) */
static uint64_t splitmix64(uint64_t *st)
{
   uint64_t z = (*st += 0x9e3779b97f4a7c15ULL);
   z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
   z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
   return z ^ (z >> 31);
}

/* first m entries of a Fisher-Yates shuffle of pool[0:n) */
static void sample_without_replacement(int *pool, int n, int m, uint64_t *st)
{
   int i;
   for (i = 0; i < m; i++)
   {
      int j = i + (int)(splitmix64(st) % (uint64_t)(n - i)), t = pool[i];
      pool[i] = pool[j];
      pool[j] = t;
   }
}

/* Upload a scaled sub-sample as its own small problem: window features (for the kernel) and all features (for FPS). */
typedef struct
{
   double *d_Xw, *d_X;
   int m;
} rank_sample;

static int rank_sample_upload(const b200_dense_view *v, const double *h_Xw, size_t h_ldw, const double *data, int ldim, int d,
                              const int *rows, int m, double scale, rank_sample *rs)
{
   int nf = v->P * v->dw, t, i;
   double *tmp = (double *)b200_xmalloc(sizeof(double) * (size_t)m * (size_t)(nf > d ? nf : d));
   rs->m = m;
   rs->d_Xw = (double *)b200_malloc(sizeof(double) * (size_t)m * nf);
   rs->d_X = (double *)b200_malloc(sizeof(double) * (size_t)m * d);
   if (!rs->d_Xw || !rs->d_X)
   {
      free(tmp);
      return -1;
   }
   for (t = 0; t < nf; t++)
      for (i = 0; i < m; i++)
         tmp[(size_t)t * m + i] = h_Xw[(size_t)t * h_ldw + rows[i]] * scale;
   if (b200_h2d(rs->d_Xw, tmp, sizeof(double) * (size_t)m * nf) || b200_sync())
   {
      free(tmp);
      return -1;
   }
   for (t = 0; t < d; t++)
      for (i = 0; i < m; i++)
         tmp[(size_t)t * m + i] = data[(size_t)t * ldim + rows[i]] * scale;
   i = b200_h2d(rs->d_X, tmp, sizeof(double) * (size_t)m * d) || b200_sync();
   free(tmp);
   return i ? -1 : 0;
}

static void rank_sample_free(rank_sample *rs)
{
   b200_free(rs->d_Xw);
   b200_free(rs->d_X);
   rs->d_Xw = rs->d_X = NULL;
}

/* host features behind a dense view: window-major for the additive kernel, the caller's data otherwise */
static const double *view_host_features(func_kernel fkernel, void *fkernel_params, const double *data, int n, int ldim, size_t *ld)
{
   if (fkernel == Nfft4GPKernelAdditiveKernel)
   {
      *ld = (size_t)n;
      return ((nfft4gp_kernel *)fkernel_params)->_buffer;
   }
   *ld = (size_t)ldim;
   return data;
}

static int centroid_to_device(const double *data, int n, int ldim, int d, const int *rows, double scale, double **d_c)
{
   double c[64];
   int t, i;
   for (t = 0; t < d; t++)
   {
      double s = 0.0;
      for (i = 0; i < n; i++)
         s += data[(size_t)t * ldim + (rows ? rows[i] : i)] * scale / n; /* ordering.c:470-474: x_mean += x / n */
      c[t] = s;
   }
   *d_c = (double *)b200_malloc(sizeof(double) * 64);
   if (!*d_c || b200_h2d(*d_c, c, sizeof(double) * (size_t)d) || b200_sync())
      return -1;
   return 0;
}

/* smallest r such that || K - K(:,1:r) K(1:r,1:r)^{-1} K(1:r,:) ||_2 / nrm < tol, K symmetric positive definite (full
 * storage, destroyed).  MATLAB bisects over r (afn_setup.m:196-207); the error is monotone in r, so one right-looking
 * Cholesky sweep that watches the trailing Schur complement finds the same r: max diag <= ||S||_2 <= trace bracket the
 * norm, the power iteration runs only when the bracket straddles the threshold. */
static int nystrom_error_rank(int m, double *K, double nrm, double tol, double *work /* 2m */)
{
   int r, i, j;
   for (r = 1; r < m; r++)
   {
      /* eliminate column r-1 */
      int c = r - 1;
      double piv = K[(size_t)c * m + c], tr = 0.0, mx = 0.0;
      if (!(piv > 0.0))
         return r; /* numerically rank deficient: the Nystrom error is at rounding level */
      for (j = r; j < m; j++)
      {
         double f = K[(size_t)c * m + j] / piv;
         double *cj = K + (size_t)j * m;
         const double *cc = K + (size_t)c * m;
         for (i = j; i < m; i++)
            cj[i] -= cc[i] * f;
      }
      for (j = r; j < m; j++)
      {
         double dj = K[(size_t)j * m + j];
         tr += dj;
         if (dj > mx)
            mx = dj;
      }
      if (mx / nrm >= tol)
         continue;
      if (tr / nrm < tol)
         return r;
      {
         /* symmetrise the trailing block (only its lower triangle is maintained) and take its 2-norm */
         int ms = m - r;
         double *S = (double *)b200_xmalloc(sizeof(double) * (size_t)ms * ms), nrm2;
         for (j = 0; j < ms; j++)
            for (i = j; i < ms; i++)
               S[(size_t)j * ms + i] = S[(size_t)i * ms + j] = K[(size_t)(j + r) * m + (i + r)];
         nrm2 = b200_spd_norm2(ms, S, ms, work);
         free(S);
         if (nrm2 / nrm < tol)
            return r;
      }
   }
   return m;
}

/* One sample of Nfft4GPRankestNysScaledEstimateRank (rankest.c:248-352) on the sample rows `rows` (the reference draws them
 * with rand(), rankest.c:278): the sample is scaled to the density of the full set (:284), ordered by farthest point sampling
 * (:287-294, here on the GPU), K without noise (:297-298, kernel block on the GPU) gets the shift nu = sqrt(n) ulp(||K||_F)
 * (:301-318), and k = 0, ngap, 2 ngap, ... (ngap = m / 50) is scanned for the first scaled Nystrom error
 * || K1' K11^{-1} K1 - (K + nu I) ||_F / || K + nu I ||_F < 0.1 (:330-345, error :196-246 with the unshifted K11), stopping early
 * once (i - 1) ngap n / m > 2 max_rank.  The reference forms the error with potrf + trtrs + gemm per candidate; here ONE
 * right-looking elimination sweep visits the candidates in order: after k columns the trailing block is the Schur complement
 * S_k = K22 - K21 K11^{-1} K12, and the error matrix is -nu I (leading block) and -(S_k + nu I) (trailing block).
 * Returns the rank scaled back by n / m (:347), or -1 on error. */
#define RANKEST_NPOINTS 50 /* rankest.c:244-246 */
static int nys_scaled_sample(const b200_dense_view *v, const double *h_Xw, size_t h_ld, const double *data, int n, int ldim, int d,
                             const int *rows, int m, int max_rank)
{
   rank_sample rs = {NULL, NULL, 0};
   b200_dense_view sv = *v;
   int *d_permf = (int *)b200_malloc(sizeof(int) * (size_t)m), i, j, c, done = 0, rank = m, ngap, rc = -1;
   double *d_K = (double *)b200_malloc(sizeof(double) * (size_t)m * m), *d_c = NULL;
   double *K = (double *)b200_xmalloc(sizeof(double) * (size_t)m * m), fro = 0.0, nu, scale = pow((double)m / n, 1.0 / d);
   if (!d_permf || !d_K || rank_sample_upload(v, h_Xw, h_ld, data, ldim, d, rows, m, scale, &rs) ||
       centroid_to_device(data, m, ldim, d, rows, scale, &d_c) || b200k_fps(rs.d_X, m, m, d, d_c, m, d_permf))
      goto out;
   sv.d_Xw = rs.d_Xw;
   sv.n = m;
   if (b200_dense_view_block(&sv, 0.0, d_permf, m, d_permf, m, 0, d_K, m) || b200_d2h(K, d_K, sizeof(double) * (size_t)m * m))
      goto out;
   for (j = 0; j < m; j++)
      for (i = 0; i < m; i++)
         fro += K[(size_t)j * m + i] * K[(size_t)j * m + i];
   fro = sqrt(fro);
   nu = sqrt((double)n) * (nextafter(fro, fro + 1.0) - fro);
   {
      double t = 0.0; /* || K + nu I ||_F */
      for (j = 0; j < m; j++)
         for (i = 0; i < m; i++)
         {
            double e = K[(size_t)j * m + i] + (i == j ? nu : 0.0);
            t += e * e;
         }
      fro = sqrt(t);
   }
   ngap = m / RANKEST_NPOINTS > 1 ? m / RANKEST_NPOINTS : 1;
   for (i = 0; i < RANKEST_NPOINTS; i++)
   {
      int k = i * ngap;
      double err2;
      if (k >= m || (int)floor((double)((i - 1) * ngap) * (double)n / m) > 2 * max_rank)
      {
         rank = (i - 1) * ngap;
         break;
      }
      /* eliminate the columns [done, k) of the lower triangle */
      for (c = done; c < k; c++)
      {
         double piv = K[(size_t)c * m + c];
         int a, b;
         if (!(piv > 0.0))
            break; /* numerically rank deficient: the error sits at the rounding level from here on */
         for (b = c + 1; b < m; b++)
         {
            double fct = K[(size_t)c * m + b] / piv;
            double *cb = K + (size_t)b * m;
            const double *cc = K + (size_t)c * m;
            for (a = b; a < m; a++)
               cb[a] -= cc[a] * fct;
         }
      }
      if (c < k)
      {
         rank = k;
         break;
      }
      done = k;
      if (k == 0)
         err2 = 1.0; /* rankest.c:202-205 */
      else
      {
         err2 = (double)k * nu * nu;
         for (j = k; j < m; j++)
         {
            double dj = K[(size_t)j * m + j] + nu;
            err2 += dj * dj;
            for (c = j + 1; c < m; c++)
               err2 += 2.0 * K[(size_t)j * m + c] * K[(size_t)j * m + c];
         }
         err2 = sqrt(err2) / fro;
      }
      if (err2 < 0.1)
      {
         rank = k;
         break;
      }
   }
   rc = (int)floor((double)rank * (double)n / m);
out:
   rank_sample_free(&rs);
   b200_free(d_c);
   b200_free(d_permf);
   b200_free(d_K);
   free(K);
   return rc;
}

/* Nfft4GPRankestNysScaled (rankest.c:354-391) for ONE given sample: exported so that the estimate can be compared with the
 * reference's compiled function on the very rows its rand() stream drew (tests/test_gpu_setup.py). */
int Nfft4GPB200RankestNysScaledSample(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, const int *rows, int m,
                                      int max_rank)
{
   b200_dense_view v;
   size_t h_ld;
   const double *h_Xw;
   int rc;
   if (!fkernel || !fkernel_params || !rows || m < 1 || m > n || d > 64 || b200_init(-1))
      return -1;
   if (b200_dense_view_acquire(fkernel, fkernel_params, data, n, ldim, d, &v))
   {
      printf("Error: the rank estimation needs one of this library's dense kernels.\n");
      return -1;
   }
   h_Xw = view_host_features(fkernel, fkernel_params, data, n, ldim, &h_ld);
   rc = nys_scaled_sample(&v, h_Xw, h_ld, data, n, ldim, d, rows, m, max_rank);
   b200_dense_view_release(&v);
   return rc;
}

/* afn_setup.m:111-293.  Returns the estimated rank (>= 0) or -1 on error. */
static int afn_rank_estimate(b200_afn *s, const b200_dense_view *v, func_kernel fkernel, void *fkernel_params, const double *data, int n,
                             int ldim, int d, int max_k)
{
   size_t h_ld;
   const double *h_Xw = view_host_features(fkernel, fkernel_params, data, n, ldim, &h_ld);
   int m = s->est_m < n ? s->est_m : n, numests = s->est_num, e, i, j, r = 0, k_scaled, rc = -1;
   int *pool = identity_perm(n), *permf = (int *)b200_xmalloc(sizeof(int) * (size_t)(m > max_k ? m : max_k));
   int *d_permf = (int *)b200_malloc(sizeof(int) * (size_t)m);
   double *Km = (double *)b200_xmalloc(sizeof(double) * (size_t)m * m), *work = (double *)b200_xmalloc(sizeof(double) * 2 * (size_t)m);
   double *d_K = (double *)b200_malloc(sizeof(double) * (size_t)m * m), *d_c = NULL;
   uint64_t st = s->est_seed;
   b200_dense_view sv = *v;
   rank_sample rs = {NULL, NULL, 0};
   if (!d_permf || !d_K)
      goto done;
   if (s->est_method == 1)
   {
      /* first stage as the reference's compiled Nfft4GPRankestNysScaled (rankest.c:354-391): mean over the samples, and "0" when
       * the estimate sits at the lower bound of the scan; the sub-samples come from this library's counter-based generator */
      long long total = 0;
      int ngap = m / RANKEST_NPOINTS > 1 ? m / RANKEST_NPOINTS : 1;
      for (e = 0; e < numests; e++)
      {
         for (i = 0; i < n; i++)
            pool[i] = i;
         sample_without_replacement(pool, n, m, &st);
         i = nys_scaled_sample(v, h_Xw, h_ld, data, n, ldim, d, pool, m, max_k);
         if (i < 0)
            goto done;
         total += i;
      }
      k_scaled = (int)floor((double)total / numests);
      if (k_scaled <= (int)floor((double)ngap * n / m))
         k_scaled = 0;
      printf("Estimate rank(scaled): %d/%d\n", k_scaled, n); /* rankest.c:386 */
      if (k_scaled >= max_k) /* afn.c:193: "not low-rank, skip nonscaled rank estimation" */
      {
         rc = k_scaled;
         goto done;
      }
      goto second_stage;
   }
   /* ---- scaled Nystrom-error estimate (afn_setup.m:163-228) */
   for (e = 0; e < numests; e++)
   {
      double nu, fro, n2;
      for (i = 0; i < n; i++)
         pool[i] = i;
      sample_without_replacement(pool, n, m, &st);
      if (rank_sample_upload(v, h_Xw, h_ld, data, ldim, d, pool, m, pow((double)m / n, 1.0 / d), &rs) ||
          centroid_to_device(data, m, ldim, d, pool, pow((double)m / n, 1.0 / d), &d_c) || b200k_fps(rs.d_X, m, m, d, d_c, m, d_permf))
         goto done;
      /* Km = K(permfps, permfps) without noise, f = 1 */
      sv.d_Xw = rs.d_Xw;
      sv.n = m;
      sv.f = 1.0;
      if (b200_dense_view_block(&sv, 0.0, d_permf, m, d_permf, m, 0, d_K, m) || b200_d2h(Km, d_K, sizeof(double) * (size_t)m * m))
         goto done;
      n2 = b200_spd_norm2(m, Km, m, work);
      nu = sqrt((double)m) * (nextafter(n2, n2 + 1.0) - n2); /* sqrt(m) * eps(norm(Km, 2)) */
      fro = 0.0;
      for (j = 0; j < m; j++)
      {
         Km[(size_t)j * m + j] += nu;
         for (i = 0; i < m; i++)
            fro += Km[(size_t)j * m + i] * Km[(size_t)j * m + i];
      }
      fro = sqrt(fro);
      i = nystrom_error_rank(m, Km, fro, 0.1, work);
      if (i > r)
         r = i;
      rank_sample_free(&rs);
      b200_free(d_c);
      d_c = NULL;
   }
   k_scaled = (int)ceil((double)r * n / m);
   if (k_scaled > max_k)
   {
      rc = k_scaled;
      goto done;
   }
second_stage:
   /* ---- numerical rank of sub-sampled kernel matrices (afn_setup.m:230-292) */
   {
      int mk = max_k < n ? max_k : n;
      double thr = 1.1 * v->mu / (v->f * v->f);
      if (mk < m)
         mk = m;
      r = 0;
      for (e = 0; e < numests; e++)
      {
         double scale = 1.0;
         int cnt = 0;
         for (i = 0; i < n; i++)
            pool[i] = i;
         sample_without_replacement(pool, n, mk, &st);
         if (mk > m)
         {
            /* sub-sample again: m of the mk points, scaled */
            for (i = 0; i < mk; i++)
               permf[i] = pool[i];
            sample_without_replacement(permf, mk, m, &st);
            scale = pow((double)m / mk, 1.0 / d);
         }
         else
            for (i = 0; i < m; i++)
               permf[i] = pool[i];
         if (rank_sample_upload(v, h_Xw, h_ld, data, ldim, d, permf, m, scale, &rs))
            goto done;
         sv = *v;
         sv.d_Xw = rs.d_Xw;
         sv.n = m;
         if (b200_dense_view_block(&sv, v->mu * v->f * v->f, NULL, m, NULL, m, 0, d_K, m) || b200_d2h(Km, d_K, sizeof(double) * (size_t)m * m))
            goto done;
         if (b200_syev_values(m, Km, work))
            goto done;
         for (i = 0; i < m; i++)
            cnt += (work[i] > thr);
         if (cnt > r)
            r = cnt;
         rank_sample_free(&rs);
      }
      rc = (int)floor((double)r * mk / m);
   }
done:
   rank_sample_free(&rs);
   b200_free(d_c);
   b200_free(d_permf);
   b200_free(d_K);
   free(pool);
   free(permf);
   free(Km);
   free(work);
   return rc;
}

/* Nfft4GPExpandPerm (utils.c:208): the k selected indices first, the remaining ones in ascending order */
static int *expand_perm(const int *sel, int k, int n)
{
   int *perm = (int *)b200_xmalloc(sizeof(int) * (size_t)n), i, pos = k;
   char *mark = (char *)b200_xcalloc((size_t)n, 1);
   for (i = 0; i < k; i++)
   {
      perm[i] = sel[i];
      mark[sel[i]] = 1;
   }
   for (i = 0; i < n; i++)
      if (!mark[i])
         perm[pos++] = i;
   free(mark);
   return perm;
}

static int afn_use_nystrom(b200_afn *s, double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int *perm, int k)
{
   void *nys = Nfft4GPPrecondNysCreate();
   int rc = nys_setup_perm((b200_nys *)nys, data, n, ldim, d, fkernel, fkernel_params, perm, k, 0);
   if (rc)
   {
      Nfft4GPPrecondNysFree(nys);
      return -1;
   }
   s->nys = nys;
   s->n = n;
   s->k = k;
   return 0;
}

int Nfft4GPPrecondAFNSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad,
                                     void *vafn_mat)
{
   b200_afn *s = (b200_afn *)vafn_mat;
   b200_dense_view v;
   double t0 = wtime();
   int own, rc = -1, k, max_k, n2, lfil, kmax, bad = 0, i;
   int *sel = NULL, *perm = NULL, *d_sel = NULL, *d_perm = NULL, *d_ia = NULL, *d_ja = NULL;
   double *d_X = NULL, *d_X2 = NULL, *d_c = NULL, *d_K11 = NULL, *d_Linv = NULL, *d_K12 = NULL, *d_GK12 = NULL, *d_a = NULL;
   long long nnz;
   if (!s || s->magic != B200_MAGIC_AFN || !fkernel || !fkernel_params || b200_init(-1))
      return -1;
   (void)require_grad; /* accepted: the derivative factors only serve Dvp / Trace / Logdet, which this library does not export */
   if (d > 64)
   {
      printf("Error: the AFN setup supports at most 64 features.\n");
      return -1;
   }
   own = b200_dense_view_acquire(fkernel, fkernel_params, data, n, ldim, d, &v);
   if (own)
   {
      if (own > 0)
         printf("Error: the AFN setup needs one of this library's dense kernels (Nfft4GPKernel{Additive,Gaussian,Matern12}Kernel).\n");
      return -1;
   }
   b200_afn_release(s);
   max_k = s->max_k < n ? s->max_k : n;
   if (max_k < 1)
      max_k = 1;
   /* 1. rank estimation (afn_setup.m:55-58) */
   if (s->est_num > 0)
   {
      k = afn_rank_estimate(s, &v, fkernel, fkernel_params, data, n, ldim, d, max_k);
      if (k < 0)
         goto done;
      s->k_est = k;
      if (k > max_k)
         k = max_k;
   }
   else
   {
      k = max_k;
      s->k_est = -1;
   }
   if (k < 1)
      k = 1;
   /* 2. landmarks = farthest point sampling on all features (afn_setup.m:61-63) */
   d_X = upload_features(data, n, ldim, d);
   d_sel = (int *)b200_malloc(sizeof(int) * (size_t)k);
   sel = (int *)b200_xmalloc(sizeof(int) * (size_t)k);
   if (!d_X || !d_sel || centroid_to_device(data, n, ldim, d, NULL, 1.0, &d_c) || b200k_fps(d_X, n, n, d, d_c, k, d_sel) ||
       b200_d2h(sel, d_sel, sizeof(int) * (size_t)k))
      goto done;
   perm = expand_perm(sel, k, n);
   /* 3. low rank: Nystrom on the FPS ordering (afn_setup.m:80-84, afn.c:294-304) */
   if (k < max_k && !s->noran)
   {
      rc = afn_use_nystrom(s, data, n, ldim, d, fkernel, fkernel_params, perm, k);
      goto done;
   }
   /* 4. K11 = L L' (with the noise term and the stabilising shift, chol_setup.m:52-56 / chol.c:439-466) */
   n2 = n - k;
   d_perm = upload_ints(perm, (size_t)n);
   d_K11 = (double *)b200_malloc(sizeof(double) * (size_t)k * k);
   d_Linv = (double *)b200_malloc(sizeof(double) * (size_t)k * k);
   if (!d_perm || !d_K11 || !d_Linv || b200_dense_view_block(&v, v.mu * v.f * v.f, d_perm, k, d_perm, k, 1, d_K11, k) ||
       landmark_cholesky(k, d_K11, d_Linv))
      goto done;
   if (n2 > 0)
   {
      /* 5. K12 = K(perm1, perm2), GK12 = L^{-1} K12 (schurCombinedKernel.m; kernels.c:3510-3523) */
      lfil = s->lfil < 1 ? 1 : s->lfil;
      kmax = lfil < n2 ? lfil : n2;
      nnz = b200k_knn_nnz(n2, lfil);
      if (nnz > 2147483647LL)
      {
         printf("Error: FSAI pattern with %lld entries exceeds the 32-bit index range of the reference's CSR format.\n", nnz);
         goto done;
      }
      d_K12 = (double *)b200_malloc(sizeof(double) * (size_t)k * n2);
      d_GK12 = (double *)b200_malloc(sizeof(double) * (size_t)k * n2);
      d_X2 = (double *)b200_malloc(sizeof(double) * (size_t)n2 * d);
      d_ia = (int *)b200_malloc(sizeof(int) * ((size_t)n2 + 1));
      d_ja = (int *)b200_malloc(sizeof(int) * (size_t)nnz);
      d_a = (double *)b200_malloc(sizeof(double) * (size_t)nnz);
      if (!d_K12 || !d_GK12 || !d_X2 || !d_ia || !d_ja || !d_a)
         goto done;
      if (b200_dense_view_block(&v, 0.0, d_perm, k, d_perm + k, n2, 0, d_K12, k) ||
          b200k_dgemm('N', 'N', k, n2, k, 1.0, d_Linv, k, d_K12, k, 0.0, d_GK12, k))
         goto done;
      /* 6. FSAI of the Schur complement S = K22 - GK12' GK12 on the non-landmark points in their order
       *    (fsai_setup.m:42-88 with the pattern size of the compiled C code: lfil entries per row incl. the diagonal) */
      if (b200k_gather_rows(n2, d, d_perm + k, d_X, n, d_X2, n2) || b200k_knn_pattern(d_X2, n2, n2, d, lfil, d_ia, d_ja) ||
          b200k_fsai_rows(v.d_Xw, v.d_dims, v.n, v.P, v.dw, v.kernel, v.f * v.f, v.l, v.mu * v.f * v.f, n2, d_ia, d_ja, d_perm + k, d_GK12, k,
                          kmax, d_a, &bad))
         goto done;
      if (bad)
      {
         /* afn_setup.m:93-98: FSAI broke down -> Nystrom with the maximum rank */
         printf("Warning: FSAI of the Schur complement broke down in %d rows; switching to the Nystrom preconditioner of rank %d.\n", bad, max_k);
         if (k < max_k)
         {
            /* landmarks for the larger rank */
            b200_free(d_sel);
            free(sel);
            free(perm);
            perm = NULL;
            d_sel = (int *)b200_malloc(sizeof(int) * (size_t)max_k);
            sel = (int *)b200_xmalloc(sizeof(int) * (size_t)max_k);
            if (!d_sel || b200k_fps(d_X, n, n, d, d_c, max_k, d_sel) || b200_d2h(sel, d_sel, sizeof(int) * (size_t)max_k))
               goto done;
            perm = expand_perm(sel, max_k, n);
         }
         rc = afn_use_nystrom(s, data, n, ldim, d, fkernel, fkernel_params, perm, max_k);
         goto done;
      }
      if (b200_fsai_set_device_factors(&s->schur, n2, (int)nnz, d_ia, d_ja, d_a))
      {
         d_ia = d_ja = NULL;
         d_a = NULL;
         goto done;
      }
      d_ia = d_ja = NULL;
      d_a = NULL;
   }
   s->n = n;
   s->k = k;
   s->d_perm = d_perm;
   s->d_Linv = d_Linv;
   s->d_K12 = d_K12;
   s->d_work = (double *)b200_malloc(sizeof(double) * (2 * (size_t)n + 2 * (size_t)k));
   d_perm = NULL;
   d_Linv = d_K12 = NULL;
   rc = s->d_work ? 0 : -1;
done:
   i = rc;
   b200_free(d_X);
   b200_free(d_X2);
   b200_free(d_c);
   b200_free(d_K11);
   b200_free(d_Linv);
   b200_free(d_K12);
   b200_free(d_GK12);
   b200_free(d_a);
   b200_free(d_sel);
   b200_free(d_perm);
   b200_free(d_ia);
   b200_free(d_ja);
   free(sel);
   free(perm);
   b200_dense_view_release(&v);
   if (!i)
      i = b200_afn_finish_setup(s); /* optional fp32 storage of K12, once the setup's work arrays are gone */
   s->tset = wtime() - t0;
   return i;
}
