/* Preconditioner handles (C host side): Nystrom (NYS), FSAI and the AFN 2x2 block preconditioner.
 * Factors live in HBM; the applies are sequences of the HBM-bound kernels of precond_kernels.cu.
 *   NYS  apply  SRC/preconds/nys.c:115-173   M^{-1} r = U S U' r + (I - U U') r / eta     (2 passes over U here, 3 in the reference)
 *   FSAI apply  SRC/preconds/fsai.c:106-123  M^{-1} r = G' (G r), G lower-triangular CSR   (G' stored explicitly: no scatter)
 *   AFN  apply  SRC/preconds/afn.c:109-136 (uncompiled in the reference) = MATLAB afn_solve.m:39-69
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

/* ------------------------------------------------------------------ staging for the host-pointer entry points */
static int stage_vectors(double **d_x, double **d_rhs, int *cap, int n)
{
   if (*cap < n)
   {
      b200_free(*d_x);
      b200_free(*d_rhs);
      *d_x = (double *)b200_malloc(sizeof(double) * (size_t)n);
      *d_rhs = (double *)b200_malloc(sizeof(double) * (size_t)n);
      *cap = n;
   }
   return (*d_x && *d_rhs) ? 0 : -1;
}

/* ================================================================== NYS */
void *Nfft4GPPrecondNysCreate(void)
{
   b200_nys *s = (b200_nys *)b200_xcalloc(1, sizeof(b200_nys));
   s->magic = B200_MAGIC_NYS;
   s->k_setup = 50; /* nys.c:8 */
   s->eta = 1.0;
   return s;
}

static void nys_release_factors(b200_nys *s)
{
   b200_free(s->d_perm);
   b200_free(s->d_U);
   b200_free(s->d_s);
   b200_free(s->d_work);
   b200_free(s->d_UT_loc);
   b200_free(s->d_sh_work);
   b200_free(s->d_gK1);
   b200_free(s->d_gdK1l);
   b200_free(s->d_gLinv);
   b200_free(s->d_gG[0]);
   b200_free(s->d_gG[1]);
   b200_free(s->d_gvec);
   s->d_gK1 = s->d_gdK1l = s->d_gLinv = s->d_gG[0] = s->d_gG[1] = s->d_gvec = NULL;
   s->has_grad = s->traces_ready = 0;
   s->d_perm = NULL;
   s->d_U = s->d_s = s->d_work = s->d_UT_loc = s->d_sh_work = NULL;
   s->work_doubles = 0;
   s->n = s->k = 0;
   s->sh_lo = s->sh_hi = 0;
}

void Nfft4GPPrecondNysReset(void *str) /* nys.c:78-98: keep permutation */
{
   b200_nys *s = (b200_nys *)str;
   if (s)
      nys_release_factors(s);
}

void Nfft4GPPrecondNysFree(void *str)
{
   b200_nys *s = (b200_nys *)str;
   if (!s)
      return;
   nys_release_factors(s);
   b200_free(s->d_x);
   b200_free(s->d_rhs);
   if (s->own_perm)
      free(s->perm);
   free(s);
}

void Nfft4GPPrecondNysSetRank(void *str, int k) { ((b200_nys *)str)->k_setup = k; }

void Nfft4GPPrecondNysSetPerm(void *str, int *perm, int own_perm)
{
   b200_nys *s = (b200_nys *)str;
   s->perm = perm;
   s->own_perm = own_perm;
}

int b200_nys_set_device_factors(b200_nys *s, int n, int k, const int *h_perm, double *d_U_owned, const double *h_s, double eta)
{
   nys_release_factors(s);
   s->n = n;
   s->k = k;
   s->eta = eta;
   s->d_U = d_U_owned;
   s->d_s = (double *)b200_malloc(sizeof(double) * (size_t)k);
   s->work_doubles = b200k_nys_work_doubles(n, k);
   s->d_work = (double *)b200_malloc(sizeof(double) * s->work_doubles);
   if (b200_h2d(s->d_s, h_s, sizeof(double) * (size_t)k))
      return -1;
   if (h_perm)
   {
      s->d_perm = (int *)b200_malloc(sizeof(int) * (size_t)n);
      if (b200_h2d(s->d_perm, h_perm, sizeof(int) * (size_t)n))
         return -1;
   }
   return b200_sync();
}

int Nfft4GPB200PrecondNysSetFactors(void *nys, int n, int k, const int *perm, const double *U, const double *sv, double eta)
{
   b200_nys *s = (b200_nys *)nys;
   double *d_U;
   if (!s || s->magic != B200_MAGIC_NYS || b200_init(-1))
      return -1;
   d_U = (double *)b200_malloc(sizeof(double) * (size_t)n * (size_t)k);
   if (b200_h2d(d_U, U, sizeof(double) * (size_t)n * (size_t)k) || b200_sync())
      return -1;
   return b200_nys_set_device_factors(s, n, k, perm, d_U, sv, eta);
}

int Nfft4GPB200PrecondNysGetFactors(void *nys, int *n, int *k, int *perm, double *U, double *sv, double *eta)
{
   b200_nys *s = (b200_nys *)nys;
   if (!s || s->magic != B200_MAGIC_NYS)
      return -1;
   if (n)
      *n = s->n;
   if (k)
      *k = s->k;
   if (eta)
      *eta = s->eta;
   if (perm && s->d_perm && b200_d2h(perm, s->d_perm, sizeof(int) * (size_t)s->n))
      return -1;
   if (U && s->d_U && b200_d2h(U, s->d_U, sizeof(double) * (size_t)s->n * (size_t)s->k))
      return -1;
   if (sv && s->d_s && b200_d2h(sv, s->d_s, sizeof(double) * (size_t)s->k))
      return -1;
   return 0;
}

/* Point-block-sharded Nystrom apply: vectors hold this rank's block [lo, hi) of the points in the original
 * order; rows of U that belong to the block are kept transposed (k x nl).  z = U'r needs ONE allreduce of k doubles. */
static int nys_solve_sharded(b200_nys *s, int nl, double *d_x, const double *d_rhs)
{
   int rank = b200_comm_rank(), size = b200_comm_size(), k = s->k;
   int lo = (int)((long long)s->n * rank / size), hi = (int)((long long)s->n * (rank + 1) / size);
   double *z, *c;
   if (nl != hi - lo)
   {
      printf("Error: sharded Nystrom apply: block of %d points expected, got %d.\n", hi - lo, nl);
      return -1;
   }
   if (!s->d_UT_loc || s->sh_lo != lo || s->sh_hi != hi)
   {
      /* positions (in the permuted order of U) of the local points, by local index */
      int *perm = (int *)b200_xmalloc(sizeof(int) * (size_t)s->n), *pos = (int *)b200_xmalloc(sizeof(int) * (size_t)(nl ? nl : 1)), i, *d_pos;
      if (s->d_perm)
      {
         if (b200_d2h(perm, s->d_perm, sizeof(int) * (size_t)s->n))
            return -1;
      }
      else
         for (i = 0; i < s->n; i++)
            perm[i] = i;
      for (i = 0; i < s->n; i++)
         if (perm[i] >= lo && perm[i] < hi)
            pos[perm[i] - lo] = i;
      b200_free(s->d_UT_loc);
      b200_free(s->d_sh_work);
      s->d_UT_loc = (double *)b200_malloc(sizeof(double) * (size_t)k * (size_t)(nl ? nl : 1));
      s->d_sh_work = (double *)b200_malloc(sizeof(double) * 2 * (size_t)k);
      d_pos = (int *)b200_malloc(sizeof(int) * (size_t)(nl ? nl : 1));
      i = !s->d_UT_loc || !s->d_sh_work || !d_pos || b200_h2d(d_pos, pos, sizeof(int) * (size_t)nl) ||
          b200k_gather_rows_T(nl, k, d_pos, s->d_U, s->n, s->d_UT_loc) || b200_sync();
      b200_free(d_pos);
      free(perm);
      free(pos);
      if (i)
         return -1;
      s->sh_lo = lo;
      s->sh_hi = hi;
   }
   z = s->d_sh_work;
   c = z + k;
   /* z = U_loc' r (k x nl times nl), allreduce, x = r / eta + U_loc ((s - 1/eta) .* z) */
   if (b200k_gemv('N', k, nl, 1.0, s->d_UT_loc, d_rhs, 0.0, z) || b200_comm_allreduce_sum(z, (size_t)k) ||
       b200k_nys_coef(k, s->d_s, 1.0 / s->eta, z, c) || b200k_copy(nl, d_rhs, d_x) || b200k_scale(nl, 1.0 / s->eta, d_x) ||
       b200k_gemv('T', k, nl, 1.0, s->d_UT_loc, c, 1.0, d_x))
      return -1;
   return 0;
}

int b200_nys_solve_dev(void *vnys, int n, double *d_x, const double *d_rhs)
{
   b200_nys *s = (b200_nys *)vnys;
   if (s && s->magic == B200_MAGIC_NYS && s->d_U && n != s->n && b200_comm_size() > 1)
      return nys_solve_sharded(s, n, d_x, d_rhs);
   if (!s || s->magic != B200_MAGIC_NYS || !s->d_U || n != s->n)
   {
      printf("Error: Nystrom preconditioner is not set up for n = %d.\n", n);
      return -1;
   }
   return b200k_nys_apply(s->n, s->k, s->d_perm, s->d_U, s->d_s, s->eta, d_rhs, d_x, s->d_work, s->work_doubles);
}

int Nfft4GPPrecondNysSolve(void *vnys_mat, int n, double *x, double *rhs) /* nys.c:115-173 */
{
   b200_nys *s = (b200_nys *)vnys_mat;
   if (!s || s->magic != B200_MAGIC_NYS || stage_vectors(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   if (b200_h2d(s->d_rhs, rhs, sizeof(double) * (size_t)n) || b200_nys_solve_dev(s, n, s->d_x, s->d_rhs))
      return -1;
   return b200_d2h(x, s->d_x, sizeof(double) * (size_t)n);
}

/* ================================================================== FSAI */
void *Nfft4GPPrecondFsaiCreate(void)
{
   b200_fsai *s = (b200_fsai *)b200_xcalloc(1, sizeof(b200_fsai));
   s->magic = B200_MAGIC_FSAI;
   s->lfil = 50; /* fsai.c:12 */
   return s;
}

void b200_fsai_release(b200_fsai *s)
{
   b200_free(s->d_ia);
   b200_free(s->d_ja);
   b200_free(s->d_a);
   b200_free(s->d_iat);
   b200_free(s->d_jat);
   b200_free(s->d_at);
   b200_free(s->d_work);
   b200_free(s->d_da);
   b200_free(s->d_dat);
   b200_free(s->d_gwork);
   s->d_ia = s->d_ja = s->d_iat = s->d_jat = NULL;
   s->d_a = s->d_at = s->d_work = NULL;
   s->d_da = s->d_dat = s->d_gwork = NULL;
   s->n = s->nnz = 0;
}

void Nfft4GPPrecondFsaiReset(void *str)
{
   b200_fsai *s = (b200_fsai *)str;
   if (s)
      b200_fsai_release(s);
}

void Nfft4GPPrecondFsaiFree(void *str)
{
   b200_fsai *s = (b200_fsai *)str;
   if (!s)
      return;
   b200_fsai_release(s);
   b200_free(s->d_x);
   b200_free(s->d_rhs);
   free(s);
}

void Nfft4GPPrecondFsaiSetLfil(void *str, int lfil) { ((b200_fsai *)str)->lfil = lfil; }

/* takes ownership of the device CSR arrays of G and builds the explicit transpose (columns ordered by row) on the device */
int b200_fsai_set_device_factors(b200_fsai *s, int n, int nnz, int *d_ia, int *d_ja, double *d_a)
{
   b200_fsai_release(s);
   s->n = n;
   s->nnz = nnz;
   s->d_ia = d_ia;
   s->d_ja = d_ja;
   s->d_a = d_a;
   s->d_iat = (int *)b200_malloc(sizeof(int) * ((size_t)n + 1));
   s->d_jat = (int *)b200_malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
   s->d_at = (double *)b200_malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
   s->d_work = (double *)b200_malloc(sizeof(double) * (size_t)(n ? n : 1));
   if (!s->d_iat || !s->d_jat || !s->d_at || !s->d_work)
      return -1;
   return b200k_csr_transpose(n, n, nnz, d_ia, d_ja, d_a, s->d_iat, s->d_jat, s->d_at) || b200_sync();
}

int b200_fsai_set_device_grad(b200_fsai *s, double *d_da)
{
   int t, *iat_tmp, *jat_tmp;
   size_t nnz = (size_t)(s->nnz ? s->nnz : 1);
   s->d_da = d_da;
   s->d_dat = (double *)b200_malloc(sizeof(double) * 3 * nnz);
   s->d_gwork = (double *)b200_malloc(sizeof(double) * 4 * (size_t)(s->n ? s->n : 1));
   iat_tmp = (int *)b200_malloc(sizeof(int) * ((size_t)s->n + 1));
   jat_tmp = (int *)b200_malloc(sizeof(int) * nnz);
   t = !s->d_dat || !s->d_gwork || !iat_tmp || !jat_tmp;
   /* same pattern, same ordering rule: the transposed values land where the entries of G' are */
   for (int q = 0; !t && q < 3; q++)
      t = b200k_csr_transpose(s->n, s->n, s->nnz, s->d_ia, s->d_ja, s->d_da + q * (size_t)s->nnz, iat_tmp, jat_tmp, s->d_dat + q * (size_t)s->nnz);
   t = t || b200_sync();
   b200_free(iat_tmp);
   b200_free(jat_tmp);
   return t ? -1 : 0;
}

/* ---- gradient side of the FSAI preconditioner M = (G'G)^{-1} (fsai.c:125-276, 675-721), device resident */
static int fsai_has_grad(const b200_fsai *s, const char *who)
{
   if (!s || s->magic != B200_MAGIC_FSAI || !s->d_a)
      return 0;
   if (!s->d_da)
   {
      printf("Setup FSAI without gradient, %s not supported.\n", who); /* fsai.c:141-145, 229-233 */
      return 0;
   }
   return 1;
}

/* fsai.c:675-698 */
int Nfft4GPPrecondFsaiInvL(void *vfsai_mat, int n, double *x, double *rhs)
{
   b200_fsai *s = (b200_fsai *)vfsai_mat;
   if (!s || s->magic != B200_MAGIC_FSAI || !s->d_a || n != s->n || stage_vectors(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   if (b200_h2d(s->d_rhs, rhs, sizeof(double) * (size_t)n) || b200k_sptrsv(0, n, s->d_ia, s->d_ja, s->d_a, s->d_rhs, s->d_x))
      return -1;
   return b200_d2h(x, s->d_x, sizeof(double) * (size_t)n);
}

/* fsai.c:700-721 */
int Nfft4GPPrecondFsaiInvLT(void *vfsai_mat, int n, double *x, double *rhs)
{
   b200_fsai *s = (b200_fsai *)vfsai_mat;
   if (!s || s->magic != B200_MAGIC_FSAI || !s->d_a || n != s->n || stage_vectors(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   if (b200_h2d(s->d_rhs, rhs, sizeof(double) * (size_t)n) || b200k_sptrsv(1, n, s->d_iat, s->d_jat, s->d_at, s->d_rhs, s->d_x))
      return -1;
   return b200_d2h(x, s->d_x, sizeof(double) * (size_t)n);
}

/* fsai.c:221-276: tr(M^{-1} dM/dtheta) = 2 sum_i (dG_ii / dtheta) / G_ii for theta = f, l, mu */
int Nfft4GPPrecondFsaiTrace(void *vfsai_mat, double **tracesp)
{
   b200_fsai *s = (b200_fsai *)vfsai_mat;
   double *tr, *d3;
   int rc;
   if (!fsai_has_grad(s, "trace"))
      return -1;
   if (!tracesp)
   {
      printf("Trace pointer cannot be NULL\n");
      return -1;
   }
   d3 = (double *)b200_malloc(sizeof(double) * 3);
   tr = *tracesp ? *tracesp : (double *)b200_xcalloc(3, sizeof(double));
   {
      double h3[3];
      rc = !d3 || b200k_fsai_trace(s->n, s->d_ia, s->d_a, s->d_da, (size_t)s->nnz, d3) || b200_d2h(h3, d3, sizeof(h3));
      if (!rc)
      {
         tr[0] += h3[0]; /* the reference accumulates into the caller's array (fsai.c:257-259) */
         tr[1] += h3[1];
         tr[2] += h3[2];
      }
   }
   b200_free(d3);
   if (rc)
   {
      if (!*tracesp)
         free(tr);
      return -1;
   }
   *tracesp = tr;
   return 0;
}

/* fsai.c:125-219: y_theta = G' ( G^{-T} dG' G^{-T} x + dG G^{-1} G^{-T} x ) for the selected hyper-parameters, stacked [f; l; mu] */
int Nfft4GPPrecondFsaiDvp(void *vfsai_mat, int n, int *mask, double *x, double **yp)
{
   b200_fsai *s = (b200_fsai *)vfsai_mat;
   double *y, *w1, *w2, *t1, *t2;
   int q, rc = 0;
   size_t nb = sizeof(double) * (size_t)n;
   if (!fsai_has_grad(s, "trace"))
      return -1;
   if (!yp)
   {
      printf("output pointer cannot be NULL\n");
      return -1;
   }
   if (n != s->n || stage_vectors(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   y = *yp ? *yp : (double *)b200_xcalloc(3 * (size_t)n, sizeof(double));
   w1 = s->d_gwork;
   w2 = w1 + n;
   t1 = w2 + n;
   t2 = t1 + n;
   rc = b200_h2d(s->d_rhs, x, nb) || b200k_sptrsv(1, n, s->d_iat, s->d_jat, s->d_at, s->d_rhs, w1); /* w1 = G^{-T} x */
   rc = rc || b200k_sptrsv(0, n, s->d_ia, s->d_ja, s->d_a, w1, t2);                                /* t2 = G^{-1} w1 */
   for (q = 0; !rc && q < 3; q++)
   {
      const double *da = s->d_da + (size_t)q * s->nnz, *dat = s->d_dat + (size_t)q * s->nnz;
      if (mask && !mask[q])
         continue;
      rc = b200k_csr_mv(n, s->d_iat, s->d_jat, dat, w1, t1) ||            /* t1 = dG' w1            */
           b200k_sptrsv(1, n, s->d_iat, s->d_jat, s->d_at, t1, w2) ||      /* w2 = G^{-T} t1         */
           b200k_csr_mv(n, s->d_ia, s->d_ja, da, t2, t1) ||                /* t1 = dG G^{-1} G^{-T} x */
           b200k_axpy(n, 1.0, t1, w2) ||                                   /* w2 += t1               */
           b200k_csr_mv(n, s->d_iat, s->d_jat, s->d_at, w2, s->d_x) ||     /* y = G' w2              */
           b200_d2h(y + (size_t)q * n, s->d_x, nb);
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

/* upload G (CSR, rows as given) */
int b200_fsai_set_host_factors(b200_fsai *s, int n, const int *ia, const int *ja, const double *a)
{
   int nnz = ia[n];
   int *d_ia = (int *)b200_malloc(sizeof(int) * ((size_t)n + 1));
   int *d_ja = (int *)b200_malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
   double *d_a = (double *)b200_malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
   if (!d_ia || !d_ja || !d_a || b200_h2d(d_ia, ia, sizeof(int) * ((size_t)n + 1)) || b200_h2d(d_ja, ja, sizeof(int) * (size_t)nnz) ||
       b200_h2d(d_a, a, sizeof(double) * (size_t)nnz) || b200_sync())
   {
      b200_free(d_ia);
      b200_free(d_ja);
      b200_free(d_a);
      return -1;
   }
   return b200_fsai_set_device_factors(s, n, nnz, d_ia, d_ja, d_a);
}

int Nfft4GPB200PrecondFsaiSetFactors(void *fsai, int n, const int *ia, const int *ja, const double *a)
{
   b200_fsai *s = (b200_fsai *)fsai;
   if (!s || s->magic != B200_MAGIC_FSAI || b200_init(-1))
      return -1;
   return b200_fsai_set_host_factors(s, n, ia, ja, a);
}

int Nfft4GPB200PrecondFsaiGetGradFactors(void *fsai, double *da)
{
   b200_fsai *s = (b200_fsai *)fsai;
   if (!s || s->magic != B200_MAGIC_FSAI || !s->d_da || !da)
      return -1;
   return b200_d2h(da, s->d_da, sizeof(double) * 3 * (size_t)s->nnz);
}

int Nfft4GPB200PrecondFsaiGetFactors(void *fsai, int *n, int *nnz, int *ia, int *ja, double *a)
{
   b200_fsai *s = (b200_fsai *)fsai;
   if (!s || s->magic != B200_MAGIC_FSAI)
      return -1;
   if (n)
      *n = s->n;
   if (nnz)
      *nnz = s->nnz;
   if (ia && s->d_ia && b200_d2h(ia, s->d_ia, sizeof(int) * ((size_t)s->n + 1)))
      return -1;
   if (ja && s->d_ja && b200_d2h(ja, s->d_ja, sizeof(int) * (size_t)s->nnz))
      return -1;
   if (a && s->d_a && b200_d2h(a, s->d_a, sizeof(double) * (size_t)s->nnz))
      return -1;
   return 0;
}

int b200_fsai_solve_dev(void *vf, int n, double *d_x, const double *d_rhs)
{
   b200_fsai *s = (b200_fsai *)vf;
   if (!s || s->magic != B200_MAGIC_FSAI || !s->d_a || n != s->n)
   {
      printf("Error: FSAI preconditioner is not set up for n = %d.\n", n);
      return -1;
   }
   /* INV = L' * L (fsai.c:112-116) */
   if (b200k_csr_mv(n, s->d_ia, s->d_ja, s->d_a, d_rhs, s->d_work))
      return -1;
   return b200k_csr_mv(n, s->d_iat, s->d_jat, s->d_at, s->d_work, d_x);
}

int Nfft4GPPrecondFsaiSolve(void *vfsai_mat, int n, double *x, double *rhs) /* fsai.c:106-123 */
{
   b200_fsai *s = (b200_fsai *)vfsai_mat;
   if (!s || s->magic != B200_MAGIC_FSAI || stage_vectors(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   if (b200_h2d(s->d_rhs, rhs, sizeof(double) * (size_t)n) || b200_fsai_solve_dev(s, n, s->d_x, s->d_rhs))
      return -1;
   return b200_d2h(x, s->d_x, sizeof(double) * (size_t)n);
}

/* log det of the FSAI preconditioner M = (G'G)^{-1}: -2 sum log G_ii, the diagonal being the last entry of every row
 * (fsai.c:278-300).  Needs no gradient data. */
double Nfft4GPPrecondFsaiLogdet(void *vfsai_mat)
{
   b200_fsai *s = (b200_fsai *)vfsai_mat;
   int *ia, i;
   double *diag, *d_diag, val = 0.0;
   int *d_idx;
   if (!s || s->magic != B200_MAGIC_FSAI || !s->d_a || s->n <= 0)
      return 0.0;
   ia = (int *)b200_xmalloc(sizeof(int) * ((size_t)s->n + 1));
   diag = (double *)b200_xmalloc(sizeof(double) * (size_t)s->n);
   d_idx = (int *)b200_malloc(sizeof(int) * (size_t)s->n);
   d_diag = (double *)b200_malloc(sizeof(double) * (size_t)s->n);
   if (d_idx && d_diag && !b200_d2h(ia, s->d_ia, sizeof(int) * ((size_t)s->n + 1)))
   {
      for (i = 0; i < s->n; i++)
         ia[i] = ia[i + 1] - 1;
      if (!b200_h2d(d_idx, ia, sizeof(int) * (size_t)s->n) && !b200k_gather_idx(s->n, d_idx, s->d_a, d_diag) &&
          !b200_d2h(diag, d_diag, sizeof(double) * (size_t)s->n))
         for (i = 0; i < s->n; i++)
            val += log(1.0 / diag[i]);
   }
   b200_free(d_idx);
   b200_free(d_diag);
   free(ia);
   free(diag);
   return 2.0 * val;
}

/* ================================================================== AFN */
void *Nfft4GPPrecondAFNCreate(void)
{
   b200_afn *s = (b200_afn *)b200_xcalloc(1, sizeof(b200_afn));
   s->magic = B200_MAGIC_AFN;
   s->max_k = 100; /* MATLAB test settings: maxrank 100-300, lfil 100, m 500, 5 estimates (test_solve_additive.m:35-40) */
   s->lfil = 100;
   s->est_m = 500;
   s->est_num = 5;
   s->est_seed = 906ULL;
   s->k_est = -1;
   s->schur.magic = B200_MAGIC_FSAI;
   return s;
}

void b200_afn_release(b200_afn *s)
{
   b200_free(s->d_perm);
   b200_free(s->d_Linv);
   b200_free(s->d_K12);
   b200_free(s->d_K12f);
   s->d_K12f = NULL;
   b200_free(s->d_work);
   b200_free(s->d_lm_idx);
   b200_free(s->d_sc_idx);
   b200_free(s->d_sh_work);
   b200_free(s->d_opz);
   b200_free(s->d_opt);
   s->d_opz = s->d_opt = NULL;
   s->k12_op = NULL; /* a new setup starts from the dense block */
   s->d_perm = NULL;
   s->d_lm_idx = s->d_sc_idx = NULL;
   s->d_Linv = s->d_K12 = s->d_work = s->d_sh_work = NULL;
   s->sh_ready = 0;
   b200_fsai_release(&s->schur);
   if (s->nys)
   {
      Nfft4GPPrecondNysFree(s->nys);
      s->nys = NULL;
   }
   s->n = s->k = 0;
}

void Nfft4GPPrecondAFNFree(void *str)
{
   b200_afn *s = (b200_afn *)str;
   if (!s)
      return;
   b200_afn_release(s);
   b200_free(s->d_x);
   b200_free(s->d_rhs);
   free(s);
}

void Nfft4GPPrecondAFNSetRank(void *str, int max_k) { ((b200_afn *)str)->max_k = max_k; }
void Nfft4GPPrecondAFNSetLfil(void *str, int lfil) { ((b200_afn *)str)->lfil = lfil; }

int Nfft4GPB200PrecondAFNSetRankEstimation(void *str, int m, int numests, int noran, unsigned long long seed)
{
   b200_afn *s = (b200_afn *)str;
   if (!s || s->magic != B200_MAGIC_AFN)
      return -1;
   s->est_m = m > 1 ? m : 2;
   s->est_num = numests;
   s->noran = noran;
   s->est_seed = seed;
   return 0;
}

/* K12 products of the apply through the NFFT operator (see b200_afn_solve_dev).  nfft_kernel = the pointer an NFFT additive kernel
 * function returned in *Kp for the SAME data and hyper-parameters as the dense kernel of the setup (NULL switches back to the dense
 * block).  release_block != 0 frees the dense K12 block (8 k (n - k) bytes; the setup needs it, the apply then does not).  The
 * operator approximates K12 to the NFFT accuracy (3e-8 relative at a resolved length scale), in fp64; the preconditioner stays
 * symmetric because the NFFT matvec is (spread = adjoint of the gather).  Single-GPU apply only. */
int Nfft4GPB200PrecondAFNSetK12Operator(void *str, void *nfft_kernel, int release_block)
{
   b200_afn *s = (b200_afn *)str;
   b200_additive *op = nfft_kernel ? b200_additive_from_handle(nfft_kernel) : NULL;
   if (!s || s->magic != B200_MAGIC_AFN || (nfft_kernel && !op))
   {
      printf("Error: Nfft4GPB200PrecondAFNSetK12Operator needs an AFN handle and an additive NFFT kernel handle.\n");
      return -1;
   }
   if (!op)
   {
      if (!s->d_K12 && !s->d_K12f && s->d_Linv && s->n > s->k)
      {
         printf("Error: the dense K12 block was released; run the setup again.\n");
         return -1;
      }
      s->k12_op = NULL;
      return 0;
   }
   if (!s->d_Linv || s->nys)
   {
      printf("Error: Nfft4GPB200PrecondAFNSetK12Operator follows the setup (block form of the AFN preconditioner).\n");
      return -1;
   }
   if (!op->built || op->n != s->n || (op->shard_mode == 2 && b200_comm_size() > 1))
   {
      printf("Error: the NFFT kernel handle is not built for the %d points of the preconditioner (unsharded).\n", s->n);
      return -1;
   }
   if (!s->d_opz)
   {
      s->d_opz = (double *)b200_malloc(sizeof(double) * (size_t)s->n);
      s->d_opt = (double *)b200_malloc(sizeof(double) * (size_t)s->n);
      if (!s->d_opz || !s->d_opt)
         return -1;
   }
   s->k12_op = op;
   if (release_block)
   {
      b200_free(s->d_K12);
      b200_free(s->d_K12f);
      s->d_K12 = NULL;
      s->d_K12f = NULL;
   }
   return 0;
}

int Nfft4GPB200PrecondAFNSetRankMethod(void *str, int method)
{
   b200_afn *s = (b200_afn *)str;
   if (!s || s->magic != B200_MAGIC_AFN || method < 0 || method > 1)
      return -1;
   s->est_method = method;
   return 0;
}

/* Storage of the K12 block (k x (n - k)): every apply reads it twice and nothing else of comparable size, so fp32 storage
 * halves the HBM traffic of a PCG iteration at large rank.  Products and vectors stay fp64; the preconditioner stays
 * symmetric (the same rounded block is used in both products).  Takes effect at the next setup / SetFactors, or
 * immediately when factors are present.  bits = 64 restores the default (needs a new setup once the fp64 block is gone). */
static int afn_apply_k12_storage(b200_afn *s)
{
   size_t cnt;
   if (s->k12_bits != 32 || !s->d_K12 || s->d_K12f || s->n <= s->k)
      return 0;
   if (s->k & 3)
   {
      printf("Warning: fp32 storage of K12 needs a landmark count that is a multiple of 4 (k = %d): keeping fp64.\n", s->k);
      return 0;
   }
   cnt = (size_t)s->k * (size_t)(s->n - s->k);
   s->d_K12f = (float *)b200_malloc(sizeof(float) * cnt);
   if (!s->d_K12f || b200k_convert_f64_f32(cnt, s->d_K12, s->d_K12f) || b200_sync())
      return -1;
   b200_free(s->d_K12);
   s->d_K12 = NULL;
   return 0;
}

int Nfft4GPB200PrecondAFNSetK12Storage(void *afn, int bits)
{
   b200_afn *s = (b200_afn *)afn;
   if (!s || s->magic != B200_MAGIC_AFN || (bits != 32 && bits != 64))
      return -1;
   if (bits == 64 && s->d_K12f)
   {
      printf("Error: the fp64 K12 block was released; run the setup again after selecting 64-bit storage.\n");
      return -1;
   }
   s->k12_bits = bits;
   return afn_apply_k12_storage(s);
}

int b200_afn_finish_setup(b200_afn *s) { return afn_apply_k12_storage(s); }

/* y = beta y + alpha op(K12(:, col0 : col0 + ncols)) x in whichever storage the handle holds */
static int afn_k12_gemv(b200_afn *s, char trans, int ncols, int col0, double alpha, const double *x, double beta, double *y)
{
   if (s->d_K12f)
      return b200k_gemv_f32(trans, s->k, ncols, alpha, s->d_K12f + (size_t)s->k * col0, x, beta, y);
   return b200k_gemv(trans, s->k, ncols, alpha, s->d_K12 + (size_t)s->k * col0, x, beta, y);
}

int Nfft4GPB200PrecondAFNGetFactors(void *afn, int *n, int *k, int *k_est, int *is_nys, int *perm, double *Linv, double *K12, int *nnz,
                                    int *ia, int *ja, double *a)
{
   b200_afn *s = (b200_afn *)afn;
   if (!s || s->magic != B200_MAGIC_AFN)
      return -1;
   if (n)
      *n = s->n;
   if (k)
      *k = s->k;
   if (k_est)
      *k_est = s->k_est;
   if (is_nys)
      *is_nys = s->nys != NULL;
   if (s->nys)
      return perm ? Nfft4GPB200PrecondNysGetFactors(s->nys, NULL, NULL, perm, NULL, NULL, NULL) : 0;
   if (perm && s->d_perm && b200_d2h(perm, s->d_perm, sizeof(int) * (size_t)s->n))
      return -1;
   if (Linv && s->d_Linv && b200_d2h(Linv, s->d_Linv, sizeof(double) * (size_t)s->k * (size_t)s->k))
      return -1;
   if (K12 && s->d_K12 && b200_d2h(K12, s->d_K12, sizeof(double) * (size_t)s->k * (size_t)(s->n - s->k)))
      return -1;
   if (K12 && s->d_K12f)
   {
      /* fp32 storage: hand back the rounded block in the fp64 layout of the interface */
      size_t cnt = (size_t)s->k * (size_t)(s->n - s->k);
      double *tmp = (double *)b200_malloc(sizeof(double) * cnt);
      int rc = !tmp || b200k_convert_f32_f64(cnt, s->d_K12f, tmp) || b200_d2h(K12, tmp, sizeof(double) * cnt);
      b200_free(tmp);
      if (rc)
         return -1;
   }
   return Nfft4GPB200PrecondFsaiGetFactors(&s->schur, NULL, nnz, ia, ja, a);
}

int b200_afn_set_device_factors(b200_afn *s, int n, int k, const int *h_perm, const double *h_L, double *d_K12_owned,
                                const int *ia, const int *ja, const double *a)
{
   double *Linv = (double *)b200_xmalloc(sizeof(double) * (size_t)k * (size_t)k);
   int rc;
   b200_afn_release(s);
   s->n = n;
   s->k = k;
   b200_tril_inverse(k, h_L, Linv);
   s->d_Linv = (double *)b200_malloc(sizeof(double) * (size_t)k * (size_t)k);
   s->d_perm = (int *)b200_malloc(sizeof(int) * (size_t)n);
   s->d_K12 = d_K12_owned;
   s->d_work = (double *)b200_malloc(sizeof(double) * (2 * (size_t)n + 2 * (size_t)k));
   rc = b200_h2d(s->d_Linv, Linv, sizeof(double) * (size_t)k * (size_t)k) || b200_h2d(s->d_perm, h_perm, sizeof(int) * (size_t)n) || b200_sync();
   free(Linv);
   if (rc)
      return -1;
   return b200_fsai_set_host_factors(&s->schur, n - k, ia, ja, a) || afn_apply_k12_storage(s);
}

int Nfft4GPB200PrecondAFNSetFactors(void *afn, int n, int k, const int *perm, const double *L, const double *K12, const int *ia,
                                    const int *ja, const double *a)
{
   b200_afn *s = (b200_afn *)afn;
   double *d_K12;
   size_t sz = sizeof(double) * (size_t)k * (size_t)(n - k);
   if (!s || s->magic != B200_MAGIC_AFN || b200_init(-1))
      return -1;
   d_K12 = (double *)b200_malloc(sz);
   if (b200_h2d(d_K12, K12, sz) || b200_sync())
      return -1;
   return b200_afn_set_device_factors(s, n, k, perm, L, d_K12, ia, ja, a);
}

/* Point-block-sharded AFN apply.  Vectors hold this rank's block [lo, hi) of the points in the original
 * order.  The k landmark unknowns are replicated (two allreduces of k doubles); the Schur unknowns are sharded -- the
 * non-landmark points are in ascending original order, so a block owns a CONTIGUOUS range [plo, phi) of Schur indices, i.e.
 * a column range of K12 and a row range of G and G' -- and the two FSAI products read the full Schur vector, assembled by
 * an allreduce of n - k doubles each.  Same arithmetic as afn_solve.m:39-69. */
static int afn_shard_prepare(b200_afn *s, int lo, int hi)
{
   int n = s->n, k = s->k, n2 = n - k, p, nl = hi - lo, n2l, rc;
   int *perm, *lm, *sc;
   if (s->sh_ready && s->sh_lo == lo && s->sh_hi == hi)
      return 0;
   perm = (int *)b200_xmalloc(sizeof(int) * (size_t)n);
   if (b200_d2h(perm, s->d_perm, sizeof(int) * (size_t)n))
   {
      free(perm);
      return -1;
   }
   lm = (int *)b200_xmalloc(sizeof(int) * (size_t)k);
   for (p = 0; p < k; p++)
      lm[p] = (perm[p] >= lo && perm[p] < hi) ? perm[p] - lo : -1;
   s->sh_plo = n2;
   s->sh_phi = n2;
   for (p = k; p < n; p++)
      if (perm[p] >= lo)
      {
         s->sh_plo = p - k;
         break;
      }
   for (p = k + s->sh_plo; p < n; p++)
      if (perm[p] >= hi)
      {
         s->sh_phi = p - k;
         break;
      }
   n2l = s->sh_phi - s->sh_plo;
   sc = (int *)b200_xmalloc(sizeof(int) * (size_t)(n2l ? n2l : 1));
   for (p = 0; p < n2l; p++)
      sc[p] = perm[k + s->sh_plo + p] - lo;
   b200_free(s->d_lm_idx);
   b200_free(s->d_sc_idx);
   b200_free(s->d_sh_work);
   s->d_lm_idx = (int *)b200_malloc(sizeof(int) * (size_t)k);
   s->d_sc_idx = (int *)b200_malloc(sizeof(int) * (size_t)(n2l ? n2l : 1));
   s->d_sh_work = (double *)b200_malloc(sizeof(double) * (4 * (size_t)k + 2 * (size_t)n2 + 2 * (size_t)n2l + 8));
   rc = !s->d_lm_idx || !s->d_sc_idx || !s->d_sh_work || b200_h2d(s->d_lm_idx, lm, sizeof(int) * (size_t)k) ||
        b200_h2d(s->d_sc_idx, sc, sizeof(int) * (size_t)n2l) || b200_sync();
   free(perm);
   free(lm);
   free(sc);
   (void)nl;
   if (rc)
      return -1;
   s->sh_lo = lo;
   s->sh_hi = hi;
   s->sh_ready = 1;
   return 0;
}

static int afn_solve_sharded(b200_afn *s, int nl, double *d_x, const double *d_rhs)
{
   int rank = b200_comm_rank(), size = b200_comm_size(), k = s->k, n2 = s->n - s->k, n2l, plo;
   int lo = (int)((long long)s->n * rank / size), hi = (int)((long long)s->n * (rank + 1) / size);
   double *rp1, *y1, *t, *u, *full, *wfull, *rp2l, *y2l;
   b200_fsai *G = &s->schur;
   if (nl != hi - lo)
   {
      printf("Error: sharded AFN apply: block of %d points expected, got %d.\n", hi - lo, nl);
      return -1;
   }
   if (afn_shard_prepare(s, lo, hi))
      return -1;
   plo = s->sh_plo;
   n2l = s->sh_phi - s->sh_plo;
   rp1 = s->d_sh_work;
   y1 = rp1 + k;
   t = y1 + k;
   u = t + k;
   full = u + k;
   wfull = full + n2;
   rp2l = wfull + n2;
   y2l = rp2l + n2l;
   /* rp1 = r(perm[0:k]) (owned entries, allreduce), y1 = A11^{-1} rp1 */
   if (b200k_gather_idx(k, s->d_lm_idx, d_rhs, rp1) || b200_comm_allreduce_sum(rp1, (size_t)k) ||
       b200k_gemv('N', k, k, 1.0, s->d_Linv, rp1, 0.0, t) || b200k_gemv('T', k, k, 1.0, s->d_Linv, t, 0.0, y1))
      return -1;
   if (n2 > 0)
   {
      /* rp2 = r(perm2) - K12' y1 on the local Schur range; assemble the full vector */
      if (b200k_gather_idx(n2l, s->d_sc_idx, d_rhs, rp2l) ||
          (n2l > 0 && afn_k12_gemv(s, 'T', n2l, plo, -1.0, y1, 1.0, rp2l)) || b200k_fill(n2, 0.0, full) ||
          (n2l > 0 && b200_d2d(full + plo, rp2l, sizeof(double) * (size_t)n2l)) || b200_comm_allreduce_sum(full, (size_t)n2))
         return -1;
      /* y2 = G' (G rp2): local rows of G, assemble, local rows of G' */
      if (b200k_fill(n2, 0.0, wfull) || (n2l > 0 && b200k_csr_mv(n2l, G->d_ia + plo, G->d_ja, G->d_a, full, wfull + plo)) ||
          b200_comm_allreduce_sum(wfull, (size_t)n2) ||
          (n2l > 0 && b200k_csr_mv(n2l, G->d_iat + plo, G->d_jat, G->d_at, wfull, y2l)))
         return -1;
      /* rp1 -= K12 y2 (local columns, allreduce), y1 = A11^{-1} rp1 */
      if (b200k_fill(k, 0.0, u) || (n2l > 0 && afn_k12_gemv(s, 'N', n2l, plo, 1.0, y2l, 0.0, u)) ||
          b200_comm_allreduce_sum(u, (size_t)k) || b200k_axpy(k, -1.0, u, rp1) || b200k_gemv('N', k, k, 1.0, s->d_Linv, rp1, 0.0, t) ||
          b200k_gemv('T', k, k, 1.0, s->d_Linv, t, 0.0, y1))
         return -1;
   }
   /* x(perm) = [y1; y2] restricted to the block */
   if (b200k_scatter_idx(k, s->d_lm_idx, y1, d_x) || (n2 > 0 && b200k_scatter_idx(n2l, s->d_sc_idx, y2l, d_x)))
      return -1;
   return 0;
}

int b200_afn_solve_dev(void *va, int n, double *d_x, const double *d_rhs)
{
   b200_afn *s = (b200_afn *)va;
   int k, n2;
   double *rp, *rp2, *y, *y2, *t;
   if (!s || s->magic != B200_MAGIC_AFN)
      return -1;
   if (s->nys) /* rank < max rank: Nystrom on the FPS ordering (afn_setup.m:80-100, afn_solve.m:41-44) */
      return b200_nys_solve_dev(s->nys, n, d_x, d_rhs);
   if (s->d_Linv && n != s->n && b200_comm_size() > 1)
   {
      if (!s->d_K12 && !s->d_K12f)
      {
         printf("Error: the sharded AFN apply needs the dense K12 block (released by Nfft4GPB200PrecondAFNSetK12Operator).\n");
         return -1;
      }
      return afn_solve_sharded(s, n, d_x, d_rhs);
   }
   if (!s->d_Linv || n != s->n)
   {
      printf("Error: AFN preconditioner is not set up for n = %d.\n", n);
      return -1;
   }
   k = s->k;
   n2 = n - k;
   rp = s->d_work;
   rp2 = rp + k;
   y = s->d_work + n;
   y2 = y + k;
   t = s->d_work + 2 * (size_t)n; /* 2k scratch */
   /* [rp; rp2] = rhs(perm) */
   if (b200k_gather_perm(n, s->d_perm, d_rhs, rp))
      return -1;
   /* y = A11 \ rp = L^{-T} (L^{-1} rp) */
   if (b200k_gemv('N', k, k, 1.0, s->d_Linv, rp, 0.0, t) || b200k_gemv('T', k, k, 1.0, s->d_Linv, t, 0.0, y))
      return -1;
   if (n2 > 0 && s->k12_op)
   {
      /* The same four steps with the K12 products done by the NFFT operator: K21 y = (K z)(perm2) for z = y scattered to the landmark
       * positions (z vanishes on the other rows, so the noise term of the operator does not reach them), and K12 y2 = (K z)(perm1)
       * for z = y2 scattered to the non-landmark positions.  Two fast matvecs (O(n)) replace two passes over the k x (n - k) block. */
      b200_additive *op = (b200_additive *)s->k12_op;
      double *z = s->d_opz, *tt = s->d_opt;
      if (b200k_fill(n, 0.0, z) || b200k_scatter_perm(k, s->d_perm, y, z) || b200_additive_matvec_dev(op, n, 1.0, z, 0.0, tt) ||
          b200k_gather_perm(n2, s->d_perm + k, tt, z) || b200k_axpy(n2, -1.0, z, rp2) || b200_fsai_solve_dev(&s->schur, n2, y2, rp2) ||
          b200k_fill(n, 0.0, z) || b200k_scatter_perm(n2, s->d_perm + k, y2, z) || b200_additive_matvec_dev(op, n, 1.0, z, 0.0, tt) ||
          b200k_gather_perm(k, s->d_perm, tt, z) || b200k_axpy(k, -1.0, z, rp) || b200k_gemv('N', k, k, 1.0, s->d_Linv, rp, 0.0, t) ||
          b200k_gemv('T', k, k, 1.0, s->d_Linv, t, 0.0, y))
         return -1;
   }
   else if (n2 > 0)
   {
      /* rp2 -= A21 y ; y2 = S^{-1} rp2 ; rp -= A12 y2 ; y = A11 \ rp */
      if (afn_k12_gemv(s, 'T', n2, 0, -1.0, y, 1.0, rp2) || b200_fsai_solve_dev(&s->schur, n2, y2, rp2) ||
          afn_k12_gemv(s, 'N', n2, 0, -1.0, y2, 1.0, rp) || b200k_gemv('N', k, k, 1.0, s->d_Linv, rp, 0.0, t) ||
          b200k_gemv('T', k, k, 1.0, s->d_Linv, t, 0.0, y))
         return -1;
   }
   return b200k_scatter_perm(n, s->d_perm, y, d_x);
}

/* afn.h:99; MATLAB afn_logdet.m: 2 (sum log diag L + sum log 1 / diag G).  The landmark factor is kept as its explicit inverse. */
double Nfft4GPPrecondAFNLogdet(void *vafn_mat)
{
   b200_afn *s = (b200_afn *)vafn_mat;
   double val = 0.0, *h;
   int i, k;
   if (!s || s->magic != B200_MAGIC_AFN)
      return 0.0;
   if (s->nys)
      return Nfft4GPPrecondNysLogdet(s->nys);
   if (!s->d_Linv)
      return 0.0;
   k = s->k;
   h = (double *)b200_xmalloc(sizeof(double) * (size_t)k * k);
   if (!b200_d2h(h, s->d_Linv, sizeof(double) * (size_t)k * k))
      for (i = 0; i < k; i++)
         val -= 2.0 * log(h[(size_t)i * k + i]); /* diag(L^{-1}) = 1 / diag(L) */
   free(h);
   if (s->n > k)
      val += Nfft4GPPrecondFsaiLogdet(&s->schur);
   return val;
}

/* afn.h:91; MATLAB afn_trace.m */
int Nfft4GPPrecondAFNTrace(void *vafn_mat, double **tracesp)
{
   b200_afn *s = (b200_afn *)vafn_mat;
   if (!s || s->magic != B200_MAGIC_AFN || !tracesp)
      return -1;
   if (s->nys)
      return Nfft4GPPrecondNysTrace(s->nys, tracesp);
   printf("Error: Nfft4GPPrecondAFNTrace: the derivative of the Schur-complement FSAI factor is not built (block form of the AFN "
          "preconditioner); only the Nystrom form carries its gradient side.\n");
   return -1;
}

int Nfft4GPPrecondAFNSolve(void *vafn_mat, int n, double *x, double *rhs)
{
   b200_afn *s = (b200_afn *)vafn_mat;
   if (!s || s->magic != B200_MAGIC_AFN || stage_vectors(&s->d_x, &s->d_rhs, &s->cap, n))
      return -1;
   if (b200_h2d(s->d_rhs, rhs, sizeof(double) * (size_t)n) || b200_afn_solve_dev(s, n, s->d_x, s->d_rhs))
      return -1;
   return b200_d2h(x, s->d_x, sizeof(double) * (size_t)n);
}

/* ================================================================== plug-point lookup for the device PCG */
static int additive_mv_dev(void *mat, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   return b200_additive_matvec_dev((b200_additive *)mat, n, alpha, d_x, beta, d_y);
}

static int window_mv_dev(void *mat, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   return b200_window_matvec_dev((b200_window *)mat, n, alpha, d_x, beta, d_y);
}

b200_dev_matvec b200_lookup_dev_matvec(void *host_fn, void *mat_data, void **dev_handle)
{
   if (host_fn == (void *)Nfft4GPAdditiveNFFTMatSymv)
   {
      b200_additive *A = b200_additive_from_handle(mat_data);
      if (A)
      {
         *dev_handle = A;
         return additive_mv_dev;
      }
   }
   if (host_fn == (void *)Nfft4GPNFFTMatSymv)
   {
      b200_window *w = (b200_window *)mat_data;
      if (w && w->magic == B200_MAGIC_WINDOW)
      {
         *dev_handle = w;
         return window_mv_dev;
      }
   }
   return NULL;
}

/* 1 when the device function is one of this library's own operators (no host callbacks inside: may be stream-captured) */
int b200_dev_fn_is_native(void *fn)
{
   return fn == (void *)additive_mv_dev || fn == (void *)window_mv_dev || fn == (void *)b200_nys_solve_dev ||
          fn == (void *)b200_fsai_solve_dev || fn == (void *)b200_afn_solve_dev || fn == (void *)b200_chol_solve_dev;
}

/* A solver was handed one of THIS library's preconditioner handles together with a function pointer that is none of this library's
 * entry points: in a non-PIE executable the address of an imported function is the executable's PLT stub, which differs from the
 * address the library (linked -Bsymbolic) compares with.  The solve still works through the host-callback path (one D2H + H2D per
 * apply); say so once instead of being silently slow. */
void b200_note_foreign_fn_on_native_handle(const void *prec_data)
{
   static int warned = 0;
   int magic;
   if (warned || !prec_data)
      return;
   memcpy(&magic, prec_data, sizeof(int));
   if (magic == B200_MAGIC_NYS || magic == B200_MAGIC_FSAI || magic == B200_MAGIC_AFN)
   {
      warned = 1;
      printf("Warning: nfft4gp_b200: a preconditioner handle of this library is applied through a function pointer that is not the "
             "library's own entry point (non-PIE executable? link with -fPIE / -fPIC so that &Nfft4GPPrecond*Solve is the library's "
             "address): using the host-callback path with one device-host round trip per apply.\n");
   }
}

b200_dev_solve b200_lookup_dev_solve(void *host_fn)
{
   if (host_fn == (void *)Nfft4GPPrecondNysSolve)
      return b200_nys_solve_dev;
   if (host_fn == (void *)Nfft4GPPrecondFsaiSolve)
      return b200_fsai_solve_dev;
   if (host_fn == (void *)Nfft4GPPrecondAFNSolve)
      return b200_afn_solve_dev;
   if (host_fn == (void *)Nfft4GPPrecondCholSolve)
      return b200_chol_solve_dev;
   return NULL;
}

int Nfft4GPB200PrecondSolveDevice(void *prec, func_solve which, int n, double *d_x, const double *d_rhs)
{
   b200_dev_solve f = b200_lookup_dev_solve((void *)which);
   if (!f)
   {
      printf("Error: Nfft4GPB200PrecondSolveDevice needs one of this library's preconditioner solve functions.\n");
      return -1;
   }
   return f(prec, n, d_x, d_rhs);
}
