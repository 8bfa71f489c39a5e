/* Gradient side of the Nystrom preconditioner (SRC/preconds/nys.c:175-516): log det, derivative-vector products and the
 * traces tr(M^{-1} dM/dtheta) that the GP loss gradient needs when it is preconditioned (TESTS/TEST3/foo3.cpp passes these
 * three functions to Nfft4GPGpProblemSetup).  In the permuted order of the setup
 *      M = K1 K11^{-1} K1' + eta I,   K11 = L L' (shifted),   U = K1 L^{-T},   eta = mu f^2,
 * and the reference differentiates the low-rank part only for theta = f, l (the cross block K1 carries no noise term) and
 * uses dM/dmu = f^2 I:
 *      dM_g = dK1_g K11^{-1} K1' + K1 K11^{-1} dK1_g' - K1 L^{-T} G_g L^{-1} K1',   G_g = L^{-1} dK11_g L^{-T}   (nys.c:246-276)
 *      Dvp:    y_g = M^{-1} dM_g x   (g = f, l),   y_mu = f^2 M^{-1} x                                             (nys.c:279-312)
 *      Trace:  tr(M^{-1} dM_g) = ( tr(dM_g) - tr( (U'U + eta I)^{-1} U' dM_g U ) ) / eta                            (nys.c:332-484)
 *      Logdet: (n - k) log eta + sum_i log(sigma_i^2 + eta)                                                        (nys.c:486-516)
 * The reference evaluates the trace with k derivative-vector products of length n; here everything that is O(n) collapses
 * into two k x k Gram matrices, S = K1'K1 and R = K1' dK1_l (fp64 tensor-core GEMMs), the rest is k x k algebra:
 *      V = L^{-1} S L^{-T} = U'U,  P = S L^{-T},  Q_g = R_g' L^{-T}  (R_f = (2/f) S),
 *      tr(dM_g) = 2 <K11^{-1}, R_g> - <V, G_g>,   U' dM_g U = Q_g' K11^{-1} P + (.)' - P' L^{-T} G_g L^{-1} P.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

static b200_nys *nys_grad_handle(void *v, const char *what)
{
   b200_nys *s = (b200_nys *)v;
   if (!s || s->magic != B200_MAGIC_NYS)
      return NULL;
   if (!s->has_grad)
   {
      printf("Setup NYS without gradient, %s not supported.\n", what);
      return NULL;
   }
   return s;
}

double Nfft4GPPrecondNysLogdet(void *vnys_mat) /* nys.c:486-516; needs no gradient data */
{
   b200_nys *s = (b200_nys *)vnys_mat;
   double *sv, val0, val;
   int i;
   if (!s || s->magic != B200_MAGIC_NYS || !s->d_s)
      return 0.0;
   sv = (double *)b200_xmalloc(sizeof(double) * (size_t)s->k);
   if (b200_d2h(sv, s->d_s, sizeof(double) * (size_t)s->k))
   {
      free(sv);
      return 0.0;
   }
   val0 = log(s->eta);
   val = val0 * (double)(s->n - s->k);
   for (i = 0; i < s->k; i++)
      val += sv[i] > 0.0 ? log(1.0 / sv[i]) : val0;
   free(sv);
   return val;
}

/* yp (device, permuted order) = dM_g xp for g = 0 (f) or 1 (l); xp device, permuted order; work: 8k doubles */
static int nys_dm_apply(b200_nys *s, int g, const double *xp, double *yp, double *wk)
{
   int n = s->n, k = s->k;
   double *a = wk, *b = wk + k, *a1 = wk + 2 * (size_t)k, *b1 = wk + 3 * (size_t)k, *c = wk + 4 * (size_t)k, *a2 = wk + 5 * (size_t)k,
          *bc = wk + 6 * (size_t)k, *t = wk + 7 * (size_t)k;
   /* a = K1' x, b = dK1_g' x */
   if (b200k_gemv('T', n, k, 1.0, s->d_gK1, xp, 0.0, a))
      return -1;
   if (g == 0)
   {
      if (b200k_copy(k, a, b) || b200k_scale(k, 2.0 / s->f, b))
         return -1;
   }
   else if (b200k_gemv('T', n, k, 1.0, s->d_gdK1l, xp, 0.0, b))
      return -1;
   /* a1 = L^{-1} a, b1 = L^{-1} b, c = G_g a1; then a2 = L^{-T} a1, bc = L^{-T} (b1 - c) */
   if (b200k_gemv('N', k, k, 1.0, s->d_gLinv, a, 0.0, a1) || b200k_gemv('N', k, k, 1.0, s->d_gLinv, b, 0.0, b1) ||
       b200k_gemv('N', k, k, 1.0, s->d_gG[g], a1, 0.0, c) || b200k_axpy(k, -1.0, c, b1) ||
       b200k_gemv('T', k, k, 1.0, s->d_gLinv, a1, 0.0, a2) || b200k_gemv('T', k, k, 1.0, s->d_gLinv, b1, 0.0, bc))
      return -1;
   (void)t;
   /* y = dK1_g a2 + K1 bc */
   if (g == 0)
   {
      /* dK1_f = (2/f) K1: y = K1 ((2/f) a2 + bc) */
      if (b200k_axpy(k, 2.0 / s->f, a2, bc) || b200k_gemv('N', n, k, 1.0, s->d_gK1, bc, 0.0, yp))
         return -1;
      return 0;
   }
   return (b200k_gemv('N', n, k, 1.0, s->d_gdK1l, a2, 0.0, yp) || b200k_gemv('N', n, k, 1.0, s->d_gK1, bc, 1.0, yp)) ? -1 : 0;
}

int Nfft4GPPrecondNysDvp(void *vnys_mat, int n, int *mask, double *x, double **yp) /* nys.c:175-330 */
{
   b200_nys *s = nys_grad_handle(vnys_mat, "dvp");
   double *y, *d_x, *d_xp, *d_yp, *d_y, *wk;
   int g, rc = 0;
   if (!s)
      return -1;
   if (!yp)
   {
      printf("output pointer cannot be NULL\n");
      return -1;
   }
   if (n != s->n)
      return -1;
   y = *yp ? *yp : (double *)b200_xcalloc((size_t)3 * n, sizeof(double));
   d_x = s->d_gvec;
   d_xp = d_x + n;
   d_yp = d_xp + n;
   d_y = d_yp + n;
   wk = d_y + n;
   rc = b200_h2d(d_x, x, sizeof(double) * (size_t)n) ||
        (s->d_perm ? b200k_gather_perm(n, s->d_perm, d_x, d_xp) : b200k_copy(n, d_x, d_xp));
   for (g = 0; g < 2 && !rc; g++)
   {
      if (mask && !mask[g])
         continue;
      /* dM_g x in the permuted order, back to the caller's order, then the full (permuting) solve */
      rc = nys_dm_apply(s, g, d_xp, d_yp, wk) ||
           (s->d_perm ? b200k_scatter_perm(n, s->d_perm, d_yp, d_y) : b200k_copy(n, d_yp, d_y)) ||
           b200_nys_solve_dev(s, n, d_yp, d_y) || b200_d2h(y + (size_t)g * n, d_yp, sizeof(double) * (size_t)n);
   }
   if (!rc && (!mask || mask[2]))
      rc = b200_nys_solve_dev(s, n, d_yp, d_x) || b200k_scale(n, s->f * s->f, d_yp) ||
           b200_d2h(y + 2 * (size_t)n, d_yp, sizeof(double) * (size_t)n);
   if (rc)
   {
      if (!*yp)
         free(y);
      return -1;
   }
   *yp = y;
   return 0;
}

static int frob(int count, const double *a, const double *b, double *d_out, double *out)
{
   return b200k_dot(count, a, b, d_out) || b200_d2h(out, d_out, sizeof(double));
}

int Nfft4GPPrecondNysTrace(void *vnys_mat, double **tracesp) /* nys.c:332-484 */
{
   b200_nys *s = nys_grad_handle(vnys_mat, "trace");
   double *tr;
   if (!s)
      return -1;
   if (!tracesp)
   {
      printf("Trace pointer cannot be NULL\n");
      return -1;
   }
   if (!s->traces_ready)
   {
      int n = s->n, k = s->k, kk = k * k, g, info = 0, rc;
      size_t kb = sizeof(double) * (size_t)kk;
      double *S = (double *)b200_malloc(kb), *R = (double *)b200_malloc(kb), *V = (double *)b200_malloc(kb), *W = (double *)b200_malloc(kb),
             *Wi = (double *)b200_malloc(kb), *Winv = (double *)b200_malloc(kb), *Kinv = (double *)b200_malloc(kb),
             *P = (double *)b200_malloc(kb), *Q = (double *)b200_malloc(kb), *T = (double *)b200_malloc(kb), *X = (double *)b200_malloc(kb),
             *Y = (double *)b200_malloc(kb), *d_out = s->d_gvec;
      double wv = 0.0;
      const double *Li = s->d_gLinv;
      rc = !S || !R || !V || !W || !Wi || !Winv || !Kinv || !P || !Q || !T || !X || !Y ||
           /* S = K1'K1, V = Li S Li', W = V + eta I, Winv = W^{-1}, Kinv = Li' Li, P = S Li' */
           b200k_dgemm('T', 'N', k, k, n, 1.0, s->d_gK1, n, s->d_gK1, n, 0.0, S, k) || b200k_dgemm('N', 'N', k, k, k, 1.0, Li, k, S, k, 0.0, X, k) ||
           b200k_dgemm('N', 'T', k, k, k, 1.0, X, k, Li, k, 0.0, V, k) || b200_d2d(W, V, kb) || b200k_add_diag(k, W, k, s->eta) ||
           b200k_potrf_inverse(k, W, Wi, &info) || info || b200k_dgemm('T', 'N', k, k, k, 1.0, Wi, k, Wi, k, 0.0, Winv, k) ||
           b200k_dgemm('T', 'N', k, k, k, 1.0, Li, k, Li, k, 0.0, Kinv, k) || b200k_dgemm('N', 'T', k, k, k, 1.0, S, k, Li, k, 0.0, P, k) ||
           frob(kk, Winv, V, d_out, &wv);
      s->traces[2] = s->f * s->f * ((double)n - wv) / s->eta;
      for (g = 0; g < 2 && !rc; g++)
      {
         double kr = 0.0, vg = 0.0, wt = 0.0;
         /* R_g = K1' dK1_g */
         if (g == 0)
            rc = b200_d2d(R, S, kb) || b200k_scale(kk, 2.0 / s->f, R);
         else
            rc = b200k_dgemm('T', 'N', k, k, n, 1.0, s->d_gK1, n, s->d_gdK1l, n, 0.0, R, k);
         /* Q = R' Li';  X = Q' Kinv P;  Y = Li P;  T = X + X' - Y' G_g Y */
         rc = rc || b200k_dgemm('T', 'T', k, k, k, 1.0, R, k, Li, k, 0.0, Q, k) || b200k_dgemm('T', 'N', k, k, k, 1.0, Q, k, Kinv, k, 0.0, Y, k) ||
              b200k_dgemm('N', 'N', k, k, k, 1.0, Y, k, P, k, 0.0, X, k) || b200k_dgemm('N', 'N', k, k, k, 1.0, Li, k, P, k, 0.0, Y, k) ||
              b200k_dgemm('T', 'N', k, k, k, 1.0, Y, k, s->d_gG[g], k, 0.0, Q, k) || b200k_dgemm('N', 'N', k, k, k, 1.0, Q, k, Y, k, 0.0, T, k) ||
              /* <Winv, X + X'> = 2 <Winv, X> (Winv symmetric) */
              frob(kk, Kinv, R, d_out, &kr) || frob(kk, V, s->d_gG[g], d_out, &vg) || frob(kk, Winv, X, d_out, &wt);
         if (!rc)
         {
            double wt2 = 0.0;
            rc = frob(kk, Winv, T, d_out, &wt2);
            s->traces[g] = ((2.0 * kr - vg) - (2.0 * wt - wt2)) / s->eta;
         }
      }
      b200_free(S);
      b200_free(R);
      b200_free(V);
      b200_free(W);
      b200_free(Wi);
      b200_free(Winv);
      b200_free(Kinv);
      b200_free(P);
      b200_free(Q);
      b200_free(T);
      b200_free(X);
      b200_free(Y);
      if (rc)
         return -1;
      s->traces_ready = 1;
   }
   tr = *tracesp ? *tracesp : (double *)b200_xcalloc(3, sizeof(double));
   tr[0] = s->traces[0];
   tr[1] = s->traces[1];
   tr[2] = s->traces[2];
   *tracesp = tr;
   return 0;
}
