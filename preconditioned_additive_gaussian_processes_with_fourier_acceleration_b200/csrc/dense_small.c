/* Small dense fp64 factorizations on the host (k x k, k = landmark count, a few hundred): the serial glue of the
 * preconditioner setups.  The reference calls LAPACK for these (dpotrf / dtrtri / dsyev: SRC/preconds/chol.c:466,
 * SRC/linearalg/matops.c:40-48, 97); this library carries no BLAS/LAPACK dependency, so they are written out here.
 * Everything O(n k^2) of the setups runs on the GPU (setup_kernels.cu); only O(k^3) work is done here.
 * All matrices are column-major. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "b200_host.h"
#include "b200_precond.h"

/* Frobenius norm of a symmetric matrix of which the lower triangle is stored (dlansy 'F','L'; chol.c:448) */
double b200_sym_fro_lower(int n, const double *A, int lda)
{
   double s = 0.0;
   int i, j;
   for (j = 0; j < n; j++)
   {
      s += A[(size_t)j * lda + j] * A[(size_t)j * lda + j];
      for (i = j + 1; i < n; i++)
         s += 2.0 * A[(size_t)j * lda + i] * A[(size_t)j * lda + i];
   }
   return sqrt(s);
}

/* the stabilising shift of the reference's Cholesky setup: sqrt(n) * ulp(||A||_F)  (chol.c:448-455) */
double b200_chol_shift(int n, double fro) { return sqrt((double)n) * (nextafter(fro, fro + 1.0) - fro); }

/* in-place lower Cholesky A = L L'; the strict upper triangle is left untouched.  Returns 0, or j+1 when the j-th
 * pivot is not positive (dpotrf convention). */
int b200_potrf_lower(int n, double *A, int lda)
{
   int i, j, k;
   for (j = 0; j < n; j++)
   {
      double *cj = A + (size_t)j * lda;
      double d = cj[j];
      /* left-looking: column j -= L(:, 0:j) * L(j, 0:j)' */
      for (k = 0; k < j; k++)
      {
         const double *ck = A + (size_t)k * lda;
         double ljk = ck[j];
         if (ljk != 0.0)
            for (i = j; i < n; i++)
               cj[i] -= ck[i] * ljk;
      }
      d = cj[j];
      if (!(d > 0.0))
         return j + 1;
      d = sqrt(d);
      cj[j] = d;
      for (i = j + 1; i < n; i++)
         cj[i] /= d;
   }
   return 0;
}

/* explicit inverse of the k x k lower Cholesky factor (column-major), upper triangle zero (dtrtri 'L','N') */
void b200_tril_inverse(int k, const double *L, double *Linv)
{
   int i, j, c;
   memset(Linv, 0, sizeof(double) * (size_t)k * (size_t)k);
   for (c = 0; c < k; c++)
   {
      /* solve L x = e_c by forward substitution */
      double *x = Linv + (size_t)c * k;
      x[c] = 1.0 / L[(size_t)c * k + c];
      for (i = c + 1; i < k; i++)
      {
         double s = 0.0;
         for (j = c; j < i; j++)
            s -= L[(size_t)j * k + i] * x[j];
         x[i] = s / L[(size_t)i * k + i];
      }
   }
}

/* ------------------------------------------------------------------ symmetric eigensolver
 * Householder tridiagonalisation followed by the implicit QL iteration with accumulated transformations.
 * In: A symmetric n x n (both triangles).  Out: columns of A = orthonormal eigenvectors, w ascending (dsyev 'V'). */
static void tridiagonalise(int n, double *V, double *d, double *e, int want_vectors)
{
   int i, j, k;
#define VV(r, c) V[(size_t)(c) * n + (r)]
   for (j = 0; j < n; j++)
      d[j] = VV(n - 1, j);
   for (i = n - 1; i > 0; i--)
   {
      double scale = 0.0, h = 0.0;
      for (k = 0; k < i; k++)
         scale += fabs(d[k]);
      if (scale == 0.0)
      {
         e[i] = d[i - 1];
         for (j = 0; j < i; j++)
         {
            d[j] = VV(i - 1, j);
            VV(i, j) = 0.0;
            VV(j, i) = 0.0;
         }
      }
      else
      {
         double f, g, hh;
         for (k = 0; k < i; k++)
         {
            d[k] /= scale;
            h += d[k] * d[k];
         }
         f = d[i - 1];
         g = sqrt(h);
         if (f > 0)
            g = -g;
         e[i] = scale * g;
         h -= f * g;
         d[i - 1] = f - g;
         for (j = 0; j < i; j++)
            e[j] = 0.0;
         for (j = 0; j < i; j++)
         {
            f = d[j];
            VV(j, i) = f;
            g = e[j] + VV(j, j) * f;
            for (k = j + 1; k <= i - 1; k++)
            {
               g += VV(k, j) * d[k];
               e[k] += VV(k, j) * f;
            }
            e[j] = g;
         }
         f = 0.0;
         for (j = 0; j < i; j++)
         {
            e[j] /= h;
            f += e[j] * d[j];
         }
         hh = f / (h + h);
         for (j = 0; j < i; j++)
            e[j] -= hh * d[j];
         for (j = 0; j < i; j++)
         {
            f = d[j];
            g = e[j];
            for (k = j; k <= i - 1; k++)
               VV(k, j) -= (f * e[k] + g * d[k]);
            d[j] = VV(i - 1, j);
            VV(i, j) = 0.0;
         }
      }
      d[i] = h;
   }
   if (!want_vectors)
   {
      for (j = 0; j < n; j++)
         d[j] = VV(j, j);
      e[0] = 0.0;
      return;
   }
   /* accumulate the transformations */
   for (i = 0; i < n - 1; i++)
   {
      double h;
      VV(n - 1, i) = VV(i, i);
      VV(i, i) = 1.0;
      h = d[i + 1];
      if (h != 0.0)
      {
         for (k = 0; k <= i; k++)
            d[k] = VV(k, i + 1) / h;
         for (j = 0; j <= i; j++)
         {
            double g = 0.0;
            for (k = 0; k <= i; k++)
               g += VV(k, i + 1) * VV(k, j);
            for (k = 0; k <= i; k++)
               VV(k, j) -= g * d[k];
         }
      }
      for (k = 0; k <= i; k++)
         VV(k, i + 1) = 0.0;
   }
   for (j = 0; j < n; j++)
   {
      d[j] = VV(n - 1, j);
      VV(n - 1, j) = 0.0;
   }
   VV(n - 1, n - 1) = 1.0;
   e[0] = 0.0;
}

static int ql_implicit(int n, double *V, double *d, double *e, int want_vectors)
{
   int i, j, k, l;
   double f = 0.0, tst1 = 0.0;
   const double eps = 2.220446049250313e-16;
   for (i = 1; i < n; i++)
      e[i - 1] = e[i];
   e[n - 1] = 0.0;
   for (l = 0; l < n; l++)
   {
      int m = l, iter = 0;
      double t = fabs(d[l]) + fabs(e[l]);
      if (t > tst1)
         tst1 = t;
      while (m < n)
      {
         if (fabs(e[m]) <= eps * tst1)
            break;
         m++;
      }
      if (m > l)
      {
         do
         {
            double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = hypot(p, 1.0), dl1, h, c, c2, c3, el1, s, s2;
            if (++iter > 300)
               return -1;
            if (p < 0)
               r = -r;
            d[l] = e[l] / (p + r);
            d[l + 1] = e[l] * (p + r);
            dl1 = d[l + 1];
            h = g - d[l];
            for (i = l + 2; i < n; i++)
               d[i] -= h;
            f += h;
            p = d[m];
            c = 1.0;
            c2 = c;
            c3 = c;
            el1 = e[l + 1];
            s = 0.0;
            s2 = 0.0;
            for (i = m - 1; i >= l; i--)
            {
               c3 = c2;
               c2 = c;
               s2 = s;
               g = c * e[i];
               h = c * p;
               r = hypot(p, e[i]);
               e[i + 1] = s * r;
               s = e[i] / r;
               c = p / r;
               p = c * d[i] - s * g;
               d[i + 1] = h + s * (c * g + s * d[i]);
               if (want_vectors)
               {
                  double *vi = V + (size_t)i * n, *vi1 = V + (size_t)(i + 1) * n;
                  for (k = 0; k < n; k++)
                  {
                     h = vi1[k];
                     vi1[k] = s * vi[k] + c * h;
                     vi[k] = c * vi[k] - s * h;
                  }
               }
            }
            p = -s * s2 * c3 * el1 * e[l] / dl1;
            e[l] = s * p;
            d[l] = c * p;
         } while (fabs(e[l]) > eps * tst1);
      }
      d[l] += f;
      e[l] = 0.0;
   }
   /* ascending order */
   for (i = 0; i < n - 1; i++)
   {
      double p = d[i];
      k = i;
      for (j = i + 1; j < n; j++)
         if (d[j] < p)
         {
            k = j;
            p = d[j];
         }
      if (k != i)
      {
         double *vi = V + (size_t)i * n, *vk = V + (size_t)k * n;
         d[k] = d[i];
         d[i] = p;
         for (j = 0; want_vectors && j < n; j++)
         {
            double tt = vi[j];
            vi[j] = vk[j];
            vk[j] = tt;
         }
      }
   }
   return 0;
#undef VV
}

int b200_syev(int n, double *A, double *w)
{
   double *e;
   int rc;
   if (n <= 0)
      return 0;
   if (n == 1)
   {
      w[0] = A[0];
      A[0] = 1.0;
      return 0;
   }
   e = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   tridiagonalise(n, A, w, e, 1);
   rc = ql_implicit(n, A, w, e, 1);
   free(e);
   return rc;
}

/* eigenvalues only (ascending); A is destroyed */
int b200_syev_values(int n, double *A, double *w)
{
   double *e;
   int rc;
   if (n <= 0)
      return 0;
   if (n == 1)
   {
      w[0] = A[0];
      return 0;
   }
   e = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   tridiagonalise(n, A, w, e, 0);
   rc = ql_implicit(n, A, w, e, 0);
   free(e);
   return rc;
}

/* largest eigenvalue of a symmetric positive semi-definite m x m matrix (full storage) by power iteration;
 * used for the 2-norm of the Nystrom error in the rank estimation (afn_setup.m:163-171). */
double b200_spd_norm2(int m, const double *S, int lds, double *work /* 2m */)
{
   double *v = work, *u = work + m, lam = 0.0, prev = -1.0;
   int i, j, it;
   if (m <= 0)
      return 0.0;
   for (i = 0; i < m; i++)
      v[i] = 1.0 / sqrt((double)m) * (1.0 + 1e-3 * (double)((i * 2654435761u) % 1000u) / 1000.0);
   for (it = 0; it < 300; it++)
   {
      double nu = 0.0;
      for (i = 0; i < m; i++)
         u[i] = 0.0;
      for (j = 0; j < m; j++)
      {
         const double *c = S + (size_t)j * lds;
         double vj = v[j];
         for (i = 0; i < m; i++)
            u[i] += c[i] * vj;
      }
      for (i = 0; i < m; i++)
         nu += u[i] * u[i];
      nu = sqrt(nu);
      if (nu == 0.0)
         return 0.0;
      lam = nu; /* ||S v|| with ||v|| = 1 */
      for (i = 0; i < m; i++)
         v[i] = u[i] / nu;
      if (fabs(lam - prev) <= 1e-7 * lam)
         break;
      prev = lam;
   }
   return lam;
}

/* symmetric tridiagonal eigen-decomposition (dstev 'V'): d[n] diagonal -> ascending eigenvalues, e[n-1] sub-diagonal
 * (destroyed), Zm n x n column-major eigenvectors.  Used by the stochastic Lanczos quadrature (lanczos.c:539). */
int b200_stev(int n, double *d, double *e, double *Zm)
{
   double *e2;
   int i, rc;
   if (n <= 0)
      return 0;
   memset(Zm, 0, sizeof(double) * (size_t)n * n);
   for (i = 0; i < n; i++)
      Zm[(size_t)i * n + i] = 1.0;
   if (n == 1)
      return 0;
   e2 = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   e2[0] = 0.0;
   for (i = 1; i < n; i++)
      e2[i] = e[i - 1]; /* ql_implicit expects the sub-diagonal in e[1..n-1] */
   rc = ql_implicit(n, Zm, d, e2, 1);
   free(e2);
   return rc;
}
