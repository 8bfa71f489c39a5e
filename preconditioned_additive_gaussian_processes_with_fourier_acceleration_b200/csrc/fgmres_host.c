/* Flexible GMRES with restarts: host control flow (C) of the reference's Nfft4GPSolverFgmres
 * (SRC/solvers/fgmres.c:3-252) -- the solver that the reference's GP loss, prediction and all TESTS drivers call
 * (SRC/optimizer/gp_loss.c:199, gp_predict.c:151) -- with the Krylov bases V, Z and every vector operation in HBM.
 * The control flow is replicated statement by statement, including its quirks:
 *   - after a restart the new residual is scaled by the ESTIMATED residual norm |rs[i]| of the previous cycle, its true norm
 *     is not recomputed (fgmres.c:126-129, 235-242);
 *   - a Givens breakdown (|gam| < eps) leaves through `label` WITHOUT adding the current cycle to x (fgmres.c:176-179);
 *   - the message at the end of every cycle is printed at print_level == 0 (fgmres.c:198-201);
 *   - *prel_res_v has maxits + 1 entries (one entry on the early exits), callee-allocated with malloc.
 * The Arnoldi step follows Nfft4GPModifiedGS (matops.c:274-346, tol_reorth = -1: no re-orthogonalisation) but keeps the
 * projection coefficients on the device: one fused pass per basis vector (k_mgs_stage1), one download of the new
 * Hessenberg column per iteration.  The Givens rotations and the small triangular solve stay on the host as in the
 * reference.  Point-block sharding (one process per GPU): every dot product is completed by an allreduce of the device
 * scalar, exactly like the PCG. */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

static int g_sharded = 0;

static int reduce_scalar(double *d_s)
{
   return g_sharded ? b200_comm_allreduce_sum(d_s, 1) : 0;
}

static int dev_norm2(int n, const double *x, double *d_tmp, double *out)
{
   double h;
   if (b200k_dot(n, x, x, d_tmp) || reduce_scalar(d_tmp) || b200_d2h(&h, d_tmp, sizeof(double)))
      return -1;
   *out = sqrt(h);
   return 0;
}

static int fgmres_device(void *mat, int n, b200_dev_matvec mv, void *prec, b200_dev_solve ps, double *d_x, const double *d_rhs, int kdim,
                         int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter, int print_level)
{
   int iter, i = 0, j, k, rc = -1;
   double normb, normr, tolr, gam, inormr, hii, hii1;
   double *c = NULL, *s = NULL, *rs = NULL, *H = NULL, *rel_res_v = NULL;
   double *V = NULL, *Z = NULL, *d_h = NULL, *v, *z, *w = NULL;
   size_t bytes = sizeof(double) * (size_t)n, ld = (size_t)n;
   const double EPSILON = DBL_EPSILON;

   if (n == 0 && !g_sharded)
   {
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      return 0;
   }
   d_h = (double *)b200_malloc(sizeof(double) * ((size_t)kdim + 4)); /* Hessenberg column + scratch scalars, device */
   if (!d_h || dev_norm2(n, d_rhs, d_h, &normb))
      goto done;
   if (normb < EPSILON)
   {
      b200k_fill(n, 0.0, d_x);
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      rc = b200_sync();
      goto done;
   }
   V = (double *)b200_malloc(bytes * ((size_t)kdim + 1));
   Z = (double *)b200_malloc(bytes * ((size_t)kdim + 1));
   H = (double *)b200_xcalloc((size_t)kdim * ((size_t)kdim + 1), sizeof(double));
   c = (double *)b200_xcalloc((size_t)kdim, sizeof(double));
   s = (double *)b200_xcalloc((size_t)kdim, sizeof(double));
   rs = (double *)b200_xcalloc((size_t)kdim + 1, sizeof(double));
   if (!V || !Z)
   {
      printf("Error: FGMRES could not allocate the Krylov bases (2 x %d x %d doubles) on the device.\n", n, kdim + 1);
      goto done;
   }
   v = V;
   if (b200_d2d(v, d_rhs, bytes) || mv(mat, n, -1.0, d_x, 1.0, v) || dev_norm2(n, v, d_h, &normr))
      goto done;
   if (normr < EPSILON)
   {
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      rc = 0;
      goto done;
   }
   tolr = atol ? tol : tol * normb;
   rel_res_v = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
   rel_res_v[0] = normr / normb;
   iter = 0;
   if (print_level > 0)
   {
      printf("--------------------------------------------------------------------------------\n");
      printf("Start FlexGMRES(%d)\n", kdim);
      printf("Residual Tol: %e\nMax number of inner iterations: %d\n", tolr, maxits);
      printf("--------------------------------------------------------------------------------\n");
      printf("Step    Residual norm  Relative res.  Convergence Rate\n");
      printf("%5d   %8e   %8e   N/A\n", 0, normr, rel_res_v[0]);
   }
   while (iter < maxits)
   {
      rs[0] = normr;
      inormr = 1.0 / normr;
      if (b200k_scale(n, inormr, v))
         goto done;
      i = 0;
      while (i < kdim && iter < maxits)
      {
         double *Hcol;
         i++;
         iter++;
         v = V + (size_t)(i - 1) * ld;
         z = Z + (size_t)(i - 1) * ld;
         w = V + (size_t)i * ld;
         if (prec)
         {
            if (ps(prec, n, z, v) || mv(mat, n, 1.0, z, 0.0, w))
               goto done;
         }
         else
         {
            if (b200_d2d(z, v, bytes) || mv(mat, n, 1.0, v, 0.0, w))
               goto done;
         }
         /* modified Gram-Schmidt against V(:, 0:i) (matops.c:301-311), coefficients stay on the device: d_h[jj] = w . v_jj */
         for (j = 0; j < i; j++)
         {
            const double *vprev = j ? V + (size_t)(j - 1) * ld : NULL;
            if (b200k_mgs_step(n, w, vprev, j ? d_h + (j - 1) : NULL, V + (size_t)j * ld, d_h + j) || reduce_scalar(d_h + j))
               goto done;
         }
         /* last update fused with ||w||^2 (matops.c:314), then t = ||w||, w /= t (matops.c:336-341) */
         if (b200k_mgs_step(n, w, V + (size_t)(i - 1) * ld, d_h + (i - 1), NULL, d_h + kdim + 1) || reduce_scalar(d_h + kdim + 1) ||
             b200k_scale_by_inv_norm(n, w, d_h + kdim + 1, d_h + i))
            goto done;
         Hcol = H + (size_t)(i - 1) * ((size_t)kdim + 1);
         if (b200_d2h(Hcol, d_h, sizeof(double) * ((size_t)i + 1)))
            goto done;
         /* least-squares problem of H (fgmres.c:163-185) */
         for (j = 1; j < i; j++)
         {
            hii = Hcol[j - 1];
            Hcol[j - 1] = c[j - 1] * hii + s[j - 1] * Hcol[j];
            Hcol[j] = -s[j - 1] * hii + c[j - 1] * Hcol[j];
         }
         hii = Hcol[i - 1];
         hii1 = Hcol[i];
         gam = sqrt(hii * hii + hii1 * hii1);
         if (fabs(gam) < EPSILON)
            goto label;
         c[i - 1] = hii / gam;
         s[i - 1] = hii1 / gam;
         rs[i] = -s[i - 1] * rs[i - 1];
         rs[i - 1] = c[i - 1] * rs[i - 1];
         Hcol[i - 1] = c[i - 1] * hii + s[i - 1] * hii1;
         normr = fabs(rs[i]);
         rel_res_v[iter] = normr / normb;
         if (print_level > 0)
            printf("%5d   %8e   %8e   %8.6f\n", iter, normr, rel_res_v[iter], rel_res_v[iter] / rel_res_v[iter - 1]);
         if (normr <= tolr)
            break;
      }
      if (print_level == 0)
         printf("Rel. residual at the end of current cycle (# of steps per cycle/total its: %d/%d): %e \n", kdim, iter, rel_res_v[iter]);
      /* upper triangular solve (fgmres.c:203-212) */
      rs[i - 1] /= H[(size_t)(i - 1) * (kdim + 1) + i - 1];
      for (k = i - 2; k >= 0; k--)
      {
         for (j = k + 1; j < i; j++)
            rs[k] -= H[(size_t)j * (kdim + 1) + k] * rs[j];
         rs[k] /= H[(size_t)k * (kdim + 1) + k];
      }
      /* x += Z(:, 0:i) rs (fgmres.c:215-219) */
      if (b200_h2d(d_h, rs, sizeof(double) * (size_t)i) || b200k_multi_axpy(n, i, Z, ld, d_h, d_x))
         goto done;
      if (normr <= tolr)
      {
         *prel_res = normr;
         break;
      }
      /* restart (fgmres.c:229-235): v = rhs - A x, w as scratch */
      v = V;
      if (b200_d2d(v, d_rhs, bytes) || mv(mat, n, 1.0, d_x, 0.0, w) || b200k_axpy(n, -1.0, w, v))
         goto done;
   }
label:
   *prel_res = normr / normb;
   *piter = iter;
   *prel_res_v = rel_res_v;
   rel_res_v = NULL;
   rc = 0;
done:
   b200_sync();
   b200_free(V);
   b200_free(Z);
   b200_free(d_h);
   free(H);
   free(c);
   free(s);
   free(rs);
   if (rc)
      free(rel_res_v);
   return rc;
}

/* foreign plug points: see pcg_host.c -- the iteration stays on the GPU, the callback gets host copies */
typedef struct
{
   func_symmatvec mv;
   func_solve ps;
   void *data;
   double *hx, *hy;
} fg_foreign;

static int fg_foreign_mv(void *vctx, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   fg_foreign *c = (fg_foreign *)vctx;
   size_t bytes = sizeof(double) * (size_t)n;
   if (b200_d2h(c->hx, d_x, bytes) || (beta != 0.0 && b200_d2h(c->hy, d_y, bytes)) || c->mv(c->data, n, alpha, c->hx, beta, c->hy))
      return -1;
   return b200_h2d(d_y, c->hy, bytes) || b200_sync();
}

static int fg_foreign_ps(void *vctx, int n, double *d_x, const double *d_rhs)
{
   fg_foreign *c = (fg_foreign *)vctx;
   size_t bytes = sizeof(double) * (size_t)n;
   if (b200_d2h(c->hx, d_rhs, bytes) || c->ps(c->data, n, c->hy, c->hx))
      return -1;
   return b200_h2d(d_x, c->hy, bytes) || b200_sync();
}

/* SRC/solvers/fgmres.c:3-252 (INC/_solvers.h:137) */
int Nfft4GPSolverFgmres(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc, double *x, double *rhs,
                        int kdim, int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter, int print_level)
{
   void *dev_mat = NULL;
   b200_dev_matvec mv = b200_lookup_dev_matvec((void *)matvec, mat_data, &dev_mat);
   b200_dev_solve ps = prec_data ? b200_lookup_dev_solve((void *)precondfunc) : NULL;
   void *dev_prec = prec_data;
   size_t bytes = sizeof(double) * (size_t)n;
   fg_foreign fmv, fps;
   double *d_x, *d_rhs;
   int rc = -1, lo = 0, hi = n;
   b200_additive *A;
   memset(&fmv, 0, sizeof(fmv));
   memset(&fps, 0, sizeof(fps));
   if (b200_init(-1))
      return -1;
   if (kdim < 1)
      kdim = 1;
   if (!mv)
   {
      fmv.mv = matvec;
      fmv.data = mat_data;
      fmv.hx = (double *)b200_host_alloc(bytes);
      fmv.hy = (double *)b200_host_alloc(bytes);
      mv = fg_foreign_mv;
      dev_mat = &fmv;
   }
   if (prec_data && !ps)
   {
      b200_note_foreign_fn_on_native_handle(prec_data);
      fps.ps = precondfunc;
      fps.data = prec_data;
      fps.hx = (double *)b200_host_alloc(bytes);
      fps.hy = (double *)b200_host_alloc(bytes);
      ps = fg_foreign_ps;
      dev_prec = &fps;
   }
   A = (matvec == Nfft4GPAdditiveNFFTMatSymv) ? b200_additive_from_handle(mat_data) : NULL;
   if (A && A->shard_mode == 2 && b200_comm_size() > 1)
   {
      if (fmv.mv || fps.ps)
      {
         printf("Error: point-block-sharded FGMRES needs this library's own matvec and preconditioner.\n");
         goto out;
      }
      lo = A->blk_lo;
      hi = A->blk_hi;
      g_sharded = 1;
   }
   bytes = sizeof(double) * (size_t)(hi - lo);
   d_x = (double *)b200_malloc(bytes);
   d_rhs = (double *)b200_malloc(bytes);
   if (d_x && d_rhs && !b200_h2d(d_x, x + lo, bytes) && !b200_h2d(d_rhs, rhs + lo, bytes))
      rc = fgmres_device(dev_mat, hi - lo, mv, dev_prec, ps, d_x, d_rhs, kdim, maxits, atol, tol, prel_res, prel_res_v, piter, print_level);
   if (!rc)
      rc = b200_d2h(x + lo, d_x, bytes);
   g_sharded = 0;
   b200_free(d_x);
   b200_free(d_rhs);
out:
   b200_host_free(fmv.hx);
   b200_host_free(fmv.hy);
   b200_host_free(fps.hx);
   b200_host_free(fps.hy);
   return rc;
}

int Nfft4GPB200SolverFgmresDevice(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc, double *d_x,
                                  const double *d_rhs, int kdim, int maxits, int atol, double tol, double *prel_res, double **prel_res_v,
                                  int *piter, int print_level)
{
   void *dev_mat = NULL;
   b200_dev_matvec mv = b200_lookup_dev_matvec((void *)matvec, mat_data, &dev_mat);
   b200_dev_solve ps = prec_data ? b200_lookup_dev_solve((void *)precondfunc) : NULL;
   b200_additive *A = (matvec == Nfft4GPAdditiveNFFTMatSymv) ? b200_additive_from_handle(mat_data) : NULL;
   int rc;
   if (!mv || (prec_data && !ps))
   {
      printf("Error: Nfft4GPB200SolverFgmresDevice needs this library's own matvec / preconditioner entry points.\n");
      return -1;
   }
   g_sharded = (A && A->shard_mode == 2 && b200_comm_size() > 1);
   rc = fgmres_device(dev_mat, n, mv, prec_data, ps, d_x, d_rhs, kdim < 1 ? 1 : kdim, maxits, atol, tol, prel_res, prel_res_v, piter, print_level);
   g_sharded = 0;
   return rc;
}
