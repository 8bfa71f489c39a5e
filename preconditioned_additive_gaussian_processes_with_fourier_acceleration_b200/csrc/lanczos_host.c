/* Preconditioned Lanczos solver and stochastic Lanczos quadrature of log det / its gradient: host control flow (C) of the
 * reference's Nfft4GPSolverLanczos (SRC/solvers/lanczos.c:3-419) and Nfft4GPLanczosQuadratureLogdet (:421-610) -- the two
 * routines behind the log-determinant term of the GP loss (SRC/optimizer/gp_loss.c:240-255) -- with the Lanczos bases V, Z and
 * every vector operation in HBM.  Replicated statement by statement, including:
 *   - the orthogonalisation Nfft4GPModifiedGS2 (matops.c:348-434): coefficients <z, v_i>, updates with z_i, against the FIRST
 *     min(iter-1, wsize) + 1 basis vectors, re-orthogonalisation while ||z|| drops below 0.7071 of its previous value;
 *   - the incremental Cholesky factor of T for the residual estimate, the final two bidiagonal solves for x (:271-291);
 *   - the second loop that keeps extending T after convergence, with rand()-filled restart vectors on breakdown (:305-394);
 *   - the end-of-iteration message printed at print_level == 0 (:263-266).
 * The projection coefficients of one orthogonalisation pass stay on the device (vec_kernels.cu:k_mgs_stage1); the host sees one
 * small download per pass.  The eigen-decomposition of T (dstev in the reference) runs on the host (dense_small.c).
 * Single GPU: under point-block sharding these two routines refuse to run (the solvers that are sharded are PCG and FGMRES). */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

static int lz_dot(int n, const double *x, const double *y, double *d_tmp, double *out)
{
   return (b200k_dot(n, x, y, d_tmp) || b200_d2h(out, d_tmp, sizeof(double))) ? -1 : 0;
}

/* Nfft4GPModifiedGS2 (matops.c:348-434) on device vectors; d_h: k + 3 doubles of device scratch, h: host mirror */
static int lz_mgs2(double *z, int n, const double *V, const double *Z, double *TD, double *TE, double *t, int k, double tol_orth,
                   double tol_reorth, double *d_h, double *h)
{
   size_t ld = (size_t)n;
   double normw = 0.0, nrm2;
   int i, first = 1;
   if (k < 0 || n == 0)
      return 0;
   if (tol_reorth > 0.0)
   {
      if (lz_dot(n, z, z, d_h, &nrm2))
         return -1;
      normw = sqrt(nrm2);
   }
   for (;;)
   {
      for (i = 0; i <= k; i++)
         if (b200k_mgs_step(n, z, i ? Z + (size_t)(i - 1) * ld : NULL, i ? d_h + (i - 1) : NULL, V + (size_t)i * ld, d_h + i))
            return -1;
      if (b200k_mgs_step(n, z, Z + (size_t)k * ld, d_h + k, NULL, d_h + k + 1) || b200_d2h(h, d_h, sizeof(double) * ((size_t)k + 2)))
         return -1;
      if (k >= 1)
         TE[0] = first ? h[k - 1] : TE[0] + h[k - 1];
      TD[0] = first ? h[k] : TD[0] + h[k];
      *t = sqrt(h[k + 1]);
      first = 0;
      if (!(*t < normw * tol_reorth && *t >= tol_orth))
         break;
      normw = *t;
   }
   return 0;
}

#include <time.h>
static double lz_now(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static int lz_timing(void)
{
   static int on = -1;
   if (on < 0)
      on = getenv("NFFT4GP_B200_SLQ_TIMING") ? 1 : 0; /* development aid: wall-clock split of a probe, printed per probe */
   return on;
}

static double *g_lz_bases[2] = {NULL, NULL};
static size_t g_lz_bases_bytes = 0;
static int g_lz_keep = 0;
static void lz_release_bases(void)
{
   b200_free(g_lz_bases[0]);
   b200_free(g_lz_bases[1]);
   g_lz_bases[0] = g_lz_bases[1] = NULL;
   g_lz_bases_bytes = 0;
}

static int lz_apply_prec(void *prec, b200_dev_solve ps, int n, double *v, const double *z)
{
   if (prec)
      return ps(prec, n, v, z);
   return b200_d2d(v, z, sizeof(double) * (size_t)n);
}

static int lanczos_device(void *mat, int n, b200_dev_matvec mv, void *prec, b200_dev_solve ps, double *d_x, const double *d_rhs, int wsize,
                          int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter, int *tsize, double **TDp,
                          double **TEp, int print_level)
{
   const double EPSILON = DBL_EPSILON;
   int iter, chol_size, k, wsizei, rc = -1, own_td = 0, own_te = 0;
   double normb, normr, normz, beta, le, ls = 0.0, tolr, t = 1.0, dotvz = 0.0, idotvz, td, te, tmp;
   double *V = NULL, *Z = NULL, *TD = NULL, *TE = NULL, *TLD = NULL, *TLE = NULL, *y = NULL, *rel_res_v = NULL, *d_h = NULL, *h = NULL, *hrand = NULL;
   double *v, *z, *w;
   double tl0 = 0.0, tl1 = 0.0, tl2 = 0.0;
   size_t bytes = sizeof(double) * (size_t)n, ld = (size_t)n;

   if (wsize <= 0)
      wsize = maxits;
   if (wsize > maxits)
      wsize = maxits;
   if (n == 0)
   {
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      return 0;
   }
   d_h = (double *)b200_malloc(sizeof(double) * ((size_t)maxits + 8));
   h = (double *)b200_xcalloc((size_t)maxits + 8, sizeof(double));
   if (!d_h || lz_dot(n, d_rhs, d_rhs, d_h, &tmp))
      goto done;
   normb = sqrt(tmp);
   if (normb < EPSILON)
   {
      b200k_fill(n, 0.0, d_x);
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      rc = b200_sync();
      goto done;
   }
   /* the two Krylov bases: (maxits + 1) n doubles each -- 1.3 GB at n = 4 10^6, maxits = 20.  The stochastic quadrature calls this
    * once per probe; allocating and releasing them every time cost more than a probe's kernels (cudaFree of that size), so inside
    * its probe loop the buffers are kept (g_lz_keep) and reused while they are large enough. */
   tl0 = lz_now();
   if (g_lz_keep && g_lz_bases[0] && g_lz_bases_bytes >= bytes * ((size_t)maxits + 1))
   {
      V = g_lz_bases[0];
      Z = g_lz_bases[1];
      g_lz_bases[0] = g_lz_bases[1] = NULL;
   }
   else
   {
      lz_release_bases();
      V = (double *)b200_malloc(bytes * ((size_t)maxits + 1));
      Z = (double *)b200_malloc(bytes * ((size_t)maxits + 1));
   }
   if (!V || !Z)
   {
      printf("Error: Lanczos could not allocate the bases (2 x %d x %d doubles) on the device.\n", n, maxits + 1);
      goto done;
   }
   TLD = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
   TLE = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
   y = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
   if (*TDp)
      TD = *TDp;
   else
   {
      TD = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
      own_td = 1;
   }
   if (*TEp)
      TE = *TEp;
   else
   {
      TE = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double)); /* the reference allocates maxits entries */
      own_te = 1;
   }
   z = Z;
   v = V;
   tl1 = lz_now();
   if (b200_d2d(z, d_rhs, bytes) || mv(mat, n, -1.0, d_x, 1.0, z) || lz_apply_prec(prec, ps, n, v, z) || lz_dot(n, z, z, d_h, &tmp))
      goto done;
   normr = sqrt(tmp);
   if (lz_dot(n, v, z, d_h, &tmp))
      goto done;
   beta = sqrt(tmp);
   if (beta < EPSILON)
   {
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      if (own_td)
         free(TD);
      if (own_te)
         free(TE);
      TD = TE = NULL;
      rc = 0;
      goto done;
   }
   tolr = atol ? tol / beta : tol;
   rel_res_v = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
   rel_res_v[0] = normr / normb;
   iter = 0;
   chol_size = 0;
   if (print_level > 0)
   {
      printf("--------------------------------------------------------------------------------\n");
      printf("Start Lanczos(%d)\n", maxits);
      printf("Residual Tol: %e\nMax number of inner iterations: %d\n", tolr, maxits);
      printf("--------------------------------------------------------------------------------\n");
      printf("Step    Residual norm  Relative res.  Convergence Rate\n");
      printf("%5d   %8e   %8e   N/A\n", 0, normr, rel_res_v[0]);
   }
   idotvz = 1.0 / beta;
   if (b200k_scale(n, idotvz, v) || b200k_scale(n, idotvz, z))
      goto done;
   while (iter < maxits)
   {
      iter++;
      z = Z + (size_t)iter * ld;
      w = V + (size_t)(iter - 1) * ld;
      v = V + (size_t)iter * ld;
      if (mv(mat, n, 1.0, w, 0.0, z))
         goto done;
      wsizei = iter - 1 < wsize ? iter - 1 : wsize;
      if (lz_mgs2(z, n, V, Z, TD + iter - 1, TE + iter - 2, &t, wsizei, EPSILON, 0.7071, d_h, h))
         goto done;
      if (t < EPSILON)
         break;
      if (lz_apply_prec(prec, ps, n, v, z) || lz_dot(n, v, z, d_h, &tmp))
         goto done;
      dotvz = sqrt(tmp);
      if (dotvz < EPSILON)
         break;
      idotvz = 1.0 / dotvz;
      if (b200k_scale(n, idotvz, v) || b200k_scale(n, idotvz, z) || lz_dot(n, z, z, d_h, &tmp))
         goto done;
      normz = sqrt(tmp);
      if (iter != 1)
      {
         TLE[iter - 2] = TE[iter - 2] / TLD[iter - 2];
         TLD[iter - 1] = sqrt(TD[iter - 1] - TLE[iter - 2] * TLE[iter - 2]);
         le = 1.0 / TLD[iter - 1];
         ls = -ls * TLE[iter - 2] * le;
         normr = fabs(le * ls) * dotvz * beta * normz;
      }
      else
      {
         TLD[0] = sqrt(TD[0]);
         ls = 1.0 / TLD[0];
         normr = dotvz / TD[0] * beta * normz;
      }
      chol_size++;
      rel_res_v[iter] = normr / normb;
      if (print_level > 0)
         printf("%5d   %8e   %8e   %8.6f\n", iter, normr, rel_res_v[iter], rel_res_v[iter] / rel_res_v[iter - 1]);
      if (normr <= tolr)
         break;
   }
   if (print_level == 0)
      printf("Rel. residual at the end of the iteration (# of its: %d): %e \n", iter, rel_res_v[iter]);
   /* the two bidiagonal solves with the Cholesky factor of T (lanczos.c:271-283), then x += V y */
   if (chol_size > 0)
   {
      y[0] = beta / TLD[0];
      for (k = 1; k < chol_size; k++)
         y[k] = (-y[k - 1] * TLE[k - 1]) / TLD[k];
      y[chol_size - 1] /= TLD[chol_size - 1];
      for (k = chol_size - 2; k >= 0; k--)
         y[k] = (y[k] - TLE[k] * y[k + 1]) / TLD[k];
      if (b200_h2d(d_h, y, sizeof(double) * (size_t)chol_size) || b200k_multi_axpy(n, chol_size, V, ld, d_h, d_x))
         goto done;
   }
   *prel_res = normr / normb;
   *piter = iter;
   /* second loop: keep extending T up to wsize rows (lanczos.c:305-394) */
   while (iter < wsize)
   {
      if (t < EPSILON || dotvz < EPSILON)
      {
         int i;
         z = Z + (size_t)iter * ld;
         v = V + (size_t)iter * ld;
         if (!hrand)
            hrand = (double *)b200_xmalloc(bytes);
         for (i = 0; i < n; i++)
            hrand[i] = (double)rand() / (double)RAND_MAX; /* Nfft4GPVecRand, vecops.c:15-25 */
         if (b200_h2d(z, hrand, bytes) || lz_mgs2(z, n, V, Z, &td, &te, &t, iter - 1, EPSILON, 0.7071, d_h, h))
            goto done;
         if (t < EPSILON)
            break;
         if (lz_apply_prec(prec, ps, n, v, z) || lz_dot(n, v, z, d_h, &tmp))
            goto done;
         dotvz = sqrt(tmp);
         if (dotvz < EPSILON)
            break;
         idotvz = 1.0 / dotvz;
         if (b200k_scale(n, idotvz, v) || b200k_scale(n, idotvz, z))
            goto done;
      }
      while (iter < wsize)
      {
         iter++;
         z = Z + (size_t)iter * ld;
         w = V + (size_t)(iter - 1) * ld;
         v = V + (size_t)iter * ld;
         if (mv(mat, n, 1.0, w, 0.0, z) || lz_mgs2(z, n, V, Z, TD + iter - 1, TE + iter - 2, &t, iter - 1, EPSILON, 0.7071, d_h, h))
            goto done;
         if (t < EPSILON)
            break;
         if (lz_apply_prec(prec, ps, n, v, z) || lz_dot(n, v, z, d_h, &tmp))
            goto done;
         dotvz = sqrt(tmp);
         if (dotvz < EPSILON)
            break;
         idotvz = 1.0 / dotvz;
         if (b200k_scale(n, idotvz, v) || b200k_scale(n, idotvz, z))
            goto done;
         if (print_level > 0)
            printf("%5d   Building T\n", iter);
      }
   }
   *tsize = iter;
   *TDp = TD;
   *TEp = TE;
   *prel_res_v = rel_res_v;
   rel_res_v = NULL;
   rc = 0;
done:
   b200_sync();
   tl2 = lz_now();
   if (lz_timing())
      printf("[slq timing] lanczos_device: bases %.1f ms, iteration + tail %.1f ms\n", 1e3 * (tl1 - tl0), 1e3 * (tl2 - tl1));
   if (g_lz_keep && V && Z && !g_lz_bases[0])
   {
      g_lz_bases[0] = V;
      g_lz_bases[1] = Z;
      g_lz_bases_bytes = bytes * ((size_t)maxits + 1);
   }
   else
   {
      b200_free(V);
      b200_free(Z);
   }
   b200_free(d_h);
   free(h);
   free(hrand);
   free(TLD);
   free(TLE);
   free(y);
   if (rc)
   {
      free(rel_res_v);
      if (own_td)
         free(TD);
      if (own_te)
         free(TE);
   }
   return rc;
}

/* foreign plug points: the iteration stays on the GPU, the callback gets host copies (as pcg_host.c) */
typedef struct
{
   func_symmatvec mv;
   func_solve ps;
   void *data;
   double *hx, *hy;
} lz_foreign;

static int lz_foreign_mv(void *vctx, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   lz_foreign *c = (lz_foreign *)vctx;
   size_t bytes = sizeof(double) * (size_t)n;
   if (b200_d2h(c->hx, d_x, bytes) || (beta != 0.0 && b200_d2h(c->hy, d_y, bytes)) || c->mv(c->data, n, alpha, c->hx, beta, c->hy))
      return -1;
   return b200_h2d(d_y, c->hy, bytes) || b200_sync();
}

static int lz_foreign_ps(void *vctx, int n, double *d_x, const double *d_rhs)
{
   lz_foreign *c = (lz_foreign *)vctx;
   size_t bytes = sizeof(double) * (size_t)n;
   if (b200_d2h(c->hx, d_rhs, bytes) || c->ps(c->data, n, c->hy, c->hx))
      return -1;
   return b200_h2d(d_x, c->hy, bytes) || b200_sync();
}

/* SRC/solvers/lanczos.c:3-419 (INC/_solvers.h:176) */
int Nfft4GPSolverLanczos(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc, double *x, double *rhs,
                         int wsize, int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter, int *tsize,
                         double **TDp, double **TEp, int print_level)
{
   void *dev_mat = NULL;
   b200_dev_matvec mv = b200_lookup_dev_matvec((void *)matvec, mat_data, &dev_mat);
   b200_dev_solve ps = prec_data ? b200_lookup_dev_solve((void *)precondfunc) : NULL;
   void *dev_prec = prec_data;
   size_t bytes = sizeof(double) * (size_t)n;
   lz_foreign fmv, fps;
   double *d_x, *d_rhs;
   b200_additive *A = (matvec == Nfft4GPAdditiveNFFTMatSymv) ? b200_additive_from_handle(mat_data) : NULL;
   int rc = -1;
   memset(&fmv, 0, sizeof(fmv));
   memset(&fps, 0, sizeof(fps));
   if (b200_init(-1))
      return -1;
   if (A && A->shard_mode == 2 && b200_comm_size() > 1)
   {
      printf("Error: Nfft4GPSolverLanczos runs on one GPU; the point-block-sharded solvers are Nfft4GPSolverPcg and Nfft4GPSolverFgmres.\n");
      return -1;
   }
   if (!mv)
   {
      fmv.mv = matvec;
      fmv.data = mat_data;
      fmv.hx = (double *)b200_host_alloc(bytes);
      fmv.hy = (double *)b200_host_alloc(bytes);
      mv = lz_foreign_mv;
      dev_mat = &fmv;
   }
   if (prec_data && !ps)
   {
      b200_note_foreign_fn_on_native_handle(prec_data);
      fps.ps = precondfunc;
      fps.data = prec_data;
      fps.hx = (double *)b200_host_alloc(bytes);
      fps.hy = (double *)b200_host_alloc(bytes);
      ps = lz_foreign_ps;
      dev_prec = &fps;
   }
   d_x = (double *)b200_malloc(bytes);
   d_rhs = (double *)b200_malloc(bytes);
   if (d_x && d_rhs && !b200_h2d(d_x, x, bytes) && !b200_h2d(d_rhs, rhs, bytes))
      rc = lanczos_device(dev_mat, n, mv, dev_prec, ps, d_x, d_rhs, wsize, maxits, atol, tol, prel_res, prel_res_v, piter, tsize, TDp, TEp,
                          print_level);
   if (!rc)
      rc = b200_d2h(x, d_x, bytes);
   b200_free(d_x);
   b200_free(d_rhs);
   b200_host_free(fmv.hx);
   b200_host_free(fmv.hy);
   b200_host_free(fps.hx);
   b200_host_free(fps.hy);
   return rc;
}

static int g_probe_sharding = 0;
/* multi-GPU stochastic Lanczos quadrature: distribute the probe vectors over the ranks of the communicator (handles replicated) */
int Nfft4GPB200SetProbeSharding(int on)
{
   g_probe_sharding = on ? 1 : 0;
   return 0;
}

/* SRC/solvers/lanczos.c:421-610 (INC/_solvers.h:216): logdet(K) / n and its three derivatives by stochastic Lanczos quadrature */
int Nfft4GPLanczosQuadratureLogdet(void *mat_data, void *dmat_data, int n, func_symmatvec matvec, func_symmatvec dmatvec, void *prec_data,
                                   func_solve precondfunc, func_trace tracefunc, func_logdet logdetfunc, func_dvp dvpfunc, int maxits, int nvecs,
                                   double *radamacher, int print_level, double *logdet, double **dlogdetp)
{
   const double EPSILON = DBL_EPSILON;
   int i, j, rc = -1;
   /* probe sharding (one process per GPU, every rank holds the whole operator): the probe vectors are independent, so rank r runs
    * the probes i = r (mod R) and the four sums are completed by ONE allreduce of four doubles at the end -- no data-path
    * collective inside the Lanczos runs.  Every rank still draws every probe, so the rand() stream (and with it the result) is
    * the one of a single-process run. */
   const int nranks = g_probe_sharding ? b200_comm_size() : 1, me = nranks > 1 ? b200_comm_rank() : 0;
   double val = 0.0, *dval, logdet_precond = 0.0, *traces_precond, *z, *dAz, *px = NULL, *x, *TV, scale;
   /* device-resident probe loop: this library's additive NFFT matvec + gradient matvec, no preconditioner (the configuration of the
    * reference's Table-1 driver and of BASELINE configs[4]).  Per probe only z goes up and four scalars come down; the reference
    * semantics path below moves z, the three gradient blocks and x through host memory and forms the dot products on the host
    * (at n = 4 10^6 that was 55 of the 128 ms per probe). */
   void *fast_mat = NULL;
   b200_dev_matvec fast_mv = prec_data ? NULL : b200_lookup_dev_matvec((void *)matvec, mat_data, &fast_mat);
   b200_additive *fast_A = (fast_mv && dmatvec == Nfft4GPAdditiveNFFTGradMatSymv) ? b200_additive_from_handle(dmat_data) : NULL;
   double *d_z = NULL, *d_xs = NULL, *d_dAz = NULL, *d_s = NULL;
   if (fast_A && (!fast_A->built || fast_A->n != n || (void *)fast_A != fast_mat || (fast_A->shard_mode == 2 && b200_comm_size() > 1)))
      fast_A = NULL;
   dval = *dlogdetp ? *dlogdetp : (double *)b200_xcalloc(3, sizeof(double));
   traces_precond = (double *)b200_xcalloc(3, sizeof(double));
   if (fast_A)
   {
      d_z = (double *)b200_malloc(sizeof(double) * (size_t)n);
      d_xs = (double *)b200_malloc(sizeof(double) * (size_t)n);
      d_dAz = (double *)b200_malloc(sizeof(double) * 3 * (size_t)n);
      d_s = (double *)b200_malloc(sizeof(double) * 8);
      if (!d_z || !d_xs || !d_dAz || !d_s)
         fast_A = NULL;
   }
   if (prec_data)
   {
      tracefunc(prec_data, &traces_precond);
      for (i = 0; i < 3; i++)
         traces_precond[i] /= (double)n;
      logdet_precond = logdetfunc(prec_data) / (double)n;
      px = (double *)b200_xmalloc(sizeof(double) * (size_t)n * 3);
   }
   z = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   x = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   dAz = (double *)b200_xmalloc(sizeof(double) * (size_t)n * 3);
   TV = (double *)b200_xmalloc(sizeof(double) * (size_t)maxits * (size_t)maxits);
   dval[0] = dval[1] = dval[2] = 0.0;
   g_lz_keep = 1; /* the probes reuse the Krylov bases (released at `done`) */
   for (i = 0; i < nvecs; i++)
   {
      double rel_res, *rel_res_v = NULL, *TD = NULL, *TE = NULL;
      int niter, tsize = 0;
      if (radamacher)
         memcpy(z, radamacher + (size_t)i * n, sizeof(double) * (size_t)n);
      else
         for (j = 0; j < n; j++) /* Nfft4GPVecRadamacher, vecops.c:27-45 */
            z[j] = ((double)rand() / (double)RAND_MAX) < 0.5 ? -1.0 : 1.0;
      if (nranks > 1 && i % nranks != me)
         continue;
      if (fast_A)
      {
         double t0 = lz_now(), t1, t2, t3;
         if (b200_h2d(d_z, z, sizeof(double) * (size_t)n) || b200_sync())
            goto done;
         t1 = lz_now();
         if (b200_additive_grad_matvec_dev(fast_A, n, 1.0, d_z, 0.0, d_dAz) || b200k_fill(n, 0.0, d_xs) || (lz_timing() && b200_sync()))
            goto done;
         t2 = lz_now();
         if (lanczos_device(fast_mat, n, fast_mv, NULL, NULL, d_xs, d_z, maxits, maxits, 0, EPSILON, &rel_res, &rel_res_v, &niter, &tsize, &TD, &TE,
                            print_level))
            goto done;
         t3 = lz_now();
         if (lz_timing())
            printf("[slq timing] probe %d: upload %.1f ms, gradient matvec %.1f ms, Lanczos %.1f ms\n", i, 1e3 * (t1 - t0), 1e3 * (t2 - t1),
                   1e3 * (t3 - t2));
      }
      else
      {
         memset(x, 0, sizeof(double) * (size_t)n);
         if (dmatvec(dmat_data, n, 1.0, z, 0.0, dAz))
            goto done;
         if (Nfft4GPSolverLanczos(mat_data, n, matvec, prec_data, precondfunc, x, z, maxits, maxits, 0, EPSILON, &rel_res, &rel_res_v, &niter,
                                  &tsize, &TD, &TE, print_level))
            goto done;
      }
      free(rel_res_v);
      while (tsize > 0 && isnan(TD[tsize - 1]))
         tsize--;
      if (tsize == 0)
      {
         printf("Warning: empty tridiagonal matrix\n");
         free(TD);
         free(TE);
         rc = 0;
         goto done;
      }
      if (b200_stev(tsize, TD, TE, TV))
      {
         printf("Warning: DSTEV failed at iteration %d/%d\n", i, nvecs);
         free(TD);
         free(TE);
         goto done;
      }
      for (j = 0; j < tsize; j++)
         val += TV[(size_t)j * tsize] * TV[(size_t)j * tsize] * log(fabs(TD[j]));
      if (prec_data)
         dvpfunc(prec_data, n, NULL, z, &px);
      for (j = 0; j < 3; j++)
      {
         double s = 0.0;
         int q;
         if (fast_A)
         {
            if (lz_dot(n, d_dAz + (size_t)j * n, d_xs, d_s, &s))
            {
               free(TD);
               free(TE);
               goto done;
            }
            dval[j] += s;
            continue;
         }
         for (q = 0; q < n; q++)
            s += dAz[(size_t)j * n + q] * x[q];
         dval[j] += s;
         if (prec_data)
         {
            s = 0.0;
            for (q = 0; q < n; q++)
               s += px[(size_t)j * n + q] * z[q];
            dval[j] -= s;
         }
      }
      free(TD);
      free(TE);
   }
   if (nranks > 1)
   {
      double h4[4] = {val, dval[0], dval[1], dval[2]}, *d4 = (double *)b200_malloc(sizeof(h4));
      int bad = !d4 || b200_h2d(d4, h4, sizeof(h4)) || b200_comm_allreduce_sum(d4, 4) || b200_d2h(h4, d4, sizeof(h4));
      b200_free(d4);
      if (bad)
         goto done;
      val = h4[0];
      dval[0] = h4[1];
      dval[1] = h4[2];
      dval[2] = h4[3];
   }
   scale = 1.0 / (double)nvecs;
   val *= scale;
   scale /= (double)n;
   for (j = 0; j < 3; j++)
      dval[j] = dval[j] * scale + traces_precond[j];
   val += logdet_precond;
   *logdet = val;
   *dlogdetp = dval;
   rc = 0;
done:
   g_lz_keep = 0;
   lz_release_bases();
   b200_free(d_z);
   b200_free(d_xs);
   b200_free(d_dAz);
   b200_free(d_s);
   free(z);
   free(x);
   free(dAz);
   free(TV);
   free(px);
   free(traces_precond);
   return rc;
}
