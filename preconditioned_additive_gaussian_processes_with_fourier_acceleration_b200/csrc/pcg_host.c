/* Preconditioned conjugate gradients: host control flow (C) of the reference's Nfft4GPSolverPcg
 * (SRC/solvers/pcg.c:3-206), with every vector operation on the GPU when the two plug points are this
 * library's own device-capable operators.  The control flow is replicated statement by statement,
 * including its quirks:
 *   - *piter is only set on convergence CONFIRMED by the true-residual recheck (pcg.c:181-193); otherwise 0;
 *   - rel_res_v[ii] is overwritten with the ABSOLUTE true residual at the recheck (pcg.c:187);
 *   - maxits is clamped to n (pcg.c:54); early exits for ||b|| < eps (pcg.c:32) and ||r0|| < tolb (pcg.c:70);
 *   - breakdown breaks rho == 0, beta == 0, pq <= 0 (pcg.c:119,136,155).
 * Per iteration the device path runs: precond apply, one dot, one fused p-update, matvec, one dot and one
 * fused x/r update + ||r||^2 -- four n-passes outside matvec/precond instead of the reference's nine. */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

/* device scalars: [0] scratch of dev_dot, [1..2] rho of odd / even iterations, [3] pq, [4] ||r||^2 */
static double *g_scal_dev = NULL;
static double *g_scal_host = NULL; /* pinned mirror */

static int scal_init(void)
{
   if (!g_scal_dev)
   {
      g_scal_dev = (double *)b200_malloc(sizeof(double) * 8);
      g_scal_host = (double *)b200_host_alloc(sizeof(double) * 8);
      if (!g_scal_dev || !g_scal_host)
         return -1;
   }
   return 0;
}

/* point-block sharding (one process per GPU): vectors hold the rank's block, every dot product is completed by an
 * allreduce of the device scalar (SURVEY.md 8(e) iii) */
static int g_pcg_sharded = 0, g_pcg_nglobal = 0;

static int dev_dot(int n, const double *x, const double *y, double *out)
{
   if (b200k_dot(n, x, y, g_scal_dev) || (g_pcg_sharded && b200_comm_allreduce_sum(g_scal_dev, 1)) ||
       b200_d2h(g_scal_host, g_scal_dev, sizeof(double)))
      return -1;
   *out = g_scal_host[0];
   return 0;
}

int b200_pcg_device(void *mat, int n, b200_dev_matvec mv, void *prec, b200_dev_solve ps, double *d_x, const double *d_rhs,
                    int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter, int print_level)
{
   int iter = 0, ii, rc = -1;
   /* no iteration graph whenever a collective can sit inside mv / ps (window sharding allreduces y inside the matvec; a captured
    * NCCL kernel next to the caller's own communicator deadlocked in a 2-GPU run, see nfft_host.c) */
   int use_graphs = !g_pcg_sharded && b200_comm_size() == 1 && !b200_prof_active() && !getenv("NFFT4GP_B200_NO_GRAPH") && b200_dev_fn_is_native((void *)mv) &&
                    (!prec || b200_dev_fn_is_native((void *)ps));
   void *it_graph[2] = {NULL, NULL};
   long long it_launches[2] = {0, 0};
   double rho = 1.0, alpha, beta, normb, normr, normr2, tolb, t;
   double *r = NULL, *z = NULL, *p = NULL, *q = NULL, *rel_res_v = NULL;
   size_t bytes = sizeof(double) * (size_t)n;

   if (scal_init())
      return -1;
   if (dev_dot(n, d_rhs, d_rhs, &t))
      return -1;
   normb = sqrt(t);
   if (normb < DBL_EPSILON)
   {
      b200k_fill(n, 0.0, d_x);
      *prel_res = 0.0;
      *piter = 0;
      *prel_res_v = (double *)b200_xcalloc(1, sizeof(double));
      return b200_sync();
   }
   tolb = atol ? tol : tol * normb;
   if (maxits > (g_pcg_sharded ? g_pcg_nglobal : n))
      maxits = g_pcg_sharded ? g_pcg_nglobal : n;
   r = (double *)b200_malloc(bytes);
   z = (double *)b200_malloc(bytes);
   p = (double *)b200_malloc(bytes);
   q = (double *)b200_malloc(bytes);

   if (b200_d2d(r, d_rhs, bytes) || mv(mat, n, -1.0, d_x, 1.0, r) || dev_dot(n, r, r, &t))
      goto done;
   normr = sqrt(t);
   if (normr < tolb)
   {
      *prel_res = normr / normb;
      *piter = 0;
      rel_res_v = (double *)b200_xmalloc(sizeof(double));
      rel_res_v[0] = *prel_res;
      *prel_res_v = rel_res_v;
      rc = 0;
      goto done;
   }
   normr2 = normr;
   rel_res_v = (double *)b200_xcalloc((size_t)maxits + 1, sizeof(double));
   rel_res_v[0] = normr / normb;
   if (print_level > 0)
   {
      printf("--------------------------------------------------------------------------------\n");
      printf("Start PCG\n");
      printf("Residual Tol: %e\nMax number of iterations: %d\n", tolb, maxits);
      printf("--------------------------------------------------------------------------------\n");
      printf("Step    Residual norm  Relative res.  Convergence Rate\n");
      printf("%5d   %8e   %8e   N/A\n", 0, normr, rel_res_v[0]);
   }
   for (ii = 1; ii <= maxits; ii++)
   {
      /* One host synchronisation per iteration: the three dot products stay on the device, the updates form
       * beta = rho / rho_prev and alpha = rho / pq themselves (k_xpby_dev, k_pcg_update_dev) and become no-ops under the
       * reference's breakdown conditions, which the host then evaluates in the reference's order (pcg.c:119,136,155). */
      double rho1, pq, rr;
      const double *zz = z;
      double *d_rho = g_scal_dev + 1 + (ii & 1), *d_rho_prev = ii == 1 ? NULL : g_scal_dev + 1 + ((ii - 1) & 1);
      double *d_pq = g_scal_dev + 3, *d_rr = g_scal_dev + 4;
      if (!prec)
         zz = r; /* z = r: no copy needed on the device path */
      rho1 = rho;
      /* iterations 2 and 3 are captured into one CUDA graph each (the scalar slots alternate with the parity of ii), later
       * ones replay them: one graph launch and one synchronisation per iteration keep the GPU busy when the host is slow
       * to launch (measured on the B200 boxes: up to 1.8 ms of launch gaps per 7 ms iteration) */
      if (use_graphs && ii >= 2 && it_graph[ii & 1])
      {
         b200_launch_count_add(it_launches[ii & 1]);
         if (b200_graph_launch(it_graph[ii & 1]) || b200_sync())
            goto done;
      }
      else
      {
         int capturing = use_graphs && ii >= 2 && b200_graph_begin() == 0, erc;
         long long before = b200_launch_count();
         if (capturing)
            b200_outer_capture_set(1);
         else if (ii >= 2)
            use_graphs = 0;
         erc = (prec && ps(prec, n, z, r)) || b200k_dot(n, zz, r, d_rho) || (g_pcg_sharded && b200_comm_allreduce_sum(d_rho, 1)) ||
               b200k_xpby_dev(n, zz, d_rho, d_rho_prev, p) || mv(mat, n, 1.0, p, 0.0, q) || b200k_dot(n, q, p, d_pq) ||
               (g_pcg_sharded && b200_comm_allreduce_sum(d_pq, 1)) || b200k_pcg_update_dev(n, d_rho, d_rho_prev, d_pq, p, q, d_x, r, d_rr) ||
               (g_pcg_sharded && b200_comm_allreduce_sum(d_rr, 1)) || b200_d2h_async(g_scal_host + 1, g_scal_dev + 1, sizeof(double) * 4);
         if (capturing)
         {
            void *exec;
            b200_outer_capture_set(0);
            exec = b200_graph_end();
            if (erc || !exec)
            {
               /* capture refused: nothing has run; enqueue this iteration plainly and stay with plain launches */
               b200_graph_destroy(exec);
               use_graphs = 0;
               ii--;
               continue;
            }
            it_graph[ii & 1] = exec;
            it_launches[ii & 1] = b200_launch_count() - before;
            if (b200_graph_launch(exec))
               goto done;
         }
         else if (erc)
            goto done;
         if (b200_sync())
            goto done;
      }
      rho = g_scal_host[1 + (ii & 1)];
      pq = g_scal_host[3];
      if (rho == 0.0)
      {
         if (print_level > 1)
            printf("rho = %.16e\n", rho);
         break;
      }
      if (ii > 1)
      {
         beta = rho / rho1;
         if (beta == 0.0)
         {
            if (print_level > 1)
               printf("beta = %.16e\n", beta);
            break;
         }
      }
      if (pq <= 0)
      {
         if (print_level > 1)
            printf("pq = %.16e\n", pq);
         break;
      }
      alpha = rho / pq;
      (void)alpha;
      rr = g_scal_host[4];
      normr = sqrt(rr);
      normr2 = normr;
      rel_res_v[ii] = normr / normb;
      if (print_level > 0)
         printf("%5d   %8e   %8e   %8.6f\n", ii, normr, rel_res_v[ii], rel_res_v[ii] / rel_res_v[ii - 1]);
      if (normr <= tolb)
      {
         if (b200_d2d(r, d_rhs, bytes) || mv(mat, n, -1.0, d_x, 1.0, r) || dev_dot(n, r, r, &t))
            goto done;
         normr2 = sqrt(t);
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
   *prel_res_v = rel_res_v;
   rel_res_v = NULL;
   rc = 0;
done:
   b200_outer_capture_set(0);
   b200_sync();
   b200_graph_destroy(it_graph[0]);
   b200_graph_destroy(it_graph[1]);
   b200_free(r);
   b200_free(z);
   b200_free(p);
   b200_free(q);
   if (rc)
      free(rel_res_v);
   return rc;
}

/* ---- FOREIGN function pointers (a caller's own CPU matvec / preconditioner, e.g. the reference's dense
 * Nfft4GPDenseMatSymv): the iteration and all its vector operations still run on the GPU; only the foreign
 * callback itself is given host copies of its operands. ---- */
typedef struct
{
   func_symmatvec mv;
   func_solve ps;
   void *data;
   double *hx, *hy;
} foreign_ctx;

static int foreign_matvec_dev(void *vctx, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   foreign_ctx *c = (foreign_ctx *)vctx;
   size_t bytes = sizeof(double) * (size_t)n;
   if (b200_d2h(c->hx, d_x, bytes))
      return -1;
   if (beta != 0.0 && b200_d2h(c->hy, d_y, bytes))
      return -1;
   if (c->mv(c->data, n, alpha, c->hx, beta, c->hy))
      return -1;
   return b200_h2d(d_y, c->hy, bytes) || b200_sync();
}

static int foreign_solve_dev(void *vctx, int n, double *d_x, const double *d_rhs)
{
   foreign_ctx *c = (foreign_ctx *)vctx;
   size_t bytes = sizeof(double) * (size_t)n;
   if (b200_d2h(c->hx, d_rhs, bytes))
      return -1;
   if (c->ps(c->data, n, c->hy, c->hx))
      return -1;
   return b200_h2d(d_x, c->hy, bytes) || b200_sync();
}

/* SRC/solvers/pcg.c:3-206 */
int Nfft4GPSolverPcg(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc, double *x,
                     double *rhs, int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter,
                     int print_level)
{
   void *dev_mat = NULL;
   b200_dev_matvec mv = b200_lookup_dev_matvec((void *)matvec, mat_data, &dev_mat);
   b200_dev_solve ps = prec_data ? b200_lookup_dev_solve((void *)precondfunc) : NULL;
   void *dev_prec = prec_data;
   size_t bytes = sizeof(double) * (size_t)n;
   foreign_ctx fmv, fps;
   double *d_x, *d_rhs;
   int rc = -1;
   memset(&fmv, 0, sizeof(fmv));
   memset(&fps, 0, sizeof(fps));
   if (b200_init(-1))
      return -1;
   if (!mv)
   {
      fmv.mv = matvec;
      fmv.data = mat_data;
      fmv.hx = (double *)b200_host_alloc(bytes);
      fmv.hy = (double *)b200_host_alloc(bytes);
      mv = foreign_matvec_dev;
      dev_mat = &fmv;
   }
   if (prec_data && !ps)
   {
      b200_note_foreign_fn_on_native_handle(prec_data);
      fps.ps = precondfunc;
      fps.data = prec_data;
      fps.hx = (double *)b200_host_alloc(bytes);
      fps.hy = (double *)b200_host_alloc(bytes);
      ps = foreign_solve_dev;
      dev_prec = &fps;
   }
   {
      /* point-block sharding: x and rhs are FULL host vectors on every rank (reference semantics); each rank moves and
       * solves for its own block [lo, hi) and fills those entries of x */
      int lo = 0, hi = n;
      b200_additive *A = (matvec == Nfft4GPAdditiveNFFTMatSymv) ? b200_additive_from_handle(mat_data) : NULL;
      if (A && A->shard_mode == 2 && b200_comm_size() > 1)
      {
         if (fmv.mv || fps.ps)
         {
            printf("Error: point-block-sharded PCG needs this library's own matvec and preconditioner.\n");
            b200_host_free(fmv.hx);
            b200_host_free(fmv.hy);
            b200_host_free(fps.hx);
            b200_host_free(fps.hy);
            return -1;
         }
         lo = A->blk_lo;
         hi = A->blk_hi;
         g_pcg_sharded = 1;
         g_pcg_nglobal = n;
      }
      bytes = sizeof(double) * (size_t)(hi - lo);
      d_x = (double *)b200_malloc(bytes);
      d_rhs = (double *)b200_malloc(bytes);
      if (!b200_h2d(d_x, x + lo, bytes) && !b200_h2d(d_rhs, rhs + lo, bytes))
         rc = b200_pcg_device(dev_mat, hi - lo, mv, dev_prec, ps, d_x, d_rhs, maxits, atol, tol, prel_res, prel_res_v, piter, print_level);
      if (!rc)
         rc = b200_d2h(x + lo, d_x, bytes);
      g_pcg_sharded = 0;
   }
   b200_free(d_x);
   b200_free(d_rhs);
   b200_host_free(fmv.hx);
   b200_host_free(fmv.hy);
   b200_host_free(fps.hx);
   b200_host_free(fps.hy);
   return rc;
}

int Nfft4GPB200SolverPcgDevice(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc,
                               double *d_x, const double *d_rhs, int maxits, int atol, double tol, double *prel_res,
                               double **prel_res_v, int *piter, int print_level)
{
   void *dev_mat = NULL;
   b200_dev_matvec mv = b200_lookup_dev_matvec((void *)matvec, mat_data, &dev_mat);
   b200_dev_solve ps = prec_data ? b200_lookup_dev_solve((void *)precondfunc) : NULL;
   if (!mv || (prec_data && !ps))
   {
      printf("Error: Nfft4GPB200SolverPcgDevice needs this library's own matvec / preconditioner entry points.\n");
      return -1;
   }
   {
      /* device vectors hold the rank's block under point-block sharding (n = block length) */
      b200_additive *A = (matvec == Nfft4GPAdditiveNFFTMatSymv) ? b200_additive_from_handle(mat_data) : NULL;
      int rc;
      if (A && A->shard_mode == 2 && b200_comm_size() > 1)
      {
         g_pcg_sharded = 1;
         g_pcg_nglobal = A->n;
      }
      rc = b200_pcg_device(dev_mat, n, mv, prec_data, ps, d_x, d_rhs, maxits, atol, tol, prel_res, prel_res_v, piter, print_level);
      g_pcg_sharded = 0;
      return rc;
   }
}
