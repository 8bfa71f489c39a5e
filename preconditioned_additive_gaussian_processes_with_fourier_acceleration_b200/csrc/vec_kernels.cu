// Fused vector kernels of the device-resident PCG (replaces the nine unfused n-passes per iteration of
// SRC/solvers/pcg.c:104-170 built on SRC/linearalg/vecops.c:3-155).  HBM-bound: every kernel streams its
// operands once with 128-bit loads; reductions are two-stage with a fixed summation order (deterministic).
#include <cuda_runtime.h>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

#define RED_BLOCKS 1184 /* 148 SMs x 8 */
#define RED_THREADS 256

__device__ double g_partial[2 * RED_BLOCKS];

__device__ __forceinline__ double block_sum(double v, double *sh)
{
   v = b200_warp_sum(v);
   int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   __syncthreads();
   if (lane == 0)
      sh[wid] = v;
   __syncthreads();
   double r = 0.0;
   if (threadIdx.x == 0)
   {
      for (int i = 0; i < (int)(blockDim.x >> 5); i++)
         r += sh[i];
   }
   return r; // valid in thread 0
}

__global__ void __launch_bounds__(RED_THREADS) k_dot_stage1(int n, const double *__restrict__ x, const double *__restrict__ y)
{
   __shared__ double sh[32];
   double s = 0.0;
   int n2 = n >> 1;
   const double2 *x2 = reinterpret_cast<const double2 *>(x), *y2 = reinterpret_cast<const double2 *>(y);
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
   {
      double2 a = x2[i], b = y2[i];
      s = fma(a.x, b.x, s);
      s = fma(a.y, b.y, s);
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
      s = fma(x[n - 1], y[n - 1], s);
   double r = block_sum(s, sh);
   if (threadIdx.x == 0)
      g_partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(RED_THREADS)
k_dot2_stage1(int n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ u, const double *__restrict__ v)
{
   __shared__ double sh[32];
   double s = 0.0, t = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      s = fma(x[i], y[i], s);
      t = fma(u[i], v[i], t);
   }
   double r = block_sum(s, sh);
   double r2 = block_sum(t, sh);
   if (threadIdx.x == 0)
   {
      g_partial[blockIdx.x] = r;
      g_partial[RED_BLOCKS + blockIdx.x] = r2;
   }
}

// x += alpha p ; r -= alpha q ; partial r.r
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_update_stage1(int n, double alpha, const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r)
{
   __shared__ double sh[32];
   double s = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      x[i] += alpha * p[i];
      double ri = r[i] - alpha * q[i];
      r[i] = ri;
      s = fma(ri, ri, s);
   }
   double t = block_sum(s, sh);
   if (threadIdx.x == 0)
      g_partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(RED_THREADS) k_reduce_stage2(int nblocks, int nvals, double *__restrict__ out)
{
   __shared__ double sh[32];
   for (int v = 0; v < nvals; v++)
   {
      double s = 0.0;
      for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
         s += g_partial[v * RED_BLOCKS + i];
      double r = block_sum(s, sh);
      if (threadIdx.x == 0)
         out[v] = r;
   }
}

__global__ void k_xpby(int n, const double *__restrict__ z, double beta, double *__restrict__ p)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      p[i] = z[i] + beta * p[i]; // pcg.c:145-146: p = beta*p ; p += z
}

// Device-scalar forms of the two PCG updates: the step lengths are formed ON THE DEVICE from the dot products of the
// same iteration (alpha = rho / pq, beta = rho / rho_prev: the same IEEE divisions the host would do), so that the host
// synchronises once per iteration instead of three times.  The reference's breakdown exits (pcg.c:119,136,155: rho == 0,
// beta == 0, pq <= 0) leave BEFORE the update they guard; the kernels therefore turn into no-ops under the same conditions
// and the host, reading the three scalars afterwards, breaks with x, r, p exactly as the reference leaves them.
__global__ void k_xpby_dev(int n, const double *__restrict__ z, const double *__restrict__ d_rho, const double *__restrict__ d_rho_prev,
                           double *__restrict__ p)
{
   const double rho = *d_rho;
   if (rho == 0.0)
      return;
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (d_rho_prev == nullptr) // first iteration: p = z
   {
      if (i < n)
         p[i] = z[i];
      return;
   }
   const double beta = rho / *d_rho_prev;
   if (beta == 0.0)
      return;
   if (i < n)
      p[i] = z[i] + beta * p[i];
}

__global__ void __launch_bounds__(RED_THREADS)
k_pcg_update_dev_stage1(int n, const double *__restrict__ d_rho, const double *__restrict__ d_rho_prev, const double *__restrict__ d_pq,
                        const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r)
{
   __shared__ double sh[32];
   const double rho = *d_rho, pq = *d_pq;
   const bool skip = rho == 0.0 || pq <= 0.0 || (d_rho_prev != nullptr && rho / *d_rho_prev == 0.0);
   const double alpha = skip ? 0.0 : rho / pq;
   double s = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      double ri = r[i];
      if (!skip)
      {
         x[i] += alpha * p[i];
         ri = ri - alpha * q[i];
         r[i] = ri;
      }
      s = fma(ri, ri, s);
   }
   double t = block_sum(s, sh);
   if (threadIdx.x == 0)
      g_partial[blockIdx.x] = t;
}

__global__ void k_copy(int n, const double *__restrict__ s, double *__restrict__ d)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      d[i] = s[i];
}

__global__ void k_fill(int n, double v, double *__restrict__ d)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      d[i] = v;
}

__global__ void k_scale(int n, double a, double *__restrict__ x)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      x[i] *= a;
}

__global__ void k_axpy(int n, double a, const double *__restrict__ x, double *__restrict__ y)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      y[i] += a * x[i];
}

static int red_blocks(int n)
{
   int b = b200_cdiv(n, RED_THREADS * 4);
   if (b < 1)
      b = 1;
   if (b > RED_BLOCKS)
      b = RED_BLOCKS;
   return b;
}


// ---- modified Gram-Schmidt without host round trips (FGMRES, SRC/linearalg/matops.c:274-346): one fused pass per basis
// vector:  w -= (*t_prev) * v_prev  (the update of the PREVIOUS step, coefficient read from device memory), then the partial
// sums of  w . v_next  (v_next == nullptr: w . w).  The coefficients never leave the device inside an Arnoldi step.
__global__ void __launch_bounds__(RED_THREADS)
k_mgs_stage1(int n, double *__restrict__ w, const double *__restrict__ v_prev, const double *__restrict__ t_prev,
             const double *__restrict__ v_next)
{
   __shared__ double sh[32];
   const double tp = v_prev ? *t_prev : 0.0;
   double s = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      double wi = w[i];
      if (v_prev)
      {
         wi = fma(-tp, v_prev[i], wi);
         w[i] = wi;
      }
      s = fma(wi, v_next ? v_next[i] : wi, s);
   }
   double r = block_sum(s, sh);
   if (threadIdx.x == 0)
      g_partial[blockIdx.x] = r;
}

// t = sqrt(*nrm2) -> *t_out ; w *= 1 / t   (matops.c:336-341)
__global__ void k_scale_by_inv_norm(int n, double *__restrict__ w, const double *__restrict__ nrm2, double *__restrict__ t_out)
{
   const double t = sqrt(*nrm2);
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i == 0)
      *t_out = t;
   if (i < n)
      w[i] *= 1.0 / t;
}

// x += Z(:, 0:m) * coef   (Z column-major with leading dimension ldz, coefficients in device memory; fgmres.c:215-219)
__global__ void k_multi_axpy(int n, int m, const double *__restrict__ Z, size_t ldz, const double *__restrict__ coef, double *__restrict__ x)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   double xi = x[i];
   for (int j = 0; j < m; j++)
      xi = fma(coef[j], Z[(size_t)j * ldz + i], xi); // same order as the reference's sequence of axpys
   x[i] = xi;
}

// ---- index gather / scatter with holes (idx < 0 = not owned by this rank), used by the point-block-sharded applies
__global__ void k_gather_idx(int m, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < m)
   {
      int j = idx[i];
      dst[i] = j >= 0 ? src[j] : 0.0;
   }
}

__global__ void k_scatter_idx(int m, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < m)
   {
      int j = idx[i];
      if (j >= 0)
         dst[j] = src[i];
   }
}

// c[i] = (s[i] - 1 / eta) * z[i]   (Nystrom apply, nys.c:141-160 folded into one coefficient vector)
__global__ void k_nys_coef(int k, const double *__restrict__ s, double inv_eta, const double *__restrict__ z, double *__restrict__ c)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < k)
      c[i] = (s[i] - inv_eta) * z[i];
}

// UT(c, j) = U(pos[j], c): local rows of the Nystrom factor, transposed (k x nl, column = point)
__global__ void k_gather_rows_T(int nl, int k, const int *__restrict__ pos, const double *__restrict__ U, int ldu, double *__restrict__ UT)
{
   size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= (size_t)nl * k)
      return;
   int c = (int)(e % k), j = (int)(e / k);
   UT[e] = U[(size_t)c * ldu + pos[j]];
}

extern "C" {

int b200k_dot(int n, const double *x, const double *y, double *d_out)
{
   B200ProfScope prof_scope_(PC_VEC);
   int nb = red_blocks(n);
   // the 128-bit path needs 16-byte aligned operands; fall back to the scalar two-vector kernel otherwise
   if ((((size_t)x | (size_t)y) & 15) == 0)
   {
      k_dot_stage1<<<nb, RED_THREADS, 0, b200_stream()>>>(n, x, y);
      B200_LAUNCHED();
      k_reduce_stage2<<<1, RED_THREADS, 0, b200_stream()>>>(nb, 1, d_out);
      B200_LAUNCHED();
   }
   else
   {
      k_dot2_stage1<<<nb, RED_THREADS, 0, b200_stream()>>>(n, x, y, x, y);
      B200_LAUNCHED();
      k_reduce_stage2<<<1, RED_THREADS, 0, b200_stream()>>>(nb, 1, d_out);
      B200_LAUNCHED();
   }
   return 0;
}

int b200k_dot2(int n, const double *x, const double *y, const double *u, const double *v, double *d_out2)
{
   B200ProfScope prof_scope_(PC_VEC);
   int nb = red_blocks(n);
   k_dot2_stage1<<<nb, RED_THREADS, 0, b200_stream()>>>(n, x, y, u, v);
   B200_LAUNCHED();
   k_reduce_stage2<<<1, RED_THREADS, 0, b200_stream()>>>(nb, 2, d_out2);
   B200_LAUNCHED();
   return 0;
}

int b200k_pcg_update(int n, double alpha, const double *p, const double *q, double *x, double *r, double *d_out)
{
   B200ProfScope prof_scope_(PC_VEC);
   int nb = red_blocks(n);
   k_pcg_update_stage1<<<nb, RED_THREADS, 0, b200_stream()>>>(n, alpha, p, q, x, r);
   B200_LAUNCHED();
   k_reduce_stage2<<<1, RED_THREADS, 0, b200_stream()>>>(nb, 1, d_out);
   B200_LAUNCHED();
   return 0;
}

int b200k_xpby_dev(int n, const double *z, const double *d_rho, const double *d_rho_prev, double *p)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_xpby_dev<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, z, d_rho, d_rho_prev, p);
   B200_LAUNCHED();
   return 0;
}

int b200k_pcg_update_dev(int n, const double *d_rho, const double *d_rho_prev, const double *d_pq, const double *p, const double *q,
                         double *x, double *r, double *d_out)
{
   B200ProfScope prof_scope_(PC_VEC);
   int nb = red_blocks(n);
   k_pcg_update_dev_stage1<<<nb, RED_THREADS, 0, b200_stream()>>>(n, d_rho, d_rho_prev, d_pq, p, q, x, r);
   B200_LAUNCHED();
   k_reduce_stage2<<<1, RED_THREADS, 0, b200_stream()>>>(nb, 1, d_out);
   B200_LAUNCHED();
   return 0;
}

int b200k_xpby(int n, const double *z, double beta, double *p)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_xpby<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, z, beta, p);
   B200_LAUNCHED();
   return 0;
}

int b200k_copy(int n, const double *src, double *dst)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_copy<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_fill(int n, double v, double *dst)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_fill<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, v, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_scale(int n, double a, double *x)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_scale<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, a, x);
   B200_LAUNCHED();
   return 0;
}

int b200k_axpy(int n, double a, const double *x, double *y)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_axpy<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, a, x, y);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather_idx(int m, const int *idx, const double *src, double *dst)
{
   B200ProfScope prof_scope_(PC_VEC);
   if (m <= 0)
      return 0;
   k_gather_idx<<<b200_cdiv(m, 256), 256, 0, b200_stream()>>>(m, idx, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_scatter_idx(int m, const int *idx, const double *src, double *dst)
{
   B200ProfScope prof_scope_(PC_VEC);
   if (m <= 0)
      return 0;
   k_scatter_idx<<<b200_cdiv(m, 256), 256, 0, b200_stream()>>>(m, idx, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_nys_coef(int k, const double *s, double inv_eta, const double *z, double *c)
{
   B200ProfScope prof_scope_(PC_NYS);
   k_nys_coef<<<b200_cdiv(k, 256), 256, 0, b200_stream()>>>(k, s, inv_eta, z, c);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather_rows_T(int nl, int k, const int *pos, const double *U, int ldu, double *UT)
{
   B200ProfScope prof_scope_(PC_OTHER);
   if (nl <= 0)
      return 0;
   k_gather_rows_T<<<b200_cdiv((long long)nl * k, 256), 256, 0, b200_stream()>>>(nl, k, pos, U, ldu, UT);
   B200_LAUNCHED();
   return 0;
}

int b200k_mgs_step(int n, double *w, const double *v_prev, const double *d_t_prev, const double *v_next, double *d_out)
{
   B200ProfScope prof_scope_(PC_VEC);
   int nb = red_blocks(n);
   k_mgs_stage1<<<nb, RED_THREADS, 0, b200_stream()>>>(n, w, v_prev, d_t_prev, v_next);
   B200_LAUNCHED();
   k_reduce_stage2<<<1, RED_THREADS, 0, b200_stream()>>>(nb, 1, d_out);
   B200_LAUNCHED();
   return 0;
}

int b200k_scale_by_inv_norm(int n, double *w, const double *d_nrm2, double *d_t_out)
{
   B200ProfScope prof_scope_(PC_VEC);
   k_scale_by_inv_norm<<<b200_cdiv(n > 0 ? n : 1, 256), 256, 0, b200_stream()>>>(n, w, d_nrm2, d_t_out);
   B200_LAUNCHED();
   return 0;
}

int b200k_multi_axpy(int n, int m, const double *Z, size_t ldz, const double *d_coef, double *x)
{
   B200ProfScope prof_scope_(PC_VEC);
   if (n <= 0 || m <= 0)
      return 0;
   k_multi_axpy<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, m, Z, ldz, d_coef, x);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
