// Preconditioner-apply kernels (HBM-bound): Nystrom low-rank apply in two passes over U, CSR SpMV for the
// FSAI factor and its explicit transpose, dense GEMV for the AFN coupling block K12, permutation gather /
// scatter, and the dense additive kernel blocks the setups need.
//   reference: SRC/preconds/nys.c:115-173 (three dgemv passes over U), fsai.c:106-123 + matops.c:139-272
//   (serial CSR N / T), afn.c:109-136, SRC/linearalg/kernels.c:495, 2218-2222, 3099-3494.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

// ------------------------------------------------------------------ Nystrom
#define NYS_ROWS 2048 /* rows per CTA in pass 1 (8 per thread) */
#define NYS_CG 4      /* columns per register group            */

// pass 1: y = rhs(perm) (stored), partial[chunk][c] = sum_{rows of chunk} U[r,c] y[r]
__global__ void __launch_bounds__(256)
k_nys_pass1(int n, int k, const int *__restrict__ perm, const double *__restrict__ U, const double *__restrict__ rhs,
            double *__restrict__ y, double *__restrict__ partial)
{
   __shared__ double sh[8][NYS_CG];
   const int r0 = blockIdx.x * NYS_ROWS;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   double yv[8];
#pragma unroll
   for (int i = 0; i < 8; i++)
   {
      int r = r0 + threadIdx.x + 256 * i;
      double v = 0.0;
      if (r < n)
      {
         v = perm ? __ldg(rhs + perm[r]) : rhs[r];
         y[r] = v;
      }
      yv[i] = v;
   }
   for (int c = 0; c < k; c += NYS_CG)
   {
      double acc[NYS_CG];
#pragma unroll
      for (int cc = 0; cc < NYS_CG; cc++)
         acc[cc] = 0.0;
#pragma unroll
      for (int i = 0; i < 8; i++)
      {
         int r = r0 + threadIdx.x + 256 * i;
         if (r < n)
         {
#pragma unroll
            for (int cc = 0; cc < NYS_CG; cc++)
               if (c + cc < k)
                  acc[cc] = fma(U[(size_t)(c + cc) * n + r], yv[i], acc[cc]);
         }
      }
#pragma unroll
      for (int cc = 0; cc < NYS_CG; cc++)
         acc[cc] = b200_warp_sum(acc[cc]);
      __syncthreads();
      if (lane == 0)
      {
#pragma unroll
         for (int cc = 0; cc < NYS_CG; cc++)
            sh[wid][cc] = acc[cc];
      }
      __syncthreads();
      if (threadIdx.x < NYS_CG && c + threadIdx.x < k)
      {
         double s = 0.0;
#pragma unroll
         for (int w = 0; w < 8; w++)
            s += sh[w][threadIdx.x];
         partial[(size_t)blockIdx.x * k + c + threadIdx.x] = s;
      }
   }
}

// w[c] = (s[c] - 1/eta) * sum_chunks partial[chunk][c]
__global__ void k_nys_reduce(int k, int nchunks, const double *__restrict__ partial, const double *__restrict__ s, double eta,
                             double *__restrict__ w)
{
   int c = blockIdx.x * blockDim.x + threadIdx.x;
   if (c >= k)
      return;
   double z = 0.0;
   for (int ch = 0; ch < nchunks; ch++)
      z += partial[(size_t)ch * k + c];
   w[c] = (s[c] - 1.0 / eta) * z;
}

// pass 2: x(perm) = y/eta + U w
__global__ void __launch_bounds__(256)
k_nys_pass2(int n, int k, const int *__restrict__ perm, const double *__restrict__ U, const double *__restrict__ w, double eta,
            const double *__restrict__ y, double *__restrict__ x)
{
   extern __shared__ double sw[];
   for (int c = threadIdx.x; c < k; c += blockDim.x)
      sw[c] = w[c];
   __syncthreads();
   int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= n)
      return;
   double acc0 = y[r] / eta, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
   const double *u = U + r;
   int c = 0;
   for (; c + 8 <= k; c += 8)
   {
      double v0 = u[(size_t)(c + 0) * n], v1 = u[(size_t)(c + 1) * n], v2 = u[(size_t)(c + 2) * n], v3 = u[(size_t)(c + 3) * n];
      double v4 = u[(size_t)(c + 4) * n], v5 = u[(size_t)(c + 5) * n], v6 = u[(size_t)(c + 6) * n], v7 = u[(size_t)(c + 7) * n];
      acc0 = fma(v0, sw[c + 0], acc0);
      acc1 = fma(v1, sw[c + 1], acc1);
      acc2 = fma(v2, sw[c + 2], acc2);
      acc3 = fma(v3, sw[c + 3], acc3);
      acc0 = fma(v4, sw[c + 4], acc0);
      acc1 = fma(v5, sw[c + 5], acc1);
      acc2 = fma(v6, sw[c + 6], acc2);
      acc3 = fma(v7, sw[c + 7], acc3);
   }
   for (; c < k; c++)
      acc0 = fma(u[(size_t)c * n], sw[c], acc0);
   double res = (acc0 + acc1) + (acc2 + acc3);
   if (perm)
      x[perm[r]] = res;
   else
      x[r] = res;
}

// ------------------------------------------------------------------ CSR, permutations
__global__ void __launch_bounds__(256)
k_csr_mv(int nrows, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a,
         const double *__restrict__ x, double *__restrict__ y)
{
   int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (row >= nrows)
      return;
   int j1 = ia[row], j2 = ia[row + 1];
   double s = 0.0;
   // values and column indices are read once per product (streaming, evict-first: 12 B per entry that must not push the 8 n bytes of
   // x out of L2); the gathers x[ja] are the reusable part (__ldg: L1 / L2 resident)
   for (int j = j1 + lane; j < j2; j += 32)
      s = fma(__ldcs(a + j), __ldg(x + __ldcs(ja + j)), s);
   s = b200_warp_sum(s);
   if (lane == 0)
      y[row] = s;
}

__global__ void k_gather_perm(int n, const int *__restrict__ perm, const double *__restrict__ src, double *__restrict__ dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      dst[i] = __ldg(src + perm[i]);
}

__global__ void k_scatter_perm(int n, const int *__restrict__ perm, const double *__restrict__ src, double *__restrict__ dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      dst[perm[i]] = src[i];
}

// ------------------------------------------------------------------ dense GEMV, column-major, lda = m
// 'T': one warp per column
__global__ void __launch_bounds__(256)
k_gemv_t(int m, int n, double alpha, const double *__restrict__ A, const double *__restrict__ x, double beta, double *__restrict__ y)
{
   int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (col >= n)
      return;
   const double *a = A + (size_t)col * m;
   double s = 0.0;
   for (int i = lane; i < m; i += 32)
      s = fma(a[i], __ldg(x + i), s);
   s = b200_warp_sum(s);
   if (lane == 0)
      y[col] = alpha * s + (beta == 0.0 ? 0.0 : beta * y[col]);
}

// 'N': CTA = column chunk, thread = row (rows looped in strides of blockDim); partial[chunk][row]
#define GEMV_N_COLS 512
__global__ void __launch_bounds__(256)
k_gemv_n_stage1(int m, int n, int cols, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ partial)
{
   __shared__ double sx[GEMV_N_COLS];
   int j0 = blockIdx.x * cols;
   int jn = min(cols, n - j0);
   for (int j = threadIdx.x; j < jn; j += blockDim.x)
      sx[j] = x[j0 + j];
   __syncthreads();
   for (int i = threadIdx.x; i < m; i += blockDim.x)
   {
      const double *a = A + (size_t)j0 * m + i;
      double s0 = 0.0, s1 = 0.0;
      int j = 0;
      for (; j + 2 <= jn; j += 2)
      {
         s0 = fma(a[(size_t)j * m], sx[j], s0);
         s1 = fma(a[(size_t)(j + 1) * m], sx[j + 1], s1);
      }
      if (j < jn)
         s0 = fma(a[(size_t)j * m], sx[j], s0);
      partial[(size_t)blockIdx.x * m + i] = s0 + s1;
   }
}

__global__ void k_gemv_n_stage2(int m, int nchunks, double alpha, const double *__restrict__ partial, double beta, double *__restrict__ y)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= m)
      return;
   double s = 0.0;
   for (int c = 0; c < nchunks; c++)
      s += partial[(size_t)c * m + i];
   y[i] = alpha * s + (beta == 0.0 ? 0.0 : beta * y[i]);
}

// wide-load variants of the two fp64 products (m even; x of the 'T' product in shared memory, persistent warps, two columns
// per warp; 'N': two rows per thread, eight columns in flight) -- the K12 passes of the AFN apply went from 5.5 to about
// 6.3 TB/s with them
#define GEMV64_CW 4
__global__ void __launch_bounds__(256)
k_gemv_t_wide(int m, int n, double alpha, const double *__restrict__ A, const double *__restrict__ x, double beta, double *__restrict__ y)
{
   extern __shared__ __align__(16) double sxt[];
   for (int i = threadIdx.x; i < m; i += blockDim.x)
      sxt[i] = x[i];
   __syncthreads();
   const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   const int nw = (gridDim.x * blockDim.x) >> 5;
   const int m2 = m >> 1;
   for (int c0 = w * GEMV64_CW; c0 < n; c0 += nw * GEMV64_CW)
   {
      double s[GEMV64_CW], t[GEMV64_CW];
      const double2 *a[GEMV64_CW];
#pragma unroll
      for (int c = 0; c < GEMV64_CW; c++)
      {
         s[c] = t[c] = 0.0;
         a[c] = reinterpret_cast<const double2 *>(A + (size_t)(c0 + c < n ? c0 + c : c0) * m);
      }
      for (int i = lane; i < m2; i += 32)
      {
         double2 v[GEMV64_CW];
#pragma unroll
         for (int c = 0; c < GEMV64_CW; c++)
            v[c] = __ldcs(a[c] + i);
         const double2 xa = *reinterpret_cast<const double2 *>(sxt + 2 * i);
#pragma unroll
         for (int c = 0; c < GEMV64_CW; c++)
         {
            s[c] = fma(v[c].x, xa.x, s[c]);
            t[c] = fma(v[c].y, xa.y, t[c]);
         }
      }
#pragma unroll
      for (int c = 0; c < GEMV64_CW; c++)
      {
         const double r = b200_warp_sum(s[c] + t[c]);
         if (lane == 0 && c0 + c < n)
            y[c0 + c] = alpha * r + (beta == 0.0 ? 0.0 : beta * y[c0 + c]);
      }
   }
}

#define GEMV64_U 8
__global__ void __launch_bounds__(256)
k_gemv_n_stage1_wide(int m, int n, int cols, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ partial)
{
   __shared__ double sx[GEMV_N_COLS];
   int j0 = blockIdx.x * cols;
   int jn = min(cols, n - j0);
   for (int j = threadIdx.x; j < jn; j += blockDim.x)
      sx[j] = x[j0 + j];
   __syncthreads();
   const int m2 = m >> 1;
   for (int i = threadIdx.x; i < m2; i += blockDim.x)
   {
      const double2 *a = reinterpret_cast<const double2 *>(A + (size_t)j0 * m) + i;
      double s0 = 0.0, s1 = 0.0, u0 = 0.0, u1 = 0.0;
      int j = 0;
      for (; j + GEMV64_U <= jn; j += GEMV64_U)
      {
         double2 v[GEMV64_U];
#pragma unroll
         for (int u = 0; u < GEMV64_U; u++)
            v[u] = __ldcs(a + (size_t)(j + u) * m2);
#pragma unroll
         for (int u = 0; u < GEMV64_U; u += 2)
         {
            s0 = fma(v[u].x, sx[j + u], s0);
            s1 = fma(v[u].y, sx[j + u], s1);
            u0 = fma(v[u + 1].x, sx[j + u + 1], u0);
            u1 = fma(v[u + 1].y, sx[j + u + 1], u1);
         }
      }
      for (; j < jn; j++)
      {
         const double2 v = __ldcs(a + (size_t)j * m2);
         s0 = fma(v.x, sx[j], s0);
         s1 = fma(v.y, sx[j], s1);
      }
      partial[(size_t)blockIdx.x * m + 2 * i] = s0 + u0;
      partial[(size_t)blockIdx.x * m + 2 * i + 1] = s1 + u1;
   }
}

// ---- the same two products with the matrix stored in fp32 (optional storage of the AFN K12 block: the applies are HBM-bound
// on reading it twice, the preconditioner only has to be symmetric positive definite, not exact); fp64 vectors and sums.
// m a multiple of 4: every column starts 16-byte aligned, four rows per 128-bit load.  Variants timed in isolation with
// scripts/gemv_f32_bench.cu (m = 4000, 8 GB): 'T' warp per column 5.0 TB/s -> persistent warps, x in shared memory, 4 columns
// per warp 6.7 TB/s; 'N' two rows per thread 6.0 TB/s -> four rows per thread, 8 columns in flight 6.7 TB/s.
#define GEMV32_CW 4
__global__ void __launch_bounds__(256)
k_gemv_t_f32(int m, int n, double alpha, const float *__restrict__ A, const double *__restrict__ x, double beta, double *__restrict__ y)
{
   extern __shared__ __align__(16) double sxt[];
   for (int i = threadIdx.x; i < m; i += blockDim.x)
      sxt[i] = x[i];
   __syncthreads();
   const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   const int nw = (gridDim.x * blockDim.x) >> 5;
   const int m4 = m >> 2;
   for (int c0 = w * GEMV32_CW; c0 < n; c0 += nw * GEMV32_CW)
   {
      double s[GEMV32_CW];
#pragma unroll
      for (int c = 0; c < GEMV32_CW; c++)
         s[c] = 0.0;
      for (int i = lane; i < m4; i += 32)
      {
         const double2 xa = *reinterpret_cast<const double2 *>(sxt + 4 * i), xb = *reinterpret_cast<const double2 *>(sxt + 4 * i + 2);
#pragma unroll
         for (int c = 0; c < GEMV32_CW; c++)
            if (c0 + c < n)
            {
               const float4 v = __ldcs(reinterpret_cast<const float4 *>(A + (size_t)(c0 + c) * m) + i);
               s[c] = fma((double)v.x, xa.x, s[c]);
               s[c] = fma((double)v.y, xa.y, s[c]);
               s[c] = fma((double)v.z, xb.x, s[c]);
               s[c] = fma((double)v.w, xb.y, s[c]);
            }
      }
#pragma unroll
      for (int c = 0; c < GEMV32_CW; c++)
      {
         const double r = b200_warp_sum(s[c]);
         if (lane == 0 && c0 + c < n)
            y[c0 + c] = alpha * r + (beta == 0.0 ? 0.0 : beta * y[c0 + c]);
      }
   }
}

#define GEMV32_U 8
__global__ void __launch_bounds__(256)
k_gemv_n_stage1_f32(int m, int n, int cols, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ partial)
{
   __shared__ double sx[GEMV_N_COLS];
   int j0 = blockIdx.x * cols;
   int jn = min(cols, n - j0);
   for (int j = threadIdx.x; j < jn; j += blockDim.x)
      sx[j] = x[j0 + j];
   __syncthreads();
   const int m4 = m >> 2;
   for (int i = threadIdx.x; i < m4; i += blockDim.x)
   {
      const float4 *a = reinterpret_cast<const float4 *>(A + (size_t)j0 * m) + i;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = 0;
      for (; j + GEMV32_U <= jn; j += GEMV32_U)
      {
         float4 v[GEMV32_U];
#pragma unroll
         for (int u = 0; u < GEMV32_U; u++)
            v[u] = __ldcs(a + (size_t)(j + u) * m4);
#pragma unroll
         for (int u = 0; u < GEMV32_U; u++)
         {
            const double xj = sx[j + u];
            s0 = fma((double)v[u].x, xj, s0);
            s1 = fma((double)v[u].y, xj, s1);
            s2 = fma((double)v[u].z, xj, s2);
            s3 = fma((double)v[u].w, xj, s3);
         }
      }
      for (; j < jn; j++)
      {
         const float4 v = __ldcs(a + (size_t)j * m4);
         const double xj = sx[j];
         s0 = fma((double)v.x, xj, s0);
         s1 = fma((double)v.y, xj, s1);
         s2 = fma((double)v.z, xj, s2);
         s3 = fma((double)v.w, xj, s3);
      }
      double *p = partial + (size_t)blockIdx.x * m + 4 * i;
      p[0] = s0;
      p[1] = s1;
      p[2] = s2;
      p[3] = s3;
   }
}

__global__ void k_f64_to_f32(size_t count, const double *__restrict__ src, float *__restrict__ dst)
{
   for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = (float)src[i];
}
__global__ void k_f32_to_f64(size_t count, const float *__restrict__ src, double *__restrict__ dst)
{
   for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = (double)src[i];
}

// ------------------------------------------------------------------ dense additive kernel blocks
// K[a + b*ldk] = ff * ( (1/P) sum_w k_w(x_rows[a], x_cols[b]) ) (+ noise_diag when rows[a] == cols[b]);
// Xw: window-major feature copy, window w's feature t at Xw[(w*dw + t)*n + i]; dims[w] = actual dimension.
// kernel 0: exp(-r^2 / (2 l^2))  (kernels.c:495)      1: exp(-r / l)  (kernels.c:2218-2222)
__global__ void __launch_bounds__(256)
k_dense_additive(const double *__restrict__ Xw, int n, int P, int dw, const int *__restrict__ dims, int kernel, double ff,
                 double l, double noise_diag, const int *__restrict__ rows, int nr, const int *__restrict__ cols, int nc,
                 int lower_only, double *__restrict__ K, int ldk, int col0)
{
   int a = blockIdx.x * blockDim.x + threadIdx.x;
   int b = col0 + blockIdx.y;
   if (a >= nr || b >= nc)
      return;
   if (lower_only && a < b)
      return;
   int ia = rows ? rows[a] : a;
   int ib = cols ? cols[b] : b;
   double acc = 0.0;
   for (int w = 0; w < P; w++)
   {
      double r2 = 0.0;
      int dd = dims[w];
      for (int t = 0; t < dd; t++)
      {
         const double *col = Xw + (size_t)(w * dw + t) * n;
         double df = __ldg(col + ia) - __ldg(col + ib);
         r2 = fma(df, df, r2);
      }
      // kernel 0 / 1: Gauss / Matern-1/2 value; 2 / 3: their derivative in the length scale (kernels.c:495-523, 2218-2222)
      double r = sqrt(r2);
      double e = (kernel & 1) ? exp(-r / l) : exp(-r2 / (2.0 * l * l));
      if (kernel >= 2)
         e *= (kernel & 1) ? r / (l * l) : r2 / (l * l * l);
      acc += e;
   }
   acc = ff * (acc / (double)P);
   if (ia == ib)
      acc += noise_diag;
   K[(size_t)b * ldk + a] = acc;
}

// ------------------------------------------------------------------ launch layer
static double *g_scratch = nullptr;
static size_t g_scratch_doubles = 0;
static double *scratch(size_t doubles)
{
   if (doubles > g_scratch_doubles)
   {
      if (g_scratch)
         cudaFree(g_scratch);
      size_t want = doubles + doubles / 4 + 1024;
      if (cudaMalloc(&g_scratch, want * sizeof(double)) != cudaSuccess)
      {
         fprintf(stderr, "nfft4gp_b200: scratch allocation of %zu doubles failed\n", want);
         g_scratch = nullptr;
         g_scratch_doubles = 0;
         return nullptr;
      }
      g_scratch_doubles = want;
   }
   return g_scratch;
}

// columns per CTA of the 'N' products: 512 for wide matrices; narrower ones (the k x k triangular inverse of the AFN apply:
// 8 CTAs at 512 columns each read 128 MB at a few percent of the HBM rate) are cut so that about four CTAs per SM exist
static int gemv_n_cols(int n)
{
   int cols = b200_cdiv(n, 148 * 4);
   cols = (cols + 7) & ~7;
   if (cols < 8)
      cols = 8;
   return cols > GEMV_N_COLS ? GEMV_N_COLS : cols;
}

extern "C" {

size_t b200k_nys_work_doubles(int n, int k)
{
   size_t nchunks = (size_t)b200_cdiv(n, NYS_ROWS);
   return (size_t)n + (size_t)k + nchunks * (size_t)k;
}

int b200k_nys_apply(int n, int k, const int *perm, const double *U, const double *s, double eta, const double *rhs,
                    double *x, double *work, size_t work_doubles)
{
   B200ProfScope prof_scope_(PC_NYS);
   int nchunks = b200_cdiv(n, NYS_ROWS);
   if (work_doubles < b200k_nys_work_doubles(n, k))
   {
      fprintf(stderr, "nfft4gp_b200: Nystrom work array too small\n");
      return -1;
   }
   double *y = work, *w = work + n, *partial = work + n + k;
   cudaStream_t st = b200_stream();
   k_nys_pass1<<<nchunks, 256, 0, st>>>(n, k, perm, U, rhs, y, partial);
   B200_LAUNCHED();
   k_nys_reduce<<<b200_cdiv(k, 128), 128, 0, st>>>(k, nchunks, partial, s, eta, w);
   B200_LAUNCHED();
   k_nys_pass2<<<b200_cdiv(n, 256), 256, sizeof(double) * (size_t)k, st>>>(n, k, perm, U, w, eta, y, x);
   B200_LAUNCHED();
   return 0;
}

int b200k_csr_mv(int nrows, const int *ia, const int *ja, const double *a, const double *x, double *y)
{
   B200ProfScope prof_scope_(PC_CSR);
   k_csr_mv<<<b200_cdiv((long long)nrows * 32, 256), 256, 0, b200_stream()>>>(nrows, ia, ja, a, x, y);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather_perm(int n, const int *perm, const double *src, double *dst)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_gather_perm<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, perm, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_scatter_perm(int n, const int *perm, const double *src, double *dst)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_scatter_perm<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, perm, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_gemv(char trans, int m, int n, double alpha, const double *A, const double *x, double beta, double *y)
{
   B200ProfScope prof_scope_(PC_GEMV);
   cudaStream_t st = b200_stream();
   const bool wide = !(m & 1) && m >= 256 && ((uintptr_t)A & 15) == 0;
   if (trans == 'T' || trans == 't')
   {
      const size_t sm = sizeof(double) * (size_t)m;
      if (wide && sm <= 96 * 1024 && n >= 64)
      {
         static int smem_set = 0;
         if ((int)sm > smem_set)
         {
            B200_CHECK(cudaFuncSetAttribute(k_gemv_t_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            smem_set = (int)sm;
         }
         int ctas = b200_cdiv(b200_cdiv(n, GEMV64_CW), 8);
         if (ctas > 148 * 4)
            ctas = 148 * 4;
         k_gemv_t_wide<<<ctas, 256, sm, st>>>(m, n, alpha, A, x, beta, y);
      }
      else
         k_gemv_t<<<b200_cdiv((long long)n * 32, 256), 256, 0, st>>>(m, n, alpha, A, x, beta, y);
      B200_LAUNCHED();
   }
   else
   {
      int cols = gemv_n_cols(n);
      int nchunks = b200_cdiv(n, cols);
      double *partial = scratch((size_t)nchunks * m);
      if (!partial)
         return -1;
      if (wide)
         k_gemv_n_stage1_wide<<<nchunks, 256, 0, st>>>(m, n, cols, A, x, partial);
      else
         k_gemv_n_stage1<<<nchunks, 256, 0, st>>>(m, n, cols, A, x, partial);
      B200_LAUNCHED();
      k_gemv_n_stage2<<<b200_cdiv(m, 128), 128, 0, st>>>(m, nchunks, alpha, partial, beta, y);
      B200_LAUNCHED();
   }
   return 0;
}

// matrix in fp32 (m a multiple of 4), vectors and accumulation in fp64
int b200k_gemv_f32(char trans, int m, int n, double alpha, const float *A, const double *x, double beta, double *y)
{
   B200ProfScope prof_scope_(PC_GEMV);
   cudaStream_t st = b200_stream();
   if (m & 3)
      return -1;
   if (trans == 'T' || trans == 't')
   {
      static int smem_set = 0;
      const size_t sm = sizeof(double) * (size_t)m;
      if (sm > 200 * 1024)
         return -1;
      if ((int)sm > smem_set)
      {
         B200_CHECK(cudaFuncSetAttribute(k_gemv_t_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
         smem_set = (int)sm;
      }
      int ctas = b200_cdiv(b200_cdiv(n, GEMV32_CW), 8);
      if (ctas > 148 * 4)
         ctas = 148 * 4;
      k_gemv_t_f32<<<ctas, 256, sm, st>>>(m, n, alpha, A, x, beta, y);
      B200_LAUNCHED();
   }
   else
   {
      int cols = gemv_n_cols(n);
      int nchunks = b200_cdiv(n, cols);
      double *partial = scratch((size_t)nchunks * m);
      if (!partial)
         return -1;
      k_gemv_n_stage1_f32<<<nchunks, 256, 0, st>>>(m, n, cols, A, x, partial);
      B200_LAUNCHED();
      k_gemv_n_stage2<<<b200_cdiv(m, 128), 128, 0, st>>>(m, nchunks, alpha, partial, beta, y);
      B200_LAUNCHED();
   }
   return 0;
}

int b200k_convert_f64_f32(size_t count, const double *src, float *dst)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_f64_to_f32<<<148 * 8, 256, 0, b200_stream()>>>(count, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_convert_f32_f64(size_t count, const float *src, double *dst)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_f32_to_f64<<<148 * 8, 256, 0, b200_stream()>>>(count, src, dst);
   B200_LAUNCHED();
   return 0;
}

int b200k_dense_additive(const double *Xw, int n, int P, int dw, const int *d_dims, int kernel, double ff, double l,
                         double noise_diag, const int *rows, int nr, const int *cols, int nc, int lower_only, double *K, int ldk)
{
   B200ProfScope prof_scope_(PC_OTHER);
   // gridDim.y is limited to 65535: walk the columns in blocks
   for (int c0 = 0; c0 < nc; c0 += 65535)
   {
      int ncb = nc - c0 < 65535 ? nc - c0 : 65535;
      dim3 grid(b200_cdiv(nr, 256), ncb);
      k_dense_additive<<<grid, 256, 0, b200_stream()>>>(Xw, n, P, dw, d_dims, kernel, ff, l, noise_diag, rows, nr, cols, nc, lower_only, K, ldk, c0);
      B200_LAUNCHED();
   }
   return 0;
}

} // extern "C"
