// Micro-benchmark (development aid): y = A' x and y = A x with A (m x n, column-major) stored in fp32, fp64 vectors and sums --
// the two HBM-bound products of the AFN apply with fp32 K12 storage.  Prints GB/s of matrix traffic per variant.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double wsum(double v)
{
   for (int o = 16; o; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
// T1: warp per column, float2 per lane
__global__ void __launch_bounds__(256) t1(int m, int n, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ y)
{
   int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if (col >= n)
      return;
   const float2 *a = (const float2 *)(A + (size_t)col * m);
   double s0 = 0, s1 = 0;
   for (int i = lane; i < m / 2; i += 32)
   {
      float2 v = a[i];
      s0 = fma((double)v.x, __ldg(x + 2 * i), s0);
      s1 = fma((double)v.y, __ldg(x + 2 * i + 1), s1);
   }
   s0 = wsum(s0 + s1);
   if (!lane)
      y[col] = s0;
}
// T3: warp per CW columns, x in shared memory (converted once per CTA), float4 per lane
template <int CW>
__global__ void __launch_bounds__(256) t3(int m, int n, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ y)
{
   extern __shared__ double sx[];
   for (int i = threadIdx.x; i < m; i += blockDim.x)
      sx[i] = x[i];
   __syncthreads();
   int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   int nw = (gridDim.x * blockDim.x) >> 5;
   const int m4 = m >> 2;
   for (int c0 = w * CW; c0 < n; c0 += nw * CW)
   {
      double s[CW];
#pragma unroll
      for (int c = 0; c < CW; c++)
         s[c] = 0.0;
      for (int i = lane; i < m4; i += 32)
      {
         const double2 xa = *(const double2 *)(sx + 4 * i), xb = *(const double2 *)(sx + 4 * i + 2);
#pragma unroll
         for (int c = 0; c < CW; c++)
            if (c0 + c < n)
            {
               const float4 v = __ldcs((const float4 *)(A + (size_t)(c0 + c) * m) + i);
               s[c] = fma((double)v.x, xa.x, s[c]);
               s[c] = fma((double)v.y, xa.y, s[c]);
               s[c] = fma((double)v.z, xb.x, s[c]);
               s[c] = fma((double)v.w, xb.y, s[c]);
            }
      }
#pragma unroll
      for (int c = 0; c < CW; c++)
      {
         double r = wsum(s[c]);
         if (!lane && c0 + c < n)
            y[c0 + c] = r;
      }
   }
}
// T4: as T3 but the products in fp32 pairs: (a.x * x0 + a.y * x1) summed in fp32 per 4 rows, then added in fp64
template <int CW>
__global__ void __launch_bounds__(256) t4(int m, int n, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ y)
{
   extern __shared__ double sxd[];
   float *sx = (float *)sxd;
   for (int i = threadIdx.x; i < m; i += blockDim.x)
      sx[i] = (float)x[i];
   __syncthreads();
   int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   int nw = (gridDim.x * blockDim.x) >> 5;
   const int m4 = m >> 2;
   for (int c0 = w * CW; c0 < n; c0 += nw * CW)
   {
      double s[CW];
#pragma unroll
      for (int c = 0; c < CW; c++)
         s[c] = 0.0;
      for (int i = lane; i < m4; i += 32)
      {
         const float4 xv = *(const float4 *)(sx + 4 * i);
#pragma unroll
         for (int c = 0; c < CW; c++)
            if (c0 + c < n)
            {
               const float4 v = __ldcs((const float4 *)(A + (size_t)(c0 + c) * m) + i);
               s[c] += (double)(v.x * xv.x + v.y * xv.y + v.z * xv.z + v.w * xv.w);
            }
      }
#pragma unroll
      for (int c = 0; c < CW; c++)
      {
         double r = wsum(s[c]);
         if (!lane && c0 + c < n)
            y[c0 + c] = r;
      }
   }
}
// N1: CTA = 512 columns, thread = 2 rows (float2), 2 columns unrolled
#define NC 512
__global__ void __launch_bounds__(256) n1(int m, int n, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ partial)
{
   __shared__ double sx[NC];
   int j0 = blockIdx.x * NC, jn = min(NC, n - j0);
   for (int j = threadIdx.x; j < jn; j += blockDim.x)
      sx[j] = x[j0 + j];
   __syncthreads();
   for (int i = threadIdx.x; i < m / 2; i += blockDim.x)
   {
      const float2 *a = (const float2 *)(A + (size_t)j0 * m) + i;
      double s0 = 0, s1 = 0, u0 = 0, u1 = 0;
      for (int j = 0; j + 2 <= jn; j += 2)
      {
         float2 v = a[(size_t)j * (m / 2)], w = a[(size_t)(j + 1) * (m / 2)];
         s0 = fma((double)v.x, sx[j], s0);
         s1 = fma((double)v.y, sx[j], s1);
         u0 = fma((double)w.x, sx[j + 1], u0);
         u1 = fma((double)w.y, sx[j + 1], u1);
      }
      partial[(size_t)blockIdx.x * m + 2 * i] = s0 + u0;
      partial[(size_t)blockIdx.x * m + 2 * i + 1] = s1 + u1;
   }
}
// N2: thread = VR float4 row groups?  simpler: thread = 4 rows (float4), U columns unrolled
template <int U, int COLS>
__global__ void __launch_bounds__(256) n2(int m, int n, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ partial)
{
   __shared__ double sx[COLS];
   int j0 = blockIdx.x * COLS, jn = min(COLS, n - j0);
   for (int j = threadIdx.x; j < jn; j += blockDim.x)
      sx[j] = x[j0 + j];
   __syncthreads();
   const int m4 = m >> 2;
   for (int i = threadIdx.x; i < m4; i += blockDim.x)
   {
      const float4 *a = (const float4 *)(A + (size_t)j0 * m) + i;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int j = 0; j + U <= jn; j += U)
      {
         float4 v[U];
#pragma unroll
         for (int u = 0; u < U; u++)
            v[u] = __ldcs(a + (size_t)(j + u) * m4);
#pragma unroll
         for (int u = 0; u < U; u++)
         {
            const double xj = sx[j + u];
            s0 = fma((double)v[u].x, xj, s0);
            s1 = fma((double)v[u].y, xj, s1);
            s2 = fma((double)v[u].z, xj, s2);
            s3 = fma((double)v[u].w, xj, s3);
         }
      }
      double *p = partial + (size_t)blockIdx.x * m + 4 * i;
      p[0] = s0, p[1] = s1, p[2] = s2, p[3] = s3;
   }
}
// N3: products in fp32 over U columns, then added in fp64
template <int U, int COLS>
__global__ void __launch_bounds__(256) n3(int m, int n, const float *__restrict__ A, const double *__restrict__ x, double *__restrict__ partial)
{
   __shared__ float sx[COLS];
   int j0 = blockIdx.x * COLS, jn = min(COLS, n - j0);
   for (int j = threadIdx.x; j < jn; j += blockDim.x)
      sx[j] = (float)x[j0 + j];
   __syncthreads();
   const int m4 = m >> 2;
   for (int i = threadIdx.x; i < m4; i += blockDim.x)
   {
      const float4 *a = (const float4 *)(A + (size_t)j0 * m) + i;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int j = 0; j + U <= jn; j += U)
      {
         float4 v[U];
#pragma unroll
         for (int u = 0; u < U; u++)
            v[u] = __ldcs(a + (size_t)(j + u) * m4);
         float f0 = 0, f1 = 0, f2 = 0, f3 = 0;
#pragma unroll
         for (int u = 0; u < U; u++)
         {
            const float xj = sx[j + u];
            f0 = fmaf(v[u].x, xj, f0);
            f1 = fmaf(v[u].y, xj, f1);
            f2 = fmaf(v[u].z, xj, f2);
            f3 = fmaf(v[u].w, xj, f3);
         }
         s0 += f0, s1 += f1, s2 += f2, s3 += f3;
      }
      double *p = partial + (size_t)blockIdx.x * m + 4 * i;
      p[0] = s0, p[1] = s1, p[2] = s2, p[3] = s3;
   }
}
template <typename F> void timeit(const char *name, double bytes, F f)
{
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   f();
   cudaEventRecord(e0);
   for (int r = 0; r < 5; r++)
      f();
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   printf("%-44s %.3f ms  %.0f GB/s  (%s)\n", name, ms / 5, bytes / (ms / 5 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
   const int m = 4000, n = 500000;
   float *A;
   double *x, *y, *xn, *partial;
   cudaMalloc(&A, (size_t)m * n * 4);
   cudaMemset(A, 0, (size_t)m * n * 4);
   cudaMalloc(&x, m * 8);
   cudaMalloc(&y, (size_t)n * 8);
   cudaMalloc(&xn, (size_t)n * 8);
   cudaMalloc(&partial, (size_t)((n + 255) / 256) * m * 8);
   cudaMemset(x, 0, m * 8);
   cudaMemset(xn, 0, (size_t)n * 8);
   const double bytes = (double)m * n * 4;
   timeit("T1 warp/column float2", bytes, [&] { t1<<<(int)(((long long)n * 32 + 255) / 256), 256>>>(m, n, A, x, y); });
   cudaFuncSetAttribute(t3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, m * 8);
   cudaFuncSetAttribute(t3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, m * 8);
   cudaFuncSetAttribute(t3<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, m * 8);
   timeit("T3 persistent, x in smem, 1 col/warp float4", bytes, [&] { t3<1><<<148 * 4, 256, m * 8>>>(m, n, A, x, y); });
   timeit("T3 persistent, x in smem, 2 col/warp float4", bytes, [&] { t3<2><<<148 * 4, 256, m * 8>>>(m, n, A, x, y); });
   timeit("T3 persistent, x in smem, 4 col/warp float4", bytes, [&] { t3<4><<<148 * 4, 256, m * 8>>>(m, n, A, x, y); });
   timeit("T4 fp32 products, 2 col/warp", bytes, [&] { t4<2><<<148 * 4, 256, m * 8>>>(m, n, A, x, y); });
   timeit("T4 fp32 products, 4 col/warp", bytes, [&] { t4<4><<<148 * 4, 256, m * 8>>>(m, n, A, x, y); });
   timeit("N1 float2, 2 columns unrolled, 512 cols/CTA", bytes, [&] { n1<<<(n + NC - 1) / NC, 256>>>(m, n, A, xn, partial); });
   timeit("N2 float4, 4 columns unrolled, 512 cols/CTA", bytes, [&] { n2<4, 512><<<(n + 511) / 512, 256>>>(m, n, A, xn, partial); });
   timeit("N2 float4, 8 columns unrolled, 512 cols/CTA", bytes, [&] { n2<8, 512><<<(n + 511) / 512, 256>>>(m, n, A, xn, partial); });
   timeit("N2 float4, 4 columns unrolled, 256 cols/CTA", bytes, [&] { n2<4, 256><<<(n + 255) / 256, 256>>>(m, n, A, xn, partial); });
   timeit("N3 fp32 products, 4 columns, 512 cols/CTA", bytes, [&] { n3<4, 512><<<(n + 511) / 512, 256>>>(m, n, A, xn, partial); });
   timeit("N3 fp32 products, 8 columns, 512 cols/CTA", bytes, [&] { n3<8, 512><<<(n + 511) / 512, 256>>>(m, n, A, xn, partial); });
   return 0;
}
