#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double *out, int iters, double a, double b)
{
   double m[CH][2];
#pragma unroll
   for (int i = 0; i < CH; i++) { m[i][0] = i; m[i][1] = -i; }
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < CH; i++) dmma(m[i][0], m[i][1], a, b);
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < CH; i++) s += m[i][0] + m[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> void run(double *out, int wps)
{
   int iters = 20000 / CH * 16;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   int threads = 32 * wps; // one CTA per SM, wps warps
   k<CH><<<148, threads>>>(out, 10, 1.0000001, 1e-9);
   cudaEventRecord(e0);
   k<CH><<<148, threads>>>(out, iters, 1.0000001, 1e-9);
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   double flops = 148.0 * wps * CH * (double)iters * 512;
   printf("chains %2d warps/SM %2d: %.2f TF/s\n", CH, wps, flops / ms / 1e9);
}
int main()
{
   double *out; cudaMalloc(&out, 148 * 1024 * 8);
   for (int w : {4, 8, 12, 16}) { run<1>(out, w); run<2>(out, w); run<4>(out, w); run<16>(out, w); }
   return 0;
}
