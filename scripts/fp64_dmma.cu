// Micro-benchmark (development aid): fp64 throughput of DFMA alone, DMMA.8x8x4 alone, and both interleaved, per SM.
// Question: do the fp64 tensor-core path and the fp64 FMA pipe of sm_100a overlap?   nvcc -arch=sm_100a -O3 fp64_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
   double f[16], m[16][2];
#pragma unroll
   for (int i = 0; i < 16; i++)
   {
      f[i] = threadIdx.x + i;
      m[i][0] = i;
      m[i][1] = -i;
   }
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < 16; i++)
      {
         if (MODE & 1)
            f[i] = fma(f[i], a, b);
         if (MODE & 2)
            dmma(m[i][0], m[i][1], a, b);
      }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 16; i++)
      s += f[i] + m[i][0] + m[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, double *out, int warps_per_sm)
{
   int iters = 4000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   int blocks = 148 * warps_per_sm / 8;
   k<MODE><<<blocks, 256>>>(out, 10, 1.0000001, 1e-9);
   cudaEventRecord(e0);
   k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   double warps = (double)blocks * 8;
   double fma_flops = (MODE & 1) ? warps * 32 * 16.0 * iters * 2 : 0;
   double mma_flops = (MODE & 2) ? warps * 16.0 * iters * 512 : 0; // 8*8*4*2 flop per warp-level mma
   printf("%-12s warps/SM %2d: %.3f ms  DFMA %.2f TF/s  DMMA %.2f TF/s  total %.2f TF/s\n", name, warps_per_sm, ms, fma_flops / ms / 1e9,
          mma_flops / ms / 1e9, (fma_flops + mma_flops) / ms / 1e9);
}
int main()
{
   double *out;
   cudaMalloc(&out, 148 * 64 * 256 * 8);
   for (int w : {8, 16, 32})
   {
      run<1>("DFMA", out, w);
      run<2>("DMMA", out, w);
      run<3>("DFMA+DMMA", out, w);
   }
   return 0;
}
