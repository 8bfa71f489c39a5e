// Micro-benchmark (development aid): the instruction pattern of the tensor-core spread in isolation -- per "k-step" 19 DMULs
// that form the A operands followed by / interleaved with 19 DMMA.8x8x4 on 19 accumulator pairs -- with 3 warps per
// sub-partition, no memory traffic.  Prints clocks per (DMUL + DMMA) pair and sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE> // 0: DMMA only (operands loop-invariant), 1: DMUL right before its DMMA, 2: 19 DMULs then 19 DMMAs, 3: DMUL only
__global__ void __launch_bounds__(128, 3) k(double *out, const double *in, int iters)
{
   double acc[19][2];
   double p[10], q = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 3)];
#pragma unroll
   for (int i = 0; i < 10; i++)
      p[i] = in[16 + i + (threadIdx.x & 3)];
#pragma unroll
   for (int i = 0; i < 19; i++)
      acc[i][0] = acc[i][1] = 0.0;
   for (int it = 0; it < iters; it++)
   {
      q += 1e-9; // products change every iteration
      if (MODE == 0)
      {
#pragma unroll
         for (int i = 0; i < 19; i++)
            dmma(acc[i][0], acc[i][1], p[i % 10], b);
      }
      else if (MODE == 1)
      {
#pragma unroll
         for (int i = 0; i < 19; i++)
            dmma(acc[i][0], acc[i][1], p[i % 10] * q, b);
      }
      else if (MODE == 2)
      {
         double af[19];
#pragma unroll
         for (int i = 0; i < 19; i++)
            af[i] = p[i % 10] * (q + i);
#pragma unroll
         for (int i = 0; i < 19; i++)
            dmma(acc[i][0], acc[i][1], af[i], b);
      }
      else
      {
#pragma unroll
         for (int i = 0; i < 19; i++)
            acc[i][0] = fma(p[i % 10], q, acc[i][0]);
      }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 19; i++)
      s += acc[i][0] + acc[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double *out, double *in)
{
   int iters = 4000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k<MODE><<<148 * 3, 128>>>(out, in, 10);
   cudaEventRecord(e0);
   k<MODE><<<148 * 3, 128>>>(out, in, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   // 3 warps per sub-partition, 19 pairs per iteration
   double clk = ms * 1e-3 * 1.965e9 / (3.0 * 19.0 * iters);
   printf("%-34s %.3f ms  %.2f clk per pair and sub-partition\n", name, ms, clk);
}
int main()
{
   double *out, *in;
   cudaMalloc(&out, 148 * 3 * 128 * 8);
   cudaMalloc(&in, 64 * 8);
   cudaMemset(in, 0, 64 * 8);
   run<0>("DMMA only", out, in);
   run<1>("DMUL right before its DMMA", out, in);
   run<2>("19 DMULs, then 19 DMMAs", out, in);
   run<3>("DFMA only (19 per iteration)", out, in);
   return 0;
}
