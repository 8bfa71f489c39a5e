// Measures the fp64 FMA pipe of the device: DFMA throughput vs resident warps per SM and per-thread ILP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
   double acc[ILP];
#pragma unroll
   for (int i = 0; i < ILP; i++)
      acc[i] = threadIdx.x * 1e-3 + i;
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < ILP; i++)
         acc[i] = fma(acc[i], a, b);
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < ILP; i++)
      s += acc[i];
   if (s == 12345.678)
      out[0] = s;
}

template <int ILP>
void run(int threads, int blocks_per_sm, int nsm, double *d)
{
   int iters = 20000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k_dfma<ILP><<<nsm * blocks_per_sm, threads>>>(d, 100, 1.0000001, 1e-9);
   cudaEventRecord(e0);
   k_dfma<ILP><<<nsm * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   double fma = (double)nsm * blocks_per_sm * threads * (double)iters * ILP;
   printf("ILP %2d threads/SM %4d : %7.2f TFLOP/s  (%.1f DFMA/clk/SM at 1965 MHz)\n", ILP, threads * blocks_per_sm, 2 * fma / ms / 1e9,
          fma / (ms * 1e-3) / nsm / 1.965e9);
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   int nsm = p.multiProcessorCount;
   printf("%s, %d SMs\n", p.name, nsm);
   double *d;
   cudaMalloc(&d, 8);
   run<1>(128, 1, nsm, d);
   run<4>(128, 1, nsm, d);
   run<16>(128, 1, nsm, d);
   run<1>(160, 2, nsm, d);
   run<4>(160, 2, nsm, d);
   run<8>(160, 2, nsm, d);
   run<16>(160, 2, nsm, d);
   run<50>(160, 2, nsm, d);
   run<4>(256, 2, nsm, d);
   run<16>(256, 2, nsm, d);
   run<4>(256, 4, nsm, d);
   run<16>(256, 4, nsm, d);
   run<2>(1024, 2, nsm, d);
   run<8>(1024, 2, nsm, d);
   return 0;
}
