// Micro-benchmark (development aid): does DMMA.8x8x4 block the sub-partition's issue port?  Per iteration and warp:
// 19 DMMAs, each followed by NI independent integer instructions (and optionally LDS); 3 warps per sub-partition.
// If the integer work overlapped with the tensor instruction the time per DMMA would stay at ~16.3 clk.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NI, int NL>
__global__ void __launch_bounds__(128, 3) k(double *out, const double *in, int iters)
{
   __shared__ double sm[512];
   for (int i = threadIdx.x; i < 512; i += blockDim.x)
      sm[i] = in[i & 63];
   __syncthreads();
   double acc[19][2];
   double a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 3)];
   int iv[8];
#pragma unroll
   for (int j = 0; j < 8; j++)
      iv[j] = threadIdx.x + j;
#pragma unroll
   for (int i = 0; i < 19; i++)
      acc[i][0] = acc[i][1] = 0.0;
   unsigned long long ld = 0ull;
   const unsigned long long *smu = reinterpret_cast<const unsigned long long *>(sm);
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < 19; i++)
      {
         dmma(acc[i][0], acc[i][1], a, b);
#pragma unroll
         for (int j = 0; j < NI; j++)
            iv[j & 7] = iv[j & 7] * 3 + it; // IMAD, 8 independent chains
#pragma unroll
         for (int j = 0; j < NL; j++)
            ld ^= smu[(threadIdx.x + 32 * j + i + it) & 511]; // conflict-free LDS.64, result consumed by the integer pipe only
      }
   }
   double s = (double)(ld & 1023ull);
#pragma unroll
   for (int i = 0; i < 19; i++)
      s += acc[i][0] + acc[i][1];
#pragma unroll
   for (int j = 0; j < 8; j++)
      s += iv[j];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NI, int NL> void run(double *out, double *in)
{
   int iters = 3000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k<NI, NL><<<148 * 3, 128>>>(out, in, 10);
   cudaEventRecord(e0);
   k<NI, NL><<<148 * 3, 128>>>(out, in, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   double clk = ms * 1e-3 * 1.965e9 / (3.0 * 19.0 * iters);
   printf("DMMA + %2d IMAD + %d LDS : %.3f ms  %.2f clk per DMMA and sub-partition (3 warps)\n", NI, NL, ms, clk);
}
int main()
{
   double *out, *in;
   cudaMalloc(&out, 148 * 3 * 128 * 8);
   cudaMalloc(&in, 64 * 8);
   cudaMemset(in, 0, 64 * 8);
   run<0, 0>(out, in);
   run<2, 0>(out, in);
   run<4, 0>(out, in);
   run<8, 0>(out, in);
   run<16, 0>(out, in);
   run<4, 1>(out, in);
   run<4, 2>(out, in);
   run<4, 4>(out, in);
   return 0;
}
