// Inner loop of the spread kernel in isolation: taps from shared memory (no global traffic), 40 accumulators.
// Variants: A = as in the kernel (7 LDS.128 per point), B = psi_3 taps kept in registers across points (upper bound
// without the broadcast loads), C = 64-bit loads for psi_1/psi_2.
#include <cuda_runtime.h>
#include <cstdio>
#define TAPS 10
#define ROW 30
#define WS 8

template <int VAR>
__device__ __forceinline__ void spread_point(double acc[2][2][TAPS], const double *__restrict__ r, double aj, int i1, int i2, const double p3r[TAPS])
{
   double2 p1, p2;
   if (VAR == 2)
   {
      p1 = make_double2(r[2 * i1], r[2 * i1 + 1]);
      p2 = make_double2(r[TAPS + 2 * i2], r[TAPS + 2 * i2 + 1]);
   }
   else
   {
      p1 = *reinterpret_cast<const double2 *>(r + 2 * i1);
      p2 = *reinterpret_cast<const double2 *>(r + TAPS + 2 * i2);
   }
   const double qa = aj * p1.x, qb = aj * p1.y;
   const double q00 = qa * p2.x, q01 = qa * p2.y, q10 = qb * p2.x, q11 = qb * p2.y;
#pragma unroll
   for (int z = 0; z < TAPS; z += 2)
   {
      double2 p3;
      if (VAR == 1)
         p3 = make_double2(p3r[z], p3r[z + 1]);
      else
         p3 = *reinterpret_cast<const double2 *>(r + 2 * TAPS + z);
      acc[0][0][z] = fma(q00, p3.x, acc[0][0][z]);
      acc[0][1][z] = fma(q01, p3.x, acc[0][1][z]);
      acc[1][0][z] = fma(q10, p3.x, acc[1][0][z]);
      acc[1][1][z] = fma(q11, p3.x, acc[1][1][z]);
      acc[0][0][z + 1] = fma(q00, p3.y, acc[0][0][z + 1]);
      acc[0][1][z + 1] = fma(q01, p3.y, acc[0][1][z + 1]);
      acc[1][0][z + 1] = fma(q10, p3.y, acc[1][0][z + 1]);
      acc[1][1][z + 1] = fma(q11, p3.y, acc[1][1][z + 1]);
   }
}

template <int VAR, int MINB>
__global__ void __launch_bounds__(128, MINB) k_mix(double *out, const double *in, int iters)
{
   __shared__ __align__(16) double sh[4][WS * ROW + WS];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   for (int i = lane; i < WS * ROW + WS; i += 32)
      sh[wid][i] = in[i];
   __syncwarp();
   const int i1 = (lane < 25 ? lane / 5 : 0), i2 = (lane < 25 ? lane % 5 : 0);
   double acc[2][2][TAPS];
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
#pragma unroll
         for (int z = 0; z < TAPS; z++)
            acc[x][y][z] = 0.0;
   double p3r[TAPS];
#pragma unroll
   for (int z = 0; z < TAPS; z++)
      p3r[z] = sh[wid][20 + z];
   for (int it = 0; it < iters; it++)
   {
      const double *row = sh[wid];
#pragma unroll
      for (int p = 0; p < WS; p++)
         spread_point<VAR>(acc, row + p * ROW, row[WS * ROW + p], i1, i2, p3r);
      __syncwarp();
   }
   double s = 0;
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
#pragma unroll
         for (int z = 0; z < TAPS; z++)
            s += acc[x][y][z];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VAR, int MINB>
void run(int nsm, double *out, double *in)
{
   int iters = 4000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k_mix<VAR, MINB><<<nsm * MINB, 128>>>(out, in, 10);
   cudaEventRecord(e0);
   k_mix<VAR, MINB><<<nsm * MINB, 128>>>(out, in, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   double pts = (double)nsm * MINB * 4 * (double)iters * 8;
   printf("variant %d warps/SM %2d : %.1f us per 1e6 points (whole GPU), %.2f clk per fp64 instr per SMSP\n", VAR, MINB * 4,
          ms * 1e3 * 1e6 / pts, ms * 1e-3 * 1.965e9 * nsm * 4 / (pts * 46));
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   int nsm = p.multiProcessorCount;
   double *out, *in;
   cudaMalloc(&out, 8 * 148 * 4 * 128 * 4);
   cudaMalloc(&in, 8 * 1024);
   double h[1024];
   for (int i = 0; i < 1024; i++)
      h[i] = 1.0 + i * 1e-3;
   cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
   run<0, 2>(nsm, out, in);
   run<0, 3>(nsm, out, in);
   run<0, 4>(nsm, out, in);
   run<1, 3>(nsm, out, in);
   run<1, 4>(nsm, out, in);
   run<2, 3>(nsm, out, in);
   run<2, 4>(nsm, out, in);
   return 0;
}
