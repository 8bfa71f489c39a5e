// Which rate can the fp64 pipe sustain for the spread kernel's inner instruction mix (6 DMUL + 40 DFMA per point,
// three register operands, 40 accumulators)?  A: operands in registers.  B: operands re-loaded from shared memory
// per point (14 doubles, as the kernel does).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>

template <bool LDS>
__global__ void __launch_bounds__(128, 3) k_mix(double *out, const double *in, int iters)
{
   __shared__ __align__(16) double sh[4][8 * 30];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   for (int i = lane; i < 240; i += 32)
      sh[wid][i] = in[i] + lane * 1e-9;
   __syncwarp();
   const int i1 = (lane < 25 ? lane / 5 : 0), i2 = (lane < 25 ? lane % 5 : 0);
   double acc[2][2][10];
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
#pragma unroll
         for (int z = 0; z < 10; z++)
            acc[x][y][z] = 0.0;
   double2 p1r = *reinterpret_cast<const double2 *>(&sh[wid][2 * i1]);
   double2 p2r = *reinterpret_cast<const double2 *>(&sh[wid][10 + 2 * i2]);
   double p3r[10];
#pragma unroll
   for (int z = 0; z < 10; z++)
      p3r[z] = sh[wid][20 + z];
   double aj = in[0];
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int p = 0; p < 8; p++)
      {
         const double *r = &sh[wid][p * 30];
         double2 p1 = p1r, p2 = p2r;
         if (LDS)
         {
            p1 = *reinterpret_cast<const double2 *>(r + 2 * i1);
            p2 = *reinterpret_cast<const double2 *>(r + 10 + 2 * i2);
         }
         const double qa = aj * p1.x, qb = aj * p1.y;
         const double q00 = qa * p2.x, q01 = qa * p2.y, q10 = qb * p2.x, q11 = qb * p2.y;
#pragma unroll
         for (int z = 0; z < 10; z += 2)
         {
            double2 p3 = make_double2(p3r[z], p3r[z + 1]);
            if (LDS)
               p3 = *reinterpret_cast<const double2 *>(r + 20 + z);
            acc[0][0][z] = fma(q00, p3.x, acc[0][0][z]);
            acc[0][1][z] = fma(q01, p3.x, acc[0][1][z]);
            acc[1][0][z] = fma(q10, p3.x, acc[1][0][z]);
            acc[1][1][z] = fma(q11, p3.x, acc[1][1][z]);
            acc[0][0][z + 1] = fma(q00, p3.y, acc[0][0][z + 1]);
            acc[0][1][z + 1] = fma(q01, p3.y, acc[0][1][z + 1]);
            acc[1][0][z + 1] = fma(q10, p3.y, acc[1][0][z + 1]);
            acc[1][1][z + 1] = fma(q11, p3.y, acc[1][1][z + 1]);
         }
         if (!LDS)
            aj += 1e-12; // keep the 6 DMULs inside the loop
      }
   }
   double s = 0;
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
#pragma unroll
         for (int z = 0; z < 10; z++)
            s += acc[x][y][z];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool LDS>
void run(int ctas_per_sm, int nsm, double *out, double *in)
{
   int iters = 4000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k_mix<LDS><<<nsm * ctas_per_sm, 128>>>(out, in, 10);
   cudaEventRecord(e0);
   k_mix<LDS><<<nsm * ctas_per_sm, 128>>>(out, in, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   double warp_instr = (double)nsm * ctas_per_sm * 4 * (double)iters * 8 * 46;
   printf("LDS %d  warps/SM %2d : %.1f fp64 warp-instr/us/SM -> %.2f clk per fp64 warp-instr per SMSP (1965 MHz), points/s/SM %.3g\n", (int)LDS,
          ctas_per_sm * 4, warp_instr / (ms * 1e3) / nsm, 1965.0 * 4 / (warp_instr / (ms * 1e3) / nsm), warp_instr / 46 / (ms * 1e-3) / nsm);
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
   for (int c = 1; c <= 3; c++)
      run<false>(c, nsm, out, in);
   for (int c = 1; c <= 3; c++)
      run<true>(c, nsm, out, in);
   return 0;
}
