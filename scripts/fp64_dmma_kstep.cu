// Micro-benchmark (development aid): the k-step of the tensor-core spread (nfft3d_mma.cu) in isolation, registers only:
// 13 "main" DMMAs sharing one B operand + 6 "edge" DMMAs (2 A x 3 B), their operand products, the boundary masks.
// 3 warps per sub-partition.  Variants isolate what keeps the fp64 unit from its DMMA-only rate (16.3 clk per instruction).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: exact k-step; 1: no masks; 2: edge replaced by 6 more main-type instructions (one B); 3: products only for the
// main part, edge operands loop-invariant; 4: no products at all (all operands loop-invariant, 4 distinct B registers);
// 5: exact k-step, two k-steps per iteration with a warp-uniform branch between them (the kernel's loop shape)
template <int MODE>
__global__ void __launch_bounds__(128, 3) k(double *out, const double *in, int iters, int brk)
{
   const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
   double acc[13][2], ace[2][3][2];
   double2 p1[5];
   double p2 = in[g], p2e = in[1 + t], p3 = in[2 + g], r1g = in[3 + g], r1e = in[4 + t], aj = in[5 + t], r3g = in[6 + g], x1[3], y2[3];
#pragma unroll
   for (int i = 0; i < 5; i++)
      p1[i] = make_double2(in[16 + 2 * i + t], in[17 + 2 * i + t]);
#pragma unroll
   for (int i = 0; i < 3; i++)
   {
      x1[i] = in[32 + i + g];
      y2[i] = in[40 + i + g];
   }
#pragma unroll
   for (int i = 0; i < 13; i++)
      acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
   for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
         ace[i][j][0] = ace[i][j][1] = 0.0;
   double b0i = in[50 + g], aei0 = in[51 + g], aei1 = in[52 + g], bei[3] = {in[53 + g], in[54 + g], in[55 + g]}, afi[13];
#pragma unroll
   for (int i = 0; i < 13; i++)
      afi[i] = in[i + t];
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int kk = 0; kk < (MODE == 5 ? 2 : 1); kk++)
      {
         if (MODE == 5 && kk == 1 && brk)
            break;
         p2 += 1e-9; // the products change every k-step
         aj += 1e-9;
         p3 += 1e-9;
         if (MODE == 4)
         {
#pragma unroll
            for (int i = 0; i < 13; i++)
               dmma(acc[i][0], acc[i][1], afi[i], b0i);
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
               dmma(ace[0][j][0], ace[0][j][1], aei0, bei[j]);
               dmma(ace[1][j][0], ace[1][j][1], aei1, bei[j]);
            }
            continue;
         }
         const double b0 = aj * r3g;
#pragma unroll
         for (int i = 0; i < 5; i++)
         {
            dmma(acc[2 * i][0], acc[2 * i][1], p1[i].x * p2, b0);
            dmma(acc[2 * i + 1][0], acc[2 * i + 1][1], p1[i].y * p2, b0);
         }
#pragma unroll
         for (int i = 0; i < 3; i++)
         {
            const double af = (MODE == 1 || i < 2 || g < 4) ? x1[i] * p2e : 0.0;
            dmma(acc[10 + i][0], acc[10 + i][1], af, b0);
         }
         if (MODE == 2)
         {
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
               dmma(ace[0][j][0], ace[0][j][1], y2[j] * p2, b0);
               dmma(ace[1][j][0], ace[1][j][1], x1[j] * p2, b0);
            }
         }
         else if (MODE == 3)
         {
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
               dmma(ace[0][j][0], ace[0][j][1], aei0, bei[j]);
               dmma(ace[1][j][0], ace[1][j][1], aei1, bei[j]);
            }
         }
         else
         {
            const double ae0 = aj * r1g;
            const double ae1 = (MODE == 1 || g < 2) ? aj * r1e : 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
               const double be = (MODE == 1 || j < 2 || g < 4) ? y2[j] * p3 : 0.0;
               dmma(ace[0][j][0], ace[0][j][1], ae0, be);
               dmma(ace[1][j][0], ace[1][j][1], ae1, be);
            }
         }
      }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 13; i++)
      s += acc[i][0] + acc[i][1];
#pragma unroll
   for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
         s += ace[i][j][0] + ace[i][j][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double *out, double *in, int brk = 0)
{
   int iters = 4000;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k<MODE><<<148 * 3, 128>>>(out, in, 10, brk);
   cudaEventRecord(e0);
   k<MODE><<<148 * 3, 128>>>(out, in, iters, brk);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   double ksteps = (MODE == 5 && !brk) ? 2.0 : 1.0;
   double clk = ms * 1e-3 * 1.965e9 / (3.0 * 19.0 * iters * ksteps);
   printf("%-60s %.3f ms  %.2f clk per DMMA and sub-partition\n", name, ms, clk);
}
int main()
{
   double *out, *in;
   cudaMalloc(&out, 148 * 3 * 128 * 8);
   cudaMalloc(&in, 128 * 8);
   cudaMemset(in, 0, 128 * 8);
   for (int pass = 0; pass < 2; pass++)
   {
      if (pass == 1)
      {
         // operand VALUES matter for the power drawn by the fp64 unit: repeat with random mantissas instead of zeros
         double h[128];
         unsigned long long st = 88172645463325252ull;
         for (int i = 0; i < 128; i++)
         {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            h[i] = 0.5 + (double)(st >> 11) / 9007199254740992.0;
         }
         cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
         printf("-- random operands in [0.5, 1.5)\n");
      }
      else
         printf("-- all operands zero (+ 1e-9 per k-step)\n");
   run<4>("no products (loop-invariant operands, 4 B registers)", out, in);
   run<3>("products for the 13 main DMMAs only", out, in);
   run<2>("19 main-type DMMAs (one B), 19 products", out, in);
   run<1>("k-step without the boundary masks", out, in);
   run<0>("exact k-step", out, in);
   run<5>("exact k-step x 2 per iteration, branch between", out, in);
   run<5>("exact k-step x 2, branch taken (1 k-step per iteration)", out, in, 1);
   }
   return 0;
}
