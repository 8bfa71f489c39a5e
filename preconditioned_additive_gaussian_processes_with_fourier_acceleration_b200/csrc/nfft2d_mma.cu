// 2-D window spread (adjoint NFFT, B^T) and gather (NFFT trafo, B) on the fp64 tensor cores of sm_100a.
//
// The DFMA stream kernels for 2-D windows (nfft12_stream.cu) move the tap rows at half the copy bandwidth because they sit on the
// SM's shared-memory pipe (ncu, profiles/r2_e: L1/shared 88 % of peak, 10.5 shared-memory wavefronts per point: every lane fetches
// its own operands for every point).  As small GEMMs the same sums need far fewer operand loads:
//   spread, per cell:  G(10 x 10) += sum_j (a_j psi1_j) psi2_j'   =  A(10 x J) B(J x 10):  M = x in 2 tiles, N = y in 2 tiles,
//           K = 4 points per mma.sync.m8n8k4.f64 -> 4 instructions and 5 operand loads per thread for 4 points;
//   gather, per cell:  T(J x 10) = Psi1(J x 10) H(10 x 10),  f_j = T(j, :) . psi2_j:  M = 8 points, K = x in 3 steps, N = y in 2
//           tiles -> 6 instructions and 5 operand loads per thread for 8 points, H in registers per cell.
// The tensor pipe needs 16 clk per instruction and sub-partition, i.e. 4 clk per point and SM for the spread (3 for the gather):
// 57 / 43 us per 4 10^6-point window against 105 us for its 688 MB of taps at the measured copy bandwidth -- the kernels become
// HBM-bound on the table traffic.  Work decomposition and staging are those of the stream kernels: a warp owns a contiguous range
// of cell-sorted points, 32 tap rows per step through shared memory with coalesced 128-bit loads (row stride 20 doubles = 4 mod 16:
// the 4 rows x 4 taps a half warp reads per operand load fall into 16 distinct 8-byte banks), the 64 x 64 grid is touched only when
// the stream enters another cell.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

#define TAPS B200_TAPS
#define NOS B200_NOS
#define QWARPS 8  /* warps per CTA */
#define QBATCH 32 /* points staged per warp and step */
#define QCTAS_PER_SM 4
#define QRD (2 * TAPS) /* doubles per tap row */

__device__ __forceinline__ void q_dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int q_tap_idx(int keyfield, int lj) { return (keyfield + 28 + lj) & 63; }
__device__ __forceinline__ void q_red_add(double *p, double v)
{
   asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

struct QSmem
{
   double rows[QWARPS][QBATCH][QRD];
   double a[QWARPS][QBATCH];
   int key[QWARPS][QBATCH];
};

// stage QBATCH tap rows (and optionally the coefficients) of the points [j0, j0 + QBATCH) of this warp; rows past `jend` are zero
__device__ __forceinline__ void q_stage(QSmem &sm, int wid, int lane, const double *__restrict__ psi, const double *__restrict__ a,
                                        const int *__restrict__ key, int j0, int jend)
{
   constexpr int V2 = QBATCH * QRD / 2; // double2 per batch
   const double2 *src = reinterpret_cast<const double2 *>(psi + (size_t)j0 * QRD);
   double2 *dst = reinterpret_cast<double2 *>(&sm.rows[wid][0][0]);
   const int valid2 = (jend - j0) * QRD / 2;
#pragma unroll
   for (int k = 0; k < V2 / 32; k++)
   {
      const int idx = lane + 32 * k;
      dst[idx] = idx < valid2 ? __ldg(src + idx) : make_double2(0.0, 0.0);
   }
   const int j = j0 + lane;
   if (a != nullptr)
      sm.a[wid][lane] = j < jend ? a[j] : 0.0;
   sm.key[wid][lane] = j < jend ? key[j] : -1;
}

// ---------------------------------------------------------------------------------------------------- spread
// accumulator tile (tx, ty): rows x = 8 tx + g, columns y = 8 ty + 2 t + {0, 1}
__device__ __forceinline__ void q_flush(double c[2][2][2], double *__restrict__ grid, int key, int g, int t)
{
   const int c1 = (key >> 6) & 63, c2 = key & 63;
#pragma unroll
   for (int tx = 0; tx < 2; tx++)
#pragma unroll
      for (int ty = 0; ty < 2; ty++)
#pragma unroll
         for (int e = 0; e < 2; e++)
         {
            const int x = 8 * tx + g, y = 8 * ty + 2 * t + e;
            if (x < TAPS && y < TAPS)
               q_red_add(grid + q_tap_idx(c1, x) * NOS + q_tap_idx(c2, y), c[tx][ty][e]);
            c[tx][ty][e] = 0.0;
         }
}

// one k-step: the four points rows[4 ks .. 4 ks + 3]; only = -1: all four, else only point `only` of the step carries weight
__device__ __forceinline__ void q_spread_step(const double (*rows)[QRD], const double *av, int ks, int g, int t, int only, double c[2][2][2])
{
   const double *row = rows[4 * ks + t];
   double aj = av[4 * ks + t];
   if (only >= 0 && t != only)
      aj = 0.0;
   const double a0 = aj * row[g];
   const double a1v = aj * row[8 + (g & 1)];
   const double b0 = row[TAPS + g];
   const double b1v = row[TAPS + 8 + (g & 1)];
   const double a1 = g < 2 ? a1v : 0.0, b1 = g < 2 ? b1v : 0.0;
   q_dmma(c[0][0][0], c[0][0][1], a0, b0);
   q_dmma(c[0][1][0], c[0][1][1], a0, b1);
   q_dmma(c[1][0][0], c[1][0][1], a1, b0);
   q_dmma(c[1][1][0], c[1][1][1], a1, b1);
}

__global__ void __launch_bounds__(QWARPS * 32, QCTAS_PER_SM)
k_spread2_mma(const double *__restrict__ psi, const int *__restrict__ key, const double *__restrict__ a, double *__restrict__ grid, int n,
              int per_warp)
{
   __shared__ __align__(16) QSmem sm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int g = lane >> 2, t = lane & 3;
   const long long b64 = (long long)(blockIdx.x * QWARPS + wid) * per_warp;
   if (b64 >= n)
      return;
   const int jb = (int)b64, je = min(n, jb + per_warp);
   double c[2][2][2];
#pragma unroll
   for (int i = 0; i < 8; i++)
      (&c[0][0][0])[i] = 0.0;
   int cur = key[jb];
   for (int j0 = jb; j0 < je; j0 += QBATCH)
   {
      __syncwarp();
      q_stage(sm, wid, lane, psi, a, key, j0, je);
      __syncwarp();
      const int cnt = min(QBATCH, je - j0);
      const int mykey = sm.key[wid][lane];
      // the common case: the whole batch lies in the current cell (rows past the end of the range are zero)
      if (__all_sync(0xffffffffu, mykey == cur || lane >= cnt))
      {
#pragma unroll
         for (int ks = 0; ks < QBATCH / 4; ks++)
            q_spread_step(sm.rows[wid], sm.a[wid], ks, g, t, -1, c);
         continue;
      }
      for (int p = 0; p < cnt; p++)
      {
         const int kp = sm.key[wid][p];
         if (kp != cur) // warp-uniform
         {
            q_flush(c, grid, cur, g, t);
            cur = kp;
         }
         q_spread_step(sm.rows[wid], sm.a[wid], p >> 2, g, t, p & 3, c);
      }
   }
   q_flush(c, grid, cur, g, t);
}

// ---------------------------------------------------------------------------------------------------- gather
// B operands of the current cell: k-step s (x = 4 s + t), tile ty (y = 8 ty + g)
__device__ __forceinline__ void q_gather_load(const double *__restrict__ h, int key, int g, int t, double hb[3][2])
{
   const int c1 = (key >> 6) & 63, c2 = key & 63;
#pragma unroll
   for (int s = 0; s < 3; s++)
#pragma unroll
      for (int ty = 0; ty < 2; ty++)
      {
         const int x = 4 * s + t, y = 8 * ty + g;
         hb[s][ty] = (x < TAPS && y < TAPS) ? __ldg(h + q_tap_idx(c1, x) * NOS + q_tap_idx(c2, y)) : 0.0;
      }
}

// the eight points rows[8 q .. 8 q + 7]: returns in the lanes with t == 0 the value of point 8 q + g
__device__ __forceinline__ double q_gather_group(const double (*rows)[QRD], int q, int g, int t, const double hb[3][2])
{
   const double *row = rows[8 * q + g];
   double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
   for (int s = 0; s < 3; s++)
   {
      const int x = 4 * s + t;
      const double ax = x < TAPS ? row[x] : 0.0;
      q_dmma(c0[0], c0[1], ax, hb[s][0]);
      q_dmma(c1[0], c1[1], ax, hb[s][1]);
   }
   // C tile ty: row = point g, columns y = 8 ty + 2 t + {0, 1}
   const double2 p2 = *reinterpret_cast<const double2 *>(row + TAPS + 2 * t);
   double v = fma(c0[1], p2.y, c0[0] * p2.x);
   if (t == 0)
   {
      const double2 p8 = *reinterpret_cast<const double2 *>(row + TAPS + 8);
      v = fma(c1[1], p8.y, fma(c1[0], p8.x, v));
   }
   v += __shfl_xor_sync(0xffffffffu, v, 1);
   v += __shfl_xor_sync(0xffffffffu, v, 2);
   return v;
}

__global__ void __launch_bounds__(QWARPS * 32, QCTAS_PER_SM)
k_gather2_mma(const double *__restrict__ psi, const int *__restrict__ key, const double *__restrict__ h, double *__restrict__ f, int n,
              int per_warp)
{
   __shared__ __align__(16) QSmem sm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int g = lane >> 2, t = lane & 3;
   const long long b64 = (long long)(blockIdx.x * QWARPS + wid) * per_warp;
   if (b64 >= n)
      return;
   const int jb = (int)b64, je = min(n, jb + per_warp);
   double hb[3][2];
   int cur = key[jb];
   q_gather_load(h, cur, g, t, hb);
   for (int j0 = jb; j0 < je; j0 += QBATCH)
   {
      __syncwarp();
      q_stage(sm, wid, lane, psi, nullptr, key, j0, je);
      __syncwarp();
      const int cnt = min(QBATCH, je - j0);
      const int mykey = sm.key[wid][lane];
      if (__all_sync(0xffffffffu, mykey == cur || lane >= cnt))
      {
#pragma unroll
         for (int q = 0; q < QBATCH / 8; q++)
         {
            const double v = q_gather_group(sm.rows[wid], q, g, t, hb);
            if (t == 0 && 8 * q + g < cnt)
               f[j0 + 8 * q + g] = v;
         }
         continue;
      }
      // a cell boundary inside the batch: one point at a time (its group of eight is evaluated with the point's cell)
      for (int p = 0; p < cnt; p++)
      {
         const int kp = sm.key[wid][p];
         if (kp != cur) // warp-uniform
         {
            cur = kp;
            q_gather_load(h, cur, g, t, hb);
         }
         const double v = q_gather_group(sm.rows[wid], p >> 3, g, t, hb);
         if (t == 0 && g == (p & 7))
            f[j0 + p] = v;
      }
   }
}

// ---------------------------------------------------------------------------------------------------- launch layer
static int q_grid(int n, int *per_warp)
{
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   long long warps = (long long)sm * QCTAS_PER_SM * QWARPS;
   long long per = (n + warps - 1) / warps;
   per = ((per + QBATCH - 1) / QBATCH) * QBATCH; // whole batches: staged rows start 16-byte aligned
   if (per < QBATCH)
      per = QBATCH;
   *per_warp = (int)per;
   long long used = (n + per - 1) / per;
   return (int)((used + QWARPS - 1) / QWARPS);
}

extern "C" {

int b200k_spread2_mma(const b200_win_dev *w)
{
   if (w->n <= 0)
      return 0;
   int per, blocks = q_grid(w->n, &per);
   k_spread2_mma<<<blocks, QWARPS * 32, 0, b200_stream()>>>(w->psi, w->key, w->a, w->grid, w->n, per);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather2_mma(const b200_win_dev *w, double *d_f)
{
   if (w->n <= 0)
      return 0;
   int per, blocks = q_grid(w->n, &per);
   k_gather2_mma<<<blocks, QWARPS * 32, 0, b200_stream()>>>(w->psi, w->key, w->grid, d_f, w->n, per);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
