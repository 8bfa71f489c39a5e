// 3-D window spread (adjoint NFFT, B^T) and gather (NFFT trafo, B) on the fp64 tensor cores of sm_100a.
//
// Per cell, the spread of J points is a rank-J update of the 10 x 10 x 10 footprint,
//      G[(x,y), z] += sum_j ( psi1_j[x] psi2_j[y] ) * ( a_j psi3_j[z] ),
// i.e. a GEMM  C(100 x 10) += A(100 x J) B(J x 10), and the gather is its transpose,  T(J x 10) = A' H,  f_j = T(j,:) . psi3_j.
// Both run on mma.sync.m8n8k4.f64 (DMMA.8x8x4 in SASS): measured on this pool's B200 (scripts/fp64_dmma.cu) the fp64
// tensor path and the DFMA pipe are ONE unit (36-37 vs 33 TFLOP/s, no overlap), but a DMMA carries 256 FMAs per warp
// instruction instead of 32, which removes the issue-slot and operand-delivery limits of the DFMA kernels
// (nfft3d_warp.cu: 57 % fp64-pipe activity at best).  Tiling: M = the 100 (x,y) pairs -> 13 tiles of 8 (96 % filled),
// N = z -> 2 tiles of 8 (10 of 16 columns used), K = points (4 per instruction).
//
// Streaming, work decomposition and the per-warp cp.async ring are those of nfft3d_warp.cu: one warp per contiguous range
// of cell-sorted points, chunks of <= 8 points of one cell, 240 B of taps per point read from HBM exactly once.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

#define TAPS B200_TAPS
#define NOS B200_NOS
#define ROW (3 * TAPS)
#define WS 8
#ifndef MWNS
#define MWNS 3 /* ring depth */
#endif
#ifndef MWARPS_PER_CTA
#define MWARPS_PER_CTA 4
#endif
#ifndef MCTAS_PER_SM
#define MCTAS_PER_SM 3
#endif
#define RS 36 /* row stride in shared memory (doubles): 36 mod 16 = 4, so the 4 point rows x 4 consecutive taps that a half
                 warp reads for an operand fragment fall into 16 distinct 8-byte banks (the natural stride 30 is 2-way conflicted) */
#define STAGE_D (WS * RS + WS)
#define CP16_PER_ROW (ROW * 8 / 16)
#define CP16_PER_STAGE (WS * CP16_PER_ROW)
#define MT 13 /* m-tiles: 13 * 8 = 104 >= 100 (x,y) pairs */
#define KS 25 /* gather k-steps: 25 * 4 = 100 (x,y) pairs */

__device__ __forceinline__ void m_dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int m_tap_idx(int keyfield, int lj) { return (keyfield + 28 + lj) & 63; }
__device__ __forceinline__ void m_cpa16(void *smem, const void *gmem)
{
   unsigned s = (unsigned)__cvta_generic_to_shared(smem);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void m_cpa8_zfill(void *smem, const void *gmem, bool valid)
{
   unsigned s = (unsigned)__cvta_generic_to_shared(smem);
   int sz = valid ? 8 : 0;
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void m_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void m_wait()
{
   asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// one warp copies the 8 tap rows starting at the chunk's first point (rows past the chunk are finite padding or the next
// points) and, for the spread, the chunk's a-values (zero-filled past the chunk) into one stage
__device__ __forceinline__ void m_stage_issue(double *st, const double *__restrict__ psi, const double *__restrict__ a, int4 d, int lane)
{
   const double *src = psi + (size_t)d.x * ROW;
#pragma unroll
   for (int k = 0; k < (CP16_PER_STAGE + 31) / 32; k++)
   {
      int idx = lane + 32 * k;
      if (idx < CP16_PER_STAGE)
      {
         int r = idx / CP16_PER_ROW, q = idx - r * CP16_PER_ROW;
         m_cpa16(st + r * RS + 2 * q, src + 2 * idx);
      }
   }
   if (a != nullptr && lane < WS)
   {
      bool ok = lane < d.y;
      m_cpa8_zfill(st + WS * RS + lane, a + d.x + (ok ? lane : 0), ok);
   }
}

// ---------------------------------------------------------------------------------------------------- spread
// C fragment of m-tile i, n-tile nt: rows m = 8 i + g <-> (x, y) = (m / 10, m % 10); columns z = 8 nt + 2 t + {0, 1}
__device__ __forceinline__ void m_spread_flush(double *__restrict__ grid, int key, int g, int t, double acc[MT][2][2])
{
   const int c1 = (key >> 12) & 63, c2 = (key >> 6) & 63, c3 = key & 63;
#pragma unroll
   for (int i = 0; i < MT; i++)
   {
      const int m = 8 * i + g;
      if (m < 100)
      {
         const int x = m / 10, y = m - 10 * x;
         double *rowp = grid + (size_t)(m_tap_idx(c1, x) * NOS + m_tap_idx(c2, y)) * NOS;
         atomicAdd(rowp + m_tap_idx(c3, 2 * t), acc[i][0][0]);
         atomicAdd(rowp + m_tap_idx(c3, 2 * t + 1), acc[i][0][1]);
         if (t == 0)
         {
            atomicAdd(rowp + m_tap_idx(c3, 8), acc[i][1][0]);
            atomicAdd(rowp + m_tap_idx(c3, 9), acc[i][1][1]);
         }
      }
      acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
   }
}

__global__ void __launch_bounds__(MWARPS_PER_CTA * 32, MCTAS_PER_SM)
k_spread3_mma(const double *__restrict__ psi, const double *__restrict__ a, const int4 *__restrict__ chunks,
              const int *__restrict__ cfirst, double *__restrict__ grid)
{
   __shared__ __align__(16) double smem[MWARPS_PER_CTA][MWNS][STAGE_D];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * MWARPS_PER_CTA + wid;
   const int g = lane >> 2, t = lane & 3;
   double(*st)[STAGE_D] = smem[wid];
   const int c0 = cfirst[gw], c1 = cfirst[gw + 1];
   if (c0 >= c1)
      return;

   // shared-memory offsets of this thread's A-operand factors: psi1[x(m)] and psi2[y(m)], m = 8 i + g (clamped for padding rows)
   int offx[MT], offy[MT];
#pragma unroll
   for (int i = 0; i < MT; i++)
   {
      int m = 8 * i + g;
      m = m < 100 ? m : 99;
      offx[i] = m / 10;
      offy[i] = TAPS + (m - 10 * (m / 10));
   }
   const double pad12 = (8 * (MT - 1) + g) < 100 ? 1.0 : 0.0; // tile 12 holds the pairs 96..99 in rows g < 4 only

   double acc[MT][2][2];
#pragma unroll
   for (int i = 0; i < MT; i++)
      acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;

   int4 dcur = chunks[c0];
   int key = dcur.z;
#pragma unroll
   for (int s = 0; s < MWNS - 1; s++)
   {
      if (c0 + s < c1)
         m_stage_issue(st[s], psi, a, chunks[c0 + s], lane);
      m_commit();
   }
   int4 dissue = chunks[c0 + MWNS - 1];
   for (int c = c0; c < c1; c++)
   {
      const int4 dnext = chunks[c + 1];
      const int4 dissue_next = chunks[c + MWNS];
      m_wait<MWNS - 2>();
      __syncwarp();
      if (c + MWNS - 1 < c1)
         m_stage_issue(st[(c - c0 + MWNS - 1) % MWNS], psi, a, dissue, lane);
      m_commit();

      if (dcur.z != key) // warp-uniform: the stream enters another cell -> write the footprint out
      {
         m_spread_flush(grid, key, g, t, acc);
         key = dcur.z;
      }

      const double *stage = st[(c - c0) % MWNS];
#pragma unroll
      for (int kk = 0; kk < 2; kk++)
      {
         if (kk == 1 && dcur.y <= 4) // warp-uniform: the second k-step holds only zero-weight points
            break;
         const double *row = stage + (4 * kk + t) * RS;   // B row k = t: the point of this thread
         const double aj = stage[WS * RS + 4 * kk + t];    // zero past the end of the chunk
         const double b0 = aj * row[2 * TAPS + g];
         const double b1 = (g < 2) ? aj * row[2 * TAPS + 8 + g] : 0.0;
#pragma unroll
         for (int i = 0; i < MT; i++)
         {
            double af = row[offx[i]] * row[offy[i]];
            if (i == MT - 1)
               af *= pad12;
            m_dmma(acc[i][0][0], acc[i][0][1], af, b0);
            m_dmma(acc[i][1][0], acc[i][1][1], af, b1);
         }
      }
      dcur = dnext;
      dissue = dissue_next;
   }
   m_spread_flush(grid, key, g, t, acc);
}

// ---------------------------------------------------------------------------------------------------- gather
// B operand (constant per cell): k-step s, row k = t <-> pair m = 4 s + t; column n = g <-> z = g (n-tile 0), 8 + g (n-tile 1, g < 2)
__device__ __forceinline__ void m_gather_load(const double *__restrict__ h, int key, int g, int t, double hb[KS][2])
{
   const int c1 = (key >> 12) & 63, c2 = (key >> 6) & 63, c3 = key & 63;
   const int z0 = m_tap_idx(c3, g), z1 = m_tap_idx(c3, 8 + (g & 1));
#pragma unroll
   for (int s = 0; s < KS; s++)
   {
      const int m = 4 * s + t;
      const int x = m / 10, y = m - 10 * x;
      const double *rowp = h + (size_t)(m_tap_idx(c1, x) * NOS + m_tap_idx(c2, y)) * NOS;
      hb[s][0] = __ldg(rowp + z0);
      hb[s][1] = (g < 2) ? __ldg(rowp + z1) : 0.0;
   }
}

__global__ void __launch_bounds__(MWARPS_PER_CTA * 32, MCTAS_PER_SM)
k_gather3_mma(const double *__restrict__ psi, const int4 *__restrict__ chunks, const int *__restrict__ cfirst,
              const double *__restrict__ h, double *__restrict__ f)
{
   __shared__ __align__(16) double smem[MWARPS_PER_CTA][MWNS][STAGE_D];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * MWARPS_PER_CTA + wid;
   const int g = lane >> 2, t = lane & 3;
   double(*st)[STAGE_D] = smem[wid];
   const unsigned full = 0xffffffffu;
   const int c0 = cfirst[gw], c1 = cfirst[gw + 1];
   if (c0 >= c1)
      return;

   int4 dcur = chunks[c0];
   int key = dcur.z;
#pragma unroll
   for (int s = 0; s < MWNS - 1; s++)
   {
      if (c0 + s < c1)
         m_stage_issue(st[s], psi, nullptr, chunks[c0 + s], lane);
      m_commit();
   }
   int4 dissue = chunks[c0 + MWNS - 1];

   double hb[KS][2];
   m_gather_load(h, key, g, t, hb);

   for (int c = c0; c < c1; c++)
   {
      const int4 dnext = chunks[c + 1];
      const int4 dissue_next = chunks[c + MWNS];
      m_wait<MWNS - 2>();
      __syncwarp();
      if (c + MWNS - 1 < c1)
         m_stage_issue(st[(c - c0 + MWNS - 1) % MWNS], psi, nullptr, dissue, lane);
      m_commit();

      if (dcur.z != key) // warp-uniform
      {
         key = dcur.z;
         m_gather_load(h, key, g, t, hb);
      }

      // A operand: row = point g of the chunk, column k = t <-> pair m = 4 s + t
      const double *row = st[(c - c0) % MWNS] + g * RS;
      double ca[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, cb[2][2] = {{0.0, 0.0}, {0.0, 0.0}}; // two chains per n-tile (even / odd s)
#pragma unroll
      for (int s = 0; s < KS; s++)
      {
         const int m = 4 * s + t;
         const int x = m / 10, y = m - 10 * x;
         const double af = row[x] * row[TAPS + y];
         if (s & 1)
         {
            m_dmma(cb[0][0], cb[0][1], af, hb[s][0]);
            m_dmma(cb[1][0], cb[1][1], af, hb[s][1]);
         }
         else
         {
            m_dmma(ca[0][0], ca[0][1], af, hb[s][0]);
            m_dmma(ca[1][0], ca[1][1], af, hb[s][1]);
         }
      }
      // C: row = point g; columns z = 2 t, 2 t + 1 (n-tile 0) and 8 + 2 t, 9 + 2 t (n-tile 1: only t = 0 is inside the footprint)
      double v = (ca[0][0] + cb[0][0]) * row[2 * TAPS + 2 * t];
      v = fma(ca[0][1] + cb[0][1], row[2 * TAPS + 2 * t + 1], v);
      if (t == 0)
      {
         v = fma(ca[1][0] + cb[1][0], row[2 * TAPS + 8], v);
         v = fma(ca[1][1] + cb[1][1], row[2 * TAPS + 9], v);
      }
      v += __shfl_xor_sync(full, v, 1);
      v += __shfl_xor_sync(full, v, 2);
      if (t == 0 && g < dcur.y)
         f[dcur.x + g] = v;
      dcur = dnext;
      dissue = dissue_next;
   }
   m_wait<0>();
}

// ---------------------------------------------------------------------------------------------------- launch layer
extern "C" {

int b200k_mma_warps_per_cta(void) { return MWARPS_PER_CTA; }
int b200k_mma_ctas_per_sm(void) { return MCTAS_PER_SM; }

int b200k_spread3_mma(const b200_win_dev *w)
{
   if (w->nchunk <= 0)
      return 0;
   k_spread3_mma<<<w->nwr / MWARPS_PER_CTA, MWARPS_PER_CTA * 32, 0, b200_stream()>>>(w->psi, w->a, (const int4 *)w->chunks, w->cfirst, w->grid);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather3_mma(const b200_win_dev *w, double *d_f)
{
   if (w->nchunk <= 0)
      return 0;
   k_gather3_mma<<<w->nwr / MWARPS_PER_CTA, MWARPS_PER_CTA * 32, 0, b200_stream()>>>(w->psi, (const int4 *)w->chunks, w->cfirst, w->grid, d_f);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
