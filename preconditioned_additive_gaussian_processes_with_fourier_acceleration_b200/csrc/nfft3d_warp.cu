// Streaming warp-per-range 3-D window spread (adjoint NFFT, B^T) and gather (NFFT trafo, B) for sm_100a.
//
// Work decomposition (built once per handle in nfft_host.c): the cell-sorted point array (sort key = packed cell index,
// c3 fastest) is cut into `nwr` contiguous RANGES of equal point count -- one per resident warp of one persistent wave.
// Every range is cut at cell boundaries and then into CHUNKS of <= 8 consecutive points of ONE oversampled-grid cell; a
// chunk descriptor {first point, count, cell key} is what the device walks, so the inner loop has no data-dependent branch.
// A warp streams the tap rows of its points (240 B/point, the PRE_PSI analogue, read from HBM once per direction)
// through a per-warp cp.async ring and never synchronises with other warps (__syncwarp only).
//
// Lane L < 25 owns the 2 x 2 x 10 block  l1 in {2 i1, 2 i1+1}, l2 in {2 i2, 2 i2+1}, all l3  (i1 = L / 5, i2 = L % 5) of the
// current cell's 10^3 window footprint in registers: 40 fp64 accumulators (spread) / 40 values of Re h (gather).
// Per point: 14 taps from shared memory feed 6 DMUL + 40 DFMA -- fp64-pipe-bound (SURVEY.md finding 4).
//
// Footprint traffic with the L2-resident grid was the fixed cost that dominated earlier versions (n-sweep: ~85 us per
// 10^7 REDs / loads, independent of n).  Consecutive cells of a range are neighbours along l3, so the register tile is
// SHIFTED instead of flushed when the stream moves from cell c3 to c3 + d (d <= 9): only the d planes that leave the
// footprint are written (spread: 100 REDs per plane) or enter it (gather: 100 loads per plane).
//
// History with ncu evidence under profiles/: (1) 160-thread CTAs, 8 segments side by side, 20 threads each: 34% of the
// warp time at block barriers; (2) warp-per-segment with size-sorted round robin: 23% occupancy lost in the tail of the
// single wave and 1000 REDs per segment; (3) cell-change test inside the unrolled point loop: register spills.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

#define TAPS B200_TAPS
#define NOS B200_NOS
#define ROW (3 * TAPS)
#define WS 8                       /* points per stage = max points per chunk */
#ifndef WNS
#define WNS 3                      /* ring depth                              */
#endif
#ifndef WARPS_PER_CTA
#define WARPS_PER_CTA 4
#endif
#ifndef CTAS_PER_SM
#define CTAS_PER_SM 3
#endif
#define STAGE_D (WS * ROW + WS)    /* doubles per stage: tap rows + a-values  */
#define CP16_PER_STAGE (WS * ROW * 8 / 16)

__device__ __forceinline__ int w_tap_idx(int keyfield, int lj) { return (keyfield + 28 + lj) & 63; }

__device__ __forceinline__ void w_cpa16(void *smem, const void *gmem)
{
   unsigned s = (unsigned)__cvta_generic_to_shared(smem);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void w_cpa8_zfill(void *smem, const void *gmem, bool valid)
{
   unsigned s = (unsigned)__cvta_generic_to_shared(smem);
   int sz = valid ? 8 : 0;
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void w_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void w_wait()
{
   asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// one warp issues the copies of a chunk (d.x = first point, d.y = count) into one stage.  Always 8 tap rows are copied:
// rows past the chunk belong to following points or to the zeroed padding (finite values); their a-values are zero-filled.
__device__ __forceinline__ void w_stage_issue(double *st, const double *__restrict__ psi, const double *__restrict__ a, int4 d, int lane)
{
   const double *src = psi + (size_t)d.x * ROW;
#pragma unroll
   for (int k = 0; k < (CP16_PER_STAGE + 31) / 32; k++)
   {
      int idx = lane + 32 * k;
      if (idx < CP16_PER_STAGE)
         w_cpa16(st + 2 * idx, src + 2 * idx);
   }
   if (a != nullptr && lane < WS)
   {
      bool ok = lane < d.y;
      w_cpa8_zfill(st + WS * ROW + lane, a + d.x + (ok ? lane : 0), ok);
   }
}

// ---------------------------------------------------------------------------------------------------- spread
__device__ __forceinline__ void spread_flush_plane(double *__restrict__ grid, int key, int i1, int i2, int z, const double v[2][2])
{
   const int c1 = (key >> 12) & 63, c2 = (key >> 6) & 63, c3 = key & 63;
   const int g3 = w_tap_idx(c3, z);
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
         atomicAdd(grid + (size_t)(w_tap_idx(c1, 2 * i1 + x) * NOS + w_tap_idx(c2, 2 * i2 + y)) * NOS + g3, v[x][y]);
}

__device__ __forceinline__ void spread_point(double acc[2][2][TAPS], const double *__restrict__ r, double aj, int i1, int i2)
{
   const double2 p1 = *reinterpret_cast<const double2 *>(r + 2 * i1);
   const double2 p2 = *reinterpret_cast<const double2 *>(r + TAPS + 2 * i2);
   const double qa = aj * p1.x, qb = aj * p1.y;
   const double q00 = qa * p2.x, q01 = qa * p2.y, q10 = qb * p2.x, q11 = qb * p2.y;
#pragma unroll
   for (int z = 0; z < TAPS; z += 2)
   {
      const double2 p3 = *reinterpret_cast<const double2 *>(r + 2 * TAPS + z);
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

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
k_spread3_warp(const double *__restrict__ psi, const double *__restrict__ a, const int4 *__restrict__ chunks,
               const int *__restrict__ cfirst, double *__restrict__ grid)
{
   __shared__ __align__(16) double smem[WARPS_PER_CTA][WNS][STAGE_D];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * WARPS_PER_CTA + wid;
   const bool act = lane < 25;
   const int i1 = act ? lane / 5 : 0, i2 = act ? lane % 5 : 0;
   double(*st)[STAGE_D] = smem[wid];
   const int c0 = cfirst[gw], c1 = cfirst[gw + 1];
   if (c0 >= c1)
      return;

   double acc[2][2][TAPS];
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
#pragma unroll
         for (int z = 0; z < TAPS; z++)
            acc[x][y][z] = 0.0;

   int4 dcur = chunks[c0]; // chunk computed in this iteration
   int key = dcur.z;
#pragma unroll
   for (int s = 0; s < WNS - 1; s++)
   {
      if (c0 + s < c1)
         w_stage_issue(st[s], psi, a, chunks[c0 + s], lane);
      w_commit();
   }
   int4 dissue = chunks[c0 + WNS - 1]; // chunk staged in this iteration (descriptor loaded one iteration ahead)
   for (int c = c0; c < c1; c++)
   {
      const int4 dnext = chunks[c + 1]; // (padded) look-ahead loads, consumed next iteration
      const int4 dissue_next = chunks[c + WNS];
      w_wait<WNS - 2>();
      __syncwarp();
      if (c + WNS - 1 < c1)
         w_stage_issue(st[(c - c0 + WNS - 1) % WNS], psi, a, dissue, lane);
      w_commit();

      if (dcur.z != key) // warp-uniform: the stream enters another cell
      {
         const int knew = dcur.z, d = knew - key;
         if ((knew >> 6) == (key >> 6) && d > 0 && d < TAPS)
         {
            // same (c1, c2) row, c3 advanced by d: write the d planes that leave the footprint, shift the register tile
            for (int t = 0; t < d; t++)
            {
               if (act)
               {
                  const double v[2][2] = {{acc[0][0][0], acc[0][1][0]}, {acc[1][0][0], acc[1][1][0]}};
                  spread_flush_plane(grid, key + t, i1, i2, 0, v);
               }
#pragma unroll
               for (int x = 0; x < 2; x++)
#pragma unroll
                  for (int y = 0; y < 2; y++)
                  {
#pragma unroll
                     for (int z = 0; z < TAPS - 1; z++)
                        acc[x][y][z] = acc[x][y][z + 1];
                     acc[x][y][TAPS - 1] = 0.0;
                  }
            }
         }
         else
         {
#pragma unroll
            for (int z = 0; z < TAPS; z++)
            {
               if (act)
               {
                  const double v[2][2] = {{acc[0][0][z], acc[0][1][z]}, {acc[1][0][z], acc[1][1][z]}};
                  spread_flush_plane(grid, key, i1, i2, z, v);
               }
#pragma unroll
               for (int x = 0; x < 2; x++)
#pragma unroll
                  for (int y = 0; y < 2; y++)
                     acc[x][y][z] = 0.0;
            }
         }
         key = knew;
      }

      const double *row = st[(c - c0) % WNS];
      if (dcur.y == WS) // warp-uniform: full chunk, straight-line code so that the points overlap in the pipes
      {
#pragma unroll
         for (int p = 0; p < WS; p++)
            spread_point(acc, row + p * ROW, row[WS * ROW + p], i1, i2);
      }
      else // short chunk (end of a cell run; sparse cells)
      {
#pragma unroll 1
         for (int p = 0; p < dcur.y; p++)
            spread_point(acc, row + p * ROW, row[WS * ROW + p], i1, i2);
      }
      dcur = dnext;
      dissue = dissue_next;
   }
   if (act)
   {
#pragma unroll
      for (int z = 0; z < TAPS; z++)
      {
         const double v[2][2] = {{acc[0][0][z], acc[0][1][z]}, {acc[1][0][z], acc[1][1][z]}};
         spread_flush_plane(grid, key, i1, i2, z, v);
      }
   }
}

// ---------------------------------------------------------------------------------------------------- gather
__device__ __forceinline__ void gather_load_plane(const double *__restrict__ h, int key, int i1, int i2, int z, bool act, double v[2][2])
{
   const int c1 = (key >> 12) & 63, c2 = (key >> 6) & 63, c3 = key & 63;
   const int g3 = w_tap_idx(c3, z);
#pragma unroll
   for (int x = 0; x < 2; x++)
#pragma unroll
      for (int y = 0; y < 2; y++)
         v[x][y] = act ? __ldg(h + (size_t)(w_tap_idx(c1, 2 * i1 + x) * NOS + w_tap_idx(c2, 2 * i2 + y)) * NOS + g3) : 0.0;
}

__device__ __forceinline__ double gather_point(const double hr[2][2][TAPS], const double *__restrict__ r, int i1, int i2)
{
   const double2 p1 = *reinterpret_cast<const double2 *>(r + 2 * i1);
   const double2 p2 = *reinterpret_cast<const double2 *>(r + TAPS + 2 * i2);
   double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
#pragma unroll
   for (int z = 0; z < TAPS; z += 2)
   {
      const double2 p3 = *reinterpret_cast<const double2 *>(r + 2 * TAPS + z);
      s00 = fma(hr[0][0][z], p3.x, s00);
      s01 = fma(hr[0][1][z], p3.x, s01);
      s10 = fma(hr[1][0][z], p3.x, s10);
      s11 = fma(hr[1][1][z], p3.x, s11);
      s00 = fma(hr[0][0][z + 1], p3.y, s00);
      s01 = fma(hr[0][1][z + 1], p3.y, s01);
      s10 = fma(hr[1][0][z + 1], p3.y, s10);
      s11 = fma(hr[1][1][z + 1], p3.y, s11);
   }
   const double t0 = fma(s01, p2.y, s00 * p2.x);
   const double t1 = fma(s11, p2.y, s10 * p2.x);
   return fma(t1, p1.y, t0 * p1.x);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
k_gather3_warp(const double *__restrict__ psi, const int4 *__restrict__ chunks, const int *__restrict__ cfirst,
               const double *__restrict__ h, double *__restrict__ f)
{
   __shared__ __align__(16) double smem[WARPS_PER_CTA][WNS][STAGE_D];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * WARPS_PER_CTA + wid;
   const bool act = lane < 25;
   const int i1 = act ? lane / 5 : 0, i2 = act ? lane % 5 : 0;
   double(*st)[STAGE_D] = smem[wid];
   const unsigned full = 0xffffffffu;
   const int c0 = cfirst[gw], c1 = cfirst[gw + 1];
   if (c0 >= c1)
      return;

   int4 dcur = chunks[c0];
   int key = dcur.z;
#pragma unroll
   for (int s = 0; s < WNS - 1; s++)
   {
      if (c0 + s < c1)
         w_stage_issue(st[s], psi, nullptr, chunks[c0 + s], lane);
      w_commit();
   }
   int4 dissue = chunks[c0 + WNS - 1];

   double hr[2][2][TAPS]; // this lane's block of Re h for the current cell; idle lanes hold exact zeros
#pragma unroll
   for (int z = 0; z < TAPS; z++)
   {
      double v[2][2];
      gather_load_plane(h, key, i1, i2, z, act, v);
      hr[0][0][z] = v[0][0];
      hr[0][1][z] = v[0][1];
      hr[1][0][z] = v[1][0];
      hr[1][1][z] = v[1][1];
   }

   for (int c = c0; c < c1; c++)
   {
      const int4 dnext = chunks[c + 1];
      const int4 dissue_next = chunks[c + WNS];
      w_wait<WNS - 2>();
      __syncwarp();
      if (c + WNS - 1 < c1)
         w_stage_issue(st[(c - c0 + WNS - 1) % WNS], psi, nullptr, dissue, lane);
      w_commit();

      if (dcur.z != key) // warp-uniform: the stream enters another cell
      {
         const int knew = dcur.z, d = knew - key;
         if ((knew >> 6) == (key >> 6) && d > 0 && d < TAPS)
         {
            for (int t = 0; t < d; t++)
            {
               double pl[2][2];
               gather_load_plane(h, key + t + 1, i1, i2, TAPS - 1, act, pl);
#pragma unroll
               for (int x = 0; x < 2; x++)
#pragma unroll
                  for (int y = 0; y < 2; y++)
                  {
#pragma unroll
                     for (int z = 0; z < TAPS - 1; z++)
                        hr[x][y][z] = hr[x][y][z + 1];
                     hr[x][y][TAPS - 1] = pl[x][y];
                  }
            }
         }
         else
         {
#pragma unroll
            for (int z = 0; z < TAPS; z++)
            {
               double pl[2][2];
               gather_load_plane(h, knew, i1, i2, z, act, pl);
               hr[0][0][z] = pl[0][0];
               hr[0][1][z] = pl[0][1];
               hr[1][0][z] = pl[1][0];
               hr[1][1][z] = pl[1][1];
            }
         }
         key = knew;
      }

      const double *row = st[(c - c0) % WNS];
      double v[WS];
      if (dcur.y == WS) // warp-uniform: full chunk, straight-line code so that the points overlap in the pipes
      {
#pragma unroll
         for (int p = 0; p < WS; p++)
            v[p] = gather_point(hr, row + p * ROW, i1, i2);
      }
      else // short chunk (end of a cell run; sparse cells)
      {
#pragma unroll
         for (int p = 0; p < WS; p++)
            v[p] = 0.0;
#pragma unroll 1
         for (int p = 0; p < dcur.y; p++)
         {
            const double val = gather_point(hr, row + p * ROW, i1, i2);
#pragma unroll
            for (int q = 0; q < WS; q++)
               if (q == p)
                  v[q] = val; // static register indices
         }
      }
      // transpose-reduce: 8 per-lane partials x 32 lanes -> the sum of point p lands in the lanes with (lane >> 2) == p
      {
         const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
         double w4[4], w2[2], w1;
#pragma unroll
         for (int i = 0; i < 4; i++)
         {
            double send = hi16 ? v[i] : v[i + 4];
            double keep = hi16 ? v[i + 4] : v[i];
            w4[i] = keep + __shfl_xor_sync(full, send, 16);
         }
#pragma unroll
         for (int i = 0; i < 2; i++)
         {
            double send = hi8 ? w4[i] : w4[i + 2];
            double keep = hi8 ? w4[i + 2] : w4[i];
            w2[i] = keep + __shfl_xor_sync(full, send, 8);
         }
         {
            double send = hi4 ? w2[0] : w2[1];
            double keep = hi4 ? w2[1] : w2[0];
            w1 = keep + __shfl_xor_sync(full, send, 4);
         }
         w1 += __shfl_xor_sync(full, w1, 2);
         w1 += __shfl_xor_sync(full, w1, 1);
         const int p = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
         if ((lane & 3) == 0 && p < dcur.y)
            f[dcur.x + p] = w1;
      }
      dcur = dnext;
      dissue = dissue_next;
   }
   w_wait<0>();
}

// ---------------------------------------------------------------------------------------------------- launch layer
static int g_sm_count = 0;
static int sm_count()
{
   if (!g_sm_count)
   {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
          g_sm_count <= 0)
         g_sm_count = 148;
   }
   return g_sm_count;
}

extern "C" {

// number of warp ranges the host must cut the sorted point array into = warps of one resident wave
int b200k_warp_ranges(void) { return sm_count() * CTAS_PER_SM * WARPS_PER_CTA; }

int b200k_spread3_warp(const b200_win_dev *w)
{
   if (w->nchunk <= 0)
      return 0;
   k_spread3_warp<<<w->nwr / WARPS_PER_CTA, WARPS_PER_CTA * 32, 0, b200_stream()>>>(w->psi, w->a, (const int4 *)w->chunks, w->cfirst, w->grid);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather3_warp(const b200_win_dev *w, double *d_f)
{
   if (w->nchunk <= 0)
      return 0;
   k_gather3_warp<<<w->nwr / WARPS_PER_CTA, WARPS_PER_CTA * 32, 0, b200_stream()>>>(w->psi, (const int4 *)w->chunks, w->cfirst, w->grid, d_f);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
