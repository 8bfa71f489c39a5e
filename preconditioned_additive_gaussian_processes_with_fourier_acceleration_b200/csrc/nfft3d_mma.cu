// 3-D window spread (adjoint NFFT, B^T) and gather (NFFT trafo, B) on the fp64 tensor cores of sm_100a.
//
// Per cell, the spread of J points is a rank-J update of the 10 x 10 x 10 footprint,
//      G[(x,y), z] += sum_j ( psi1_j[x] psi2_j[y] ) * ( a_j psi3_j[z] ),
// i.e. a GEMM  C(100 x 10) += A(100 x J) B(J x 10), and the gather is its transpose,  T(J x 10) = A' H,  f_j = T(j,:) . psi3_j.
// Both run on mma.sync.m8n8k4.f64 (DMMA.8x8x4 in SASS): measured on this pool's B200 (scripts/fp64_dmma.cu) the fp64
// tensor path and the DFMA pipe are ONE unit (36-37 vs 33 TFLOP/s, no overlap), but a DMMA carries 256 FMAs per warp
// instruction instead of 32, which removes the issue-slot and operand-delivery limits of the DFMA kernels
// (nfft3d_warp.cu: 57 % fp64-pipe activity at best).  The footprint is tiled in two parts so that the awkward extent 10 = 8 + 2
// does not waste half of the instructions:
//   main  z in 0..7 : M = the 100 (x,y) pairs -> 13 tiles of 8 (96 % filled), N = z (one tile), K = points (4 per instruction);
//                     tiles 0..9 hold (x = i, y = 0..7), tiles 10..12 the pairs with y in {8, 9}: a thread's operand factors are
//                     then psi2[g] (one load per 4 points) and psi1[0..9] (warp-broadcast loads), which keeps the shared-memory
//                     pipe (128 B/clk per SM, shared by the 4 tensor pipes) off the critical path
//   edge  z in 8..9 : M = x -> 2 tiles, N = the 20 (y, z) pairs -> 3 tiles, K = points   (spread; 6 instead of 13 instructions)
//                     M = points, N = (y, z) -> 3 tiles, K = x -> 3 steps                (gather; 9 instead of 25 instructions)
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
#ifndef MX_CLK
#define MX_CLK 0 /* development aid: in-kernel clocks of the spread, printed by a few warps */
#endif
#define MT 13 /* m-tiles: 13 * 8 = 104 >= 100 (x,y) pairs */
#define KS 25 /* gather k-steps: 25 * 4 = 100 (x,y) pairs */

__device__ __forceinline__ void m_dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// fire-and-forget fp64 add to GLOBAL memory (RED.E.ADD.F64): with the grid pointer coming out of a descriptor in memory the
// compiler cannot prove the address space and atomicAdd() turns into a generic atomic with shared / local fall-back loops
__device__ __forceinline__ void m_red_add(double *p, double v)
{
   asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
// fp64 product as a volatile PTX instruction: keeps its place in front of the tensor instructions in the PTX stream
__device__ __forceinline__ double m_vmul(double a, double b)
{
   double r;
   asm volatile("mul.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
   return r;
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

// range (round k, slot q) of the fused launches -> window, chunk interval; returns false when the slot is empty
__device__ __forceinline__ bool m_range_lookup(const b200_mma_win *__restrict__ W, int P, int total, int r, int &w, int &c0, int &c1)
{
   const int k = r / total, q = r - k * total;
   w = 0;
   while (w + 1 < P && q >= W[w + 1].warp_begin)
      w++;
   const int j = q - W[w].warp_begin;
   if (j >= W[w].nwarps)
      return false;
   const int *cf = W[w].cfirst + k * W[w].nwarps + j;
   c0 = cf[0];
   c1 = cf[1];
   return true;
}
static __device__ int g_mma_range_counter; // next dynamic range of the running fused launch of this translation unit
__device__ __forceinline__ int m_range_next(int *counter, int total, int lane)
{
   int nx = 0;
   if (lane == 0)
      nx = atomicAdd(counter, 1);
   return total + __shfl_sync(0xffffffffu, nx, 0);
}

#ifndef MMA_PART
#define MMA_PART 0 /* 1: spread only, 2: gather only (the Makefile builds the two halves with different ptxas levels) */
#endif

#if MMA_PART != 2
// ---------------------------------------------------------------------------------------------------- spread
// Writing a cell's footprint straight to the L2-resident grid costs 1040 fp64 REDs per cell change and warp -- 21 M per
// launch at n = 10^6, independent of n, and the L2 atomic units made that 190 of the 338 us of the fused spread.
// Consecutive cells of a range are neighbours along l3, so every warp keeps a private 10-plane ring of the footprint in
// shared memory (plane u = c3 + z lives in slot u mod 10): on a cell change the register tile is added to the ring and only
// the d planes that LEAVE the footprint go to global memory (100 REDs per plane, as in the DFMA kernels' register shift).
#define CB_PLANE 100
#define CB_DOUBLES (TAPS * CB_PLANE)

// main accumulators: tile i < 10: rows y = g, x = i; tile 10 + i': rows (x = 4 i' + g / 2, y = 8 + (g & 1)); columns z = 2 t + {0, 1}
// edge accumulators: tile (i, j): rows x = 8 i + g; columns 8 j + 2 t + {0, 1} <-> y = 4 j + t, z = 8 + {0, 1}
// The register tile IS the running total of the cell's footprint: on a cell change it is stored to the ring (LOAD = false),
// the planes that leave are written out, and the tile of the next cell is loaded back (LOAD = true) -- no fp64 additions,
// which would queue behind the tensor instructions of the other warps on the one fp64 unit.
template <bool LOAD>
__device__ __forceinline__ void m_acc_ring(double *__restrict__ cb, int key, int g, int t, double acc[MT][2], double ace[2][3][2])
{
   const int c3 = key & 63;
   const int s0 = ((c3 + 2 * t) % TAPS) * CB_PLANE, s1 = ((c3 + 2 * t + 1) % TAPS) * CB_PLANE;
   const int s8 = ((c3 + 8) % TAPS) * CB_PLANE, s9 = ((c3 + 9) % TAPS) * CB_PLANE;
#pragma unroll
   for (int i = 0; i < MT; i++)
   {
      const int x = i < 10 ? i : 4 * (i - 10) + (g >> 1);
      const int y = i < 10 ? g : 8 + (g & 1);
      if (x < TAPS)
      {
         if (LOAD)
         {
            acc[i][0] = cb[s0 + x * TAPS + y];
            acc[i][1] = cb[s1 + x * TAPS + y];
         }
         else
         {
            cb[s0 + x * TAPS + y] = acc[i][0];
            cb[s1 + x * TAPS + y] = acc[i][1];
         }
      }
      else if (LOAD)
         acc[i][0] = acc[i][1] = 0.0;
   }
#pragma unroll
   for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
      {
         const int x = 8 * i + g, y = 4 * j + t;
         if (x < TAPS && y < TAPS)
         {
            if (LOAD)
            {
               ace[i][j][0] = cb[s8 + x * TAPS + y];
               ace[i][j][1] = cb[s9 + x * TAPS + y];
            }
            else
            {
               cb[s8 + x * TAPS + y] = ace[i][j][0];
               cb[s9 + x * TAPS + y] = ace[i][j][1];
            }
         }
         else if (LOAD)
            ace[i][j][0] = ace[i][j][1] = 0.0;
      }
}

// Per (c1, c2) column of cells: the grid offsets of the (up to) four (x, y) footprint entries e = lane + 32 k that this lane
// writes out when a plane leaves the ring.  Computed once per column change, so that flushing a plane is 4 x (LDS, RED, STS).
struct m_col_offsets
{
   int o[4];
};
__device__ __forceinline__ m_col_offsets m_column_offsets(int key, int lane)
{
   const int c1 = (key >> 12) & 63, c2 = (key >> 6) & 63;
   m_col_offsets r;
#pragma unroll
   for (int k = 0; k < 4; k++)
   {
      const int e = lane + 32 * k;
      const int x = e / TAPS, y = e - TAPS * x;
      r.o[k] = e < CB_PLANE ? (m_tap_idx(c1, x) * NOS + m_tap_idx(c2, y)) * NOS : -1;
   }
   return r;
}

// planes z in [0, nz) of the footprint of cell `key` leave the ring: add them to the grid and clear their slots
__device__ __forceinline__ void m_ring_flush(double *__restrict__ cb, double *__restrict__ grid, int key, int nz, int lane, const m_col_offsets &co)
{
   const int c3 = key & 63;
   for (int z = 0; z < nz; z++)
   {
      double *plane = cb + ((c3 + z) % TAPS) * CB_PLANE + lane;
      double *gp = grid + m_tap_idx(c3, z);
#pragma unroll
      for (int k = 0; k < 4; k++)
         if (co.o[k] >= 0)
         {
            m_red_add(gp + co.o[k], plane[32 * k]);
            plane[32 * k] = 0.0;
         }
   }
}

// one warp streams the chunks [c0, c1) of one window
__device__ __forceinline__ void m_spread_range(const double *__restrict__ psi, const double *__restrict__ a, const int4 *__restrict__ chunks,
                                               int c0, int c1, double *__restrict__ grid, double (*st)[STAGE_D], double *__restrict__ cb,
                                               int lane)
{
   const int g = lane >> 2, t = lane & 3;
   if (c0 >= c1)
      return;
   for (int e = lane; e < CB_DOUBLES; e += 32)
      cb[e] = 0.0;
   m_col_offsets co;

   double acc[MT][2], ace[2][3][2];
#pragma unroll
   for (int i = 0; i < MT; i++)
      acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
   for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
         ace[i][j][0] = ace[i][j][1] = 0.0;

   int4 dcur = chunks[c0];
   int key = dcur.z;
   co = m_column_offsets(key, lane);
#pragma unroll
   for (int s = 0; s < MWNS - 1; s++)
   {
      if (c0 + s < c1)
         m_stage_issue(st[s], psi, a, chunks[c0 + s], lane);
      m_commit();
   }
   int4 dissue = chunks[c0 + MWNS - 1];
#if MX_CLK
   long long tk_stage = 0, tk_flush = 0, tk_comp = 0, tk_n = 0, tk_cells = 0, tk_f[3] = {0, 0, 0};
   const long long tk_begin = clock64();
   unsigned long long tk_ns0;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tk_ns0));
#endif
   for (int c = c0; c < c1; c++)
   {
#if MX_CLK
      const long long tk0 = clock64();
#endif
      const int4 dnext = chunks[c + 1];
      const int4 dissue_next = chunks[c + MWNS];
      m_wait<MWNS - 2>();
      __syncwarp();
      if (c + MWNS - 1 < c1)
         m_stage_issue(st[(c - c0 + MWNS - 1) % MWNS], psi, a, dissue, lane);
      m_commit();
#if MX_CLK
      const long long tk1 = clock64();
      tk_cells += dcur.z != key;
#endif

      if (dcur.z != key) // warp-uniform: the stream enters another cell
      {
         const int knew = dcur.z, d = knew - key;
         const bool same_column = (knew >> 6) == (key >> 6);
#if MX_CLK
         const long long tf0 = clock64();
#endif
         m_acc_ring<false>(cb, key, g, t, acc, ace);
         __syncwarp();
#if MX_CLK
         const long long tf1 = clock64();
#endif
         // same (c1, c2) column, c3 advanced by d: only the d planes that leave the footprint are written out
         m_ring_flush(cb, grid, key, (same_column && d > 0 && d < TAPS) ? d : TAPS, lane, co);
         __syncwarp();
         if (!same_column)
            co = m_column_offsets(knew, lane);
         key = knew;
#if MX_CLK
         const long long tf2 = clock64();
#endif
         m_acc_ring<true>(cb, key, g, t, acc, ace); // planes that stay keep their totals, the new ones were cleared
#if MX_CLK
         tk_f[0] += tf1 - tf0;
         tk_f[1] += tf2 - tf1;
         tk_f[2] += clock64() - tf2;
#endif
      }

#if MX_CLK
      const long long tk2 = clock64();
#endif
      const double *stage = st[(c - c0) % MWNS];
#pragma unroll
      for (int kk = 0; kk < 2; kk++)
      {
         if (kk == 1 && dcur.y <= 4) // warp-uniform: the second k-step holds only zero-weight points
            break;
         const double *row = stage + (4 * kk + t) * RS;   // operand column / row k = t: the point of this thread
         // ---- all shared-memory operands of the k-step first (18 loads in flight together: their latency is paid once, not
         // in front of every product -- ncu showed the products stalled on the short scoreboard)
         const double aj = stage[WS * RS + 4 * kk + t];    // zero past the end of the chunk
         const double r3g = row[2 * TAPS + g];
         const double p2 = row[TAPS + g];
         const double p2e = row[TAPS + 8 + (g & 1)];
         const double p3 = row[2 * TAPS + 8 + (g & 1)];
         const double r1g = row[g];
         const double r1e = row[8 + (g & 1)];
         double2 p1[5];
#pragma unroll
         for (int i = 0; i < 5; i++)
            p1[i] = *reinterpret_cast<const double2 *>(row + 2 * i);
         double x1[3], y2[3];
#pragma unroll
         for (int i = 0; i < 3; i++)
         {
            const int x = 4 * i + (g >> 1), y = 4 * i + (g >> 1);
            x1[i] = row[x < TAPS ? x : TAPS - 1];
            y2[i] = row[TAPS + (y < TAPS ? y : TAPS - 1)];
         }
         // ---- all operand products of the k-step first, then its 19 tensor instructions.  The spread half of this file is
         // built with `-Xptxas -O1` (Makefile), which keeps this order; at the default level ptxas sinks every load and product
         // next to its DMMA, and the warp then pays the shared-memory and product latencies once per tensor instruction
         // instead of once per k-step (measured: 327 -> 321 us per fused launch at n = 10^6, 119 -> 114 us at 2.5 10^5; the gather
         // loses 6 us with -O1 and keeps the default level).
         // main part: B(k, n = g) = a psi3[g];  A(m, k) = psi1[x] psi2[y] with (x, y) = (i, g) for the tiles i < 10 and
         // (4 i' + g / 2, 8 + (g & 1)) for the tiles 10 + i'
         // edge part: A(x, k) = a psi1[x], B(k, n = g of tile j) = psi2[4 j + g / 2] psi3[8 + (g & 1)]
         const double b0 = m_vmul(aj, r3g);
         double af[MT];
#pragma unroll
         for (int i = 0; i < 5; i++)
         {
            af[2 * i] = m_vmul(p1[i].x, p2);
            af[2 * i + 1] = m_vmul(p1[i].y, p2);
         }
#pragma unroll
         for (int i = 0; i < 3; i++)
         {
            const double v = m_vmul(x1[i], p2e);
            af[10 + i] = (i < 2 || g < 4) ? v : 0.0;
         }
         const double ae0 = m_vmul(aj, r1g);
         const double ae1v = m_vmul(aj, r1e);
         const double ae1 = (g < 2) ? ae1v : 0.0;
         double be[3];
#pragma unroll
         for (int j = 0; j < 3; j++)
         {
            const double v = m_vmul(y2[j], p3);
            be[j] = (j < 2 || g < 4) ? v : 0.0;
         }
#pragma unroll
         for (int i = 0; i < MT; i++)
            m_dmma(acc[i][0], acc[i][1], af[i], b0);
#pragma unroll
         for (int j = 0; j < 3; j++)
         {
            m_dmma(ace[0][j][0], ace[0][j][1], ae0, be[j]);
            m_dmma(ace[1][j][0], ace[1][j][1], ae1, be[j]);
         }
      }
#if MX_CLK
      {
         const long long tk3 = clock64();
         tk_stage += tk1 - tk0;
         tk_flush += tk2 - tk1;
         tk_comp += tk3 - tk2;
         tk_n++;
      }
#endif
      dcur = dnext;
      dissue = dissue_next;
   }
#if MX_CLK
   if (lane == 0 && ((blockIdx.x * MWARPS_PER_CTA + (threadIdx.x >> 5)) % 97) == 0)
   {
      unsigned long long tk_ns1;
      unsigned smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tk_ns1));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const long long tot = clock64() - tk_begin;
      printf("warp %4d sm %3u: [%7llu .. %7llu ns] chunks %3lld cells %2lld total %lld clk %.3f GHz | per chunk: stage %lld flush %lld compute %lld | per cell: store %lld planes %lld load %lld\n",
             blockIdx.x * MWARPS_PER_CTA + (threadIdx.x >> 5), smid, tk_ns0 % 10000000ull, tk_ns1 % 10000000ull, tk_n, tk_cells, tot,
             (double)tot / (double)(tk_ns1 - tk_ns0), tk_stage / tk_n, tk_flush / tk_n, tk_comp / tk_n, tk_f[0] / (tk_cells + 1), tk_f[1] / (tk_cells + 1),
             tk_f[2] / (tk_cells + 1));
   }
#endif
   m_acc_ring<false>(cb, key, g, t, acc, ace);
   __syncwarp();
   m_ring_flush(cb, grid, key, TAPS, lane, co);
   __syncwarp();
}

// dynamic shared memory of the spread kernels: the cp.async stages of the CTA's warps, then their footprint rings
#define SPREAD_SMEM_BYTES (sizeof(double) * (size_t)MWARPS_PER_CTA * (MWNS * STAGE_D + CB_DOUBLES))

__global__ void __launch_bounds__(MWARPS_PER_CTA * 32, MCTAS_PER_SM)
k_spread3_mma(const double *__restrict__ psi, const double *__restrict__ a, const int4 *__restrict__ chunks,
              const int *__restrict__ cfirst, double *__restrict__ grid)
{
   extern __shared__ __align__(16) double dsm[];
   double(*smem)[MWNS][STAGE_D] = reinterpret_cast<double(*)[MWNS][STAGE_D]>(dsm);
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * MWARPS_PER_CTA + wid;
   m_spread_range(psi, a, chunks, cfirst[gw], cfirst[gw + 1], grid, smem[wid], dsm + MWARPS_PER_CTA * MWNS * STAGE_D + wid * CB_DOUBLES, lane);
}

// all windows of an additive kernel in ONE persistent wave: warp gw belongs to the window w with
// wbeg[w] <= gw < wbeg[w + 1] and streams range gw - wbeg[w] of that window's own partition
__global__ void __launch_bounds__(MWARPS_PER_CTA * 32, MCTAS_PER_SM)
k_spread3_mma_batch(const b200_mma_win *__restrict__ W, int P, int rounds)
{
   extern __shared__ __align__(16) double dsm[];
   double(*smem)[MWNS][STAGE_D] = reinterpret_cast<double(*)[MWNS][STAGE_D]>(dsm);
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int total = gridDim.x * MWARPS_PER_CTA;
   int r = blockIdx.x * MWARPS_PER_CTA + wid;
   while (r < total * rounds)
   {
      int w, c0, c1;
      if (m_range_lookup(W, P, total, r, w, c0, c1))
         m_spread_range(W[w].psi, W[w].a, (const int4 *)W[w].chunks, c0, c1, W[w].grid, smem[wid],
                        dsm + MWARPS_PER_CTA * MWNS * STAGE_D + wid * CB_DOUBLES, lane);
      if (rounds == 1)
         break;
      r = m_range_next(&g_mma_range_counter, total, lane);
   }
}

#endif // MMA_PART != 2

#if MMA_PART != 1
// ---------------------------------------------------------------------------------------------------- gather
// B operands (constant per cell).  main k-steps: s < 20: row k = t <-> (x = s / 2, y = 4 (s & 1) + t); s = 20 + s':
// (x = 2 s' + t / 2, y = 8 + (t & 1)); column n = g <-> z = g.
// edge: k-step s (3), row k = t <-> x = 4 s + t; tile j (3), column n = g <-> y = 4 j + g / 2, z = 8 + (g & 1).
__device__ __forceinline__ void m_gather_load(const double *__restrict__ h, int key, int g, int t, double hb[KS], double he[3][3])
{
   const int c1 = (key >> 12) & 63, c2 = (key >> 6) & 63, c3 = key & 63;
   const int z0 = m_tap_idx(c3, g), z1 = m_tap_idx(c3, 8 + (g & 1));
#pragma unroll
   for (int s = 0; s < KS; s++)
   {
      const int x = s < 20 ? (s >> 1) : 2 * (s - 20) + (t >> 1);
      const int y = s < 20 ? 4 * (s & 1) + t : 8 + (t & 1);
      hb[s] = __ldg(h + (size_t)(m_tap_idx(c1, x) * NOS + m_tap_idx(c2, y)) * NOS + z0);
   }
#pragma unroll
   for (int s = 0; s < 3; s++)
#pragma unroll
      for (int j = 0; j < 3; j++)
      {
         const int x = 4 * s + t, y = 4 * j + (g >> 1);
         he[s][j] = (x < TAPS && y < TAPS) ? __ldg(h + (size_t)(m_tap_idx(c1, x) * NOS + m_tap_idx(c2, y)) * NOS + z1) : 0.0;
      }
}

__device__ __forceinline__ void m_gather_range(const double *__restrict__ psi, const int4 *__restrict__ chunks, int c0, int c1,
                                               const double *__restrict__ h, double *__restrict__ f, double (*st)[STAGE_D], int lane)
{
   const int g = lane >> 2, t = lane & 3;
   const unsigned full = 0xffffffffu;
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

   double hb[KS], he[3][3];
   m_gather_load(h, key, g, t, hb, he);

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
         m_gather_load(h, key, g, t, hb, he);
      }

      // A operand: row = point g of the chunk, column k = t <-> pair m = 4 s + t
      const double *row = st[(c - c0) % MWNS] + g * RS;
      // main part, four dependency chains.  A(point g, k = t) = psi1[x] psi2[y] with the pair of k-step s (see m_gather_load)
      double ca[2] = {0.0, 0.0}, cb[2] = {0.0, 0.0}, cc[2] = {0.0, 0.0}, cd[2] = {0.0, 0.0};
      {
         const double q0 = row[TAPS + t], q1 = row[TAPS + 4 + t], q2 = row[TAPS + 8 + (t & 1)];
#pragma unroll
         for (int x = 0; x < 10; x += 2)
         {
            const double2 p1 = *reinterpret_cast<const double2 *>(row + x);
            m_dmma(ca[0], ca[1], p1.x * q0, hb[2 * x]);
            m_dmma(cb[0], cb[1], p1.x * q1, hb[2 * x + 1]);
            m_dmma(cc[0], cc[1], p1.y * q0, hb[2 * x + 2]);
            m_dmma(cd[0], cd[1], p1.y * q1, hb[2 * x + 3]);
         }
#pragma unroll
         for (int s = 0; s < 5; s++)
         {
            const double af = row[2 * s + (t >> 1)] * q2;
            if ((s & 3) == 0)
               m_dmma(ca[0], ca[1], af, hb[20 + s]);
            else if ((s & 3) == 1)
               m_dmma(cb[0], cb[1], af, hb[20 + s]);
            else if ((s & 3) == 2)
               m_dmma(cc[0], cc[1], af, hb[20 + s]);
            else
               m_dmma(cd[0], cd[1], af, hb[20 + s]);
         }
         ca[0] += cb[0] + (cc[0] + cd[0]);
         ca[1] += cb[1] + (cc[1] + cd[1]);
      }
      // edge part: T'(point, (y, z)) = sum_x psi1[x] H[x, y, z], z in {8, 9}
      double ce[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int s = 0; s < 3; s++)
      {
         const int x = 4 * s + t;
         const double ax = (x < TAPS) ? row[x] : 0.0;
#pragma unroll
         for (int j = 0; j < 3; j++)
            m_dmma(ce[j][0], ce[j][1], ax, he[s][j]);
      }
      // main C: row = point g; columns z = 2 t, 2 t + 1.  edge C: columns 8 j + 2 t + {0, 1} <-> y = 4 j + t, z = 8 + {0, 1}
      double v = ca[0] * row[2 * TAPS + 2 * t];
      v = fma(ca[1], row[2 * TAPS + 2 * t + 1], v);
      {
         const double p8 = row[2 * TAPS + 8], p9 = row[2 * TAPS + 9];
#pragma unroll
         for (int j = 0; j < 3; j++)
         {
            const int y = 4 * j + t;
            if (y < TAPS)
               v = fma(row[TAPS + y], fma(ce[j][1], p9, ce[j][0] * p8), v);
         }
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

__global__ void __launch_bounds__(MWARPS_PER_CTA * 32, MCTAS_PER_SM)
k_gather3_mma(const double *__restrict__ psi, const int4 *__restrict__ chunks, const int *__restrict__ cfirst,
              const double *__restrict__ h, double *__restrict__ f)
{
   __shared__ __align__(16) double smem[MWARPS_PER_CTA][MWNS][STAGE_D];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * MWARPS_PER_CTA + wid;
   m_gather_range(psi, chunks, cfirst[gw], cfirst[gw + 1], h, f, smem[wid], lane);
}

__global__ void __launch_bounds__(MWARPS_PER_CTA * 32, MCTAS_PER_SM)
k_gather3_mma_batch(const b200_mma_win *__restrict__ W, int P, int rounds)
{
   __shared__ __align__(16) double smem[MWARPS_PER_CTA][MWNS][STAGE_D];
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int total = gridDim.x * MWARPS_PER_CTA;
   int r = blockIdx.x * MWARPS_PER_CTA + wid;
   while (r < total * rounds)
   {
      int w, c0, c1;
      if (m_range_lookup(W, P, total, r, w, c0, c1))
         m_gather_range(W[w].psi, (const int4 *)W[w].chunks, c0, c1, W[w].grid, W[w].f, smem[wid], lane);
      if (rounds == 1)
         break;
      __syncwarp(); // the stages are reused by the next range
      r = m_range_next(&g_mma_range_counter, total, lane);
   }
}

#endif // MMA_PART != 1

// ---------------------------------------------------------------------------------------------------- launch layer
extern "C" {

static int mma_counter_reset(void)
{
   static int *addr = nullptr;
   if (!addr)
      B200_CHECK(cudaGetSymbolAddress((void **)&addr, g_mma_range_counter));
   B200_CHECK(cudaMemsetAsync(addr, 0, sizeof(int), b200_stream())); // a memset node when the matvec is being captured
   return 0;
}

#if MMA_PART != 2
// warps of one resident wave of the tensor-core kernels (the fused launches size their partitions with it)
int b200k_mma_warp_ranges(void)
{
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   return sm * MCTAS_PER_SM * MWARPS_PER_CTA;
}

int b200k_spread3_mma(const b200_win_dev *w)
{
   if (w->nchunk <= 0)
      return 0;
   B200_CHECK(cudaFuncSetAttribute(k_spread3_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPREAD_SMEM_BYTES));
   k_spread3_mma<<<w->nwr / MWARPS_PER_CTA, MWARPS_PER_CTA * 32, SPREAD_SMEM_BYTES, b200_stream()>>>(w->psi, w->a, (const int4 *)w->chunks,
                                                                                                   w->cfirst, w->grid);
   B200_LAUNCHED();
   return 0;
}

#endif // MMA_PART != 2
#if MMA_PART != 1
int b200k_gather3_mma(const b200_win_dev *w, double *d_f)
{
   if (w->nchunk <= 0)
      return 0;
   k_gather3_mma<<<w->nwr / MWARPS_PER_CTA, MWARPS_PER_CTA * 32, 0, b200_stream()>>>(w->psi, (const int4 *)w->chunks, w->cfirst, w->grid, d_f);
   B200_LAUNCHED();
   return 0;
}

// fused launches over the P windows described by d_batch (device array); total_warps = sum of nwarps, a multiple of the CTA size
#endif // MMA_PART != 1
#if MMA_PART != 2
int b200k_spread3_mma_batch(const b200_mma_win *d_batch, int P, int total_warps, int rounds)
{
   static bool attr_set = false; // once: this call may sit inside a stream capture (CUDA graph of the matvec)
   if (!attr_set)
   {
      B200_CHECK(cudaFuncSetAttribute(k_spread3_mma_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPREAD_SMEM_BYTES));
      attr_set = true;
   }
   if (rounds > 1 && mma_counter_reset())
      return -1;
   k_spread3_mma_batch<<<total_warps / MWARPS_PER_CTA, MWARPS_PER_CTA * 32, SPREAD_SMEM_BYTES, b200_stream()>>>(d_batch, P, rounds);
   B200_LAUNCHED();
   return 0;
}

#endif // MMA_PART != 2
#if MMA_PART != 1
int b200k_gather3_mma_batch(const b200_mma_win *d_batch, int P, int total_warps, int rounds)
{
   if (rounds > 1 && mma_counter_reset())
      return -1;
   k_gather3_mma_batch<<<total_warps / MWARPS_PER_CTA, MWARPS_PER_CTA * 32, 0, b200_stream()>>>(d_batch, P, rounds);
   B200_LAUNCHED();
   return 0;
}

#endif // MMA_PART != 1

} // extern "C"
