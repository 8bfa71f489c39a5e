// 1-D feature windows (BASELINE config C4: additive kernel with many 1-D windows): spread (adjoint NFFT, B^T) and gather (NFFT
// trafo, B) WITHOUT tap tables, atomics or a global permutation of the points.
//
// For a 1-D window the oversampled grid has 64 cells and a point with fractional offset u in [0, 1) inside its cell c touches
// the ten taps psi_t(u) = phi(u + 4 - t).  On a cell every psi_t is one smooth piece of the Kaiser-Bessel window; its degree-14
// interpolant p_t(v), v = 2u - 1, is exact to 4.6e-16 phi(0) in double (kb_window_coeffs.h, generated in 40-digit arithmetic).
// Hence
//   gather:  f_i = sum_t h[c_i - 4 + t] p_t(v_i) = C_{c_i}(v_i),  C_c = sum_t h[c - 4 + t] p_t   -- ONE Horner evaluation
//            (14 FMAs) per point from a 64 x 15 coefficient table per window instead of ten taps;
//   spread:  g[c - 4 + t] += sum_k p_t[k] M_c[k],  M_c[k] = sum_{i in cell c} a_i v_i^k            -- 15 moments per point.
// Per point and window the kernels read the in-cell offset v (8 B, exact) and a 2-byte index instead of 80 B of taps + 4 B key + 8 B of
// permuted coefficient, and the prep / combine passes with their random 8-byte accesses disappear: points are sorted by cell
// only INSIDE blocks of T consecutive points of the ORIGINAL order (done once per handle, nfft_host.c:b200_poly1d_block_sort), so x
// and y are moved as coalesced blocks through shared memory and the "permutation" is a shared-memory index.
// Two generations of kernels live in this file:
//   direct   (k_poly1d_spread, k_poly1d_gather<NCH>): 2 CTAs x 16 warps per SM, offsets / indices loaded from global memory per run,
//            one transposing butterfly per run.  Kept for the gradient gather (NCH = 2), for a vector that is not 16-byte aligned
//            and as a cross-check (NFFT4GP_B200_POLY1D_ASYNC=0).
//   staged   (k_poly1d_spread_async, k_poly1d_gather_async): the production pair, described where it is defined below.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"
#include "kb_window_coeffs.h"

#define TAPS B200_TAPS
#define PNC (KB_POLY_DEG + 1) /* 15 coefficients */
#define PLD 16                /* leading dimension of the per-cell rows (moments / coefficients) */
#define PWARPS 16
#define PTHREADS (PWARPS * 32)
#ifndef PU
#define PU 4 /* points per lane whose loads are issued together */
#endif

__constant__ double c_p1[TAPS][PNC] = KB_POLY_TABLE;

__device__ __forceinline__ int p_tap_idx(int key, int t) { return (key + 28 + t) & 63; } // grid index of tap t of the cell with key c + 32
__device__ __forceinline__ void p_red_add(double *p, double v)
{
   asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

// sum of m[q] over the 32 lanes for q = 0..15 with 8 + 4 + 2 + 1 + 1 shuffles: afterwards every lane holds the total of the value
// index q = 8 b4 + 4 b3 + 2 b2 + b1 (b_i = bit i of the lane id)
__device__ __forceinline__ double p_transpose_reduce16(const double m[16], int lane, int &q)
{
   const unsigned full = 0xffffffffu;
   const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
   double w8[8], w4[4], w2[2], w1;
#pragma unroll
   for (int i = 0; i < 8; i++)
   {
      const double send = b4 ? m[i] : m[i + 8], keep = b4 ? m[i + 8] : m[i];
      w8[i] = keep + __shfl_xor_sync(full, send, 16);
   }
#pragma unroll
   for (int i = 0; i < 4; i++)
   {
      const double send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
      w4[i] = keep + __shfl_xor_sync(full, send, 8);
   }
#pragma unroll
   for (int i = 0; i < 2; i++)
   {
      const double send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
      w2[i] = keep + __shfl_xor_sync(full, send, 4);
   }
   {
      const double send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
      w1 = keep + __shfl_xor_sync(full, send, 2);
   }
   w1 += __shfl_xor_sync(full, w1, 1);
   q = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
   return w1;
}

// ---------------------------------------------------------------------------------------------------- spread: moments
// moments of one point: m[k] += a v^k, k = 0..14, from ONE chain in v^4: e = a v^(4q); m[4q] += e, m[4q+r] = fma(e, v^r, m[4q+r]) --
// 21 fp64 instructions per point (3 + 3 products, 4 additions, 11 FMAs) instead of 30 with separate power and sum steps
__device__ __forceinline__ void p_moments(double m[16], double a, double v)
{
   const double v2 = v * v, v3 = v2 * v, v4 = v2 * v2;
   double e = a;
#pragma unroll
   for (int k = 0; k < PNC; k += 4)
   {
      m[k] += e;
      m[k + 1] = fma(e, v, m[k + 1]);
      m[k + 2] = fma(e, v2, m[k + 2]);
      if (k + 3 < PNC)
      {
         m[k + 3] = fma(e, v3, m[k + 3]);
         e *= v4;
      }
   }
}
// the stored coordinate IS v = 2 u - 1 (u = offset of the point inside its cell; the cell is the run the point stands in)
__device__ __forceinline__ double p_offset(double xt) { return xt; }

// The work items (window w, block b), w-major, are cut into gridDim.x equal consecutive ranges: ONE wave of CTAs whatever the
// number of windows (a grid of G x nw CTAs left a second, nearly empty wave behind: 304 CTAs on 296 slots doubled the time).
// A CTA keeps the moments of its current window in shared memory and adds them to the window's global table when it moves on.
__global__ void __launch_bounds__(PTHREADS)
k_poly1d_spread(b200_poly1d_dev p, const double *__restrict__ x, double s)
{
   extern __shared__ __align__(16) double psm[];
   double *xs = psm, *msm = psm + p.T;
   unsigned short *rs = reinterpret_cast<unsigned short *>(msm + 64 * PLD);
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   const long long items = (long long)p.nw * p.nblk;
   const long long it0 = items * blockIdx.x / gridDim.x, it1 = items * (blockIdx.x + 1) / gridDim.x;
   int wcur = -1;
   for (long long it = it0; it < it1; it++)
   {
      const int w = (int)(it / p.nblk), b = (int)(it - (long long)w * p.nblk);
      const size_t wbase = (size_t)w * p.nblk * p.T;
      const double *__restrict__ xt = p.xt + wbase;
      const unsigned short *__restrict__ lp = p.lp + wbase;
      const int base = b * p.T, cnt = min(p.T, p.n - base);
      __syncthreads(); // the previous item's runs are done with xs / rs / msm
      if (w != wcur)
      {
         if (wcur >= 0)
         {
            double *mg = p.mom + (size_t)wcur * 64 * PLD;
            for (int e = tid; e < 64 * PLD; e += PTHREADS)
               if (msm[e] != 0.0)
                  p_red_add(mg + e, msm[e]);
         }
         for (int e = tid; e < 64 * PLD; e += PTHREADS)
            msm[e] = 0.0;
         wcur = w;
      }
      for (int j = tid; j < p.T; j += PTHREADS)
         xs[j] = j < cnt ? s * __ldg(x + base + j) : 0.0;
      if (tid < 65)
         rs[tid] = p.rs[((size_t)w * p.nblk + b) * B200_POLY_RS + tid];
      __syncthreads();
      for (int c = wid; c < 64; c += PWARPS)
      {
         const int j0 = rs[c], j1 = rs[c + 1];
         if (j0 == j1) // warp-uniform
            continue;
         double m[16];
#pragma unroll
         for (int k = 0; k < 16; k++)
            m[k] = 0.0;
         // all loads of (a chunk of) the run first -- up to PU points per lane in flight, so the latency of the coordinate / index loads is
         // paid once per run instead of once per point -- then the arithmetic, PU independent chains
         for (int jc = j0; jc < j1; jc += 32 * PU)
         {
            double xv[PU];
            int lv[PU];
#pragma unroll
            for (int u = 0; u < PU; u++)
            {
               const int j = jc + lane + 32 * u;
               const bool ok = j < j1;
               xv[u] = ok ? __ldg(xt + base + j) : 0.0;
               lv[u] = ok ? (int)lp[base + j] : -1;
            }
#pragma unroll
            for (int u = 0; u < PU; u++)
               if (lv[u] >= 0)
                  p_moments(m, xs[lv[u]], p_offset(xv[u]));
         }
         int q;
         const double tot = p_transpose_reduce16(m, lane, q);
         if (!(lane & 1) && q < PNC)
            msm[c * PLD + q] += tot; // this warp is the only one that touches cell c of this CTA
      }
   }
   __syncthreads();
   if (wcur >= 0)
   {
      double *mg = p.mom + (size_t)wcur * 64 * PLD;
      for (int e = tid; e < 64 * PLD; e += PTHREADS)
         if (msm[e] != 0.0)
            p_red_add(mg + e, msm[e]);
   }
}

// moments -> oversampled grid: g[l] = sum_t sum_k p_t[k] M[cell with tap t at l][k]; one CTA per window, one thread per grid cell
__global__ void k_poly1d_mom2grid(const double *__restrict__ mom, double *__restrict__ grid, size_t grid_stride)
{
   const int w = blockIdx.x, l = threadIdx.x;
   const double *m = mom + (size_t)w * 64 * PLD;
   double g = 0.0;
#pragma unroll
   for (int t = 0; t < TAPS; t++)
   {
      const double *mc = m + ((l - 28 - t) & 63) * PLD;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < PNC; k++)
         acc = fma(c_p1[t][k], mc[k], acc);
      g += acc;
   }
   grid[(size_t)w * grid_stride + l] = g;
}

// Re h (64 values per window) -> polynomial coefficients per cell: C[c][k] = scale_w * sum_t h[tap t of cell c] p_t[k]
__global__ void k_poly1d_grid2coef(const double *__restrict__ grid, size_t grid_stride, const double *__restrict__ wscale, double *__restrict__ ctab)
{
   const int w = blockIdx.x, c = threadIdx.x >> 4, k = threadIdx.x & 15;
   const double *h = grid + (size_t)w * grid_stride;
   double acc = 0.0;
   if (k < PNC)
   {
#pragma unroll
      for (int t = 0; t < TAPS; t++)
         acc = fma(h[p_tap_idx(c, t)], c_p1[t][k], acc);
   }
   ctab[(size_t)w * 64 * PLD + c * PLD + k] = (wscale ? wscale[w] : 1.0) * acc;
}

// ---------------------------------------------------------------------------------------------------- gather: Horner + epilogue
// degree-14 polynomial as E(v^2) + v O(v^2): two independent FMA chains of 7 instead of one of 14 (Estrin's scheme -- depth 5, three more
// products -- was measured slower in the staged gather: 641 vs 606 us for 64 windows)
__device__ __forceinline__ double p_horner(const double cf[PNC], double v)
{
   const double v2 = v * v;
   double e = cf[14], o = cf[13];
#pragma unroll
   for (int k = 12; k >= 0; k -= 2)
   {
      e = fma(e, v2, cf[k]);
      if (k - 1 >= 0)
         o = fma(o, v2, cf[k - 1]);
   }
   return fma(o, v, e);
}

// NCH = 1: y = beta y + c0 (sum_w f_w + nw mu a)                                     (nfft_interface.c:428-494, 803-814)
// NCH = 2: gradient blocks y0 = beta y0 + f2 (sum_w f_w + nw mu a), y1 = beta y1 + sum_w c1_w fd_w (c1_w folded into the table),
//          y2 = beta y2 + nw ff a                                                       (nfft_interface.c:539-617, 826-837)
template <int NCH>
__global__ void __launch_bounds__(PTHREADS)
k_poly1d_gather(b200_poly1d_dev p, const double *__restrict__ x, double s, double c0, double mu, double ff, double beta, double *__restrict__ y,
                size_t ystride)
{
   extern __shared__ __align__(16) double psm[];
   double *acc = psm;                        // [NCH][T]
   double *ct = psm + (size_t)NCH * p.T;     // [NCH][64][PLD]
   unsigned short *rs = reinterpret_cast<unsigned short *>(ct + NCH * 64 * PLD);
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b = blockIdx.x;
   const int base = b * p.T, cnt = min(p.T, p.n - base);
   for (int j = tid; j < NCH * p.T; j += PTHREADS)
      acc[j] = 0.0;
   for (int w = 0; w < p.nw; w++)
   {
      const size_t wbase = (size_t)w * p.nblk * p.T + base;
      const double *__restrict__ xt = p.xt + wbase;
      const unsigned short *__restrict__ lp = p.lp + wbase;
      __syncthreads(); // the previous window's runs are done with ct / rs (and acc is zeroed)
#pragma unroll
      for (int ch = 0; ch < NCH; ch++)
         for (int e = tid; e < 64 * PLD; e += PTHREADS)
            ct[ch * 64 * PLD + e] = __ldg(p.ctab + ((size_t)ch * p.nw + w) * 64 * PLD + e);
      if (tid < 65)
         rs[tid] = p.rs[((size_t)w * p.nblk + b) * B200_POLY_RS + tid];
      __syncthreads();
      for (int c = wid; c < 64; c += PWARPS)
      {
         const int j0 = rs[c], j1 = rs[c + 1];
         if (j0 == j1) // warp-uniform
            continue;
         double cf[NCH][PNC];
#pragma unroll
         for (int ch = 0; ch < NCH; ch++)
#pragma unroll
            for (int k = 0; k < PNC; k++)
               cf[ch][k] = ct[ch * 64 * PLD + c * PLD + k];
         // all loads of (a chunk of) the run first, then the polynomials (PU independent pairs of chains in v^2 per lane)
         for (int jc = j0; jc < j1; jc += 32 * PU)
         {
            double xv[PU];
            int lv[PU];
#pragma unroll
            for (int u = 0; u < PU; u++)
            {
               const int j = jc + lane + 32 * u;
               const bool ok = j < j1;
               xv[u] = ok ? __ldg(xt + j) : 0.0;
               lv[u] = ok ? (int)lp[j] : -1;
            }
#pragma unroll
            for (int u = 0; u < PU; u++)
               if (lv[u] >= 0)
               {
                  const double v = p_offset(xv[u]);
#pragma unroll
                  for (int ch = 0; ch < NCH; ch++)
                     acc[ch * p.T + lv[u]] += p_horner(cf[ch], v); // every point of the block belongs to exactly one run of this window
               }
         }
      }
   }
   __syncthreads();
   const double nwmu = (double)p.nw * mu;
   for (int j = tid; j < cnt; j += PTHREADS)
   {
      const size_t i = (size_t)base + j;
      const double ai = s * __ldg(x + i);
      if (NCH == 1)
         y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + c0 * (acc[j] + nwmu * ai);
      else
      {
         double *y0 = y, *y1 = y + ystride, *y2 = y + 2 * ystride;
         const double a0 = c0 * (acc[j] + nwmu * ai), a1 = acc[p.T + j], a2 = (double)p.nw * ff * ai;
         y0[i] = (beta == 0.0 ? 0.0 : beta * y0[i]) + a0;
         y1[i] = (beta == 0.0 ? 0.0 : beta * y1[i]) + a1;
         y2[i] = (beta == 0.0 ? 0.0 : beta * y2[i]) + a2;
      }
   }
}

// ---------------------------------------------------------------------------------------------------- staged variants (bulk copies)
// The kernels above read a run's offsets / indices straight from global memory and reduce the lanes' moments once per run.  ncu of
// that structure (profiles/r2_l_*): DP instructions are only 25 % of the instructions issued -- 35 % go into the per-run butterfly,
// 22 % into per-item address arithmetic and staging, the rest into per-slot branches that also keep the compiler from interleaving
// the points' dependent chains.  The staged kernels remove all three:
//  * ONE CTA per SM; one thread streams the whole NEXT (window, block) item -- offsets, indices, run starts, the vector block (spread)
//    or the coefficient table (gather) -- into the other half of a double buffer with four bulk copies (cp.async.bulk -> UBLKCP,
//    completion on an mbarrier) while the CTA works on the current one from shared memory;
//  * runs are walked in full chunks of 32 PU points without any predicate (PU independent chains per lane) plus one clamped tail;
//  * spread: the scaled points lie in [-1/4, 1/4], i.e. in the 32 cells with keys 16..47 (key 48 only for a point exactly on the
//    upper end).  Warp w owns the cells 16 + w + PAW h for the whole kernel and keeps their lane-private moments in REGISTERS across
//    the items of a window: the butterfly runs once per window and warp instead of once per run, and its 15 totals go straight to
//    the window's global table as fp64 REDs.  Any other key is flushed around its run.
//  * no CTA barrier per item: the last warp to finish an item re-arms its buffer with the item after the next, so warps drift up
//    to one item apart (the gather therefore accumulates even and odd windows in two arrays).
#ifndef PAW
#define PAW 32 /* warps of the staged kernels */
#endif
#ifndef PUA
#define PUA 2 /* points per lane and chunk of the staged kernels */
#endif
#define PATHREADS (PAW * 32)
#define PNH (32 / PAW) /* home cells per warp */

__device__ __forceinline__ unsigned p_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p_mbar_init(unsigned long long *bar)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(p_saddr(bar)) : "memory");
}
__device__ __forceinline__ void p_mbar_expect(unsigned long long *bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(p_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_mbar_wait(unsigned long long *bar, unsigned parity)
{
   asm volatile("{\n"
                ".reg .pred ok;\n"
                "PWAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 ok, [%0], %1;\n"
                "@ok bra PDONE_%=;\n"
                "bra PWAIT_%=;\n"
                "PDONE_%=:\n"
                "}\n" ::"r"(p_saddr(bar)),
                "r"(parity)
                : "memory");
}
// bytes: a multiple of 16; both addresses 16-byte aligned.  policy: L2 eviction class of the lines read -- the offsets / indices stream
// through once per matvec (evict first), the vector block is read again by every window (evict last; without the hints 25 % of the
// vector reads of the spread came from DRAM: ncu r2_o)
__device__ __forceinline__ void p_bulk(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long policy)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(p_saddr(dst)),
                "l"(src), "r"(bytes), "r"(p_saddr(bar)), "l"(policy)
                : "memory");
}
__device__ __forceinline__ unsigned long long p_policy_stream()
{
   unsigned long long pol;
   asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
   return pol;
}
__device__ __forceinline__ unsigned long long p_policy_keep()
{
   unsigned long long pol;
   asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
   return pol;
}

// double-buffered part of the shared memory of the staged kernels; T is a multiple of 32
struct PStage
{
   double *xt0;               // [2][T]
   unsigned short *lp0, *rs0; // [2][T], [2][B200_POLY_RS]
   unsigned long long *bar;   // [2]
   int T;
   __device__ __forceinline__ double *xt(int buf) const { return xt0 + (size_t)buf * T; }
   __device__ __forceinline__ unsigned short *lp(int buf) const { return lp0 + (size_t)buf * T; }
   __device__ __forceinline__ unsigned short *rs(int buf) const { return rs0 + buf * B200_POLY_RS; }
};
#define PSTAGE_ITEM_BYTES(T) ((unsigned)(T) * 10u + B200_POLY_RS * 2u)
// issued by ONE thread after the mbarrier of the buffer was armed with the byte count of everything that goes into it
__device__ __forceinline__ void p_stage_item(const b200_poly1d_dev &p, const PStage &st, int buf, int w, int b)
{
   const size_t off = (size_t)w * p.nblk * p.T + (size_t)b * p.T;
   const unsigned long long pol = p_policy_stream();
   p_bulk(st.xt(buf), p.xt + off, (unsigned)p.T * 8u, st.bar + buf, pol);
   p_bulk(st.lp(buf), p.lp + off, (unsigned)p.T * 2u, st.bar + buf, pol);
   p_bulk(st.rs(buf), p.rs + ((size_t)w * p.nblk + b) * B200_POLY_RS, B200_POLY_RS * 2u, st.bar + buf, pol);
}

// A warp that has finished an item counts itself off; the LAST one re-arms the buffer with the item after the next.  No CTA barrier per
// item: a fast warp runs up to one item ahead of a slow one (run lengths differ), the data of item k + 2 cannot overwrite item k
// before every warp is done with it.
__device__ __forceinline__ bool p_last_finisher(unsigned *done, int lane)
{
   unsigned old = 0;
   __syncwarp();
   if (lane == 0)
   {
      old = atomicAdd(done, 1u);
      if (old == PAW - 1)
         *done = 0; // nobody counts on this buffer again before it has been refilled and consumed
   }
   return lane == 0 && old == PAW - 1;
}

__global__ void __launch_bounds__(PATHREADS, 1)
k_poly1d_gather_async(b200_poly1d_dev p, const double *__restrict__ x, double s, double c0, double mu, double beta, double *__restrict__ y)
{
   extern __shared__ __align__(16) double psm[];
   double *acc = psm;           // [2][T]: windows of even / odd number (neighbouring windows may be in flight together)
   double *ctb = acc + 2 * (size_t)p.T; // [2][64][PLD]
   PStage st;
   st.T = p.T;
   st.xt0 = ctb + 2 * 64 * PLD;
   st.lp0 = reinterpret_cast<unsigned short *>(st.xt0 + 2 * (size_t)p.T);
   st.rs0 = st.lp0 + 2 * (size_t)p.T;
   st.bar = reinterpret_cast<unsigned long long *>(st.rs0 + 2 * B200_POLY_RS);
   unsigned *done = reinterpret_cast<unsigned *>(st.bar + 2);
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b = blockIdx.x;
   const int base = b * p.T, cnt = min(p.T, p.n - base);
   const unsigned item_bytes = PSTAGE_ITEM_BYTES(p.T) + 64 * PLD * 8;
   auto stage = [&](int w, int buf) // one thread
   {
      p_mbar_expect(st.bar + buf, item_bytes);
      p_stage_item(p, st, buf, w, b);
      p_bulk(ctb + buf * 64 * PLD, p.ctab + (size_t)w * 64 * PLD, 64 * PLD * 8, st.bar + buf, p_policy_keep());
   };
   if (tid == 0)
   {
      p_mbar_init(st.bar);
      p_mbar_init(st.bar + 1);
      done[0] = done[1] = 0;
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      stage(0, 0);
      if (p.nw > 1)
         stage(1, 1);
   }
   for (int j = tid; j < 2 * p.T; j += PATHREADS)
      acc[j] = 0.0;
   __syncthreads();
   for (int w = 0; w < p.nw; w++)
   {
      const int buf = w & 1;
      p_mbar_wait(st.bar + buf, (unsigned)(w >> 1) & 1u);
      const double *xt = st.xt(buf);
      const unsigned short *lp = st.lp(buf), *rs = st.rs(buf);
      const double *ct = ctb + buf * 64 * PLD;
      double *ac = acc + (size_t)buf * p.T;
      for (int c = wid; c < 64; c += PAW)
      {
         const int j0 = rs[c], j1 = rs[c + 1];
         if (j0 == j1) // warp-uniform
            continue;
         double cf[PLD]; // the row is 128 bytes, aligned: 8 x LDS.128 (entry 15 is padding)
#pragma unroll
         for (int k = 0; k < PLD; k += 2)
         {
            const double2 t = *reinterpret_cast<const double2 *>(ct + c * PLD + k);
            cf[k] = t.x;
            cf[k + 1] = t.y;
         }
         int jc = j0;
         for (; jc + 32 * PUA <= j1; jc += 32 * PUA) // full chunks: no predicate, PUA independent Horner chains per lane
         {
            double v[PUA], f[PUA];
            int l[PUA];
#pragma unroll
            for (int u = 0; u < PUA; u++)
            {
               v[u] = xt[jc + lane + 32 * u];
               l[u] = lp[jc + lane + 32 * u];
            }
#pragma unroll
            for (int u = 0; u < PUA; u++)
               f[u] = p_horner(cf, v[u]);
#pragma unroll
            for (int u = 0; u < PUA; u++)
               ac[l[u]] += f[u]; // every point of the block belongs to exactly one run of this window
         }
         for (; jc < j1; jc += 32) // tail: clamped loads, one predicated update
         {
            const int j = jc + lane, jj = min(j, j1 - 1);
            const double f = p_horner(cf, xt[jj]);
            if (j < j1)
               ac[lp[jj]] += f;
         }
      }
      if (p_last_finisher(done + buf, lane) && w + 2 < p.nw)
         stage(w + 2, buf);
   }
   __syncthreads();
   const double nwmu = (double)p.nw * mu;
   for (int j = tid; j < cnt; j += PATHREADS)
   {
      const size_t i = (size_t)base + j;
      const double ai = s * __ldg(x + i);
      y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + c0 * (acc[j] + acc[p.T + j] + nwmu * ai);
   }
}

// moments of the run [j0, j1) of one cell into the lane-private m
__device__ __forceinline__ void p_run_moments(double m[16], const double *xs, const double *xt, const unsigned short *lp, int j0, int j1, int lane)
{
   int jc = j0;
   for (; jc + 32 * PUA <= j1; jc += 32 * PUA)
   {
      double v[PUA], a[PUA];
#pragma unroll
      for (int u = 0; u < PUA; u++)
      {
         v[u] = xt[jc + lane + 32 * u];
         a[u] = xs[lp[jc + lane + 32 * u]];
      }
#pragma unroll
      for (int u = 0; u < PUA; u++)
         p_moments(m, a[u], v[u]);
   }
   for (; jc < j1; jc += 32)
   {
      const int j = jc + lane, jj = min(j, j1 - 1);
      const double a = xs[lp[jj]];
      p_moments(m, j < j1 ? a : 0.0, xt[jj]);
   }
}
// lanes' moments of one cell -> the window's global table (fp64 REDs: 15 per warp), scaled; m is cleared
__device__ __forceinline__ void p_flush_moments(double m[16], double *mom_w, int key, double s, int lane)
{
   int q;
   const double tot = p_transpose_reduce16(m, lane, q);
   if (!(lane & 1) && q < PNC && tot != 0.0)
      p_red_add(mom_w + key * PLD + q, s * tot);
#pragma unroll
   for (int k = 0; k < 16; k++)
      m[k] = 0.0;
}

// the block of the vector is staged raw (x has to be 16-byte aligned: checked at launch); the factor s goes onto the moments
__global__ void __launch_bounds__(PATHREADS, 1)
k_poly1d_spread_async(b200_poly1d_dev p, const double *__restrict__ x, double s)
{
   extern __shared__ __align__(16) double psm[];
   double *xs0 = psm; // [2][T]
   PStage st;
   st.T = p.T;
   st.xt0 = xs0 + 2 * (size_t)p.T;
   st.lp0 = reinterpret_cast<unsigned short *>(st.xt0 + 2 * (size_t)p.T);
   st.rs0 = st.lp0 + 2 * (size_t)p.T;
   st.bar = reinterpret_cast<unsigned long long *>(st.rs0 + 2 * B200_POLY_RS);
   unsigned *done = reinterpret_cast<unsigned *>(st.bar + 2);
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   const long long items = (long long)p.nw * p.nblk;
   const long long it0 = items * blockIdx.x / gridDim.x, it1 = items * (blockIdx.x + 1) / gridDim.x;
   const int nit = (int)(it1 - it0);
   int w = (int)(it0 / p.nblk), b = (int)(it0 - (long long)w * p.nblk), wcur = -1;
   double m[PNH][16];
#pragma unroll
   for (int h = 0; h < PNH; h++)
#pragma unroll
      for (int k = 0; k < 16; k++)
         m[h][k] = 0.0;
   // one thread: everything of the item (w, b) into buffer `buf`
   auto stage = [&](int sw, int sb, int buf)
   {
      const int cnt = min(p.T, p.n - sb * p.T), even = cnt & ~1;
      double *xs = xs0 + (size_t)buf * p.T;
      if (cnt & 1)
         xs[cnt - 1] = __ldg(x + (size_t)sb * p.T + cnt - 1); // released to the CTA by the arrive below
      p_mbar_expect(st.bar + buf, PSTAGE_ITEM_BYTES(p.T) + (unsigned)even * 8u);
      p_stage_item(p, st, buf, sw, sb);
      if (even)
         p_bulk(xs, x + (size_t)sb * p.T, (unsigned)even * 8u, st.bar + buf, p_policy_keep());
   };
   // the item `ahead` places after (w, b)
   auto item_after = [&](int ahead, int &sw, int &sb)
   {
      sw = w;
      sb = b + ahead;
      while (sb >= p.nblk)
      {
         sb -= p.nblk;
         sw++;
      }
   };
   if (tid == 0)
   {
      p_mbar_init(st.bar);
      p_mbar_init(st.bar + 1);
      done[0] = done[1] = 0;
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      if (nit > 0)
         stage(w, b, 0);
      if (nit > 1)
      {
         int sw, sb;
         item_after(1, sw, sb);
         stage(sw, sb, 1);
      }
   }
   __syncthreads();
   for (int k = 0; k < nit; k++)
   {
      const int buf = k & 1;
      if (w != wcur) // per warp: its home cells' moments of the finished window go to the global table
      {
         if (wcur >= 0)
         {
#pragma unroll
            for (int h = 0; h < PNH; h++)
               p_flush_moments(m[h], p.mom + (size_t)wcur * 64 * PLD, 16 + wid + PAW * h, s, lane);
         }
         wcur = w;
      }
      p_mbar_wait(st.bar + buf, (unsigned)(k >> 1) & 1u);
      const double *xs = xs0 + (size_t)buf * p.T, *xt = st.xt(buf);
      const unsigned short *lp = st.lp(buf), *rs = st.rs(buf);
#pragma unroll
      for (int h = 0; h < PNH; h++) // the warp's home cells: moments stay in registers
      {
         const int c = 16 + wid + PAW * h;
         p_run_moments(m[h], xs, xt, lp, rs[c], rs[c + 1], lane);
      }
      for (int c = wid; c < 32; c += PAW) // keys 0..15 and 48..63: empty unless a point sits exactly on the upper end
      {
         const int key = c < 16 ? c : c + 32;
         const int j0 = rs[key], j1 = rs[key + 1];
         if (j0 == j1) // warp-uniform
            continue;
         double *mom_w = p.mom + (size_t)w * 64 * PLD;
         p_flush_moments(m[0], mom_w, 16 + wid, s, lane);
         p_run_moments(m[0], xs, xt, lp, j0, j1, lane);
         p_flush_moments(m[0], mom_w, key, s, lane);
      }
      if (p_last_finisher(done + buf, lane) && k + 2 < nit)
      {
         int sw, sb;
         item_after(2, sw, sb);
         stage(sw, sb, buf);
      }
      if (++b == p.nblk)
      {
         b = 0;
         w++;
      }
   }
   if (wcur >= 0)
   {
#pragma unroll
      for (int h = 0; h < PNH; h++)
         p_flush_moments(m[h], p.mom + (size_t)wcur * 64 * PLD, 16 + wid + PAW * h, s, lane);
   }
}

// ---------------------------------------------------------------------------------------------------- launch layer
extern "C" {

static size_t poly_smem_spread(int T) { return sizeof(double) * ((size_t)T + 64 * PLD) + 2 * 80; }
// staged variants: acc [2][T] and the coefficient tables [2] (gather) or the vector block [2][T] (spread) + two buffers of (T doubles + T + B200_POLY_RS
// shorts) + 2 mbarriers + 2 counters
static size_t poly_smem_spread_async(int T) { return sizeof(double) * (2 * (size_t)T + 2 * (size_t)T) + 2 * 2 * ((size_t)T + B200_POLY_RS) + 16 + 8; }
static size_t poly_smem_gather_async(int T) { return sizeof(double) * (2 * (size_t)T + 2 * 64 * PLD + 2 * (size_t)T) + 2 * 2 * ((size_t)T + B200_POLY_RS) + 16 + 8; }
static bool poly_async_enabled(void)
{
   static int on = -1;
   if (on < 0)
   {
      const char *e = getenv("NFFT4GP_B200_POLY1D_ASYNC"); // 0: the direct-load kernels (A/B, cross-check)
      on = (e && atoi(e) == 0) ? 0 : 1;
   }
   return on != 0;
}
static size_t poly_smem_gather(int T, int nch) { return sizeof(double) * (size_t)nch * ((size_t)T + 64 * PLD) + 2 * 80; }

// largest block size (points per locally sorted block) the gather kernel of the gradient matvec still fits into shared memory with
// largest block the kernels hold in shared memory: 5984 points is where the staged gather (36 B per point + tables) still fits 227 KB
int b200k_poly1d_max_block(void) { return poly_async_enabled() ? 5984 : 8192; }
// CTAs of one resident wave of the spread / gather kernels (staged: one per SM; direct: two per SM, registers and shared memory)
int b200k_poly1d_wave_slots(void)
{
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   return poly_async_enabled() ? sm : 2 * sm;
}

// x (n values, original order) -> oversampled grids of the nw local windows (64 doubles each, stride grid_stride)
int b200k_poly1d_spread(const b200_poly1d_dev *p, const double *d_x, double s, double *d_grids, size_t grid_stride)
{
   B200ProfScope prof_scope_(PC_SPREAD);
   static bool attr_set = false;
   if (!attr_set)
   {
      B200_CHECK(cudaFuncSetAttribute(k_poly1d_spread, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poly_smem_spread(b200k_poly1d_max_block())));
      attr_set = true;
   }
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   long long items = (long long)p->nw * p->nblk;
   int G = 2 * sm; // one wave: two CTAs per SM (registers and shared memory allow exactly two)
   if (G > items)
      G = (int)items;
   if (G < 1)
      G = 1;
   B200_CHECK(cudaMemsetAsync(p->mom, 0, sizeof(double) * (size_t)p->nw * 64 * PLD, b200_stream()));
   if (poly_async_enabled() && poly_smem_spread_async(p->T) <= 227 * 1024 && ((uintptr_t)d_x & 15) == 0)
   {
      static bool attr2 = false;
      if (!attr2)
      {
         B200_CHECK(cudaFuncSetAttribute(k_poly1d_spread_async, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
         attr2 = true;
      }
      int G1 = sm < items ? sm : (int)items; // one CTA per SM
      k_poly1d_spread_async<<<G1, PATHREADS, poly_smem_spread_async(p->T), b200_stream()>>>(*p, d_x, s);
   }
   else
      k_poly1d_spread<<<G, PTHREADS, poly_smem_spread(p->T), b200_stream()>>>(*p, d_x, s);
   B200_LAUNCHED();
   k_poly1d_mom2grid<<<p->nw, 64, 0, b200_stream()>>>(p->mom, d_grids, grid_stride);
   B200_LAUNCHED();
   return 0;
}

// channel ch (0 kernel, 1 derivative kernel) of the coefficient tables from Re h of the nw local windows; d_wscale: optional per-window factor
int b200k_poly1d_coef(const b200_poly1d_dev *p, int ch, const double *d_grids, size_t grid_stride, const double *d_wscale)
{
   B200ProfScope prof_scope_(PC_GATHER);
   k_poly1d_grid2coef<<<p->nw, 64 * PLD, 0, b200_stream()>>>(d_grids, grid_stride, d_wscale, p->ctab + (size_t)ch * p->nw * 64 * PLD);
   B200_LAUNCHED();
   return 0;
}

// nch = 1: matvec epilogue into y (n values); nch = 2: gradient epilogue into y (3 blocks of ystride values)
int b200k_poly1d_gather(const b200_poly1d_dev *p, int nch, const double *d_x, double s, double c0, double mu, double ff, double beta, double *d_y,
                        size_t ystride)
{
   B200ProfScope prof_scope_(PC_GATHER);
   static bool attr_set = false;
   if (!attr_set)
   {
      B200_CHECK(cudaFuncSetAttribute(k_poly1d_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poly_smem_gather(b200k_poly1d_max_block(), 1)));
      B200_CHECK(cudaFuncSetAttribute(k_poly1d_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poly_smem_gather(b200k_poly1d_max_block(), 2)));
      attr_set = true;
   }
   if (nch == 1 && poly_async_enabled() && poly_smem_gather_async(p->T) <= 227 * 1024)
   {
      static bool attr2 = false;
      if (!attr2)
      {
         B200_CHECK(cudaFuncSetAttribute(k_poly1d_gather_async, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
         attr2 = true;
      }
      k_poly1d_gather_async<<<p->nblk, PATHREADS, poly_smem_gather_async(p->T), b200_stream()>>>(*p, d_x, s, c0, mu, beta, d_y);
   }
   else if (nch == 1)
      k_poly1d_gather<1><<<p->nblk, PTHREADS, poly_smem_gather(p->T, 1), b200_stream()>>>(*p, d_x, s, c0, mu, ff, beta, d_y, ystride);
   else
      k_poly1d_gather<2><<<p->nblk, PTHREADS, poly_smem_gather(p->T, 2), b200_stream()>>>(*p, d_x, s, c0, mu, ff, beta, d_y, ystride);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
