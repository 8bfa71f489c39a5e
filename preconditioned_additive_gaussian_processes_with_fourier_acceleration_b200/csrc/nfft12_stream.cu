// Streaming spread (adjoint NFFT, B^T) and gather (NFFT trafo, B) for 1-D and 2-D windows (sm_100a).
//
// These windows are HBM-bound (1.25 / 8.3 flop per byte): per point the kernels must move the tap row
// (80 / 160 B, the PRE_PSI analogue) and one coefficient, and do 10 / 100 FMAs.  The oversampled grid is tiny (64 or 64^2
// values) while thousands of cell-sorted points share a cell, so the thread-per-point kernels (nfft_kernels.cu) serialise
// on same-address atomics: 1.14 ms per 4 * 10^6-point 2-D window instead of the 0.1 ms the table read takes.
// Here the roles are transposed, as in the 3-D kernels: a warp owns a contiguous range of sorted points and its LANES OWN
// THE OUTPUTS of the current cell's footprint (2-D: lane L < 25 owns the 2 x 2 block (2 i1 + {0,1}, 2 i2 + {0,1}); 1-D: lane
// L < 10 owns tap L) in registers; tap rows are staged 32 points at a time through shared memory with fully coalesced
// 128-bit loads, and the footprint touches the L2-resident grid only when the stream enters another cell.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"
#include "kb_window_coeffs.h"

#define TAPS B200_TAPS
#define NOS B200_NOS
#define SWARPS 8     /* warps per CTA */
#define SBATCH 32    /* points staged per warp and step */
#define SCTAS_PER_SM 4

__device__ __forceinline__ int s_tap_idx(int keyfield, int lj) { return (keyfield + 28 + lj) & 63; }

template <int D>
struct StreamSmem
{
   double rows[SWARPS][SBATCH][D * TAPS];
   double a[SWARPS][SBATCH];
   int key[SWARPS][SBATCH];
};

// stage SBATCH tap rows (and optionally the coefficients) of the points [j0, j0 + SBATCH) of this warp; rows past `jend` are zero
template <int D>
__device__ __forceinline__ void s_stage(StreamSmem<D> &sm, int wid, int lane, const double *__restrict__ psi, const double *__restrict__ a,
                                        const int *__restrict__ key, int j0, int jend)
{
   constexpr int RD = D * TAPS;              // doubles per row (10 or 20: even, rows are 16-byte aligned)
   constexpr int V2 = SBATCH * RD / 2;       // double2 per batch
   const double2 *src = reinterpret_cast<const double2 *>(psi + (size_t)j0 * RD);
   double2 *dst = reinterpret_cast<double2 *>(&sm.rows[wid][0][0]);
   const int valid2 = (jend - j0) * RD / 2;  // double2 inside the range
#pragma unroll
   for (int k = 0; k < V2 / 32; k++)
   {
      int idx = lane + 32 * k;
      dst[idx] = idx < valid2 ? __ldg(src + idx) : make_double2(0.0, 0.0);
   }
   const int j = j0 + lane;
   if (a != nullptr)
      sm.a[wid][lane] = j < jend ? a[j] : 0.0;
   sm.key[wid][lane] = j < jend ? key[j] : -1;
}

// Opt-in alternative to reading the tabulated taps (NFFT4GP_B200_WINDOW_EVAL=1): evaluate the Kaiser-Bessel window here.
// Per tap j the argument t = u + 4 - j stays inside ONE unit interval while the point's fractional grid offset u runs over
// [0, 1), so psi is ten smooth pieces; each is a degree-14 polynomial in v = 2u - 1 (kb_window_coeffs.h, 4.6e-16 phi(0)).
// Lane = point of the batch: 8 bytes per dimension from HBM instead of 80, ~150 FMAs per dimension, coefficients in constant
// memory (all lanes read the same entry).  The staged rows are the same as the tabulated ones to rounding.
__constant__ double c_kb_poly[TAPS][KB_POLY_DEG + 1] = KB_POLY_TABLE;

template <int D>
__device__ __forceinline__ void s_stage_eval(StreamSmem<D> &sm, int wid, int lane, const double *__restrict__ xs, const double *__restrict__ a,
                                             const int *__restrict__ key, int j0, int jend)
{
   const int j = j0 + lane;
   const bool ok = j < jend;
#pragma unroll
   for (int t = 0; t < D; t++)
   {
      const double xj = ok ? xs[(size_t)j * D + t] : 0.0;
      const double xn = xj * (double)NOS;
      const double v = 2.0 * (xn - floor(xn)) - 1.0;
#pragma unroll
      for (int lj = 0; lj < TAPS; lj++)
      {
         double acc = c_kb_poly[lj][KB_POLY_DEG];
#pragma unroll
         for (int k = KB_POLY_DEG - 1; k >= 0; k--)
            acc = fma(acc, v, c_kb_poly[lj][k]);
         sm.rows[wid][lane][t * TAPS + lj] = ok ? acc : 0.0;
      }
   }
   if (a != nullptr)
      sm.a[wid][lane] = ok ? a[j] : 0.0;
   sm.key[wid][lane] = ok ? key[j] : -1;
}

// ---------------------------------------------------------------------------------------------------- spread
template <int D, bool EVAL>
__global__ void __launch_bounds__(SWARPS * 32, SCTAS_PER_SM)
k_spread_stream(const double *__restrict__ psi, const int *__restrict__ key, const double *__restrict__ a, double *__restrict__ grid, int n,
                int per_warp)
{
   __shared__ __align__(16) StreamSmem<D> sm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * SWARPS + wid;
   const long long b64 = (long long)gw * per_warp;
   if (b64 >= n)
      return;
   const int jb = (int)b64, je = min(n, jb + per_warp);
   // 2-D: lane L < 25 owns (x, y) in {2 i1, 2 i1 + 1} x {2 i2, 2 i2 + 1};  1-D: lane L < 10 owns x = L
   const bool act = D == 2 ? lane < 25 : lane < TAPS;
   const int i1 = D == 2 ? (act ? lane / 5 : 0) : (act ? lane : 0);
   const int i2 = D == 2 ? (act ? lane % 5 : 0) : 0;
   double acc00 = 0.0, acc01 = 0.0, acc10 = 0.0, acc11 = 0.0;
   int cur = key[jb];

   auto flush = [&](int k) {
      if (!act)
         return;
      if (D == 2)
      {
         const int c1 = (k >> 6) & 63, c2 = k & 63;
         double *r0 = grid + s_tap_idx(c1, 2 * i1) * NOS, *r1 = grid + s_tap_idx(c1, 2 * i1 + 1) * NOS;
         const int y0 = s_tap_idx(c2, 2 * i2), y1 = s_tap_idx(c2, 2 * i2 + 1);
         atomicAdd(r0 + y0, acc00);
         atomicAdd(r0 + y1, acc01);
         atomicAdd(r1 + y0, acc10);
         atomicAdd(r1 + y1, acc11);
      }
      else
         atomicAdd(grid + s_tap_idx(k & 63, i1), acc00);
   };

   for (int j0 = jb; j0 < je; j0 += SBATCH)
   {
      __syncwarp();
      if (EVAL)
         s_stage_eval<D>(sm, wid, lane, psi /* = sorted scaled coordinates */, a, key, j0, je);
      else
         s_stage<D>(sm, wid, lane, psi, a, key, j0, je);
      __syncwarp();
      const int cnt = min(SBATCH, je - j0);
      // fast path (the common case: thousands of sorted points per cell): the whole batch lies in the current cell
      const int mykey = sm.key[wid][lane];
      if (cnt == SBATCH && __all_sync(0xffffffffu, mykey == cur))
      {
#pragma unroll 8
         for (int p = 0; p < SBATCH; p++)
         {
            const double *row = sm.rows[wid][p];
            const double aj = sm.a[wid][p];
            if (D == 2)
            {
               const double2 p1 = *reinterpret_cast<const double2 *>(row + 2 * i1);
               const double2 p2 = *reinterpret_cast<const double2 *>(row + TAPS + 2 * i2);
               const double q0 = aj * p1.x, q1 = aj * p1.y;
               acc00 = fma(q0, p2.x, acc00);
               acc01 = fma(q0, p2.y, acc01);
               acc10 = fma(q1, p2.x, acc10);
               acc11 = fma(q1, p2.y, acc11);
            }
            else
               acc00 = fma(aj, row[i1], acc00);
         }
         continue;
      }
      for (int p = 0; p < cnt; p++)
      {
         const int kp = sm.key[wid][p];
         if (kp != cur) // warp-uniform
         {
            flush(cur);
            acc00 = acc01 = acc10 = acc11 = 0.0;
            cur = kp;
         }
         const double *row = sm.rows[wid][p];
         const double aj = sm.a[wid][p];
         if (D == 2)
         {
            const double2 p1 = *reinterpret_cast<const double2 *>(row + 2 * i1);
            const double2 p2 = *reinterpret_cast<const double2 *>(row + TAPS + 2 * i2);
            const double q0 = aj * p1.x, q1 = aj * p1.y;
            acc00 = fma(q0, p2.x, acc00);
            acc01 = fma(q0, p2.y, acc01);
            acc10 = fma(q1, p2.x, acc10);
            acc11 = fma(q1, p2.y, acc11);
         }
         else
            acc00 = fma(aj, row[i1], acc00);
      }
   }
   flush(cur);
}

// ---------------------------------------------------------------------------------------------------- gather
template <int D, bool EVAL>
__global__ void __launch_bounds__(SWARPS * 32, SCTAS_PER_SM)
k_gather_stream(const double *__restrict__ psi, const int *__restrict__ key, const double *__restrict__ h, double *__restrict__ f, int n,
                int per_warp)
{
   __shared__ __align__(16) StreamSmem<D> sm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int gw = blockIdx.x * SWARPS + wid;
   const unsigned full = 0xffffffffu;
   const long long b64 = (long long)gw * per_warp;
   if (b64 >= n)
      return;
   const int jb = (int)b64, je = min(n, jb + per_warp);
   const bool act = D == 2 ? lane < 25 : lane < TAPS;
   const int i1 = D == 2 ? (act ? lane / 5 : 0) : (act ? lane : 0);
   const int i2 = D == 2 ? (act ? lane % 5 : 0) : 0;
   double h00 = 0.0, h01 = 0.0, h10 = 0.0, h11 = 0.0; // this lane's block of Re h for the current cell (idle lanes: zeros)
   int cur = -2;

   auto load = [&](int k) {
      if (!act)
         return;
      if (D == 2)
      {
         const int c1 = (k >> 6) & 63, c2 = k & 63;
         const double *r0 = h + s_tap_idx(c1, 2 * i1) * NOS, *r1 = h + s_tap_idx(c1, 2 * i1 + 1) * NOS;
         const int y0 = s_tap_idx(c2, 2 * i2), y1 = s_tap_idx(c2, 2 * i2 + 1);
         h00 = __ldg(r0 + y0);
         h01 = __ldg(r0 + y1);
         h10 = __ldg(r1 + y0);
         h11 = __ldg(r1 + y1);
      }
      else
         h00 = __ldg(h + s_tap_idx(k & 63, i1));
   };

   for (int j0 = jb; j0 < je; j0 += SBATCH)
   {
      __syncwarp();
      if (EVAL)
         s_stage_eval<D>(sm, wid, lane, psi /* = sorted scaled coordinates */, nullptr, key, j0, je);
      else
         s_stage<D>(sm, wid, lane, psi, nullptr, key, j0, je);
      __syncwarp();
      const int cnt = min(SBATCH, je - j0);
      for (int p0 = 0; p0 < cnt; p0 += 8)
      {
         double v[8];
#pragma unroll
         for (int q = 0; q < 8; q++)
         {
            const int p = p0 + q;
            const int kp = sm.key[wid][p < SBATCH ? p : SBATCH - 1];
            if (p < cnt && kp != cur) // warp-uniform
            {
               cur = kp;
               load(cur);
            }
            const double *row = sm.rows[wid][p < SBATCH ? p : SBATCH - 1];
            if (D == 2)
            {
               const double2 p1 = *reinterpret_cast<const double2 *>(row + 2 * i1);
               const double2 p2 = *reinterpret_cast<const double2 *>(row + TAPS + 2 * i2);
               const double t0 = fma(h01, p2.y, h00 * p2.x), t1 = fma(h11, p2.y, h10 * p2.x);
               v[q] = fma(t1, p1.y, t0 * p1.x);
            }
            else
               v[q] = h00 * row[i1];
         }
         // transpose-reduce: 8 per-lane partials x 32 lanes -> the sum of point q lands in the lanes with (lane >> 2) == q'
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
         const int q = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
         if ((lane & 3) == 0 && p0 + q < cnt)
            f[j0 + p0 + q] = w1;
      }
   }
}

// ---------------------------------------------------------------------------------------------------- launch layer
static int stream_grid(int n, int *per_warp)
{
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   long long warps = (long long)sm * SCTAS_PER_SM * SWARPS;
   long long per = (n + warps - 1) / warps;
   per = ((per + SBATCH - 1) / SBATCH) * SBATCH; // whole batches: staged rows start 16-byte aligned
   if (per < SBATCH)
      per = SBATCH;
   *per_warp = (int)per;
   long long used = (n + per - 1) / per;
   return (int)((used + SWARPS - 1) / SWARPS);
}

extern "C" {

int b200k_spread_stream(const b200_win_dev *w)
{
   if (w->n <= 0)
      return 0;
   int per, blocks = stream_grid(w->n, &per);
   if (w->xs && w->d == 1)
      k_spread_stream<1, true><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->xs, w->key, w->a, w->grid, w->n, per);
   else if (w->xs)
      k_spread_stream<2, true><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->xs, w->key, w->a, w->grid, w->n, per);
   else if (w->d == 1)
      k_spread_stream<1, false><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->psi, w->key, w->a, w->grid, w->n, per);
   else
      k_spread_stream<2, false><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->psi, w->key, w->a, w->grid, w->n, per);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather_stream(const b200_win_dev *w, double *d_f)
{
   if (w->n <= 0)
      return 0;
   int per, blocks = stream_grid(w->n, &per);
   if (w->xs && w->d == 1)
      k_gather_stream<1, true><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->xs, w->key, w->grid, d_f, w->n, per);
   else if (w->xs)
      k_gather_stream<2, true><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->xs, w->key, w->grid, d_f, w->n, per);
   else if (w->d == 1)
      k_gather_stream<1, false><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->psi, w->key, w->grid, d_f, w->n, per);
   else
      k_gather_stream<2, false><<<blocks, SWARPS * 32, 0, b200_stream()>>>(w->psi, w->key, w->grid, d_f, w->n, per);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
