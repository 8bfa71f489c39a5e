// 2-D feature windows WITHOUT tap tables, on the fp64 tensor cores of sm_100a.
//
// The tap-table kernels for 2-D windows (nfft2d_mma.cu) are bound by their table: 172 B per point and direction against 24 B of
// algorithmic traffic (ncu r2_h: 79 - 86 % of the copy bandwidth on the table bytes).  On a grid cell every Kaiser-Bessel tap is a
// polynomial of degree 14 in the in-cell offset v = 2 u - 1 (kb_window_coeffs.h, exact to 4.6e-16 phi(0); the 1-D path of
// nfft1d_poly.cu rests on the same fact), so for a point in the 2-D cell (c1, c2)
//   gather:  f_j = sum_{t1,t2} h[c1-4+t1][c2-4+t2] p_t1(v1) p_t2(v2) = sum_{k1,k2} C_cell[k1][k2] v1^k1 v2^k2,  C_cell = P' H_cell P
//   spread:  g[c1-4+t1][c2-4+t2] += sum_{k1,k2} P[t1][k1] M_cell[k1][k2] P[t2][k2],  M_cell[k1][k2] = sum_{j in cell} a_j v1_j^k1 v2_j^k2
// with P the 10 x 15 coefficient matrix of the taps.  Both sums over the points of a cell are GEMMs with Vandermonde operands:
//   spread:  M(16 x 16) += A(16 x J) B(J x 16),  A[k1][j] = a_j v1_j^k1, B[j][k2] = v2_j^k2:  4 mma.sync.m8n8k4.f64 per 4 points;
//   gather:  T(J x 16) = V1(J x 16) C(16 x 16), f_j = sum_k2 T[j][k2] v2_j^k2:                8 mma.sync.m8n8k4.f64 per 8 points,
// i.e. ONE tensor instruction per point and direction (256 FMA slots for the 225 products of the bivariate polynomial) plus ~4 fp64
// products per thread for the powers: a thread always needs the SAME power (its row / column of the fragment) of varying points.
// Per point the kernels read the two offsets (16 B), the cell key (4 B) and the coefficient (8 B): 28 B instead of 172 B; the fp64
// unit becomes the bound (~22 clk of a sub-partition's pipe per point and direction).  The moments / coefficients of the 33 x 33
// occupied cells are converted from / to the 64 x 64 grid by two small kernels (one CTA per occupied cell, separable 15 <-> 10
// transforms).  Work decomposition as in nfft2d_mma.cu: a warp owns a contiguous range of cell-sorted points, 32 points per step.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"
#include "kb_window_coeffs.h"

#define TAPS B200_TAPS
#define NOS B200_NOS
#define PNC (KB_POLY_DEG + 1) /* 15 coefficients */
#define PLD 16                /* leading dimension of the per-cell 16 x 16 tables */
#define RWARPS 8              /* warps per CTA */
#define RBATCH 32             /* points per warp and step */
#define RCTAS_PER_SM 4

__constant__ double c_p2[TAPS][PNC] = KB_POLY_TABLE;

__device__ __forceinline__ void r_dmma(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int r_tap_idx(int keyfield, int lj) { return (keyfield + 28 + lj) & 63; }
__device__ __forceinline__ void r_red_add(double *p, double v)
{
   asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
// v^e for the thread-constant exponent e in 0..7, given v^2 and v^4
__device__ __forceinline__ double r_pow8(double v, double v2, double v4, int e)
{
   double p = (e & 1) ? v : 1.0;
   if (e & 2)
      p *= v2;
   if (e & 4)
      p *= v4;
   return p;
}

// per warp: the 32 points of a step.  The repeated squares of the offsets are formed ONCE per point when the step is staged (lane j for
// point j) -- the 8 threads that share a point in a fragment would each recompute them
struct RSmem
{
   double x[4][RWARPS][RBATCH], y[4][RWARPS][RBATCH]; // [i]: v^(2^i) = v, v^2, v^4, v^8
   double a[RWARPS][RBATCH];
   int key[RWARPS][RBATCH], out[RWARPS][RBATCH]; // out: where the gather writes the point's value
};
__device__ __forceinline__ void r_stage(RSmem &sm, int wid, int lane, double2 v, double a, int key, int out = 0)
{
   sm.out[wid][lane] = out;
   const double x2 = v.x * v.x, x4 = x2 * x2, y2 = v.y * v.y, y4 = y2 * y2;
   sm.x[0][wid][lane] = v.x;
   sm.x[1][wid][lane] = x2;
   sm.x[2][wid][lane] = x4;
   sm.x[3][wid][lane] = x4 * x4;
   sm.y[0][wid][lane] = v.y;
   sm.y[1][wid][lane] = y2;
   sm.y[2][wid][lane] = y4;
   sm.y[3][wid][lane] = y4 * y4;
   sm.a[wid][lane] = a;
   sm.key[wid][lane] = key;
}

// ---------------------------------------------------------------------------------------------------- offsets and keys
// sorted scaled coordinates -> in-cell offsets (exact: 64 x is exact, so are its fractional part u and v = 2 u - 1) and the packed
// cell key of k_psi (nfft_kernels.cu)
__global__ void k_poly2d_offsets(const double *__restrict__ xs, double2 *__restrict__ pv, int *__restrict__ key, int n)
{
   const int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= n)
      return;
   double v[2];
   int k = 0;
#pragma unroll
   for (int t = 0; t < 2; t++)
   {
      const double tt = xs[(size_t)j * 2 + t] * (double)NOS, fl = floor(tt);
      k = (k << 6) | (((int)fl + 32) & 63);
      v[t] = 2.0 * (tt - fl) - 1.0;
   }
   pv[j] = make_double2(v[0], v[1]);
   key[j] = k;
}

// ---------------------------------------------------------------------------------------------------- spread: moments on DMMA
// accumulator tile (tx, ty): rows k1 = 8 tx + g, columns k2 = 8 ty + 2 t + {0, 1}
__device__ __forceinline__ void r_flush(double c[2][2][2], double *__restrict__ mom, int key, int g, int t)
{
   double *m = mom + (size_t)key * (PLD * PLD);
#pragma unroll
   for (int tx = 0; tx < 2; tx++)
#pragma unroll
      for (int ty = 0; ty < 2; ty++)
#pragma unroll
         for (int e = 0; e < 2; e++)
         {
            const int k1 = 8 * tx + g, k2 = 8 * ty + 2 * t + e;
            if (k1 < PNC && k2 < PNC && c[tx][ty][e] != 0.0)
               r_red_add(m + k1 * PLD + k2, c[tx][ty][e]);
            c[tx][ty][e] = 0.0;
         }
}

// one k-step: the four points 4 ks .. 4 ks + 3 of the batch; only = -1: all four, else only point `only` of the step carries weight
__device__ __forceinline__ void r_spread_step(const RSmem &sm, int wid, int ks, int g, int t, int only, double c[2][2][2])
{
   const int idx = 4 * ks + t;
   double aj = sm.a[wid][idx];
   if (only >= 0 && t != only)
      aj = 0.0;
   const double a0 = aj * r_pow8(sm.x[0][wid][idx], sm.x[1][wid][idx], sm.x[2][wid][idx], g), a1 = a0 * sm.x[3][wid][idx];
   const double b0 = r_pow8(sm.y[0][wid][idx], sm.y[1][wid][idx], sm.y[2][wid][idx], g), b1 = b0 * sm.y[3][wid][idx];
   r_dmma(c[0][0][0], c[0][0][1], a0, b0);
   r_dmma(c[0][1][0], c[0][1][1], a0, b1);
   r_dmma(c[1][0][0], c[1][0][1], a1, b0);
   r_dmma(c[1][1][0], c[1][1][1], a1, b1);
}

// FUSED: the coefficient of a sorted point is s x[perm[j]], fetched here one step ahead (the random 8-byte reads of the vector, which is
// L2-resident, hide behind the tensor work of the current step) instead of a separate pass that writes it to a[] in sorted order
template <bool FUSED>
__global__ void __launch_bounds__(RWARPS * 32, RCTAS_PER_SM)
k_spread2_poly(const double2 *__restrict__ pv, const int *__restrict__ key, const double *__restrict__ a, const int *__restrict__ perm, double s,
               double *__restrict__ mom, int n, int per_warp)
{
   __shared__ RSmem sm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int g = lane >> 2, t = lane & 3;
   const long long b64 = (long long)(blockIdx.x * RWARPS + wid) * per_warp;
   if (b64 >= n)
      return;
   const int jb = (int)b64, je = min(n, jb + per_warp);
   // coefficient of the sorted point j (this lane's point of the step that is fetched now).  FUSED: `pnext` holds perm[j], loaded one
   // step earlier; the vector entry is read now and the index of the step after is requested, so neither load waits for the other
   int pnext = -1;
   auto coef = [&](int j, bool ok) -> double
   {
      if (!FUSED)
         return ok ? __ldg(a + j) : 0.0;
      const double v = pnext >= 0 ? s * __ldg(a + pnext) : 0.0;
      pnext = j + RBATCH < je ? __ldg(perm + j + RBATCH) : -1;
      return v;
   };
   double c[2][2][2];
#pragma unroll
   for (int i = 0; i < 8; i++)
      (&c[0][0][0])[i] = 0.0;
   int cur = key[jb];
   // the next step's point is fetched while the current step is computed (rows past the end of the range carry no weight)
   bool okn = jb + lane < je;
   double2 vn = okn ? __ldg(pv + jb + lane) : make_double2(0.0, 0.0);
   if (FUSED && okn)
      pnext = __ldg(perm + jb + lane);
   double an = coef(jb + lane, okn);
   int kn = okn ? __ldg(key + jb + lane) : -1;
   for (int j0 = jb; j0 < je; j0 += RBATCH)
   {
      const bool ok = okn;
      const int mykey = kn;
      __syncwarp();
      r_stage(sm, wid, lane, vn, an, kn);
      __syncwarp();
      {
         const int j = j0 + RBATCH + lane;
         okn = j < je;
         vn = okn ? __ldg(pv + j) : make_double2(0.0, 0.0);
         an = coef(j, okn);
         kn = okn ? __ldg(key + j) : -1;
      }
      const int cnt = min(RBATCH, je - j0);
      if (__all_sync(0xffffffffu, mykey == cur || !ok)) // the common case: the whole batch lies in the current cell
      {
#pragma unroll
         for (int ks = 0; ks < RBATCH / 4; ks++)
            r_spread_step(sm, wid, ks, g, t, -1, c);
         continue;
      }
      // a cell boundary inside the batch: k-steps that lie in one cell run whole, the others point by point (everything warp-uniform)
      for (int ks = 0; 4 * ks < cnt; ks++)
      {
         const int p0 = 4 * ks, p1 = min(p0 + 4, cnt), k0 = sm.key[wid][p0];
         bool same = true;
         for (int p = p0 + 1; p < p1; p++)
            same = same && sm.key[wid][p] == k0;
         if (same)
         {
            if (k0 != cur)
            {
               r_flush(c, mom, cur, g, t);
               cur = k0;
            }
            r_spread_step(sm, wid, ks, g, t, -1, c);
            continue;
         }
         for (int p = p0; p < p1; p++)
         {
            const int kp = sm.key[wid][p];
            if (kp != cur)
            {
               r_flush(c, mom, cur, g, t);
               cur = kp;
            }
            r_spread_step(sm, wid, ks, g, t, p & 3, c);
         }
      }
   }
   r_flush(c, mom, cur, g, t);
}

// moments of the occupied cells -> oversampled grid.  One CTA per occupied cell: Q = P M P' (10 x 10) in two separable steps, then
// 100 fp64 REDs into the cell's footprint
__global__ void __launch_bounds__(256)
k_poly2d_mom2grid(const int *__restrict__ cells, double *__restrict__ mom, double *__restrict__ grid)
{
   __shared__ double P[TAPS][PLD], M[PLD][PLD + 1], R[TAPS][PLD + 1];
   const int cell = cells[blockIdx.x], tid = threadIdx.x;
   if (tid < TAPS * PLD)
      P[tid >> 4][tid & 15] = (tid & 15) < PNC ? c_p2[tid >> 4][tid & 15] : 0.0;
   M[tid >> 4][tid & 15] = mom[(size_t)cell * (PLD * PLD) + tid];
   mom[(size_t)cell * (PLD * PLD) + tid] = 0.0; // the scratch table is all zero again when the next window's spread starts
   __syncthreads();
   if (tid < TAPS * PLD) // R[t1][k2] = sum_k1 P[t1][k1] M[k1][k2]
   {
      const int t1 = tid >> 4, k2 = tid & 15;
      double acc = 0.0;
#pragma unroll
      for (int k1 = 0; k1 < PNC; k1++)
         acc = fma(P[t1][k1], M[k1][k2], acc);
      R[t1][k2] = acc;
   }
   __syncthreads();
   if (tid < TAPS * TAPS) // Q[t1][t2] = sum_k2 R[t1][k2] P[t2][k2]
   {
      const int t1 = tid / TAPS, t2 = tid - t1 * TAPS;
      double acc = 0.0;
#pragma unroll
      for (int k2 = 0; k2 < PNC; k2++)
         acc = fma(R[t1][k2], P[t2][k2], acc);
      if (acc != 0.0)
         r_red_add(grid + r_tap_idx((cell >> 6) & 63, t1) * NOS + r_tap_idx(cell & 63, t2), acc);
   }
}

// ---------------------------------------------------------------------------------------------------- gather
// Re h (64 x 64) -> bivariate coefficients of the occupied cells: C = P' H P (16 x 16, row / column 15 zero)
__global__ void __launch_bounds__(256)
k_poly2d_grid2coef(const int *__restrict__ cells, const double *__restrict__ grid, double *__restrict__ ctab)
{
   __shared__ double P[TAPS][PLD], H[TAPS][TAPS + 1], R[PLD][TAPS + 1];
   const int cell = cells[blockIdx.x], tid = threadIdx.x;
   if (tid < TAPS * PLD)
      P[tid >> 4][tid & 15] = (tid & 15) < PNC ? c_p2[tid >> 4][tid & 15] : 0.0;
   if (tid < TAPS * TAPS)
   {
      const int t1 = tid / TAPS, t2 = tid - t1 * TAPS;
      H[t1][t2] = grid[r_tap_idx((cell >> 6) & 63, t1) * NOS + r_tap_idx(cell & 63, t2)];
   }
   __syncthreads();
   if (tid < PLD * TAPS) // R[k1][t2] = sum_t1 P[t1][k1] H[t1][t2]
   {
      const int k1 = tid / TAPS, t2 = tid - k1 * TAPS;
      double acc = 0.0;
#pragma unroll
      for (int t1 = 0; t1 < TAPS; t1++)
         acc = fma(P[t1][k1], H[t1][t2], acc);
      R[k1][t2] = acc;
   }
   __syncthreads();
   {
      const int k1 = tid >> 4, k2 = tid & 15; // C[k1][k2] = sum_t2 R[k1][t2] P[t2][k2]
      double acc = 0.0;
#pragma unroll
      for (int t2 = 0; t2 < TAPS; t2++)
         acc = fma(R[k1][t2], P[t2][k2], acc);
      ctab[(size_t)cell * (PLD * PLD) + tid] = acc;
   }
}

// B operands of the current cell: k-step s (k1 = 4 s + t), tile ty (k2 = 8 ty + g)
__device__ __forceinline__ void r_gather_load(const double *__restrict__ ctab, int key, int g, int t, double hb[4][2])
{
   const double *cc = ctab + (size_t)key * (PLD * PLD);
#pragma unroll
   for (int s = 0; s < 4; s++)
#pragma unroll
      for (int ty = 0; ty < 2; ty++)
         hb[s][ty] = __ldg(cc + (4 * s + t) * PLD + 8 * ty + g);
}

// the eight points 8 q .. 8 q + 7 of the batch: returns in the lanes with t == 0 the value of point 8 q + g
__device__ __forceinline__ double r_gather_group(const RSmem &sm, int wid, int q, int g, int t, const double hb[4][2])
{
   const int idx = 8 * q + g;
   const double y = sm.y[0][wid][idx], x4 = sm.x[2][wid][idx];
   double p = (t & 1) ? sm.x[0][wid][idx] : 1.0; // x^t, then x^(4 s + t)
   if (t & 2)
      p *= sm.x[1][wid][idx];
   double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
   for (int s = 0; s < 4; s++)
   {
      r_dmma(c0[0], c0[1], p, hb[s][0]);
      r_dmma(c1[0], c1[1], p, hb[s][1]);
      p *= x4;
   }
   // C tile ty: row = point g, columns k2 = 8 ty + 2 t + {0, 1}:  f = sum_k2 T[k2] y^k2
   double w = (t & 1) ? sm.y[1][wid][idx] : 1.0; // y^(2 t)
   if (t & 2)
      w *= sm.y[2][wid][idx];
   double v = w * fma(sm.y[3][wid][idx], fma(c1[1], y, c1[0]), fma(c0[1], y, c0[0]));
   v += __shfl_xor_sync(0xffffffffu, v, 1);
   v += __shfl_xor_sync(0xffffffffu, v, 2);
   return v;
}

__global__ void __launch_bounds__(RWARPS * 32, RCTAS_PER_SM)
k_gather2_poly(const double2 *__restrict__ pv, const int *__restrict__ key, const double *__restrict__ ctab, const int *__restrict__ perm,
               double *__restrict__ f, int n, int per_warp)
{
   // perm != NULL: the value of the sorted point j goes to f[perm[j]] -- the ORIGINAL point order, so that the combine pass reads the
   // window's results as a stream instead of gathering them through the inverse permutation (the scattered 8-byte stores ride on the
   // tensor work here; the gathers cost a pass of their own there)
   __shared__ RSmem sm;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int g = lane >> 2, t = lane & 3;
   const long long b64 = (long long)(blockIdx.x * RWARPS + wid) * per_warp;
   if (b64 >= n)
      return;
   const int jb = (int)b64, je = min(n, jb + per_warp);
   double hb[4][2];
   int cur = key[jb];
   r_gather_load(ctab, cur, g, t, hb);
   bool okn = jb + lane < je;
   double2 vn = okn ? __ldg(pv + jb + lane) : make_double2(0.0, 0.0);
   int kn = okn ? __ldg(key + jb + lane) : -1;
   int on = okn ? (perm ? __ldg(perm + jb + lane) : jb + lane) : 0;
   for (int j0 = jb; j0 < je; j0 += RBATCH)
   {
      const bool ok = okn;
      const int mykey = kn;
      __syncwarp();
      r_stage(sm, wid, lane, vn, 0.0, kn, on);
      __syncwarp();
      {
         const int j = j0 + RBATCH + lane; // the next step's point is fetched while this step is computed
         okn = j < je;
         vn = okn ? __ldg(pv + j) : make_double2(0.0, 0.0);
         kn = okn ? __ldg(key + j) : -1;
         on = okn ? (perm ? __ldg(perm + j) : j) : 0;
      }
      const int cnt = min(RBATCH, je - j0);
      if (__all_sync(0xffffffffu, mykey == cur || !ok))
      {
#pragma unroll
         for (int q = 0; q < RBATCH / 8; q++)
         {
            const double val = r_gather_group(sm, wid, q, g, t, hb);
            if (t == 0 && 8 * q + g < cnt)
               f[sm.out[wid][8 * q + g]] = val;
         }
         continue;
      }
      // a cell boundary inside the batch: groups of eight that lie in one cell run whole, the others point by point (a point's group
      // is then evaluated with that point's cell and only its value is kept); everything warp-uniform
      for (int q = 0; 8 * q < cnt; q++)
      {
         const int p0 = 8 * q, p1 = min(p0 + 8, cnt), k0 = sm.key[wid][p0];
         bool same = true;
         for (int p = p0 + 1; p < p1; p++)
            same = same && sm.key[wid][p] == k0;
         if (same)
         {
            if (k0 != cur)
            {
               cur = k0;
               r_gather_load(ctab, cur, g, t, hb);
            }
            const double val = r_gather_group(sm, wid, q, g, t, hb);
            if (t == 0 && p0 + g < cnt)
               f[sm.out[wid][p0 + g]] = val;
            continue;
         }
         for (int p = p0; p < p1; p++)
         {
            const int kp = sm.key[wid][p];
            if (kp != cur)
            {
               cur = kp;
               r_gather_load(ctab, cur, g, t, hb);
            }
            const double val = r_gather_group(sm, wid, q, g, t, hb);
            if (t == 0 && g == (p & 7))
               f[sm.out[wid][p]] = val;
         }
      }
   }
}

// ---------------------------------------------------------------------------------------------------- launch layer
static int r_grid(int n, int *per_warp)
{
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   long long warps = (long long)sm * RCTAS_PER_SM * RWARPS;
   long long per = (n + warps - 1) / warps;
   per = ((per + RBATCH - 1) / RBATCH) * RBATCH;
   if (per < RBATCH)
      per = RBATCH;
   *per_warp = (int)per;
   long long used = (n + per - 1) / per;
   return (int)((used + RWARPS - 1) / RWARPS);
}

// per-process scratch: moments and coefficients of all 64 x 64 cells of ONE window (the windows of a matvec follow each other on the
// library stream).  Allocated at plan time (never inside a graph capture).
static double *g_mom2 = nullptr, *g_ctab2 = nullptr;

extern "C" {

int b200k_poly2d_prepare(void)
{
   const size_t bytes = sizeof(double) * (size_t)NOS * NOS * PLD * PLD;
   if (!g_mom2)
   {
      B200_CHECK(cudaMalloc((void **)&g_mom2, bytes));
      B200_CHECK(cudaMemset(g_mom2, 0, bytes)); // k_poly2d_mom2grid clears what it has consumed
   }
   if (!g_ctab2)
   {
      B200_CHECK(cudaMalloc((void **)&g_ctab2, bytes));
      B200_CHECK(cudaMemset(g_ctab2, 0, bytes));
   }
   return 0;
}

// sorted scaled coordinates -> w->pv, w->key
int b200k_poly2d_offsets(const b200_win_dev *w, const double *d_xs)
{
   B200ProfScope prof_scope_(PC_OTHER);
   if (w->n <= 0)
      return 0;
   k_poly2d_offsets<<<b200_cdiv(w->n, 128), 128, 0, b200_stream()>>>(d_xs, reinterpret_cast<double2 *>(w->pv), w->key, w->n);
   B200_LAUNCHED();
   return 0;
}

// w->a (sorted coefficients; d_x == NULL) or s * d_x[perm] (fused prep) -> w->grid (zeroed by the caller)
int b200k_spread2_poly(const b200_win_dev *w, const double *d_x, double s)
{
   if (w->n <= 0 || w->ncell <= 0)
      return 0;
   if (!g_mom2)
      return -1;
   int per, blocks = r_grid(w->n, &per);
   if (d_x)
      k_spread2_poly<true><<<blocks, RWARPS * 32, 0, b200_stream()>>>(reinterpret_cast<const double2 *>(w->pv), w->key, d_x, w->perm, s, g_mom2,
                                                                      w->n, per);
   else
      k_spread2_poly<false><<<blocks, RWARPS * 32, 0, b200_stream()>>>(reinterpret_cast<const double2 *>(w->pv), w->key, w->a, nullptr, 1.0, g_mom2,
                                                                       w->n, per);
   B200_LAUNCHED();
   k_poly2d_mom2grid<<<w->ncell, 256, 0, b200_stream()>>>(w->cells, g_mom2, w->grid);
   B200_LAUNCHED();
   return 0;
}

// w->grid (Re h) -> d_f (original point order if w->f_orig, else sorted order)
int b200k_gather2_poly(const b200_win_dev *w, double *d_f)
{
   if (w->n <= 0 || w->ncell <= 0)
      return 0;
   if (!g_ctab2)
      return -1;
   int per, blocks = r_grid(w->n, &per);
   k_poly2d_grid2coef<<<w->ncell, 256, 0, b200_stream()>>>(w->cells, w->grid, g_ctab2);
   B200_LAUNCHED();
   k_gather2_poly<<<blocks, RWARPS * 32, 0, b200_stream()>>>(reinterpret_cast<const double2 *>(w->pv), w->key, g_ctab2, w->f_orig ? w->perm : nullptr,
                                                             d_f, w->n, per);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
