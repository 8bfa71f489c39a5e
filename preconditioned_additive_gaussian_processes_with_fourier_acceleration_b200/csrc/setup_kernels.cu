// Setup-side kernels of the preconditioners (sm_100a): everything that is O(n) or larger in
//   Nystrom setup   SRC/preconds/nys.c:518-660      landmark GEMMs  U = K1 * inv(L)',  U'U          -> k_dgemm_dmma (fp64 tensor cores)
//   FSAI setup      SRC/preconds/fsai.c:302-673     KNN pattern (kernels.c:121-277) + one small Cholesky solve per row
//   AFN setup       MATLAB afn_setup.m:30-109       farthest point sampling (ordering.c:422-711), L^{-1} K12, FSAI of the Schur complement
// The O(k^3) factorizations of the k x k landmark block stay on the host (dense_small.c).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

// =====================================================================================================================
// fp64 GEMM on the tensor cores: C = alpha * op(A) * op(B) + beta * C, column-major.
// mma.sync.m8n8k4.f64 (DMMA.8x8x4 in SASS) is the only fp64 tensor-core path (tcgen05 has no fp64 kind).
// CTA tile 128 x 64, k-step 16, 8 warps as 4 (m) x 2 (n), warp tile 32 x 32 = 4 x 4 mma tiles.
// Operands are addressed through element strides, so that 'N' / 'T' are the same kernel.
// =====================================================================================================================
#define GBM 128
#define GBN 64
#define GBK 16
#define GPADA 4 /* (GBM + 4) % 16 == 4: the 16 lanes of a half warp hit 16 distinct 8-byte banks */
#define GPADB 4

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A(m, k) = A[m * sAm + k * sAk];  B(k, n) = B[k * sBk + n * sBn].  Split-K over blockIdx.z: slice z accumulates
// k in [z * kchunk, min(K, (z+1) * kchunk)) and writes its partial tile to Cpart + z * M * N (ld M) when nsplit > 1.
__global__ void __launch_bounds__(256)
k_dgemm_dmma(int M, int N, int K, double alpha, const double *__restrict__ A, long long sAm, long long sAk,
             const double *__restrict__ B, long long sBk, long long sBn, double beta, double *__restrict__ C, int ldc,
             int kchunk, int nsplit, double *__restrict__ Cpart)
{
   __shared__ double As[GBK][GBM + GPADA];
   __shared__ double Bs[GBK][GBN + GPADB];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int wm = warp & 3, wn = warp >> 2;
   const int g = lane >> 2, t = lane & 3;
   const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
   const int kbeg = blockIdx.z * kchunk;
   const int kend = min(K, kbeg + kchunk);
   double acc[4][4][2];
#pragma unroll
   for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
         acc[i][j][0] = acc[i][j][1] = 0.0;

   for (int k0 = kbeg; k0 < kend; k0 += GBK)
   {
      // ---- stage the A tile (GBM x GBK): the unit-stride dimension runs fastest over the threads
      if (sAm == 1)
      {
         for (int e = tid; e < GBM * GBK; e += 256)
         {
            int m = e % GBM, kk = e / GBM;
            int gm = m0 + m, gk = k0 + kk;
            As[kk][m] = (gm < M && gk < kend) ? __ldg(A + (long long)gm * sAm + (long long)gk * sAk) : 0.0;
         }
      }
      else
      {
         for (int e = tid; e < GBM * GBK; e += 256)
         {
            int kk = e % GBK, m = e / GBK;
            int gm = m0 + m, gk = k0 + kk;
            As[kk][m] = (gm < M && gk < kend) ? __ldg(A + (long long)gm * sAm + (long long)gk * sAk) : 0.0;
         }
      }
      if (sBn == 1)
      {
         for (int e = tid; e < GBN * GBK; e += 256)
         {
            int nn = e % GBN, kk = e / GBN;
            int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < kend) ? __ldg(B + (long long)gk * sBk + (long long)gn * sBn) : 0.0;
         }
      }
      else
      {
         for (int e = tid; e < GBN * GBK; e += 256)
         {
            int kk = e % GBK, nn = e / GBK;
            int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < kend) ? __ldg(B + (long long)gk * sBk + (long long)gn * sBn) : 0.0;
         }
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < GBK; ks += 4)
      {
         double af[4], bf[4];
#pragma unroll
         for (int i = 0; i < 4; i++)
            af[i] = As[ks + t][wm * 32 + i * 8 + g];
#pragma unroll
         for (int j = 0; j < 4; j++)
            bf[j] = Bs[ks + t][wn * 32 + j * 8 + g];
#pragma unroll
         for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
               dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      __syncthreads();
   }
   // ---- epilogue: c0, c1 of tile (i, j) are (row g, cols 2t, 2t+1)
#pragma unroll
   for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
         for (int c = 0; c < 2; c++)
         {
            int gm = m0 + wm * 32 + i * 8 + g;
            int gn = n0 + wn * 32 + j * 8 + 2 * t + c;
            if (gm < M && gn < N)
            {
               if (nsplit > 1)
                  Cpart[(size_t)blockIdx.z * M * N + (size_t)gn * M + gm] = acc[i][j][c];
               else
               {
                  double *p = C + (size_t)gn * ldc + gm;
                  *p = (beta == 0.0) ? alpha * acc[i][j][c] : alpha * acc[i][j][c] + beta * *p;
               }
            }
         }
}

// deterministic reduction of the split-K partial tiles, in slice order
__global__ void k_splitk_reduce(int M, int N, int nsplit, const double *__restrict__ Cpart, double alpha, double beta,
                                double *__restrict__ C, int ldc)
{
   size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= (size_t)M * N)
      return;
   double s = 0.0;
   for (int z = 0; z < nsplit; z++)
      s += Cpart[(size_t)z * M * N + e];
   int gm = (int)(e % M), gn = (int)(e / M);
   double *p = C + (size_t)gn * ldc + gm;
   *p = (beta == 0.0) ? alpha * s : alpha * s + beta * *p;
}

// =====================================================================================================================
// Farthest point sampling (ordering.c:422-711 "parallel algorithm 1"; MATLAB fps_par.m): first point = nearest to the
// centroid, then repeatedly the point with the largest distance to the selected set.  X feature-major, leading dim ld.
// One fused kernel per selected point: update the running minimum distance with the last selected point and reduce the
// arg max per block; the last block to finish reduces the block results (ties: lowest index).
// =====================================================================================================================
struct FpsState
{
   unsigned int ticket;
   int last; // index selected by the previous step
};

__global__ void __launch_bounds__(256)
k_fps_step(const double *__restrict__ X, int n, int ld, int d, const double *__restrict__ center /* NULL: use X[:, last] */,
           int find_min, double *__restrict__ mind, double *__restrict__ bval, int *__restrict__ bidx, FpsState *st,
           int *__restrict__ perm_out, int step)
{
   __shared__ double sval[256];
   __shared__ int sidx[256];
   __shared__ double sc[64];
   __shared__ bool is_last;
   const int tid = threadIdx.x;
   const int last = st->last;
   for (int t = tid; t < d; t += 256)
      sc[t] = center ? center[t] : X[(size_t)t * ld + last];
   __syncthreads();
   double best = find_min ? INFINITY : -1.0;
   int besti = 0x7fffffff;
   for (int i = blockIdx.x * 256 + tid; i < n; i += gridDim.x * 256)
   {
      double r2 = 0.0;
      for (int t = 0; t < d; t++)
      {
         double df = X[(size_t)t * ld + i] - sc[t];
         r2 = fma(df, df, r2);
      }
      double v;
      if (find_min)
         v = r2;
      else
      {
         v = fmin(mind[i], r2);
         mind[i] = v;
      }
      bool better = find_min ? (v < best) : (v > best);
      if (better)
      {
         best = v;
         besti = i;
      }
   }
   sval[tid] = best;
   sidx[tid] = besti;
   __syncthreads();
   for (int o = 128; o > 0; o >>= 1)
   {
      if (tid < o)
      {
         double v2 = sval[tid + o];
         int i2 = sidx[tid + o];
         bool better = find_min ? (v2 < sval[tid] || (v2 == sval[tid] && i2 < sidx[tid]))
                                : (v2 > sval[tid] || (v2 == sval[tid] && i2 < sidx[tid]));
         if (better)
         {
            sval[tid] = v2;
            sidx[tid] = i2;
         }
      }
      __syncthreads();
   }
   if (tid == 0)
   {
      bval[blockIdx.x] = sval[0];
      bidx[blockIdx.x] = sidx[0];
      __threadfence();
      unsigned int tk = atomicAdd(&st->ticket, 1u);
      is_last = (tk == gridDim.x - 1);
   }
   __syncthreads();
   if (!is_last)
      return;
   __threadfence();
   double v = find_min ? INFINITY : -1.0;
   int vi = 0x7fffffff;
   for (int b = tid; b < (int)gridDim.x; b += 256)
   {
      double v2 = ((volatile double *)bval)[b];
      int i2 = ((volatile int *)bidx)[b];
      bool better = find_min ? (v2 < v || (v2 == v && i2 < vi)) : (v2 > v || (v2 == v && i2 < vi));
      if (better)
      {
         v = v2;
         vi = i2;
      }
   }
   sval[tid] = v;
   sidx[tid] = vi;
   __syncthreads();
   for (int o = 128; o > 0; o >>= 1)
   {
      if (tid < o)
      {
         double v2 = sval[tid + o];
         int i2 = sidx[tid + o];
         bool better = find_min ? (v2 < sval[tid] || (v2 == sval[tid] && i2 < sidx[tid]))
                                : (v2 > sval[tid] || (v2 == sval[tid] && i2 < sidx[tid]));
         if (better)
         {
            sval[tid] = v2;
            sidx[tid] = i2;
         }
      }
      __syncthreads();
   }
   if (tid == 0)
   {
      perm_out[step] = sidx[0];
      st->last = sidx[0];
      st->ticket = 0;
   }
}

__global__ void k_fill_double(size_t n, double v, double *p)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      p[i] = v;
}

// =====================================================================================================================
// KNN pattern of the FSAI (kernels.c:121-277): row i >= lfil keeps its lfil-1 nearest neighbours among the points
// 0..i-1 (Euclidean distance over ALL d features) plus the diagonal, stored LAST; rows i < lfil are dense lower.
// Brute force, as the reference, but tiled: a CTA owns 128 query rows (one per thread, coordinates in registers),
// streams the candidate points through shared memory (point-major, broadcast double2 reads) and keeps each query's
// current lfil-1 best in a max-heap (slot-major in global memory, root = current threshold).
// =====================================================================================================================
#define KNN_QB 128
#define KNN_TILE 128

template <int DR>
__global__ void __launch_bounds__(KNN_QB)
k_knn_bruteforce(const double *__restrict__ X, int n, int ld, int d, int lfil, double *__restrict__ heap_d,
                 int *__restrict__ heap_i, int nblk)
{
   extern __shared__ double smem[];
   const int DP = DR;            // padded dimension (even)
   double *sc = smem;            // [KNN_TILE][DP]
   double *syy = smem + (size_t)KNN_TILE * DP;
   const int tid = threadIdx.x;
   const int blk = nblk - 1 - blockIdx.x; // heaviest blocks first
   const int q0 = blk * KNN_QB;
   const int i = q0 + tid;
   const int K1 = lfil - 1;
   const bool active = (i < n) && (i >= lfil) && K1 > 0;
   double q[DR];
   double xx = 0.0;
#pragma unroll
   for (int t = 0; t < DR; t++)
   {
      q[t] = (t < d && i < n) ? X[(size_t)t * ld + i] : 0.0;
      xx = fma(q[t], q[t], xx);
   }
   double *hd = heap_d + (size_t)blk * KNN_QB * K1 + tid; // slot s at hd[s * KNN_QB]
   int *hi = heap_i + (size_t)blk * KNN_QB * K1 + tid;
   int cnt = 0;
   double thr = INFINITY;
   const int jmax = min(n, q0 + KNN_QB) - 1; // candidates 0 .. jmax-1 serve some query of this block
   for (int j0 = 0; j0 < jmax; j0 += KNN_TILE)
   {
      __syncthreads();
      // stage the candidate tile point-major
      for (int e = tid; e < KNN_TILE * DP; e += KNN_QB)
      {
         int t = e / KNN_TILE, jj = e % KNN_TILE;
         int j = j0 + jj;
         sc[(size_t)jj * DP + t] = (t < d && j < n) ? X[(size_t)t * ld + j] : 0.0;
      }
      __syncthreads();
      {
         int jj = tid;
         double s = 0.0;
         for (int t = 0; t < DP; t++)
            s = fma(sc[(size_t)jj * DP + t], sc[(size_t)jj * DP + t], s);
         syy[jj] = s;
      }
      __syncthreads();
      if (!active)
         continue;
      const int jlim = min(KNN_TILE, i - j0); // only candidates j < i
      for (int jj = 0; jj < jlim; jj++)
      {
         const double2 *cp = reinterpret_cast<const double2 *>(sc + (size_t)jj * DP);
         double dot = 0.0;
#pragma unroll
         for (int t = 0; t < DR / 2; t++)
         {
            double2 c = cp[t];
            dot = fma(q[2 * t], c.x, dot);
            dot = fma(q[2 * t + 1], c.y, dot);
         }
         double dist = (xx + syy[jj]) - 2.0 * dot;
         if (cnt < K1)
         {
            // sift up
            int s = cnt++;
            while (s > 0)
            {
               int p = (s - 1) >> 1;
               double pd = hd[(size_t)p * KNN_QB];
               if (pd >= dist)
                  break;
               hd[(size_t)s * KNN_QB] = pd;
               hi[(size_t)s * KNN_QB] = hi[(size_t)p * KNN_QB];
               s = p;
            }
            hd[(size_t)s * KNN_QB] = dist;
            hi[(size_t)s * KNN_QB] = j0 + jj;
            if (cnt == K1)
               thr = hd[0];
         }
         else if (dist < thr)
         {
            // replace the root and sift down
            int s = 0;
            for (;;)
            {
               int l = 2 * s + 1, r = l + 1, c = l;
               if (l >= K1)
                  break;
               double cd = hd[(size_t)l * KNN_QB];
               if (r < K1)
               {
                  double rd = hd[(size_t)r * KNN_QB];
                  if (rd > cd)
                  {
                     cd = rd;
                     c = r;
                  }
               }
               if (cd <= dist)
                  break;
               hd[(size_t)s * KNN_QB] = cd;
               hi[(size_t)s * KNN_QB] = hi[(size_t)c * KNN_QB];
               s = c;
            }
            hd[(size_t)s * KNN_QB] = dist;
            hi[(size_t)s * KNN_QB] = j0 + jj;
            thr = hd[0];
         }
      }
   }
}

// generic dimension (d > 16): query coordinates in shared memory
__global__ void __launch_bounds__(KNN_QB)
k_knn_bruteforce_generic(const double *__restrict__ X, int n, int ld, int d, int lfil, double *__restrict__ heap_d,
                         int *__restrict__ heap_i, int nblk)
{
   extern __shared__ double smem[];
   const int DP = d;
   double *sc = smem;                              // [KNN_TILE][DP]
   double *syy = sc + (size_t)KNN_TILE * DP;       // [KNN_TILE]
   double *sq = syy + KNN_TILE;                    // [DP][KNN_QB]
   const int tid = threadIdx.x;
   const int blk = nblk - 1 - blockIdx.x;
   const int q0 = blk * KNN_QB;
   const int i = q0 + tid;
   const int K1 = lfil - 1;
   const bool active = (i < n) && (i >= lfil) && K1 > 0;
   double xx = 0.0;
   for (int t = 0; t < d; t++)
   {
      double v = (i < n) ? X[(size_t)t * ld + i] : 0.0;
      sq[(size_t)t * KNN_QB + tid] = v;
      xx = fma(v, v, xx);
   }
   double *hd = heap_d + (size_t)blk * KNN_QB * K1 + tid;
   int *hi = heap_i + (size_t)blk * KNN_QB * K1 + tid;
   int cnt = 0;
   double thr = INFINITY;
   const int jmax = min(n, q0 + KNN_QB) - 1;
   for (int j0 = 0; j0 < jmax; j0 += KNN_TILE)
   {
      __syncthreads();
      for (int e = tid; e < KNN_TILE * DP; e += KNN_QB)
      {
         int t = e / KNN_TILE, jj = e % KNN_TILE;
         int j = j0 + jj;
         sc[(size_t)jj * DP + t] = (j < n) ? X[(size_t)t * ld + j] : 0.0;
      }
      __syncthreads();
      {
         int jj = tid;
         double s = 0.0;
         for (int t = 0; t < DP; t++)
            s = fma(sc[(size_t)jj * DP + t], sc[(size_t)jj * DP + t], s);
         syy[jj] = s;
      }
      __syncthreads();
      if (!active)
         continue;
      const int jlim = min(KNN_TILE, i - j0);
      for (int jj = 0; jj < jlim; jj++)
      {
         double dot = 0.0;
         for (int t = 0; t < d; t++)
            dot = fma(sq[(size_t)t * KNN_QB + tid], sc[(size_t)jj * DP + t], dot);
         double dist = (xx + syy[jj]) - 2.0 * dot;
         if (cnt < K1)
         {
            int s = cnt++;
            while (s > 0)
            {
               int p = (s - 1) >> 1;
               double pd = hd[(size_t)p * KNN_QB];
               if (pd >= dist)
                  break;
               hd[(size_t)s * KNN_QB] = pd;
               hi[(size_t)s * KNN_QB] = hi[(size_t)p * KNN_QB];
               s = p;
            }
            hd[(size_t)s * KNN_QB] = dist;
            hi[(size_t)s * KNN_QB] = j0 + jj;
            if (cnt == K1)
               thr = hd[0];
         }
         else if (dist < thr)
         {
            int s = 0;
            for (;;)
            {
               int l = 2 * s + 1, r = l + 1, c = l;
               if (l >= K1)
                  break;
               double cd = hd[(size_t)l * KNN_QB];
               if (r < K1)
               {
                  double rd = hd[(size_t)r * KNN_QB];
                  if (rd > cd)
                  {
                     cd = rd;
                     c = r;
                  }
               }
               if (cd <= dist)
                  break;
               hd[(size_t)s * KNN_QB] = cd;
               hi[(size_t)s * KNN_QB] = hi[(size_t)c * KNN_QB];
               s = c;
            }
            hd[(size_t)s * KNN_QB] = dist;
            hi[(size_t)s * KNN_QB] = j0 + jj;
            thr = hd[0];
         }
      }
   }
}

// Row pointers of the pattern: row i has min(i + 1, lfil) entries (kernels.c:131-166).
__global__ void k_knn_rowptr(int n, int lfil, int *__restrict__ ia)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i > n)
      return;
   long long v;
   if (n <= lfil || i <= lfil)
      v = (long long)i * (i + 1) / 2;
   else
      v = (long long)lfil * (lfil + 1) / 2 + (long long)(i - lfil) * lfil;
   ia[i] = (int)v;
}

// One warp per row: neighbours sorted by column index (rank sort; indices are distinct), diagonal last.
__global__ void __launch_bounds__(128)
k_knn_emit(int n, int lfil, const int *__restrict__ ia, const int *__restrict__ heap_i, int *__restrict__ ja)
{
   int row = blockIdx.x * 4 + (threadIdx.x >> 5);
   int lane = threadIdx.x & 31;
   if (row >= n)
      return;
   int base = ia[row];
   if (n <= lfil || row < lfil)
   {
      for (int j = lane; j <= row; j += 32)
         ja[base + j] = j;
      return;
   }
   const int K1 = lfil - 1;
   const int blk = row / KNN_QB, tid = row % KNN_QB;
   const int *hi = heap_i + (size_t)blk * KNN_QB * K1 + tid;
   for (int e = lane; e < K1; e += 32)
   {
      int v = hi[(size_t)e * KNN_QB];
      int rank = 0;
      for (int f = 0; f < K1; f++)
         rank += (hi[(size_t)f * KNN_QB] < v);
      ja[base + rank] = v;
   }
   if (lane == 0)
      ja[base + K1] = row;
}

// =====================================================================================================================
// FSAI rows (fsai.c:353-399): for row i with pattern J (diagonal last), A = K(J, J) [+ noise on the diagonal,
// - Z' Z for a Schur complement, Z = (L^{-1} K12)(:, J)], Cholesky, g = A^{-1} e_last / sqrt(e_last' A^{-1} e_last).
// One CTA (128 threads) per row, the k x k block in shared memory (k <= lfil <= 128).
// =====================================================================================================================
struct DenseDesc
{
   const double *Xw; // window-major features [P][dw][n]
   const int *dims;  // [P]
   int n, P, dw;
   int kernel; // 0 Gauss, 1 Matern-1/2
   double ff, l, noise_diag;
};

__device__ __forceinline__ double additive_kernel_value(const double *pa, const double *pb, int P, int dw, const int *sdims, int kernel,
                                                        double l, double ff)
{
   double acc = 0.0;
   for (int w = 0; w < P; w++)
   {
      double r2 = 0.0;
      int dd = sdims[w];
      for (int t = 0; t < dd; t++)
      {
         double df = pa[w * dw + t] - pb[w * dw + t];
         r2 = fma(df, df, r2);
      }
      acc += (kernel == 0) ? exp(-r2 / (2.0 * l * l)) : exp(-sqrt(r2) / l);
   }
   return ff * (acc / (double)P);
}

// d/dl of additive_kernel_value (kernels.c:522, 2220): Gauss exp(-r^2 / 2 l^2) r^2 / l^3, Matern-1/2 exp(-r / l) r / l^2
__device__ __forceinline__ double additive_kernel_dl_value(const double *pa, const double *pb, int P, int dw, const int *sdims, int kernel,
                                                           double l, double ff)
{
   double acc = 0.0;
   for (int w = 0; w < P; w++)
   {
      double r2 = 0.0;
      int dd = sdims[w];
      for (int t = 0; t < dd; t++)
      {
         double df = pa[w * dw + t] - pb[w * dw + t];
         r2 = fma(df, df, r2);
      }
      const double r = sqrt(r2);
      acc += (kernel == 0) ? exp(-r2 / (2.0 * l * l)) * r2 / (l * l * l) : exp(-r / l) * r / (l * l);
   }
   return ff * (acc / (double)P);
}

// v <- (L L')^{-1} v with the Cholesky factor in A (lower, element (r, c) at A[c * lda + r]); all 128 threads of the CTA call it
__device__ __forceinline__ void fs_chol_solve(const double *A, int lda, int k, double *v, int tid)
{
   for (int b = 0; b < k; b++) // L y = v, column oriented
   {
      __syncthreads();
      const double yb = v[b] / A[b * lda + b];
      __syncthreads();
      if (tid == 0)
         v[b] = yb;
      for (int r = b + 1 + tid; r < k; r += 128)
         v[r] -= A[b * lda + r] * yb;
   }
   for (int b = k - 1; b >= 0; b--) // L' x = y
   {
      __syncthreads();
      const double xb = v[b] / A[b * lda + b];
      __syncthreads();
      if (tid == 0)
         v[b] = xb;
      for (int r = tid; r < b; r += 128)
         v[r] -= A[r * lda + b] * xb;
   }
   __syncthreads();
}

#define FS_CH 16 /* rows of Z staged per pass of the Schur Gram */

// GRAD (plain rows only): additionally the derivatives of the row in f, l and mu (fsai.c:495-560): with g the scaled row,
// w = -K^{-1} dK g and dg = w - 0.5 w_last dd g, dd = 1 / sqrt(e' K^{-1} e); dK/df = (2 / f) K gives dg = -g / f in closed form,
// dK/dmu = f^2 I, dK/dl from the points that are still in shared memory.  da = [3][da_stride].
template <bool SCHUR, bool GRAD>
__global__ void __launch_bounds__(128)
k_fsai_rows(DenseDesc D, int nrows, const int *__restrict__ ia, const int *__restrict__ ja, const int *__restrict__ rowmap,
            const double *__restrict__ Z, int kz, double *__restrict__ a, int *__restrict__ breakdown, int kmax, double *__restrict__ da,
            size_t da_stride)
{
   static_assert(!(SCHUR && GRAD), "the derivative rows are built for plain kernels");
   extern __shared__ double sm[];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int row = blockIdx.x;
   if (row >= nrows)
      return;
   const int j1 = ia[row], k = ia[row + 1] - j1;
   const int lda = kmax | 1; // odd leading dimension: column and row walks are both conflict-free
   const int dtot = D.P * D.dw;
   double *A = sm;                               // [kmax][lda]  (element (r, c) at A[c * lda + r], lower used)
   double *pts = A + (size_t)kmax * lda;         // [kmax][dtot]
   double *x = pts + (size_t)kmax * dtot;        // [kmax]
   double *zc = x + kmax;                        // [FS_CH][kpad]  (Schur only)
   __shared__ int sidx[128];
   __shared__ int sdims[64];
   __shared__ int bad;
   if (tid == 0)
      bad = 0;
   for (int w = tid; w < D.P; w += 128)
      sdims[w] = D.dims[w];
   for (int e = tid; e < k; e += 128)
   {
      int j = ja[j1 + e];
      sidx[e] = j;
   }
   __syncthreads();
   // gather the coordinates of the row's points (through rowmap: position in the FSAI ordering -> original point)
   for (int e = tid; e < k * dtot; e += 128)
   {
      int p = e / dtot, c = e % dtot;
      int orig = rowmap ? rowmap[sidx[p]] : sidx[p];
      pts[p * dtot + c] = D.Xw[(size_t)c * D.n + orig];
   }
   __syncthreads();
   // kernel block, lower triangle
   for (int e = tid; e < k * k; e += 128)
   {
      int r = e % k, c = e / k;
      if (r < c)
         continue;
      double v = additive_kernel_value(pts + r * dtot, pts + c * dtot, D.P, D.dw, sdims, D.kernel, D.l, D.ff);
      if (r == c)
         v += D.noise_diag;
      A[c * lda + r] = v;
   }
   __syncthreads();
   if (SCHUR)
   {
      // A -= Z(:, J)' Z(:, J) with fp64 tensor-core tiles: warp w owns the 8 x 8 lower tiles w, w+4, ...
      const int T = (k + 7) >> 3;
      const int kpad = (T << 3) + 4; // (kpad % 16) in {4, 12}
      const int ntile = T * (T + 1) / 2;
      const int g = lane >> 2, t = lane & 3;
      // up to 36 tiles per warp (kmax = 128 -> T = 16 -> 136 tiles / 4 warps = 34); tile index -> (tr, tc), tr >= tc
      double acc[36][2];
      int toff[36];
#pragma unroll
      for (int q = 0; q < 36; q++)
      {
         acc[q][0] = acc[q][1] = 0.0;
         int tl = warp + 4 * q;
         int tr = (int)((sqrtf(8.0f * tl + 1.0f) - 1.0f) * 0.5f);
         while (tr * (tr + 1) / 2 > tl)
            tr--;
         while ((tr + 1) * (tr + 2) / 2 <= tl)
            tr++;
         int tc = tl - tr * (tr + 1) / 2;
         toff[q] = (tl < ntile) ? ((tr * 8) | ((tc * 8) << 16)) : -1;
      }
      for (int c0 = 0; c0 < kz; c0 += FS_CH)
      {
         __syncthreads();
         for (int e = tid; e < FS_CH * (T << 3); e += 128)
         {
            int cc = e % FS_CH, p = e / FS_CH;
            double v = 0.0;
            if (p < k && c0 + cc < kz)
               v = __ldg(Z + (size_t)sidx[p] * kz + c0 + cc);
            zc[cc * kpad + p] = v;
         }
         __syncthreads();
#pragma unroll
         for (int q = 0; q < 36; q++)
         {
            if (toff[q] >= 0)
            {
               const int ro = toff[q] & 0xffff, co = toff[q] >> 16;
#pragma unroll
               for (int ks = 0; ks < FS_CH; ks += 4)
               {
                  double af = zc[(ks + t) * kpad + ro + g];
                  double bf = zc[(ks + t) * kpad + co + g];
                  dmma_8x8x4(acc[q][0], acc[q][1], af, bf);
               }
            }
         }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 36; q++)
      {
         if (toff[q] >= 0)
         {
            const int ro = toff[q] & 0xffff, co = toff[q] >> 16;
#pragma unroll
            for (int c = 0; c < 2; c++)
            {
               int r = ro + g, cc = co + 2 * t + c;
               if (r < k && cc <= r)
                  A[cc * lda + r] -= acc[q][c];
            }
         }
      }
      __syncthreads();
   }
   // ---- right-looking Cholesky in shared memory
   for (int j = 0; j < k; j++)
   {
      double piv = A[j * lda + j];
      if (!(piv > 0.0))
      {
         if (tid == 0)
            bad = 1;
         piv = 1.0; // keep going with finite numbers; the row is flagged
      }
      double dj = sqrt(piv);
      __syncthreads();
      for (int r = j + tid; r < k; r += 128)
         A[j * lda + r] = (r == j) ? dj : A[j * lda + r] / dj;
      __syncthreads();
      // trailing update: A(r, c) -= L(r, j) L(c, j), j < c <= r
      const int m = k - j - 1;
      for (int e = tid; e < m * m; e += 128)
      {
         int r = j + 1 + e % m, c = j + 1 + e / m;
         if (r >= c)
            A[c * lda + r] -= A[j * lda + r] * A[j * lda + c];
      }
      __syncthreads();
   }
   // ---- L y = e_k  =>  y = e_k / L_kk ;  L' x = y  (column-oriented back substitution)
   for (int e = tid; e < k; e += 128)
      x[e] = (e == k - 1) ? 1.0 / A[(k - 1) * lda + (k - 1)] : 0.0;
   __syncthreads();
   for (int b = k - 1; b >= 0; b--)
   {
      double xb = x[b] / A[b * lda + b];
      __syncthreads();
      if (tid == 0)
         x[b] = xb;
      for (int r = tid; r < b; r += 128)
         x[r] -= A[r * lda + b] * xb;
      __syncthreads();
   }
   // scale by 1 / sqrt(e' A^{-1} e) = 1 / sqrt(x_last)   (fsai.c:396)
   double sc = 1.0 / sqrt(x[k - 1]);
   for (int e = tid; e < k; e += 128)
      a[j1 + e] = x[e] * sc;
   if (GRAD)
   {
      double *g = x, *tv = zc; // zc: [kmax] work vector behind x (the Schur staging area is not used by plain rows)
      __syncthreads();
      for (int e = tid; e < k; e += 128)
      {
         g[e] *= sc;
         da[j1 + e] = -g[e] / sqrt(D.ff);            // d/df
         tv[e] = -D.ff * g[e];                       // -dK/dmu g
      }
      __syncthreads();
      fs_chol_solve(A, lda, k, tv, tid);
      {
         const double half = 0.5 * tv[k - 1] * sc;
         for (int e = tid; e < k; e += 128)
            da[2 * da_stride + j1 + e] = tv[e] - half * g[e];
      }
      __syncthreads();
      for (int r = tid; r < k; r += 128)            // -dK/dl g
      {
         double acc = 0.0;
         for (int c = 0; c < k; c++)
            acc = fma(additive_kernel_dl_value(pts + r * dtot, pts + c * dtot, D.P, D.dw, sdims, D.kernel, D.l, D.ff), g[c], acc);
         tv[r] = -acc;
      }
      __syncthreads();
      fs_chol_solve(A, lda, k, tv, tid);
      {
         const double half = 0.5 * tv[k - 1] * sc;
         for (int e = tid; e < k; e += 128)
            da[da_stride + j1 + e] = tv[e] - half * g[e];
      }
   }
   if (tid == 0 && bad)
      atomicAdd(breakdown, 1);
}

// ---- sparse triangular solves with the FSAI factor (fsai.c:675-721), synchronisation-free: a warp takes the next row from a ticket
// counter, waits for the rows it depends on (their flags) and publishes its own.  Warps get their tickets in row order and a warp
// that holds a ticket is resident, so the lowest unfinished row always makes progress.
// lower: rows of G in CSR, diagonal LAST in every row; x_i = (rhs_i - sum_{j<i} G_ij x_j) / G_ii, tickets ascending.
// upper: rows of G' in CSR (columns ascending, so the diagonal comes FIRST); x_i = (rhs_i - sum_{j>i} G'_ij x_j) / G'_ii, descending.
template <bool UPPER>
__global__ void __launch_bounds__(256)
k_sptrsv(int n, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a, const double *__restrict__ rhs,
         volatile double *x, volatile int *ready, int *ticket)
{
   const int lane = threadIdx.x & 31;
   for (;;)
   {
      int tk = 0;
      if (lane == 0)
         tk = atomicAdd(ticket, 1);
      tk = __shfl_sync(0xffffffffu, tk, 0);
      if (tk >= n)
         return;
      const int row = UPPER ? n - 1 - tk : tk;
      const int j1 = ia[row], j2 = ia[row + 1];
      const int jd = UPPER ? j1 : j2 - 1;            // position of the diagonal
      double s = 0.0;
      for (int j = (UPPER ? j1 + 1 : j1) + lane; j < (UPPER ? j2 : j2 - 1); j += 32)
      {
         const int c = ja[j];
         while (ready[c] == 0)
         {
         }
         __threadfence();
         s = fma(a[j], x[c], s);
      }
      s = b200_warp_sum(s);
      if (lane == 0)
      {
         x[row] = (rhs[row] - s) / a[jd];
         __threadfence();
         ready[row] = 1;
      }
   }
}

__global__ void k_fsai_trace(int n, const int *__restrict__ ia, const double *__restrict__ a, const double *__restrict__ da, size_t stride,
                             double *__restrict__ out3)
{
   // 2 sum_i dG_ii / G_ii for the three hyper-parameters (fsai.c:255-263); one CTA, deterministic order
   __shared__ double sh[3][256];
   double acc[3] = {0.0, 0.0, 0.0};
   for (int i = threadIdx.x; i < n; i += 256)
   {
      const int jd = ia[i + 1] - 1;
      const double inv = 1.0 / a[jd];
      for (int t = 0; t < 3; t++)
         acc[t] += da[t * stride + jd] * inv;
   }
   for (int t = 0; t < 3; t++)
      sh[t][threadIdx.x] = acc[t];
   __syncthreads();
   for (int o = 128; o > 0; o >>= 1)
   {
      if (threadIdx.x < o)
         for (int t = 0; t < 3; t++)
            sh[t][threadIdx.x] += sh[t][threadIdx.x + o];
      __syncthreads();
   }
   if (threadIdx.x < 3)
      out3[threadIdx.x] = 2.0 * sh[threadIdx.x][0];
}

// =====================================================================================================================
// CSR transpose (for the scatter-free G' product): count, scan, fill, then order every column by row index.
// =====================================================================================================================
__global__ void k_count_cols(int nnz, const int *__restrict__ ja, int *__restrict__ cnt)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e < nnz)
      atomicAdd(cnt + ja[e] + 1, 1);
}

// single-block exclusive-to-inclusive scan in place over v[0..n] (v[0] = 0 on entry)
__global__ void __launch_bounds__(1024)
k_scan_inplace(int n1, int *__restrict__ v)
{
   __shared__ int ssum[1024];
   __shared__ int carry;
   const int tid = threadIdx.x;
   if (tid == 0)
      carry = 0;
   __syncthreads();
   for (int b0 = 0; b0 < n1; b0 += 1024)
   {
      int i = b0 + tid;
      int x = (i < n1) ? v[i] : 0;
      ssum[tid] = x;
      __syncthreads();
      for (int o = 1; o < 1024; o <<= 1)
      {
         int y = (tid >= o) ? ssum[tid - o] : 0;
         __syncthreads();
         ssum[tid] += y;
         __syncthreads();
      }
      if (i < n1)
         v[i] = ssum[tid] + carry;
      __syncthreads();
      if (tid == 1023)
         carry += ssum[1023];
      __syncthreads();
   }
}

__global__ void k_fill_transpose(int nrows, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a,
                                 const int *__restrict__ iat, int *__restrict__ cursor, int *__restrict__ jtmp, double *__restrict__ atmp)
{
   int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   int lane = threadIdx.x & 31;
   if (row >= nrows)
      return;
   for (int e = ia[row] + lane; e < ia[row + 1]; e += 32)
   {
      int c = ja[e];
      int pos = iat[c] + atomicAdd(cursor + c, 1);
      jtmp[pos] = row;
      atmp[pos] = a[e];
   }
}

// one warp per column: rank sort by row index (distinct within a column) into the final arrays
__global__ void __launch_bounds__(128)
k_sort_segments(int nseg, const int *__restrict__ ip, const int *__restrict__ jin, const double *__restrict__ ain,
                int *__restrict__ jout, double *__restrict__ aout)
{
   int seg = blockIdx.x * 4 + (threadIdx.x >> 5);
   int lane = threadIdx.x & 31;
   if (seg >= nseg)
      return;
   int b = ip[seg], len = ip[seg + 1] - b;
   for (int e = lane; e < len; e += 32)
   {
      int v = jin[b + e];
      int rank = 0;
      for (int f = 0; f < len; f++)
         rank += (jin[b + f] < v);
      jout[b + rank] = v;
      aout[b + rank] = ain[b + e];
   }
}

// column-major k x n2 block: Y(:, j) = Linv * Y(:, j) is done with k_dgemm_dmma; helper to permute feature columns
__global__ void k_gather_rows(int n, int ncol, const int *__restrict__ perm, const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd)
{
   size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= (size_t)n * ncol)
      return;
   int i = (int)(e % n), c = (int)(e / n);
   dst[(size_t)c * ldd + i] = src[(size_t)c * lds + perm[i]];
}

// U(:, c) *= scale[c]
__global__ void k_scale_columns(int n, int k, const double *__restrict__ scale, double *__restrict__ U)
{
   size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= (size_t)n * k)
      return;
   U[e] *= scale[e / n];
}


// =====================================================================================================================
// Blocked Cholesky of the k x k landmark block and the explicit inverse of its factor, on the device (k up to a few
// thousand: the O(k^3) work is GEMM-shaped and goes through k_dgemm_dmma; the 64 x 64 diagonal blocks are factored and
// inverted by one CTA in shared memory).
// =====================================================================================================================
#define CH_NB 64

// A: nb x nb diagonal block (lower, leading dimension lda) -> L11 in place; W (nb x nb, ld nb, upper zero) = inv(L11)
__global__ void __launch_bounds__(256)
k_chol_diag_block(double *__restrict__ A, int lda, int nb, double *__restrict__ W, int ldw, int *__restrict__ info, int j0)
{
   extern __shared__ double chs[];
   double (*S)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(chs);
   double (*X)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(chs + CH_NB * (CH_NB + 1));
   __shared__ int bad;
   const int tid = threadIdx.x;
   if (tid == 0)
      bad = 0;
   for (int e = tid; e < nb * nb; e += 256)
   {
      int r = e % nb, c = e / nb;
      S[c][r] = (r >= c) ? A[(size_t)c * lda + r] : 0.0;
      X[c][r] = 0.0;
   }
   __syncthreads();
   for (int j = 0; j < nb; j++)
   {
      double piv = S[j][j];
      if (!(piv > 0.0))
      {
         if (tid == 0 && !bad)
            bad = j0 + j + 1;
         piv = 1.0;
      }
      double dj = sqrt(piv);
      __syncthreads();
      for (int r = j + tid; r < nb; r += 256)
         S[j][r] = (r == j) ? dj : S[j][r] / dj;
      __syncthreads();
      const int m = nb - j - 1;
      for (int e = tid; e < m * m; e += 256)
      {
         int r = j + 1 + e % m, c = j + 1 + e / m;
         if (r >= c)
            S[c][r] -= S[j][r] * S[j][c];
      }
      __syncthreads();
   }
   // inverse: thread c solves L x = e_c by forward substitution
   if (tid < nb)
   {
      int c = tid;
      X[c][c] = 1.0 / S[c][c];
      for (int i = c + 1; i < nb; i++)
      {
         double s = 0.0;
         for (int j = c; j < i; j++)
            s -= S[j][i] * X[c][j];
         X[c][i] = s / S[i][i];
      }
   }
   __syncthreads();
   for (int e = tid; e < nb * nb; e += 256)
   {
      int r = e % nb, c = e / nb;
      if (r >= c)
         A[(size_t)c * lda + r] = S[c][r];
      W[(size_t)c * ldw + r] = X[c][r];
   }
   if (tid == 0 && bad)
      atomicCAS(info, 0, bad);
}

__global__ void k_add_diag(int n, double *__restrict__ A, int lda, double v)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      A[(size_t)i * lda + i] += v;
}

__global__ void k_copy_block(int m, int n, const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd)
{
   size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= (size_t)m * n)
      return;
   int r = (int)(e % m), c = (int)(e / m);
   dst[(size_t)c * ldd + r] = src[(size_t)c * lds + r];
}

// ---------------------------------------------------------------------------------------------------------------------
static double *g_part = nullptr;
static size_t g_part_doubles = 0;

extern "C" {

int b200k_dgemm(char transa, char transb, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                double beta, double *C, int ldc)
{
   B200ProfScope prof_scope_(PC_OTHER);
   if (M <= 0 || N <= 0)
      return 0;
   long long sAm = (transa == 'N') ? 1 : lda, sAk = (transa == 'N') ? lda : 1;
   long long sBk = (transb == 'N') ? 1 : ldb, sBn = (transb == 'N') ? ldb : 1;
   int gx = b200_cdiv(M, GBM), gy = b200_cdiv(N, GBN);
   int nsplit = 1, kchunk = K;
   long long tiles = (long long)gx * gy;
   if (tiles < 148 && K >= 4096)
   {
      nsplit = (int)((2 * 148 + tiles - 1) / tiles);
      int maxsplit = b200_cdiv(K, 1024);
      if (nsplit > maxsplit)
         nsplit = maxsplit;
      if (nsplit > 256)
         nsplit = 256;
      kchunk = b200_cdiv(b200_cdiv(K, nsplit), GBK) * GBK;
      nsplit = b200_cdiv(K, kchunk);
   }
   double *part = nullptr;
   if (nsplit > 1)
   {
      size_t need = (size_t)nsplit * M * N;
      if (need > g_part_doubles)
      {
         if (g_part)
            cudaFree(g_part);
         B200_CHECK(cudaMalloc(&g_part, need * sizeof(double)));
         g_part_doubles = need;
      }
      part = g_part;
   }
   dim3 grid(gx, gy, nsplit);
   k_dgemm_dmma<<<grid, 256, 0, b200_stream()>>>(M, N, K, alpha, A, sAm, sAk, B, sBk, sBn, beta, C, ldc, kchunk, nsplit, part);
   B200_LAUNCHED();
   if (nsplit > 1)
   {
      k_splitk_reduce<<<b200_cdiv((long long)M * N, 256), 256, 0, b200_stream()>>>(M, N, nsplit, part, alpha, beta, C, ldc);
      B200_LAUNCHED();
   }
   return 0;
}

// Farthest point sampling of k points; d_perm_out[k] (device).  d_center: centroid (device, d doubles).
int b200k_fps(const double *d_X, int n, int ld, int d, const double *d_center, int k, int *d_perm_out)
{
   B200ProfScope prof_scope_(PC_OTHER);
   if (d > 64)
   {
      fprintf(stderr, "nfft4gp_b200: farthest point sampling supports at most 64 features\n");
      return -1;
   }
   int nb = b200_cdiv(n, 256);
   if (nb > 148 * 8)
      nb = 148 * 8;
   double *mind = nullptr, *bval = nullptr;
   int *bidx = nullptr;
   FpsState *st = nullptr;
   B200_CHECK(cudaMalloc(&mind, sizeof(double) * (size_t)n));
   B200_CHECK(cudaMalloc(&bval, sizeof(double) * (size_t)nb));
   B200_CHECK(cudaMalloc(&bidx, sizeof(int) * (size_t)nb));
   B200_CHECK(cudaMalloc(&st, sizeof(FpsState)));
   B200_CHECK(cudaMemsetAsync(st, 0, sizeof(FpsState), b200_stream()));
   k_fill_double<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>((size_t)n, INFINITY, mind);
   B200_LAUNCHED();
   // step 0: nearest to the centroid
   k_fps_step<<<nb, 256, 0, b200_stream()>>>(d_X, n, ld, d, d_center, 1, mind, bval, bidx, st, d_perm_out, 0);
   B200_LAUNCHED();
   for (int s = 1; s < k; s++)
   {
      k_fps_step<<<nb, 256, 0, b200_stream()>>>(d_X, n, ld, d, nullptr, 0, mind, bval, bidx, st, d_perm_out, s);
      B200_LAUNCHED();
   }
   B200_CHECK(cudaStreamSynchronize(b200_stream()));
   cudaFree(mind);
   cudaFree(bval);
   cudaFree(bidx);
   cudaFree(st);
   return 0;
}

// KNN pattern: fills d_ia[n+1] and d_ja[nnz]; nnz = d_ia[n] is returned through *nnz_out.
long long b200k_knn_nnz(int n, int lfil)
{
   if (n <= lfil)
      return (long long)n * (n + 1) / 2;
   return (long long)lfil * (lfil + 1) / 2 + (long long)(n - lfil) * lfil;
}

int b200k_knn_pattern(const double *d_X, int n, int ld, int d, int lfil, int *d_ia, int *d_ja)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_knn_rowptr<<<b200_cdiv(n + 1, 256), 256, 0, b200_stream()>>>(n, lfil, d_ia);
   B200_LAUNCHED();
   double *heap_d = nullptr;
   int *heap_i = nullptr;
   int nblk = b200_cdiv(n, KNN_QB);
   const int K1 = lfil - 1;
   if (n > lfil && K1 > 0)
   {
      size_t slots = (size_t)nblk * KNN_QB * K1;
      B200_CHECK(cudaMalloc(&heap_d, sizeof(double) * slots));
      B200_CHECK(cudaMalloc(&heap_i, sizeof(int) * slots));
      if (d <= 16)
      {
         int DR = d <= 4 ? 4 : d <= 8 ? 8 : d <= 12 ? 12 : 16;
         size_t smem = sizeof(double) * ((size_t)KNN_TILE * DR + KNN_TILE);
         if (DR == 4)
            k_knn_bruteforce<4><<<nblk, KNN_QB, smem, b200_stream()>>>(d_X, n, ld, d, lfil, heap_d, heap_i, nblk);
         else if (DR == 8)
            k_knn_bruteforce<8><<<nblk, KNN_QB, smem, b200_stream()>>>(d_X, n, ld, d, lfil, heap_d, heap_i, nblk);
         else if (DR == 12)
            k_knn_bruteforce<12><<<nblk, KNN_QB, smem, b200_stream()>>>(d_X, n, ld, d, lfil, heap_d, heap_i, nblk);
         else
            k_knn_bruteforce<16><<<nblk, KNN_QB, smem, b200_stream()>>>(d_X, n, ld, d, lfil, heap_d, heap_i, nblk);
      }
      else
      {
         size_t smem = sizeof(double) * ((size_t)KNN_TILE * d + KNN_TILE + (size_t)d * KNN_QB);
         if (smem > 200 * 1024)
         {
            fprintf(stderr, "nfft4gp_b200: KNN pattern supports at most %d features\n", (int)(200 * 1024 / 8 / (KNN_TILE + KNN_QB)));
            return -1;
         }
         B200_CHECK(cudaFuncSetAttribute(k_knn_bruteforce_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
         k_knn_bruteforce_generic<<<nblk, KNN_QB, smem, b200_stream()>>>(d_X, n, ld, d, lfil, heap_d, heap_i, nblk);
      }
      B200_LAUNCHED();
   }
   k_knn_emit<<<b200_cdiv(n, 4), 128, 0, b200_stream()>>>(n, lfil, d_ia, heap_i, d_ja);
   B200_LAUNCHED();
   B200_CHECK(cudaStreamSynchronize(b200_stream()));
   if (heap_d)
      cudaFree(heap_d);
   if (heap_i)
      cudaFree(heap_i);
   return 0;
}

// FSAI values for the rows of a pattern.  rowmap (device, may be NULL): pattern index -> original point (AFN: perm2).
// d_Z (may be NULL): kz x nrows column-major L^{-1} K12 of the Schur complement.  *breakdown_out = rows whose Cholesky hit
// a non-positive pivot.
static int fsai_rows_impl(const double *d_Xw, const int *d_dims, int n_points, int P, int dw, int kernel, double ff, double l,
                          double noise_diag, int nrows, const int *d_ia, const int *d_ja, const int *d_rowmap, const double *d_Z, int kz,
                          int kmax, double *d_a, int *breakdown_out, double *d_da, size_t da_stride);

int b200k_fsai_rows(const double *d_Xw, const int *d_dims, int n_points, int P, int dw, int kernel, double ff, double l,
                    double noise_diag, int nrows, const int *d_ia, const int *d_ja, const int *d_rowmap, const double *d_Z, int kz,
                    int kmax, double *d_a, int *breakdown_out)
{
   return fsai_rows_impl(d_Xw, d_dims, n_points, P, dw, kernel, ff, l, noise_diag, nrows, d_ia, d_ja, d_rowmap, d_Z, kz, kmax, d_a, breakdown_out,
                         nullptr, 0);
}

// the same rows with their derivatives in f, l, mu: d_da = [3][da_stride] (plain kernels, no Schur term)
int b200k_fsai_rows_grad(const double *d_Xw, const int *d_dims, int n_points, int P, int dw, int kernel, double ff, double l,
                         double noise_diag, int nrows, const int *d_ia, const int *d_ja, int kmax, double *d_a, double *d_da, size_t da_stride,
                         int *breakdown_out)
{
   return fsai_rows_impl(d_Xw, d_dims, n_points, P, dw, kernel, ff, l, noise_diag, nrows, d_ia, d_ja, nullptr, nullptr, 0, kmax, d_a, breakdown_out,
                         d_da, da_stride);
}

// x = G^{-1} rhs (upper = 0: rows of G, diagonal last) or x = G^{-T} rhs (upper = 1: rows of G', diagonal first)
int b200k_sptrsv(int upper, int n, const int *d_ia, const int *d_ja, const double *d_a, const double *d_rhs, double *d_x)
{
   B200ProfScope prof_scope_(PC_CSR);
   int *flags = nullptr;
   B200_CHECK(cudaMalloc(&flags, sizeof(int) * ((size_t)n + 1)));
   B200_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t)n + 1), b200_stream()));
   int sm = 148, dev = 0;
   if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   int blocks = b200_cdiv(n, 8) < sm * 4 ? b200_cdiv(n, 8) : sm * 4; // all resident: a waiting warp must never block an unscheduled one
   if (upper)
      k_sptrsv<true><<<blocks, 256, 0, b200_stream()>>>(n, d_ia, d_ja, d_a, d_rhs, d_x, flags, flags + n);
   else
      k_sptrsv<false><<<blocks, 256, 0, b200_stream()>>>(n, d_ia, d_ja, d_a, d_rhs, d_x, flags, flags + n);
   B200_LAUNCHED();
   B200_CHECK(cudaStreamSynchronize(b200_stream()));
   cudaFree(flags);
   return 0;
}

int b200k_fsai_trace(int n, const int *d_ia, const double *d_a, const double *d_da, size_t stride, double *d_out3)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_fsai_trace<<<1, 256, 0, b200_stream()>>>(n, d_ia, d_a, d_da, stride, d_out3);
   B200_LAUNCHED();
   return 0;
}

static int fsai_rows_impl(const double *d_Xw, const int *d_dims, int n_points, int P, int dw, int kernel, double ff, double l,
                          double noise_diag, int nrows, const int *d_ia, const int *d_ja, const int *d_rowmap, const double *d_Z, int kz,
                          int kmax, double *d_a, int *breakdown_out, double *d_da, size_t da_stride)
{
   B200ProfScope prof_scope_(PC_OTHER);
   if (kmax > 128 || P > 64)
   {
      fprintf(stderr, "nfft4gp_b200: FSAI setup supports lfil <= 128 and at most 64 windows (lfil = %d, windows = %d)\n", kmax, P);
      return -1;
   }
   DenseDesc D;
   D.Xw = d_Xw;
   D.dims = d_dims;
   D.n = n_points;
   D.P = P;
   D.dw = dw;
   D.kernel = kernel;
   D.ff = ff;
   D.l = l;
   D.noise_diag = noise_diag;
   int lda = kmax | 1;
   int T = (kmax + 7) >> 3, kpad = (T << 3) + 4;
   size_t smem = sizeof(double) * ((size_t)kmax * lda + (size_t)kmax * P * dw + kmax + (d_Z ? (size_t)FS_CH * kpad : (d_da ? (size_t)kmax : 0)));
   if (smem > 220 * 1024)
   {
      fprintf(stderr, "nfft4gp_b200: FSAI row block does not fit shared memory (%zu bytes)\n", smem);
      return -1;
   }
   int *d_bad = nullptr;
   B200_CHECK(cudaMalloc(&d_bad, sizeof(int)));
   B200_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), b200_stream()));
   if (d_Z)
   {
      B200_CHECK(cudaFuncSetAttribute(k_fsai_rows<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_fsai_rows<true, false><<<nrows, 128, smem, b200_stream()>>>(D, nrows, d_ia, d_ja, d_rowmap, d_Z, kz, d_a, d_bad, kmax, nullptr, 0);
   }
   else if (d_da)
   {
      B200_CHECK(cudaFuncSetAttribute(k_fsai_rows<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_fsai_rows<false, true><<<nrows, 128, smem, b200_stream()>>>(D, nrows, d_ia, d_ja, d_rowmap, d_Z, kz, d_a, d_bad, kmax, d_da, da_stride);
   }
   else
   {
      B200_CHECK(cudaFuncSetAttribute(k_fsai_rows<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_fsai_rows<false, false><<<nrows, 128, smem, b200_stream()>>>(D, nrows, d_ia, d_ja, d_rowmap, d_Z, kz, d_a, d_bad, kmax, nullptr, 0);
   }
   B200_LAUNCHED();
   int bad = 0;
   B200_CHECK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, b200_stream()));
   B200_CHECK(cudaStreamSynchronize(b200_stream()));
   cudaFree(d_bad);
   if (breakdown_out)
      *breakdown_out = bad;
   return 0;
}

// Explicit transpose of a CSR matrix with nrows rows and ncols columns, columns ordered by row index.
int b200k_csr_transpose(int nrows, int ncols, int nnz, const int *d_ia, const int *d_ja, const double *d_a, int *d_iat, int *d_jat, double *d_at)
{
   B200ProfScope prof_scope_(PC_OTHER);
   int *cursor = nullptr, *jtmp = nullptr;
   double *atmp = nullptr;
   B200_CHECK(cudaMemsetAsync(d_iat, 0, sizeof(int) * ((size_t)ncols + 1), b200_stream()));
   if (nnz > 0)
   {
      B200_CHECK(cudaMalloc(&cursor, sizeof(int) * (size_t)ncols));
      B200_CHECK(cudaMalloc(&jtmp, sizeof(int) * (size_t)nnz));
      B200_CHECK(cudaMalloc(&atmp, sizeof(double) * (size_t)nnz));
      B200_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)ncols, b200_stream()));
      k_count_cols<<<b200_cdiv(nnz, 256), 256, 0, b200_stream()>>>(nnz, d_ja, d_iat);
      B200_LAUNCHED();
   }
   k_scan_inplace<<<1, 1024, 0, b200_stream()>>>(ncols + 1, d_iat);
   B200_LAUNCHED();
   if (nnz > 0)
   {
      k_fill_transpose<<<b200_cdiv(nrows, 8), 256, 0, b200_stream()>>>(nrows, d_ia, d_ja, d_a, d_iat, cursor, jtmp, atmp);
      B200_LAUNCHED();
      k_sort_segments<<<b200_cdiv(ncols, 4), 128, 0, b200_stream()>>>(ncols, d_iat, jtmp, atmp, d_jat, d_at);
      B200_LAUNCHED();
      B200_CHECK(cudaStreamSynchronize(b200_stream()));
      cudaFree(cursor);
      cudaFree(jtmp);
      cudaFree(atmp);
   }
   return 0;
}

int b200k_gather_rows(int n, int ncol, const int *d_perm, const double *d_src, int lds, double *d_dst, int ldd)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_gather_rows<<<b200_cdiv((long long)n * ncol, 256), 256, 0, b200_stream()>>>(n, ncol, d_perm, d_src, lds, d_dst, ldd);
   B200_LAUNCHED();
   return 0;
}

int b200k_scale_columns(int n, int k, const double *d_scale, double *d_U)
{
   B200ProfScope prof_scope_(PC_OTHER);
   k_scale_columns<<<b200_cdiv((long long)n * k, 256), 256, 0, b200_stream()>>>(n, k, d_scale, d_U);
   B200_LAUNCHED();
   return 0;
}

int b200k_add_diag(int n, double *d_A, int lda, double v)
{
   k_add_diag<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, d_A, lda, v);
   B200_LAUNCHED();
   return 0;
}

// In: d_A lower triangle of an SPD n x n matrix (ld n).  Out: d_A = its Cholesky factor L (lower), d_Linv = inv(L) (full
// storage, upper zero).  *info = 0 or the 1-based index of the first non-positive pivot (dpotrf convention).
int b200k_potrf_inverse(int n, double *d_A, double *d_Linv, int *info)
{
   B200ProfScope prof_scope_(PC_OTHER);
   cudaStream_t st = b200_stream();
   double *panel = nullptr, *T = nullptr;
   int *d_info = nullptr;
   *info = 0;
   B200_CHECK(cudaMalloc(&panel, sizeof(double) * (size_t)n * CH_NB));
   B200_CHECK(cudaMalloc(&T, sizeof(double) * (size_t)n * CH_NB));
   B200_CHECK(cudaMalloc(&d_info, sizeof(int)));
   B200_CHECK(cudaMemsetAsync(d_info, 0, sizeof(int), st));
   B200_CHECK(cudaMemsetAsync(d_Linv, 0, sizeof(double) * (size_t)n * n, st));
   const size_t chsm = sizeof(double) * 2 * CH_NB * (CH_NB + 1);
   B200_CHECK(cudaFuncSetAttribute(k_chol_diag_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chsm));
   for (int j0 = 0; j0 < n; j0 += CH_NB)
   {
      int nb = n - j0 < CH_NB ? n - j0 : CH_NB;
      int m = n - j0 - nb;
      double *Ajj = d_A + (size_t)j0 * n + j0;
      double *Wjj = d_Linv + (size_t)j0 * n + j0; // inv(L_jj) lands directly in the diagonal block of Linv
      k_chol_diag_block<<<1, 256, chsm, st>>>(Ajj, n, nb, Wjj, n, d_info, j0);
      B200_LAUNCHED();
      if (m > 0)
      {
         double *A21 = Ajj + nb;
         // panel = A21 * inv(L_jj)'  ;  A21 <- panel  ;  A22 -= panel * panel'
         if (b200k_dgemm('N', 'T', m, nb, nb, 1.0, A21, n, Wjj, n, 0.0, panel, m))
            return -1;
         k_copy_block<<<b200_cdiv((long long)m * nb, 256), 256, 0, st>>>(m, nb, panel, m, A21, n);
         B200_LAUNCHED();
         if (b200k_dgemm('N', 'T', m, m, nb, -1.0, panel, m, panel, m, 1.0, Ajj + (size_t)nb * n + nb, n))
            return -1;
      }
   }
   // inv(L) by block rows: X(i, 0:i0) = -inv(L_ii) * ( L(i, 0:i0) * X(0:i0, 0:i0) )
   for (int i0 = CH_NB; i0 < n; i0 += CH_NB)
   {
      int nb = n - i0 < CH_NB ? n - i0 : CH_NB;
      if (b200k_dgemm('N', 'N', nb, i0, i0, 1.0, d_A + i0, n, d_Linv, n, 0.0, T, nb) ||
          b200k_dgemm('N', 'N', nb, i0, nb, -1.0, d_Linv + (size_t)i0 * n + i0, n, T, nb, 0.0, d_Linv + i0, n))
         return -1;
   }
   B200_CHECK(cudaMemcpyAsync(info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
   B200_CHECK(cudaStreamSynchronize(st));
   cudaFree(panel);
   cudaFree(T);
   cudaFree(d_info);
   return 0;
}

} // extern "C"
