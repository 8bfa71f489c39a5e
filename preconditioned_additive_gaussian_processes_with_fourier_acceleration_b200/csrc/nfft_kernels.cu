// Hand-written sm_100a kernels of the NFFT window pipeline (one feature window of the additive kernel):
//
//   prep      a_j = (alpha/P) x[perm_j]                                  (nfft_interface.c:409-420)
//   spread    g[l] += a_j psi_1 psi_2 psi_3      adjoint NFFT, B^T       (NFFT3 fastsum_trafo step 1)
//   fft fwd   g_hat_k = sum_l g[l] e^{+2 pi i k l/64},  k in [-16,16)^d   (pruned, shared-memory radix-8x8)
//   scale     f_hat_k = g_hat_k * b_k / prod phi_hat(k_t)^2               (fused into the first inverse pass)
//   fft inv   h[l] = Re sum_k f_hat_k e^{-2 pi i k l/64}                  (zero-padded, real part only)
//   gather    f_j = sum_taps h[l] psi_1 psi_2 psi_3    NFFT trafo, B
//   combine   y = beta y + f^2 ( sum_w f_w + mu a )                       (nfft_interface.c:428-494, 803-814)
//
// Points are bin-sorted by oversampled-grid cell at setup.  The hot 3-D spread / gather kernels live in
// nfft3d_warp.cu (warp-per-segment, register-tiled); this file holds the generic warp-aggregated kernels used for
// 1-D / 2-D windows (and as the on-device cross-check of the 3-D kernels, variant 0), the shared-memory FFT passes and
// the epilogues.  No cuFFT, no CPU fallback.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdint>
#include <cstdio>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

#define TAPS B200_TAPS
#define NOS B200_NOS
#define NB B200_NBAND

// ---------------------------------------------------------------------------------------------------
// Kaiser-Bessel taps (PRE_PSI analogue).  Same operation order as the oracle's shim_kb_phi.
// ---------------------------------------------------------------------------------------------------
template <int D>
__global__ void k_psi(const double *__restrict__ xs, double *__restrict__ psi, int *__restrict__ key, int n)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= n)
      return;
   const double pi = 3.14159265358979323846264338327950288;
   const double b = pi * (2.0 - 1.0 / 2.0);
   int k = 0;
#pragma unroll
   for (int t = 0; t < D; t++)
   {
      double xj = xs[(size_t)j * D + t];
      int c = (int)floor(xj * (double)NOS);
      int u = c - B200_MCUT;
      k = (k << 6) | ((c + 32) & 63);
      for (int lj = 0; lj < TAPS; lj++)
      {
         double x = xj - (double)(u + lj) / (double)NOS;
         double tt = x * (double)NOS;
         double d2 = 16.0 - tt * tt;
         double v;
         if (d2 > 0.0)
         {
            double s = sqrt(d2);
            v = sinh(b * s) / (pi * s);
         }
         else if (d2 < 0.0)
         {
            double s = sqrt(-d2);
            v = sin(b * s) / (pi * s);
         }
         else
            v = b / pi;
         psi[(size_t)j * (D * TAPS) + t * TAPS + lj] = v;
      }
   }
   key[j] = k;
}

// a_j = s * x[perm_j]
__global__ void k_prep(const double *__restrict__ x, const int *__restrict__ perm, double s, double *__restrict__ a, int n)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j < n)
      a[j] = s * __ldg(x + perm[j]);
}

// ---------------------------------------------------------------------------------------------------
// generic spread / gather (any D in 1..3): one thread per cell-sorted point, warp-level reduction of
// same-cell contributions before the RED.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tap_index(int keyfield, int lj) { return (keyfield + 28 + lj) & 63; }

template <int D>
__global__ void k_spread_generic(const double *__restrict__ psi, const int *__restrict__ key, const double *__restrict__ a,
                                 double *__restrict__ grid, int n)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   bool valid = j < n;
   int jj = valid ? j : n - 1;
   int k = key[jj];
   double av = valid ? a[jj] : 0.0;
   const double *p = psi + (size_t)jj * (D * TAPS);
   const unsigned full = 0xffffffffu;
   int lane = threadIdx.x & 31;
   int k0 = __shfl_sync(full, k, 0);
   bool uniform = __all_sync(full, k == k0);
   if (D == 1)
   {
      int c1 = k & 63;
      double mine = 0.0;
      for (int l = 0; l < TAPS; l++)
      {
         double v = av * p[l];
         if (uniform)
         {
            v = b200_warp_sum(v);
            if (lane == l)
               mine = v;
         }
         else if (valid)
            atomicAdd(grid + tap_index(c1, l), v);
      }
      if (uniform && lane < TAPS)
         atomicAdd(grid + tap_index(c1, lane), mine);
   }
   else if (D == 2)
   {
      int c1 = (k >> 6) & 63, c2 = k & 63;
      for (int l1 = 0; l1 < TAPS; l1++)
      {
         double a1 = av * p[l1];
         int row = tap_index(c1, l1) * NOS;
         double mine = 0.0;
         for (int l2 = 0; l2 < TAPS; l2++)
         {
            double v = a1 * p[TAPS + l2];
            if (uniform)
            {
               v = b200_warp_sum(v);
               if (lane == l2)
                  mine = v;
            }
            else if (valid)
               atomicAdd(grid + row + tap_index(c2, l2), v);
         }
         if (uniform && lane < TAPS)
            atomicAdd(grid + row + tap_index(c2, lane), mine);
      }
   }
   else
   {
      int c1 = (k >> 12) & 63, c2 = (k >> 6) & 63, c3 = k & 63;
      for (int l1 = 0; l1 < TAPS; l1++)
      {
         double a1 = av * p[l1];
         int i1 = tap_index(c1, l1);
         for (int l2 = 0; l2 < TAPS; l2++)
         {
            double a2 = a1 * p[TAPS + l2];
            int row = (i1 * NOS + tap_index(c2, l2)) * NOS;
            double mine = 0.0;
            for (int l3 = 0; l3 < TAPS; l3++)
            {
               double v = a2 * p[2 * TAPS + l3];
               if (uniform)
               {
                  v = b200_warp_sum(v);
                  if (lane == l3)
                     mine = v;
               }
               else if (valid)
                  atomicAdd(grid + row + tap_index(c3, l3), v);
            }
            if (uniform && lane < TAPS)
               atomicAdd(grid + row + tap_index(c3, lane), mine);
         }
      }
   }
}

template <int D>
__global__ void k_gather_generic(const double *__restrict__ psi, const int *__restrict__ key, const double *__restrict__ h,
                                 double *__restrict__ f, int n)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= n)
      return;
   int k = key[j];
   const double *p = psi + (size_t)j * (D * TAPS);
   double acc = 0.0;
   if (D == 1)
   {
      int c1 = k & 63;
      for (int l = 0; l < TAPS; l++)
         acc += __ldg(h + tap_index(c1, l)) * p[l];
   }
   else if (D == 2)
   {
      int c1 = (k >> 6) & 63, c2 = k & 63;
      for (int l1 = 0; l1 < TAPS; l1++)
      {
         int row = tap_index(c1, l1) * NOS;
         double s1 = 0.0;
         for (int l2 = 0; l2 < TAPS; l2++)
            s1 += __ldg(h + row + tap_index(c2, l2)) * p[TAPS + l2];
         acc += s1 * p[l1];
      }
   }
   else
   {
      int c1 = (k >> 12) & 63, c2 = (k >> 6) & 63, c3 = k & 63;
      double p3[TAPS];
      int i3[TAPS];
#pragma unroll
      for (int l3 = 0; l3 < TAPS; l3++)
      {
         p3[l3] = p[2 * TAPS + l3];
         i3[l3] = tap_index(c3, l3);
      }
      for (int l1 = 0; l1 < TAPS; l1++)
      {
         int i1 = tap_index(c1, l1);
         double s1 = 0.0;
         for (int l2 = 0; l2 < TAPS; l2++)
         {
            const double *hr = h + (size_t)(i1 * NOS + tap_index(c2, l2)) * NOS;
            double s2 = 0.0;
#pragma unroll
            for (int l3 = 0; l3 < TAPS; l3++)
               s2 += __ldg(hr + i3[l3]) * p3[l3];
            s1 += s2 * p[TAPS + l2];
         }
         acc += s1 * p[l1];
      }
   }
   f[j] = acc;
}

// ---------------------------------------------------------------------------------------------------
// batched shared-memory FFT, length 64 = 8 x 8, one pencil per 8 threads, pruned to / from the 32-mode band.
// Band storage index kidx = k + 16 for k in [-16, 16); FFT bin of k is k mod 64 = (kidx + 48) & 63.
// ---------------------------------------------------------------------------------------------------
__constant__ double c_tw_re[64];
__constant__ double c_tw_im[64];

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
   return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by S*i
template <int S>
__device__ __forceinline__ double2 cmul_si(double2 a)
{
   return S > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// in-register 8-point DFT: v[r] <- sum_q v[q] e^{S 2 pi i q r / 8}
template <int S>
__device__ __forceinline__ void dft8(double2 v[8])
{
   const double h = 0.70710678118654752440084436210485;
   double2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
   double2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
   double2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
   double2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
   // odd branch twiddles w8^q = e^{S 2 pi i q/8}
   b1 = cmul(b1, make_double2(h, S * h));
   b2 = cmul_si<S>(b2);
   b3 = cmul(b3, make_double2(-h, S * h));
   // DFT4 of a -> even outputs
   {
      double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_si<S>(csub(a1, a3));
      v[0] = cadd(t0, t2);
      v[4] = csub(t0, t2);
      v[2] = cadd(t1, t3);
      v[6] = csub(t1, t3);
   }
   // DFT4 of b -> odd outputs
   {
      double2 t0 = cadd(b0, b2), t1 = csub(b0, b2), t2 = cadd(b1, b3), t3 = cmul_si<S>(csub(b1, b3));
      v[1] = cadd(t0, t2);
      v[5] = csub(t0, t2);
      v[3] = cadd(t1, t3);
      v[7] = csub(t1, t3);
   }
}

template <int S>
__device__ __forceinline__ double2 tw64(int j)
{
   j &= 63;
   return make_double2(c_tw_re[j], S * c_tw_im[j]);
}

// 64-point FFT shared by 8 threads.  In: v[q] = x[t + 8q].  Out: v[u] = X[t + 8u].
// ex points at this pencil's exchange region; element (r, t) lives at ex[r*sr + t*st_].
template <int S>
__device__ __forceinline__ void fft64(double2 v[8], int t, double2 *ex, int sr, int st_)
{
   dft8<S>(v);
#pragma unroll
   for (int r = 1; r < 8; r++)
      v[r] = cmul(v[r], tw64<S>(t * r));
#pragma unroll
   for (int r = 0; r < 8; r++)
      ex[r * sr + t * st_] = v[r];
   __syncthreads();
#pragma unroll
   for (int q = 0; q < 8; q++)
      v[q] = ex[t * sr + q * st_];
   dft8<S>(v);
}

// contiguous-axis forward pass: real rows [rows][64] -> complex band [rows][32]   (sign +)
// blockIdx.y = window of a batch of equally shaped windows stored with strides in_ws / out_ws (elements)
__global__ void __launch_bounds__(256)
k_fft_rows_fwd(const double *__restrict__ in, double2 *__restrict__ out, int rows, size_t in_ws, size_t out_ws)
{
   __shared__ double2 ex[32 * 8 * 9];
   const int p = threadIdx.x >> 3, t = threadIdx.x & 7;
   in += blockIdx.y * in_ws;
   out += blockIdx.y * out_ws;
   const int row = blockIdx.x * 32 + p;
   const bool ok = row < rows;
   double2 v[8];
#pragma unroll
   for (int q = 0; q < 8; q++)
      v[q] = make_double2(ok ? in[(size_t)row * NOS + t + 8 * q] : 0.0, 0.0);
   fft64<1>(v, t, ex + p * 72, 9, 1);
   if (ok)
   {
      double2 *o = out + (size_t)row * NB;
      o[t + 16] = v[0]; // k = t
      o[t + 24] = v[1]; // k = t + 8
      o[t] = v[6];      // k = t + 48 -> t - 16
      o[t + 8] = v[7];  // k = t + 56 -> t - 8
   }
}

// contiguous-axis inverse pass: complex band [rows][32] (optionally * coef) -> real rows [rows][64]   (sign -)
__global__ void __launch_bounds__(256)
k_fft_rows_inv(const double2 *__restrict__ in, const double2 *__restrict__ coef, double *__restrict__ out, int rows, size_t in_ws,
               size_t coef_ws, size_t out_ws)
{
   __shared__ double2 ex[32 * 8 * 9];
   const int p = threadIdx.x >> 3, t = threadIdx.x & 7;
   in += blockIdx.y * in_ws;
   out += blockIdx.y * out_ws;
   if (coef)
      coef += blockIdx.y * coef_ws;
   const int row = blockIdx.x * 32 + p;
   const bool ok = row < rows;
   double2 v[8];
   const double2 z = make_double2(0.0, 0.0);
#pragma unroll
   for (int q = 0; q < 8; q++)
      v[q] = z;
   if (ok)
   {
      const double2 *i_ = in + (size_t)row * NB;
      v[0] = i_[t + 16];
      v[1] = i_[t + 24];
      v[6] = i_[t];
      v[7] = i_[t + 8];
      if (coef)
      {
         const double2 *c = coef + (size_t)row * NB;
         v[0] = cmul(v[0], c[t + 16]);
         v[1] = cmul(v[1], c[t + 24]);
         v[6] = cmul(v[6], c[t]);
         v[7] = cmul(v[7], c[t + 8]);
      }
   }
   fft64<-1>(v, t, ex + p * 72, 9, 1);
   if (ok)
   {
      double *o = out + (size_t)row * NOS;
#pragma unroll
      for (int u = 0; u < 8; u++)
         o[t + 8 * u] = v[u].x;
   }
}

// strided-axis passes on slabs of 32 contiguous columns.  One CTA per slab: 32 columns x 8 threads.
// FWD: 64 rows in -> 32 band rows out (sign +).  INV: 32 band rows in (* coef) -> 64 rows out (sign -).
template <bool FWD>
__global__ void __launch_bounds__(256)
k_fft_cols(const double2 *__restrict__ in, const double2 *__restrict__ coef, double2 *__restrict__ out,
           int in_slab, int in_row, int out_slab, int out_row, size_t in_ws, size_t coef_ws, size_t out_ws)
{
   __shared__ double2 ex[8 * 8 * 32];
   const int c = threadIdx.x & 31, t = threadIdx.x >> 5;
   in += blockIdx.y * in_ws;
   out += blockIdx.y * out_ws;
   if (coef)
      coef += blockIdx.y * coef_ws;
   const double2 *i_ = in + (size_t)blockIdx.x * in_slab + c;
   double2 *o = out + (size_t)blockIdx.x * out_slab + c;
   double2 v[8];
   if (FWD)
   {
#pragma unroll
      for (int q = 0; q < 8; q++)
         v[q] = i_[(size_t)(t + 8 * q) * in_row];
      fft64<1>(v, t, ex + c, 8 * 32, 32);
      o[(size_t)(t + 16) * out_row] = v[0];
      o[(size_t)(t + 24) * out_row] = v[1];
      o[(size_t)t * out_row] = v[6];
      o[(size_t)(t + 8) * out_row] = v[7];
   }
   else
   {
      const double2 z = make_double2(0.0, 0.0);
#pragma unroll
      for (int q = 0; q < 8; q++)
         v[q] = z;
      v[0] = i_[(size_t)(t + 16) * in_row];
      v[1] = i_[(size_t)(t + 24) * in_row];
      v[6] = i_[(size_t)t * in_row];
      v[7] = i_[(size_t)(t + 8) * in_row];
      if (coef)
      {
         const double2 *cc = coef + (size_t)blockIdx.x * in_slab + c;
         v[0] = cmul(v[0], cc[(size_t)(t + 16) * in_row]);
         v[1] = cmul(v[1], cc[(size_t)(t + 24) * in_row]);
         v[6] = cmul(v[6], cc[(size_t)t * in_row]);
         v[7] = cmul(v[7], cc[(size_t)(t + 8) * in_row]);
      }
      fft64<-1>(v, t, ex + c, 8 * 32, 32);
#pragma unroll
      for (int u = 0; u < 8; u++)
         o[(size_t)(t + 8 * u) * out_row] = v[u];
   }
}

// ---------------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------------
#define B200_MAXW_INLINE 8

__global__ void k_combine(int n, int P, const double *const *__restrict__ fptr, const int *const *__restrict__ invptr,
                          const double *__restrict__ x, double s, double c0, double mu, double beta, double *__restrict__ y)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   const double ai = s * x[i];
   double acc = 0.0;
   for (int w = 0; w < P; w++)
   {
      const int *iv = invptr[w]; // NULL: the window's results are already in the original order (table-free 2-D path)
      acc += c0 * (__ldg(fptr[w] + (iv ? iv[i] : i)) + mu * ai);
   }
   y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + acc;
}

__global__ void k_combine_grad(int n, int P, const double *const *__restrict__ fptr, const double *const *__restrict__ f2ptr,
                               const int *const *__restrict__ invptr, const double *__restrict__ x, double s, double f2,
                               double ff, const double *__restrict__ c1w, double mu, double beta, double *__restrict__ y)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   const double ai = s * x[i];
   double a0 = 0.0, a1 = 0.0, a2 = 0.0;
   for (int w = 0; w < P; w++)
   {
      const int *iv = invptr[w];
      int j = iv ? iv[i] : i;
      a0 += f2 * (__ldg(fptr[w] + j) + mu * ai);
      a1 += c1w[w] * __ldg(f2ptr[w] + j);
      a2 += ff * ai;
   }
   double *y0 = y, *y1 = y + n, *y2 = y + 2 * (size_t)n;
   if (beta == 0.0)
   {
      y0[i] = a0;
      y1[i] = a1;
      y2[i] = a2;
   }
   else
   {
      y0[i] = beta * y0[i] + a0;
      y1[i] = beta * y1[i] + a1;
      y2[i] = beta * y2[i] + a2;
   }
}

// ---------------------------------------------------------------------------------------------------
// launch layer
// ---------------------------------------------------------------------------------------------------
static bool g_tw_ready = false;
static int ensure_twiddles()
{
   if (g_tw_ready)
      return 0;
   double re[64], im[64];
   for (int j = 0; j < 64; j++)
   {
      long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)j / 64.0L;
      re[j] = (double)cosl(a);
      im[j] = (double)sinl(a);
   }
   B200_CHECK(cudaMemcpyToSymbol(c_tw_re, re, sizeof(re)));
   B200_CHECK(cudaMemcpyToSymbol(c_tw_im, im, sizeof(im)));
   g_tw_ready = true;
   return 0;
}

extern "C" {

int b200k_psi(const b200_win_dev *w, const double *d_xs)
{
   B200ProfScope prof_scope_(PC_OTHER);
   if (w->n <= 0)
      return 0;
   int grid = b200_cdiv(w->n, 128);
   if (w->d == 1)
      k_psi<1><<<grid, 128, 0, b200_stream()>>>(d_xs, w->psi, w->key, w->n);
   else if (w->d == 2)
      k_psi<2><<<grid, 128, 0, b200_stream()>>>(d_xs, w->psi, w->key, w->n);
   else
      k_psi<3><<<grid, 128, 0, b200_stream()>>>(d_xs, w->psi, w->key, w->n);
   B200_LAUNCHED();
   return 0;
}

__global__ void k_prep_batch(const int *const *__restrict__ perm, double *const *__restrict__ a, const int *__restrict__ nn,
                             const double *__restrict__ x, double s)
{
   const int w = blockIdx.y;
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j < nn[w])
      a[w][j] = s * __ldg(x + perm[w][j]);
}

int b200k_prep_batch(int nb, const int *const *d_perm, double *const *d_a, const int *d_n, int nmax, const double *d_x, double s)
{
   B200ProfScope prof_scope_(PC_PREP);
   dim3 grid(b200_cdiv(nmax, 256), nb);
   k_prep_batch<<<grid, 256, 0, b200_stream()>>>(d_perm, d_a, d_n, d_x, s);
   B200_LAUNCHED();
   return 0;
}

int b200k_prep(const b200_win_dev *w, const double *d_x, double scale)
{
   B200ProfScope prof_scope_(PC_PREP);
   k_prep<<<b200_cdiv(w->n, 256), 256, 0, b200_stream()>>>(d_x, w->perm, scale, w->a, w->n);
   B200_LAUNCHED();
   return 0;
}

// profiling brackets of the fused launches (the batch kernels live in nfft3d_mma.cu)
int b200k_spread_batch_begin(double *d_grids, size_t doubles)
{
   b200_prof_begin(PC_SPREAD);
   B200_CHECK(cudaMemsetAsync(d_grids, 0, sizeof(double) * doubles, b200_stream()));
   return 0;
}
int b200k_spread_batch_end(void)
{
   b200_prof_end();
   return 0;
}
int b200k_gather_batch_begin(void)
{
   b200_prof_begin(PC_GATHER);
   return 0;
}
int b200k_gather_batch_end(void)
{
   b200_prof_end();
   return 0;
}

static size_t grid_elems(int d) { return d == 1 ? 64 : (d == 2 ? 64 * 64 : 64 * 64 * 64); }

// prep + spread in one launch for the paths that fetch their coefficient themselves (table-free 2-D windows); 1 = not such a window
int b200k_spread_fused_prep(const b200_win_dev *w, const double *d_x, double s)
{
   if (!(w->d == 2 && w->pv))
      return 1;
   B200ProfScope prof_scope_(PC_SPREAD);
   B200_CHECK(cudaMemsetAsync(w->grid, 0, sizeof(double) * grid_elems(w->d), b200_stream()));
   return b200k_spread2_poly(w, d_x, s) ? -1 : 0;
}

int b200k_spread(const b200_win_dev *w, int variant)
{
   B200ProfScope prof_scope_(PC_SPREAD);
   B200_CHECK(cudaMemsetAsync(w->grid, 0, sizeof(double) * grid_elems(w->d), b200_stream()));
   if (variant < 0) // automatic: tensor-core kernels for dense cells, DFMA kernels with register-tile shifting for sparse ones
      variant = (w->d == 2 || (w->nseg > 0 && w->n / w->nseg >= B200_MMA_MIN_POINTS_PER_CELL)) ? 5 : 1;
   if (w->d == 3 && variant == 5) // fp64 tensor-core kernels (nfft3d_mma.cu)
      return b200k_spread3_mma(w);
   if (w->d == 3 && variant != 0) // DFMA kernels (nfft3d_warp.cu); variant 0 = generic kernel (cross-check)
      return b200k_spread3_warp(w);
   if (w->d == 2 && w->pv) // 2-D without tap tables (nfft2d_poly.cu): the plan holds offsets only, whatever variant is asked for
      return b200k_spread2_poly(w, nullptr, 1.0);
   if (w->d == 2 && variant != 0 && variant != 1 && !w->xs) // 2-D: fp64 tensor-core kernels (nfft2d_mma.cu); variant 1 = DFMA stream kernels
      return b200k_spread2_mma(w);
   if (w->d < 3 && variant != 0) // 1-D / 2-D: streaming warp-per-range kernels (nfft12_stream.cu)
      return b200k_spread_stream(w);
   int grid = b200_cdiv(w->n, 128);
   if (w->d == 1)
      k_spread_generic<1><<<grid, 128, 0, b200_stream()>>>(w->psi, w->key, w->a, w->grid, w->n);
   else if (w->d == 2)
      k_spread_generic<2><<<grid, 128, 0, b200_stream()>>>(w->psi, w->key, w->a, w->grid, w->n);
   else
      k_spread_generic<3><<<grid, 128, 0, b200_stream()>>>(w->psi, w->key, w->a, w->grid, w->n);
   B200_LAUNCHED();
   return 0;
}

int b200k_gather(const b200_win_dev *w, double *d_f, int variant)
{
   B200ProfScope prof_scope_(PC_GATHER);
   if (variant < 0)
      variant = (w->d == 2 || (w->nseg > 0 && w->n / w->nseg >= B200_MMA_MIN_POINTS_PER_CELL)) ? 5 : 1;
   if (w->d == 3 && variant == 5)
      return b200k_gather3_mma(w, d_f);
   if (w->d == 3 && variant != 0)
      return b200k_gather3_warp(w, d_f);
   if (w->d == 2 && w->pv) // 2-D without tap tables (nfft2d_poly.cu)
      return b200k_gather2_poly(w, d_f);
   if (w->d == 2 && variant != 0 && variant != 1 && !w->xs) // 2-D: fp64 tensor-core kernel (nfft2d_mma.cu); variant 1 = DFMA stream kernel
      return b200k_gather2_mma(w, d_f);
   if (w->d == 2 && variant != 0) // 1-D stays thread-per-point: 10 taps per point need no cross-lane reduction (measured 62 vs 152 us)
      return b200k_gather_stream(w, d_f);
   int grid = b200_cdiv(w->n, 128);
   if (w->d == 1)
      k_gather_generic<1><<<grid, 128, 0, b200_stream()>>>(w->psi, w->key, w->grid, d_f, w->n);
   else if (w->d == 2)
      k_gather_generic<2><<<grid, 128, 0, b200_stream()>>>(w->psi, w->key, w->grid, d_f, w->n);
   else
      k_gather_generic<3><<<grid, 128, 0, b200_stream()>>>(w->psi, w->key, w->grid, d_f, w->n);
   B200_LAUNCHED();
   return 0;
}

// strides (in elements of the respective type) between consecutive windows of a batch; see nfft_host.c (the additive
// handle stores the work arrays of all windows contiguously with the slot size of the common window dimension)
static size_t ws_grid(int d) { return grid_elems(d); }                     // double
static size_t ws_spec(int d) { return grid_elems(d) / 2; }                 // double2 (specA / specB slots hold 64^d doubles)
static size_t ws_band(int d) { return d == 1 ? 32 : (d == 2 ? 32 * 32 : 32 * 32 * 32); } // double2

// grid (real 64^d) -> band (complex 32^d), unscaled, e^{+}; nb equally shaped windows starting at w
int b200k_forward_fft_batch(const b200_win_dev *w, int nb)
{
   B200ProfScope prof_scope_(PC_FFT_FWD);
   if (ensure_twiddles())
      return -1;
   cudaStream_t s = b200_stream();
   const int d = w->d;
   double2 *A = (double2 *)w->specA, *B = (double2 *)w->specB, *band = (double2 *)w->band;
   const size_t gs = ws_grid(d), ss = ws_spec(d), bs = ws_band(d);
   if (d == 1)
   {
      k_fft_rows_fwd<<<dim3(1, nb), 256, 0, s>>>(w->grid, band, 1, gs, bs);
      B200_LAUNCHED();
   }
   else if (d == 2)
   {
      k_fft_rows_fwd<<<dim3(2, nb), 256, 0, s>>>(w->grid, A, 64, gs, ss); // [64 l1][32 k2]
      B200_LAUNCHED();
      k_fft_cols<true><<<dim3(1, nb), 256, 0, s>>>(A, nullptr, band, 0, 32, 0, 32, ss, 0, bs); // along l1 -> [32 k1][32 k2]
      B200_LAUNCHED();
   }
   else
   {
      k_fft_rows_fwd<<<dim3(128, nb), 256, 0, s>>>(w->grid, A, 4096, gs, ss); // [64 l1][64 l2][32 k3]
      B200_LAUNCHED();
      k_fft_cols<true><<<dim3(64, nb), 256, 0, s>>>(A, nullptr, B, 64 * 32, 32, 32 * 32, 32, ss, 0, ss); // along l2 -> [64 l1][32 k2][32 k3]
      B200_LAUNCHED();
      k_fft_cols<true><<<dim3(32, nb), 256, 0, s>>>(B, nullptr, band, 32, 32 * 32, 32, 32 * 32, ss, 0, bs); // along l1 -> [32 k1][32 k2][32 k3]
      B200_LAUNCHED();
   }
   return 0;
}

// band * coef -> grid (real part of the zero-padded e^{-} transform); which = 0 kernel, 1 derivative-kernel coefficients
int b200k_backward_fft_batch(const b200_win_dev *w, int nb, int which)
{
   B200ProfScope prof_scope_(PC_FFT_INV);
   if (ensure_twiddles())
      return -1;
   cudaStream_t s = b200_stream();
   const int d = w->d;
   double2 *A = (double2 *)w->specA, *B = (double2 *)w->specB, *band = (double2 *)w->band;
   const double2 *coef = (const double2 *)(which ? w->dcoef : w->coef);
   const size_t gs = ws_grid(d), ss = ws_spec(d), bs = ws_band(d);
   if (d == 1)
   {
      k_fft_rows_inv<<<dim3(1, nb), 256, 0, s>>>(band, coef, w->grid, 1, bs, bs, gs);
      B200_LAUNCHED();
   }
   else if (d == 2)
   {
      k_fft_cols<false><<<dim3(1, nb), 256, 0, s>>>(band, coef, A, 0, 32, 0, 32, bs, bs, ss); // along k1 -> [64 l1][32 k2]
      B200_LAUNCHED();
      k_fft_rows_inv<<<dim3(2, nb), 256, 0, s>>>(A, nullptr, w->grid, 64, ss, 0, gs);
      B200_LAUNCHED();
   }
   else
   {
      k_fft_cols<false><<<dim3(32, nb), 256, 0, s>>>(band, coef, B, 32, 32 * 32, 32, 32 * 32, bs, bs, ss); // along k1 -> [64 l1][32 k2][32 k3]
      B200_LAUNCHED();
      k_fft_cols<false><<<dim3(64, nb), 256, 0, s>>>(B, nullptr, A, 32 * 32, 32, 64 * 32, 32, ss, 0, ss); // along k2 -> [64 l1][64 l2][32 k3]
      B200_LAUNCHED();
      k_fft_rows_inv<<<dim3(128, nb), 256, 0, s>>>(A, nullptr, w->grid, 4096, ss, 0, gs);
      B200_LAUNCHED();
   }
   return 0;
}

int b200k_forward_fft(const b200_win_dev *w) { return b200k_forward_fft_batch(w, 1); }

int b200k_backward_fft(const b200_win_dev *w, const double *d_coef)
{
   return b200k_backward_fft_batch(w, 1, d_coef == w->dcoef ? 1 : 0);
}

int b200k_combine(int n, int P, const double *const *d_fptr, const int *const *d_invptr, const double *d_x, double s,
                  double c0, double mu, double beta, double *d_y)
{
   B200ProfScope prof_scope_(PC_COMBINE);
   k_combine<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, P, d_fptr, d_invptr, d_x, s, c0, mu, beta, d_y);
   B200_LAUNCHED();
   return 0;
}

int b200k_combine_grad(int n, int P, const double *const *d_fptr, const double *const *d_f2ptr, const int *const *d_invptr,
                       const double *d_x, double s, double f2, double ff, const double *d_c1w, double mu, double beta, double *d_y)
{
   B200ProfScope prof_scope_(PC_COMBINE);
   k_combine_grad<<<b200_cdiv(n, 256), 256, 0, b200_stream()>>>(n, P, d_fptr, d_f2ptr, d_invptr, d_x, s, f2, ff, d_c1w, mu, beta, d_y);
   B200_LAUNCHED();
   return 0;
}

} // extern "C"
