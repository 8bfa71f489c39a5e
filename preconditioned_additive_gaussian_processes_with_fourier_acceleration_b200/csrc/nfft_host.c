/* Host side (C) of the NFFT interface of libnfft4gp_b200: handle life cycle, per-window point scaling and
 * cell sort, kernel Fourier coefficients, and the matvec drivers that sequence the CUDA kernels.
 * Mirrors the operator interface of the reference's SRC/external/nfft_interface.c (same names, argument
 * meaning, return codes) -- each function cites the lines it replaces.  The arithmetic of the hot path
 * (NFFT3 fastsum in the reference) runs entirely on the GPU; there is no CPU fallback. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

int g_variant_spread3 = -1, g_variant_gather3 = -1; /* -1 = automatic (b200k_spread / b200k_gather) */

/* ------------------------------------------------------------------ small host utilities */
void *b200_xmalloc(size_t bytes)
{
   void *p = malloc(bytes ? bytes : 1);
   if (!p)
   {
      printf("Error: nfft4gp_b200 host allocation of %zu bytes failed.\n", bytes);
      exit(EXIT_FAILURE); /* SRC/utils/memory.h:85-95 */
   }
   return p;
}

void *b200_xcalloc(size_t n, size_t sz)
{
   void *p = calloc(n ? n : 1, sz ? sz : 1);
   if (!p)
   {
      printf("Error: nfft4gp_b200 host allocation failed.\n");
      exit(EXIT_FAILURE);
   }
   return p;
}

static size_t ipow(int b, int e)
{
   size_t r = 1;
   while (e-- > 0)
      r *= (size_t)b;
   return r;
}

/* modified Bessel function I0, power series (all terms positive) */
static double bessel_i0(double x)
{
   double q = 0.25 * x * x, term = 1.0, sum = 1.0;
   int k;
   for (k = 1; k < 1000; k++)
   {
      term *= q / ((double)k * (double)k);
      sum += term;
      if (term < 1e-20 * sum)
         break;
   }
   return sum;
}

/* phi_hat(k) of the Kaiser-Bessel window, sigma = 2, b = pi (2 - 1/sigma) */
static double kb_phi_hat(int k)
{
   double b = M_PI * (2.0 - 0.5);
   double w = 2.0 * M_PI * (double)k / (double)B200_NOS;
   return bessel_i0((double)B200_MCUT * sqrt(b * b - w * w));
}

/* in-place Stockham radix-2 DFT of length n (power of two) with stride, sign -1: e^{-2 pi i jk/n} */
static void host_fft_strided(double *re, double *im, int n, size_t stride, int sign, double *wr, double *wi, double *tr, double *ti)
{
   int half = n / 2, l, m, j, k;
   double *ar = wr + n, *ai = wi + n; /* scratch copies live after the twiddles */
   for (j = 0; j < n; j++)
   {
      ar[j] = re[(size_t)j * stride];
      ai[j] = im[(size_t)j * stride];
   }
   /* Stockham autosort: l = number of sub-transforms, m = their current length */
   for (l = half, m = 1; l >= 1; l >>= 1, m <<= 1)
   {
      for (j = 0; j < l; j++)
      {
         double c = wr[j * m], s = sign * wi[j * m];
         for (k = 0; k < m; k++)
         {
            double x0r = ar[k + j * m], x0i = ai[k + j * m];
            double x1r = ar[k + j * m + l * m], x1i = ai[k + j * m + l * m];
            double dr = x0r - x1r, di = x0i - x1i;
            tr[k + 2 * j * m] = x0r + x1r;
            ti[k + 2 * j * m] = x0i + x1i;
            tr[k + 2 * j * m + m] = dr * c - di * s;
            ti[k + 2 * j * m + m] = dr * s + di * c;
         }
      }
      memcpy(ar, tr, sizeof(double) * (size_t)n);
      memcpy(ai, ti, sizeof(double) * (size_t)n);
   }
   for (j = 0; j < n; j++)
   {
      re[(size_t)j * stride] = ar[j];
      im[(size_t)j * stride] = ai[j];
   }
}

/* d-dimensional unnormalised DFT of an n^d row-major array (split re / im), sign as above */
static void host_fft_nd(double *re, double *im, int d, int n, int sign)
{
   size_t total = ipow(n, d), npencil = total / (size_t)n, p;
   double *wr = (double *)b200_xmalloc(sizeof(double) * 2 * (size_t)n);
   double *wi = (double *)b200_xmalloc(sizeof(double) * 2 * (size_t)n);
   double *tr = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   double *ti = (double *)b200_xmalloc(sizeof(double) * (size_t)n);
   int t, j;
   for (j = 0; j < n; j++)
   {
      long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
      wr[j] = (double)cosl(a);
      wi[j] = (double)sinl(a);
   }
   for (t = 0; t < d; t++)
   {
      size_t stride = ipow(n, d - 1 - t);
      for (p = 0; p < npencil; p++)
      {
         size_t hi = p / stride, lo = p % stride;
         size_t base = hi * stride * (size_t)n + lo;
         host_fft_strided(re + base, im + base, n, stride, sign, wr, wi, tr, ti);
      }
   }
   free(wr);
   free(wi);
   free(tr);
   free(ti);
}

/* swap halves along every dimension (n even) */
static void host_fftshift_nd(double *re, double *im, int d, int n)
{
   size_t total = ipow(n, d), idx;
   double *r2 = (double *)b200_xmalloc(sizeof(double) * total);
   double *i2 = (double *)b200_xmalloc(sizeof(double) * total);
   int t;
   for (idx = 0; idx < total; idx++)
   {
      size_t rem = idx, dst = 0, mul = 1;
      for (t = d - 1; t >= 0; t--)
      {
         int c = (int)(rem % (size_t)n);
         rem /= (size_t)n;
         dst += (size_t)((c + n / 2) % n) * mul;
         mul *= (size_t)n;
      }
      r2[dst] = re[idx];
      i2[dst] = im[idx];
   }
   memcpy(re, r2, sizeof(double) * total);
   memcpy(im, i2, sizeof(double) * total);
   free(r2);
   free(i2);
}

/* radial kernels of the reference's four fastsum plans (nfft_interface.c:227,230,363,366); which = 1: derivative kernel */
static double radial_kernel(int kernel, int which, double r, double c)
{
   if (kernel == 0)
   {
      double q = r * r / (c * c);
      return which ? q * exp(-q) : exp(-q);
   }
   else
   {
      double q = fabs(r) / c;
      return which ? q * exp(-q) : exp(-q);
   }
}

/* coef_k = b_k / prod_t phi_hat(k_t)^2 with b = fftshift(FFT_forward(fftshift(K(min(|j/N - 1/2|, 1/2)) / N^d)))
 * (NFFT3 fastsum_precompute_kernel with eps_I = eps_B = 0).  Interleaved complex out. */
static void build_coef(int d, int kernel, int which, double c, double *out)
{
   int N = B200_NBAND, t;
   size_t total = ipow(N, d), j;
   double *re = (double *)b200_xmalloc(sizeof(double) * total);
   double *im = (double *)b200_xcalloc(total, sizeof(double));
   double ph[B200_NBAND];
   for (j = 0; j < total; j++)
   {
      double r;
      if (d == 1)
      {
         r = -((double)j / (double)N - 0.5);
         if (r < -0.5)
            r = -0.5;
         if (r > 0.5)
            r = 0.5;
      }
      else
      {
         size_t kk = j;
         double r2 = 0.0;
         for (t = 0; t < d; t++)
         {
            double cc = (double)(kk % (size_t)N) / (double)N - 0.5;
            r2 += cc * cc;
            kk /= (size_t)N;
         }
         r = sqrt(r2);
         if (r >= 0.5)
            r = 0.5;
      }
      re[j] = radial_kernel(kernel, which, r, c) / (double)total;
   }
   host_fftshift_nd(re, im, d, N);
   host_fft_nd(re, im, d, N, -1);
   host_fftshift_nd(re, im, d, N);
   for (t = 0; t < N; t++)
      ph[t] = 1.0 / kb_phi_hat(t - N / 2);
   for (j = 0; j < total; j++)
   {
      size_t rem = j;
      double s = 1.0;
      for (t = 0; t < d; t++)
      {
         s *= ph[rem % (size_t)N];
         rem /= (size_t)N;
      }
      s *= s;
      out[2 * j] = re[j] * s;
      out[2 * j + 1] = im[j] * s;
   }
   free(re);
   free(im);
}

/* ------------------------------------------------------------------ window (str_adj analogue) */
static b200_window *window_create(int max_n, int dim)
{
   b200_window *w = (b200_window *)b200_xcalloc(1, sizeof(b200_window));
   w->magic = B200_MAGIC_WINDOW;
   w->d = dim;
   w->N = B200_NBAND; /* nfft_interface.c:18-21 */
   w->p = 1;
   w->m = B200_MCUT;
   w->eps = 0.0;
   w->NN = B200_NOS; /* nfft_interface.c:25-27 */
   w->scale = -1.0;
   w->kernel_scale = 1.0;
   w->max_n = max_n;
   w->n = 0;
   return w;
}

static void window_free_device(b200_window *w)
{
   b200_win_dev *v = &w->dev;
   b200_free(v->psi);
   b200_free(v->pv);
   b200_free(v->cells);
   b200_free(v->xs);
   b200_free(v->key);
   b200_free(v->perm);
   b200_free(v->inv);
   b200_free(v->chunks);
   b200_free(v->cfirst);
   b200_free(v->a);
   b200_free(v->f);
   b200_free(v->f2);
   if (v->grid != w->ext_grid)
      b200_free(v->grid);
   if (v->specA != w->ext_specA)
      b200_free(v->specA);
   if (v->specB != w->ext_specB)
      b200_free(v->specB);
   if (v->band != w->ext_band)
      b200_free(v->band);
   if (v->coef != w->ext_coef)
      b200_free(v->coef);
   if (v->dcoef != w->ext_dcoef)
      b200_free(v->dcoef);
   memset(v, 0, sizeof(*v));
}

static void window_free(b200_window *w)
{
   if (!w)
      return;
   window_free_device(w);
   free(w->h_perm);
   free(w->h_chunks);
   free(w);
}

/* First-call part of nfft_interface.c:150-213: centre every feature, r_max, scale to radius 0.25 (unless
 * r_max is already in [0.125, 0.25]), point-major transpose -- same sequential fp64 arithmetic as the reference
 * so that `scale` is bit-identical -- followed by this library's own bin sort by oversampled-grid cell.
 * `cols` holds the window's d feature columns, n contiguous values each (feature-major, ldim = n). */
/* centre, r_max and scale of a window in the reference's sequential fp64 arithmetic (nfft_interface.c:162-196): returns the centred,
 * scaled feature columns (feature-major, n_all values each; caller frees) and sets w->scale */
static double *window_scale_points(b200_window *w, const double *cols, int n_all)
{
   int d = w->d, i, j;
   double radius = 0.0;
   double *tmp = (double *)b200_xmalloc(sizeof(double) * (size_t)n_all * d);
   memcpy(tmp, cols, sizeof(double) * (size_t)n_all * d);
   for (i = 0; i < d; i++)
   {
      double c = 0.0;
      for (j = 0; j < n_all; j++)
         c += tmp[(size_t)i * n_all + j];
      c /= (double)n_all;
      for (j = 0; j < n_all; j++)
         tmp[(size_t)i * n_all + j] -= c;
   }
   for (j = 0; j < n_all; j++)
   {
      double r = 0.0;
      for (i = 0; i < d; i++)
         r += tmp[(size_t)i * n_all + j] * tmp[(size_t)i * n_all + j];
      r = sqrt(r);
      if (r > radius)
         radius = r;
   }
   if (radius > 0.25 || radius < 0.125)
   {
      w->scale = 0.25 / radius;
      for (j = 0; j < n_all * d; j++)
         tmp[j] *= w->scale;
   }
   else
      w->scale = 1.0;
   return tmp;
}

/* 2-D windows take the table-free path of nfft2d_poly.cu from POLY2D_MIN_N points (about 60 per occupied cell: below that the
 * per-cell work of that path -- 256 REDs / 8 coefficient loads per thread at every cell change -- outweighs the table traffic);
 * NFFT4GP_B200_POLY2D=0 never, =1 always */
#define POLY2D_MIN_N 65536
static int window_wants_poly2d(int d, int n)
{
   const char *e = getenv("NFFT4GP_B200_POLY2D"), *ev = getenv("NFFT4GP_B200_WINDOW_EVAL");
   if (d != 2 || (e && atoi(e) == 0) || (ev && atoi(ev) > 0))
      return 0;
   return (e && atoi(e) > 0) || n >= POLY2D_MIN_N;
}

static int window_setup_points(b200_window *w, const double *cols, int n_all, int lo, int hi)
{
   /* n_all points define centre and scale (identical on every rank); this rank tabulates the block [lo, hi) only
    * (the whole range unless the handle is sharded by point block) */
   const int n = hi - lo;
   int d = w->d, i, j, t;
   size_t nbins = ipow(B200_NOS, d);
   double *tmp, *xs;
   int *keys, *binstart, *perm, *inv;
   b200_win_dev *v = &w->dev;

   if (d < 1 || d > 3)
   {
      printf("Error: nfft4gp_b200 supports window dimensions 1..3, got %d.\n", d);
      return -1;
   }
   if (b200_init(-1))
      return -1;
   tmp = window_scale_points(w, cols, n_all);

   /* cell keys, stable counting sort */
   keys = (int *)b200_xmalloc(sizeof(int) * (size_t)n);
   binstart = (int *)b200_xcalloc(nbins + 1, sizeof(int));
   for (j = 0; j < n; j++)
   {
      int k = 0;
      for (t = 0; t < d; t++)
      {
         int c = (int)floor(tmp[(size_t)t * n_all + lo + j] * (double)B200_NOS);
         k = (k << 6) | ((c + 32) & 63);
      }
      keys[j] = k;
      binstart[k + 1]++;
   }
   for (i = 0; i < (int)nbins; i++)
      binstart[i + 1] += binstart[i];
   perm = (int *)b200_xmalloc(sizeof(int) * (size_t)n);
   inv = (int *)b200_xmalloc(sizeof(int) * (size_t)n);
   {
      int *cursor = (int *)b200_xmalloc(sizeof(int) * nbins);
      memcpy(cursor, binstart, sizeof(int) * nbins);
      for (j = 0; j < n; j++)
      {
         int pos = cursor[keys[j]]++;
         perm[pos] = j;
         inv[j] = pos;
      }
      free(cursor);
   }
   /* sorted, point-major scaled coordinates */
   xs = (double *)b200_xmalloc(sizeof(double) * (size_t)n * d);
   for (j = 0; j < n; j++)
      for (t = 0; t < d; t++)
         xs[(size_t)j * d + t] = tmp[(size_t)t * n_all + lo + perm[j]];
   free(tmp);

   /* device state */
   window_free_device(w);
   v->d = d;
   v->n = n;
   {
      size_t pad = 16, gsz = ipow(B200_NOS, d), bsz = ipow(B200_NBAND, d);
      double *d_xs = (double *)b200_malloc(sizeof(double) * (size_t)n * d);
      const int poly2d = window_wants_poly2d(d, n);
      if (poly2d)
      {
         /* table-free 2-D path: 16 B of offsets per point instead of 160 B of taps, and the list of occupied cells */
         int *hcells = (int *)b200_xmalloc(sizeof(int) * nbins), nc = 0;
         for (i = 0; i < (int)nbins; i++)
            if (binstart[i + 1] > binstart[i])
               hcells[nc++] = i;
         v->pv = (double *)b200_malloc(sizeof(double) * ((size_t)n + pad) * 2);
         v->cells = (int *)b200_malloc(sizeof(int) * (size_t)(nc ? nc : 1));
         v->ncell = nc;
         v->f_orig = 1;
         if (!v->pv || !v->cells || b200k_poly2d_prepare() || b200_h2d(v->cells, hcells, sizeof(int) * (size_t)nc))
         {
            free(hcells);
            return -1;
         }
         free(hcells);
      }
      else
      {
         v->psi = (double *)b200_malloc(sizeof(double) * ((size_t)n + pad) * d * B200_TAPS);
         b200_memset(v->psi, 0, sizeof(double) * ((size_t)n + pad) * d * B200_TAPS);
      }
      v->key = (int *)b200_malloc(sizeof(int) * (size_t)n);
      v->perm = (int *)b200_malloc(sizeof(int) * (size_t)n);
      v->inv = (int *)b200_malloc(sizeof(int) * (size_t)n);
      v->a = (double *)b200_malloc(sizeof(double) * ((size_t)n + pad));
      b200_memset(v->a, 0, sizeof(double) * ((size_t)n + pad));
      v->f = (double *)b200_malloc(sizeof(double) * (size_t)n);
      v->f2 = NULL; /* allocated on the first gradient matvec */
      v->grid = w->ext_grid ? w->ext_grid : (double *)b200_malloc(sizeof(double) * gsz);
      v->specA = w->ext_specA ? w->ext_specA : (double *)b200_malloc(sizeof(double) * gsz); /* complex [64^(d-1)][32] */
      v->specB = w->ext_specB ? w->ext_specB : (double *)b200_malloc(sizeof(double) * gsz);
      v->band = w->ext_band ? w->ext_band : (double *)b200_malloc(sizeof(double) * 2 * bsz);
      v->coef = w->ext_coef ? w->ext_coef : (double *)b200_malloc(sizeof(double) * 2 * bsz);
      v->dcoef = w->ext_dcoef ? w->ext_dcoef : (double *)b200_malloc(sizeof(double) * 2 * bsz);
      b200_h2d(d_xs, xs, sizeof(double) * (size_t)n * d);
      b200_h2d(v->perm, perm, sizeof(int) * (size_t)n);
      b200_h2d(v->inv, inv, sizeof(int) * (size_t)n);

      /* segments: runs of one cell, split at B200_SEG_MAX, sorted by size (descending) so that the 8 jobs of a CTA
       * of the tiled kernels carry similar work and the longest CTAs start first */
      if (d == 3)
      {
         /* Work decomposition of the streaming 3-D kernels: the cell-sorted point array is cut into nwr CONTIGUOUS
          * ranges of (almost) equal point count, one per resident warp of one persistent wave; each range is cut at cell
          * boundaries and then into CHUNKS of <= 8 consecutive points of one cell -- the unit the device pipeline
          * stages and computes without any data-dependent branch.  Equal point counts give equal fp64 work per warp
          * whatever the density of the cloud (a heavy cell is simply shared by several warps). */
         int nwr = b200k_warp_ranges();
         int nseg = 0, nch = 0, cap = 4096, r = 0, b, prev = -1;
         int *hc = (int *)b200_xmalloc(sizeof(int) * 4 * (size_t)cap);
         int *hw = (int *)b200_xcalloc((size_t)nwr + 1, sizeof(int));
         /* cost model in point-equivalents of the fp64 pipe: 1 per point, 1 per chunk (ring bookkeeping), and per cell
          * change 2 per plane shifted in the same (c1,c2) row or 24 for a full footprint flush / reload */
         double total = 0.0, acc_cost = 0.0, target;
         for (b = 0; b < (int)nbins; b++)
         {
            int cnt = binstart[b + 1] - binstart[b];
            if (cnt == 0)
               continue;
            total += (double)cnt + (double)((cnt + 7) / 8);
            total += (prev >= 0 && (prev >> 6) == (b >> 6) && b - prev < B200_TAPS) ? 2.0 * (b - prev) : 24.0;
            prev = b;
         }
         target = total / (double)nwr;
         prev = -1;
         hw[0] = 0;
         for (b = 0; b < (int)nbins; b++)
         {
            int s = binstart[b], e = binstart[b + 1];
            if (s == e)
               continue;
            nseg++;
            acc_cost += (prev >= 0 && (prev >> 6) == (b >> 6) && b - prev < B200_TAPS) ? 2.0 * (b - prev) : 24.0;
            prev = b;
            while (s < e)
            {
               int c = e - s > 8 ? 8 : e - s;
               /* open the next warp range when this one has its share of the cost */
               while (r + 1 < nwr && acc_cost >= target * (double)(r + 1))
               {
                  r++;
                  hw[r] = nch;
               }
               if (nch + 8 >= cap)
               {
                  cap *= 2;
                  hc = (int *)realloc(hc, sizeof(int) * 4 * (size_t)cap);
                  if (!hc)
                     exit(EXIT_FAILURE);
               }
               hc[4 * nch + 0] = s;
               hc[4 * nch + 1] = c;
               hc[4 * nch + 2] = b;
               hc[4 * nch + 3] = 0;
               nch++;
               s += c;
               acc_cost += (double)c + 1.0;
            }
         }
         while (r + 1 < nwr)
         {
            r++;
            hw[r] = nch;
         }
         hw[nwr] = nch;
         for (i = 0; i < 4; i++) /* padding descriptors (count 0) so that look-ahead loads stay in bounds */
         {
            hc[4 * (nch + i) + 0] = 0;
            hc[4 * (nch + i) + 1] = 0;
            hc[4 * (nch + i) + 2] = 0;
            hc[4 * (nch + i) + 3] = 0;
         }
         v->nseg = nseg;
         v->nchunk = nch;
         v->nwr = nwr;
         v->chunks = (int *)b200_malloc(sizeof(int) * 4 * ((size_t)nch + 4));
         v->cfirst = (int *)b200_malloc(sizeof(int) * ((size_t)nwr + 1));
         b200_h2d(v->chunks, hc, sizeof(int) * 4 * ((size_t)nch + 4));
         b200_h2d(v->cfirst, hw, sizeof(int) * ((size_t)nwr + 1));
         b200_sync();
         free(w->h_chunks);
         w->h_chunks = hc; /* kept: fused launches cut the same chunk list into a different number of ranges */
         w->h_nchunk = nch;
         free(hw);
      }
      if ((poly2d ? b200k_poly2d_offsets(v, d_xs) : b200k_psi(v, d_xs)) || b200_sync())
         return -1;
      /* opt-in: 1-D / 2-D windows evaluate the window in the kernels from the coordinates (nfft12_stream.cu) */
      if (!poly2d && d < 3 && getenv("NFFT4GP_B200_WINDOW_EVAL") && atoi(getenv("NFFT4GP_B200_WINDOW_EVAL")) > 0)
         v->xs = d_xs;
      else
         b200_free(d_xs);
   }
   free(w->h_perm);
   w->h_perm = perm;
   free(inv);
   free(xs);
   free(keys);
   free(binstart);
   w->n = n;
   w->n_global = n_all;
   return 0;
}

/* Every-call part of nfft_interface.c:215-259 (Gauss) / 351-395 (Matern-1/2): new length scale -> new Fourier
 * coefficient tables for the kernel and its derivative kernel.  The node-dependent tables are NOT rebuilt. */
static int window_build(b200_window *w, int kernel, double f, double l, double mu)
{
   size_t bsz = ipow(B200_NBAND, w->d);
   double *c = (double *)b200_xmalloc(sizeof(double) * 2 * bsz);
   w->kernel = kernel;
   w->sigma[0] = kernel == 0 ? l * w->scale * sqrt(2.0) : l * w->scale; /* :219 / :355 */
   w->sigma[1] = w->sigma[0];
   w->mu = mu;
   build_coef(w->d, kernel, 0, w->sigma[0], c);
   if (b200_h2d(w->dev.coef, c, sizeof(double) * 2 * bsz))
      return -1;
   build_coef(w->d, kernel, 1, w->sigma[1], c);
   if (b200_h2d(w->dev.dcoef, c, sizeof(double) * 2 * bsz) || b200_sync())
      return -1;
   free(c);
   w->kernel_scale = f; /* :256 */
   w->built = 1;
   return 0;
}

/* first half of a window's matvec: a = s * x[perm], adjoint-NFFT spread onto the oversampled grid */
static int window_spread(b200_window *w, const double *d_x, double s)
{
   b200_win_dev *v = &w->dev;
   int rc = b200k_spread_fused_prep(v, d_x, s); /* table-free 2-D windows fetch s * x[perm] inside the spread */
   if (rc <= 0)
      return rc;
   return (b200k_prep(v, d_x, s) || b200k_spread(v, g_variant_spread3)) ? -1 : 0;
}

/* second half: FFT, multiplication by the kernel's Fourier coefficients, inverse FFT, interpolation -> dev.f (sorted order) */
static int window_finish(b200_window *w, int with_derivative)
{
   b200_win_dev *v = &w->dev;
   if (b200k_forward_fft(v) || b200k_backward_fft(v, v->coef) || b200k_gather(v, v->f, g_variant_gather3))
      return -1;
   if (with_derivative)
   {
      if (!v->f2)
         v->f2 = (double *)b200_malloc(sizeof(double) * (size_t)(v->n ? v->n : 1));
      /* the adjoint is shared between kernel and derivative kernel (the reference runs it twice, :533-534) */
      if (b200k_backward_fft(v, v->dcoef) || b200k_gather(v, v->f2, g_variant_gather3))
         return -1;
   }
   return 0;
}

static int window_apply(b200_window *w, const double *d_x, double s, int with_derivative)
{
   return (window_spread(w, d_x, s) || window_finish(w, with_derivative)) ? -1 : 0;
}

/* gradient scale factor of nfft_interface.c:536-537 without the f^2 */
static double window_grad_scale(const b200_window *w)
{
   return w->kernel == 0 ? 2.0 * w->scale * sqrt(2.0) / w->sigma[1] : w->scale / w->sigma[1];
}

/* ------------------------------------------------------------------ single-window API */
static nfft4gp_kernel *kernel_shell(int max_n)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)b200_xcalloc(1, sizeof(nfft4gp_kernel));
   k->_max_n = max_n;
   k->_noise_level = 0.0;
   return k;
}

typedef struct
{
   int magic;
   b200_window *win;
   double *d_x, *d_y;     /* staging for the host-pointer API: n and 3n */
   const double **d_fptr; /* device arrays with one entry */
   const double **d_f2ptr;
   const int **d_invptr;
   double *d_c1w;
} b200_single;

/* nfft_interface.c:3-42 */
void *Nfft4GPNFFTKernelParamCreate(int max_n, int dim)
{
   nfft4gp_kernel *k = kernel_shell(max_n);
   b200_single *s = (b200_single *)b200_xcalloc(1, sizeof(b200_single));
   s->magic = B200_MAGIC_SINGLE;
   s->win = window_create(max_n, dim);
   s->win->owner = s;
   k->_external = s;
   printf("Initializing NFFT kernel with N = %d, p = %d, m = %d, eps = %4.2f\n", s->win->N, s->win->p, s->win->m, s->win->eps);
   printf("Problem size: %d, problem dimension: %d\n", max_n, dim);
   return k;
}

static void single_free_device(b200_single *s)
{
   b200_free(s->d_x);
   b200_free(s->d_y);
   b200_free((void *)s->d_fptr);
   b200_free((void *)s->d_f2ptr);
   b200_free((void *)s->d_invptr);
   b200_free(s->d_c1w);
   s->d_x = s->d_y = s->d_c1w = NULL;
   s->d_fptr = s->d_f2ptr = NULL;
   s->d_invptr = NULL;
}

/* nfft_interface.c:44-77 */
void Nfft4GPNFFTKernelParamFree(void *kernel)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)kernel;
   if (!k)
      return;
   b200_single *s = (b200_single *)k->_external;
   if (s)
   {
      single_free_device(s);
      window_free(s->win);
      free(s);
   }
   free(k);
}

/* nfft_interface.c:79-82 */
void Nfft4GPNFFTKernelFree(void *str) { (void)str; }

/* nfft_interface.c:101-128: forget the points of a single-window handle (the next kernel build takes new data) */
int Nfft4GPNFFTKernelParamRemovePoints(void *kernel)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)kernel;
   b200_single *s = k ? (b200_single *)k->_external : NULL;
   if (!s || s->magic != B200_MAGIC_SINGLE)
      return -1;
   /* the staging of the host-pointer API caches the window's device pointers (f, inv) and is sized for the old n */
   single_free_device(s);
   window_free_device(s->win);
   free(s->win->h_perm);
   free(s->win->h_chunks);
   s->win->h_perm = NULL;
   s->win->h_chunks = NULL;
   s->win->h_nchunk = 0;
   s->win->n = s->win->n_global = 0;
   s->win->scale = -1.0;
   s->win->built = 0;
   return 0;
}

/* nfft_interface.c:84-99: additionally drop the plans (here: nothing survives the points but the handle itself) */
int Nfft4GPNFFTKernelParamFreeNFFTKernel(void *kernel) { return Nfft4GPNFFTKernelParamRemovePoints(kernel); }

static int single_kernel_build(void *str, int kernel, double *data, int n, int ldim, int d, double **Kp, double **dKp)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)str;
   b200_single *s = (b200_single *)k->_external;
   b200_window *w = s->win;
   if (Kp == NULL || dKp == NULL)
   {
      printf("Error: NFFT kernel requires Kp and dKp to be not NULL.\n"); /* :136-140 */
      return -1;
   }
   if (n != ldim)
   {
      printf("Error: NFFT kernel requires n == ldim.\n"); /* :142-146 */
      return -1;
   }
   if (d != w->d)
   {
      printf("Error: NFFT kernel created for dimension %d, got %d.\n", w->d, d);
      return -1;
   }
   if (w->scale < 0.0) /* :150 */
   {
      if (window_setup_points(w, data, n, 0, n))
         return -1;
   }
   if (window_build(w, kernel, k->_params[0], k->_params[1], k->_noise_level))
      return -1;
   *Kp = (double *)w; /* :258-259 */
   *dKp = (double *)w;
   return 0;
}

/* nfft_interface.c:129-263 */
int Nfft4GPNFFTKernelGaussianKernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc,
                                    double **Kp, double **dKp)
{
   (void)permr; (void)kr; (void)permc; (void)kc;
   return single_kernel_build(str, 0, data, n, ldim, d, Kp, dKp);
}

/* nfft_interface.c:265-398 */
int Nfft4GPNFFTKernelMatern12Kernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc,
                                    double **Kp, double **dKp)
{
   (void)permr; (void)kr; (void)permc; (void)kc;
   return single_kernel_build(str, 1, data, n, ldim, d, Kp, dKp);
}

static int single_ensure_staging(b200_single *s, int n)
{
   if (!s->d_x)
   {
      const double *fp;
      const int *ip;
      s->d_x = (double *)b200_malloc(sizeof(double) * (size_t)n);
      s->d_y = (double *)b200_malloc(sizeof(double) * 3 * (size_t)n);
      s->d_fptr = (const double **)b200_malloc(sizeof(double *));
      s->d_f2ptr = (const double **)b200_malloc(sizeof(double *));
      s->d_invptr = (const int **)b200_malloc(sizeof(int *));
      s->d_c1w = (double *)b200_malloc(sizeof(double));
      fp = s->win->dev.f;
      ip = s->win->dev.f_orig ? NULL : s->win->dev.inv;
      if (b200_h2d((void *)s->d_fptr, &fp, sizeof(fp)) || b200_h2d((void *)s->d_invptr, &ip, sizeof(ip)) || b200_sync())
         return -1;
   }
   return 0;
}

int b200_window_matvec_dev(b200_window *w, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   b200_single *s = (b200_single *)w->owner;
   double ff = w->kernel_scale * w->kernel_scale;
   if (!w->built || n != w->n)
   {
      printf("Error: NFFT matvec on a kernel that was not built for n = %d.\n", n);
      return -1;
   }
   if (!s || single_ensure_staging(s, n))
      return -1;
   if (window_apply(w, d_x, alpha, 0))
      return -1;
   return b200k_combine(n, 1, s->d_fptr, s->d_invptr, d_x, alpha, ff, w->mu, beta, d_y);
}

int b200_window_grad_matvec_dev(b200_window *w, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   b200_single *s = (b200_single *)w->owner;
   double ff = w->kernel_scale * w->kernel_scale, f2 = w->kernel_scale * 2.0, c1;
   const double *fp;
   if (!w->built || n != w->n)
   {
      printf("Error: NFFT matvec on a kernel that was not built for n = %d.\n", n);
      return -1;
   }
   if (!s || single_ensure_staging(s, n))
      return -1;
   if (window_apply(w, d_x, alpha, 1))
      return -1;
   fp = w->dev.f2;
   c1 = ff * window_grad_scale(w);
   if (b200_h2d((void *)s->d_f2ptr, &fp, sizeof(fp)) || b200_h2d(s->d_c1w, &c1, sizeof(c1)) || b200_sync())
      return -1;
   return b200k_combine_grad(n, 1, s->d_fptr, s->d_f2ptr, s->d_invptr, d_x, alpha, f2, ff, s->d_c1w, w->mu, beta, d_y);
}

static int single_host_call(void *data, int n, double alpha, double *x, double beta, double *y, int grad)
{
   b200_window *w = (b200_window *)data;
   b200_single *s;
   size_t ny = grad ? 3 * (size_t)n : (size_t)n;
   int rc;
   if (!w || w->magic != B200_MAGIC_WINDOW)
   {
      printf("Error: Nfft4GPNFFTMatSymv needs the pointer returned by the NFFT kernel function.\n");
      return -1;
   }
   s = (b200_single *)w->owner;
   if (!w->built || n != w->n) /* before the staging is sized with it */
   {
      printf("Error: NFFT matvec on a kernel that was not built for n = %d.\n", n);
      return -1;
   }
   if (!s || single_ensure_staging(s, n))
      return -1;
   if (b200_h2d(s->d_x, x, sizeof(double) * (size_t)n))
      return -1;
   if (beta != 0.0 && b200_h2d(s->d_y, y, sizeof(double) * ny))
      return -1;
   rc = grad ? b200_window_grad_matvec_dev(w, n, alpha, s->d_x, beta, s->d_y) : b200_window_matvec_dev(w, n, alpha, s->d_x, beta, s->d_y);
   if (rc)
      return -1;
   return b200_d2h(y, s->d_y, sizeof(double) * ny);
}

/* nfft_interface.c:400-497 */
int Nfft4GPNFFTMatSymv(void *data, int n, double alpha, double *x, double beta, double *y)
{
   return single_host_call(data, n, alpha, x, beta, y, 0);
}

/* nfft_interface.c:499-620 */
int Nfft4GPNFFTGradMatSymv(void *data, int n, double alpha, double *x, double beta, double *y)
{
   return single_host_call(data, n, alpha, x, beta, y, 1);
}

/* ------------------------------------------------------------------ additive (multi-window) API */

/* nfft_interface.c:622-674 */
void *Nfft4GPNFFTAdditiveKernelParamCreate(double *data, int n, int ldim, int d, int *windows, int nwindows, int dwindows)
{
   nfft4gp_kernel *k = kernel_shell(3 * n);
   b200_additive *A = (b200_additive *)b200_xcalloc(1, sizeof(b200_additive));
   int skip = 1, i, j;
   int *fw = windows;
   double *dst;
   (void)d;
   A->magic = B200_MAGIC_ADDITIVE;
   A->n = n;
   A->P = nwindows;
   A->dw = dwindows;
   k->_iparams[0] = nwindows;
   k->_iparams[1] = dwindows;
   k->_iparams[2] = 0;
   while (skip < dwindows && windows[nwindows * dwindows - skip] < 0)
   {
      skip++;
      k->_iparams[2]++;
   }
   A->skip_last = k->_iparams[2];
   k->_ibufferp = (int **)b200_xmalloc(sizeof(int *));
   k->_ibufferp[0] = windows; /* caller-owned, kept by pointer (:638-639) */
   k->_buffer = (double *)b200_xmalloc(sizeof(double) * (size_t)n * nwindows * dwindows);
   k->_own_buffer = 1;
   A->win = (b200_window **)b200_xcalloc((size_t)nwindows, sizeof(b200_window *));
   A->dims = (int *)b200_xcalloc((size_t)nwindows, sizeof(int));
   dst = k->_buffer;
   for (i = 0; i < nwindows; i++)
   {
      int actual = 0;
      for (j = 0; j < dwindows; j++)
      {
         if (fw[0] >= 0)
         {
            memcpy(dst, data + (size_t)fw[0] * ldim, sizeof(double) * (size_t)n);
            fw++;
            dst += n;
            actual++;
         }
      }
      A->win[i] = window_create(n, actual);
      A->dims[i] = actual;
      printf("Initializing NFFT kernel with N = %d, p = %d, m = %d, eps = %4.2f\n", B200_NBAND, 1, B200_MCUT, 0.0);
      printf("Problem size: %d, problem dimension: %d\n", n, actual);
   }
   k->_external = A;
   return k;
}

static int additive_ensure_device(b200_additive *A)
{
   int P = A->P, i;
   if (A->d_x)
      return 0;
   A->d_x = (double *)b200_malloc(sizeof(double) * (size_t)A->n);
   A->d_y = (double *)b200_malloc(sizeof(double) * 3 * (size_t)A->n);
   A->d_fptr = (const double **)b200_malloc(sizeof(double *) * (size_t)P);
   A->d_f2ptr = (const double **)b200_malloc(sizeof(double *) * (size_t)P);
   A->d_invptr = (const int **)b200_malloc(sizeof(int *) * (size_t)P);
   A->d_c1w = (double *)b200_malloc(sizeof(double) * (size_t)P);
   {
      const double **fp = (const double **)b200_xmalloc(sizeof(double *) * (size_t)P);
      const int **ip = (const int **)b200_xmalloc(sizeof(int *) * (size_t)P);
      for (i = 0; i < P; i++)
      {
         fp[i] = A->win[i]->dev.f;
         ip[i] = A->win[i]->dev.f_orig ? NULL : A->win[i]->dev.inv;
      }
      if (b200_h2d((void *)A->d_fptr, fp, sizeof(double *) * (size_t)P) ||
          b200_h2d((void *)A->d_invptr, ip, sizeof(int *) * (size_t)P) || b200_sync())
         return -1;
      free(fp);
      free(ip);
   }
   return 0;
}

static void additive_graphs_release(b200_additive *A);

/* ------------------------------------------------------------------ 1-D windows without tap tables (nfft1d_poly.cu)
 * An additive kernel whose windows are all 1-D (BASELINE config C4) takes the polynomial path: per window the handle keeps the
 * scaled coordinates sorted by cell INSIDE blocks of T consecutive points (8 B + a 2-byte block-local index per point and
 * window, against 80 B of taps + key + permutations of the table path), the spread accumulates 15 moments per cell and the
 * gather evaluates one degree-14 polynomial per point.  Under window sharding a rank builds only the windows it owns. */
#define POLY1D_MIN_N 32768 /* below this the per-window kernels are launch-bound either way and the debug taps stay available */
static int additive_wants_poly1d(const b200_additive *A, int n)
{
   const char *e = getenv("NFFT4GP_B200_POLY1D");
   int i;
   if (e && atoi(e) == 0)
      return 0;
   if (A->shard_mode == 2 && b200_comm_size() > 1)
      return 0; /* point-block sharding keeps the table path */
   for (i = 0; i < A->P; i++)
      if (A->win[i]->d != 1)
         return 0;
   return (e && atoi(e) > 0) || n >= POLY1D_MIN_N;
}

static void additive_poly_release(b200_additive *A)
{
   b200_free((void *)A->poly.xt);
   b200_free((void *)A->poly.lp);
   b200_free((void *)A->poly.rs);
   b200_free(A->poly.mom);
   b200_free(A->poly.ctab);
   b200_free(A->d_poly_wscale);
   free(A->poly_win);
   memset(&A->poly, 0, sizeof(A->poly));
   A->poly_win = NULL;
   A->d_poly_wscale = NULL;
   A->poly1d = 0;
}

/* small per-window device state of the polynomial path: grid / band / coefficient slots only (no point tables) */
static int window_setup_small(b200_window *w, int n)
{
   b200_win_dev *v = &w->dev;
   size_t gsz = ipow(B200_NOS, w->d), bsz = ipow(B200_NBAND, w->d);
   window_free_device(w);
   v->d = w->d;
   v->n = n;
   v->grid = w->ext_grid ? w->ext_grid : (double *)b200_malloc(sizeof(double) * gsz);
   v->specA = w->ext_specA ? w->ext_specA : (double *)b200_malloc(sizeof(double) * gsz);
   v->specB = w->ext_specB ? w->ext_specB : (double *)b200_malloc(sizeof(double) * gsz);
   v->band = w->ext_band ? w->ext_band : (double *)b200_malloc(sizeof(double) * 2 * bsz);
   v->coef = w->ext_coef ? w->ext_coef : (double *)b200_malloc(sizeof(double) * 2 * bsz);
   v->dcoef = w->ext_dcoef ? w->ext_dcoef : (double *)b200_malloc(sizeof(double) * 2 * bsz);
   w->n = w->n_global = n;
   return (v->grid && v->specA && v->specB && v->band && v->coef && v->dcoef) ? 0 : -1;
}

/* Order the scaled coordinates xt[0..n) of one 1-D window by grid cell INSIDE blocks of T consecutive points (pure host code, exported
 * for the CPU tests as Nfft4GPB200DebugPoly1dBlockSort): out_xt / out_lp [nblk * T] = offset v = 2 u - 1 in [-1, 1) inside its cell
 * (u = fractional part of 64 x: exact in double, so is v) and position inside its block of the sorted point -- the cell itself is
 * the run the point stands in --, out_rs [nblk][B200_POLY_RS] = the 65 run starts per cell key (c + 32) & 63.
 * Inside a run the points are dealt out by shared-memory bank: the kernels index acc[lp] / xs[lp] (8-byte words, 16 per wavefront)
 * with 16 consecutive entries of a run per half-warp, so the run is emitted in rounds over the 16 classes lp mod 16 (each class in
 * original order): as long as all classes have points left, every group of 16 is conflict-free (ncu of the unordered runs: 3.0
 * wavefronts per ideal one, the shared-memory pipe 70 % busy).  Deterministic: every sum the kernels form has a fixed order. */
void b200_poly1d_block_sort(const double *xt, int n, int T, double *out_xt, unsigned short *out_lp, unsigned short *out_rs)
{
   const int nblk = (n + T - 1) / T;
   int b, j;
   unsigned short *tmp = (unsigned short *)b200_xcalloc((size_t)T, sizeof(unsigned short));
   unsigned char *key = (unsigned char *)b200_xcalloc((size_t)T, 1);
   for (b = 0; b < nblk; b++)
   {
      const int base = b * T, cnt = n - base < T ? n - base : T;
      unsigned short *rs = out_rs + B200_POLY_RS * (size_t)b;
      int count[65], c;
      memset(count, 0, sizeof(count));
      for (j = 0; j < cnt; j++)
      {
         key[j] = (unsigned char)(((int)floor(xt[base + j] * (double)B200_NOS) + 32) & 63);
         count[key[j] + 1]++;
      }
      for (c = 0; c < 64; c++)
         count[c + 1] += count[c];
      for (c = 0; c <= 64; c++)
         rs[c] = (unsigned short)count[c];
      for (j = 0; j < cnt; j++) /* stable counting sort by cell */
         tmp[count[key[j]]++] = (unsigned short)j;
      for (c = 0; c < 64; c++)
      {
         const int r0 = rs[c], r1 = rs[c + 1];
         int len[16], start[16], q, r, pos = r0, longest = 0;
         if (r0 == r1)
            continue;
         memset(len, 0, sizeof(len));
         for (j = r0; j < r1; j++)
            len[tmp[j] & 15]++;
         start[0] = 0;
         for (q = 0; q < 16; q++)
         {
            if (q)
               start[q] = start[q - 1] + len[q - 1];
            if (len[q] > longest)
               longest = len[q];
         }
         {
            /* bucket the run by bank class; this run's slice of out_lp serves as scratch (it is written for good further down) */
            unsigned short *bucket = out_lp + base + r0;
            int fill[16];
            memcpy(fill, start, sizeof(fill));
            for (j = r0; j < r1; j++)
               bucket[fill[tmp[j] & 15]++] = tmp[j];
            /* emit round r: the r-th point of every class that still has one; tmp[r0..r1) receives the final order */
            for (r = 0; r < longest; r++)
               for (q = 0; q < 16; q++)
                  if (r < len[q])
                     tmp[pos++] = bucket[start[q] + r];
         }
      }
      for (j = 0; j < cnt; j++)
      {
         const double t = xt[base + tmp[j]] * (double)B200_NOS;
         out_xt[base + j] = 2.0 * (t - floor(t)) - 1.0;
         out_lp[base + j] = tmp[j];
      }
   }
   free(tmp);
   free(key);
}

int Nfft4GPB200DebugPoly1dBlockSort(const double *xt, int n, int T, double *out_xt, unsigned short *out_lp, unsigned short *out_rs)
{
   if (!xt || n < 1 || T < 32 || T > 65535 || !out_xt || !out_lp || !out_rs)
      return -1;
   b200_poly1d_block_sort(xt, n, T, out_xt, out_lp, out_rs);
   return 0;
}

/* first kernel build of a polynomial-path handle: contiguous work-array slots for the owned windows, scaling of every owned
 * window in the reference's arithmetic, block-local counting sort by cell, upload */
static int additive_poly_setup(b200_additive *A, const double *cols_all, int n)
{
   const int P = A->P;
   int i, li, nw = 0, T, nblk, rc = -1;
   size_t per_w, gsz = B200_NOS, bsz = B200_NBAND;
   double *h_xt = NULL;
   unsigned short *h_lp = NULL, *h_rs = NULL;
   if (b200_init(-1))
      return -1;
   A->poly_win = (int *)b200_xcalloc((size_t)P, sizeof(int));
   for (i = 0; i < P; i++)
      if (b200_comm_owns_window(A, i))
         A->poly_win[nw++] = i;
   /* block size: the gather runs one CTA per block, two CTAs per SM, so the number of blocks should fill whole waves of
    * `slots` CTAs: T = n / (slots m) for the smallest m that brings T under the shared-memory limit (a multiple of 32, >= 1024) */
   {
      const int slots = b200k_poly1d_wave_slots(), tmax = b200k_poly1d_max_block();
      int m = 1;
      while (((long long)n + (long long)slots * m - 1) / ((long long)slots * m) > tmax - 31)
         m++;
      T = (int)(((long long)n + (long long)slots * m - 1) / ((long long)slots * m));
      T = (T + 31) & ~31;
      if (T < 1024)
         T = 1024;
      {
         const char *e = getenv("NFFT4GP_B200_POLY1D_T"); /* tuning aid: block size, a multiple of 32 */
         if (e && atoi(e) >= 32 && atoi(e) <= tmax)
            T = atoi(e) & ~31;
      }
   }
   nblk = (n + T - 1) / T;
   per_w = (size_t)nblk * T;
   A->poly.n = n;
   A->poly.T = T;
   A->poly.nblk = nblk;
   A->poly.nw = nw;
   A->nbatch = nw;
   A->grids_doubles = gsz * (size_t)(nw ? nw : 1);
   A->d_grids = (double *)b200_malloc(sizeof(double) * A->grids_doubles);
   A->d_specA = (double *)b200_malloc(sizeof(double) * gsz * (size_t)(nw ? nw : 1));
   A->d_specB = (double *)b200_malloc(sizeof(double) * gsz * (size_t)(nw ? nw : 1));
   A->d_band = (double *)b200_malloc(sizeof(double) * 2 * bsz * (size_t)(nw ? nw : 1));
   A->d_coef = (double *)b200_malloc(sizeof(double) * 2 * bsz * (size_t)(nw ? nw : 1));
   A->d_dcoef = (double *)b200_malloc(sizeof(double) * 2 * bsz * (size_t)(nw ? nw : 1));
   if (!A->d_grids || !A->d_specA || !A->d_specB || !A->d_band || !A->d_coef || !A->d_dcoef)
      return -1;
   if (nw == 0)
   {
      A->poly1d = 1;
      return 0;
   }
   A->poly.xt = (const double *)b200_malloc(sizeof(double) * per_w * nw);
   A->poly.lp = (const unsigned short *)b200_malloc(sizeof(unsigned short) * per_w * nw);
   A->poly.rs = (const unsigned short *)b200_malloc(sizeof(unsigned short) * B200_POLY_RS * (size_t)nblk * nw);
   A->poly.mom = (double *)b200_malloc(sizeof(double) * 64 * 16 * (size_t)nw);
   A->poly.ctab = (double *)b200_malloc(sizeof(double) * 2 * 64 * 16 * (size_t)nw);
   A->d_poly_wscale = (double *)b200_malloc(sizeof(double) * (size_t)nw);
   if (!A->poly.xt || !A->poly.lp || !A->poly.rs || !A->poly.mom || !A->poly.ctab || !A->d_poly_wscale)
      return -1;
   h_xt = (double *)b200_xcalloc(per_w, sizeof(double));
   h_lp = (unsigned short *)b200_xcalloc(per_w, sizeof(unsigned short));
   h_rs = (unsigned short *)b200_xcalloc(B200_POLY_RS * (size_t)nblk, sizeof(unsigned short));
   for (li = 0; li < nw; li++)
   {
      b200_window *w = A->win[A->poly_win[li]];
      double *xt = window_scale_points(w, cols_all + (size_t)A->poly_win[li] * n * A->dw, n);
      w->ext_grid = A->d_grids + gsz * (size_t)li;
      w->ext_specA = A->d_specA + gsz * (size_t)li;
      w->ext_specB = A->d_specB + gsz * (size_t)li;
      w->ext_band = A->d_band + 2 * bsz * (size_t)li;
      w->ext_coef = A->d_coef + 2 * bsz * (size_t)li;
      w->ext_dcoef = A->d_dcoef + 2 * bsz * (size_t)li;
      if (window_setup_small(w, n))
      {
         free(xt);
         goto done;
      }
      b200_poly1d_block_sort(xt, n, T, h_xt, h_lp, h_rs);
      free(xt);
      if (b200_h2d((void *)(A->poly.xt + per_w * li), h_xt, sizeof(double) * per_w) ||
          b200_h2d((void *)(A->poly.lp + per_w * li), h_lp, sizeof(unsigned short) * per_w) ||
          b200_h2d((void *)(A->poly.rs + B200_POLY_RS * (size_t)nblk * li), h_rs, sizeof(unsigned short) * B200_POLY_RS * (size_t)nblk) || b200_sync())
         goto done;
   }
   A->poly1d = 1;
   rc = 0;
done:
   free(h_xt);
   free(h_lp);
   free(h_rs);
   return rc;
}


static int additive_kernel_build(void *str, int kernel, double *data, int n, int ldim, int d, double **Kp, double **dKp)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)str;
   b200_additive *A = (b200_additive *)k->_external;
   double *cols = k->_buffer;
   int i;
   (void)data; (void)d;
   if (Kp == NULL || dKp == NULL)
   {
      printf("Error: NFFT kernel requires Kp and dKp to be not NULL.\n");
      return -1;
   }
   if (n != ldim || n != A->n)
   {
      printf("Error: NFFT kernel requires n == ldim.\n");
      return -1;
   }
   if (!A->d_grids && A->poly1d == 0)
      A->poly1d = additive_wants_poly1d(A, n) ? 2 : -1; /* 2 = wanted, set up below */
   if (A->poly1d == 2)
   {
      A->blk_lo = 0;
      A->blk_hi = n;
      if (additive_poly_setup(A, cols, n))
      {
         printf("Error: set-up of the 1-D polynomial path failed.\n");
         return -1;
      }
   }
   if (A->poly1d == 1)
   {
      double *c1 = (double *)b200_xmalloc(sizeof(double) * (size_t)(A->poly.nw ? A->poly.nw : 1));
      double ff = k->_params[0] * k->_params[0];
      for (i = 0; i < A->poly.nw; i++)
      {
         b200_window *w = A->win[A->poly_win[i]];
         if (window_build(w, kernel, k->_params[0], k->_params[1], k->_noise_level))
         {
            free(c1);
            return -1;
         }
         c1[i] = ff * window_grad_scale(w);
      }
      if (A->poly.nw && (b200_h2d(A->d_poly_wscale, c1, sizeof(double) * (size_t)A->poly.nw) || b200_sync()))
      {
         free(c1);
         return -1;
      }
      free(c1);
      /* the hyper-parameters of the epilogue are read from any built window */
      for (i = 0; i < A->P; i++)
      {
         A->win[i]->kernel_scale = k->_params[0];
         A->win[i]->mu = k->_noise_level;
      }
      A->kernel = kernel;
      A->built = 1;
      additive_graphs_release(A);
      if (additive_ensure_device(A))
         return -1;
      if (A->shard_mode == 1 && b200_comm_size() > 1 && !A->d_ypart)
         A->d_ypart = (double *)b200_malloc(sizeof(double) * (size_t)A->n);
      *Kp = (double *)str;
      *dKp = (double *)str;
      return 0;
   }
   if (!A->d_grids)
   {
      /* one contiguous buffer for the oversampled grids of all windows (the allreduce buffer of point-block sharding) */
      size_t off = 0;
      for (i = 0; i < A->P; i++)
         off += ipow(B200_NOS, A->win[i]->d);
      if (b200_init(-1))
         return -1;
      A->grids_doubles = off;
      A->d_grids = (double *)b200_malloc(sizeof(double) * off);
      off = 0;
      for (i = 0; i < A->P; i++)
      {
         A->win[i]->ext_grid = A->d_grids + off;
         off += ipow(B200_NOS, A->win[i]->d);
      }
      {
         /* the leading windows of equal dimension share contiguous work arrays: their FFT passes run as one batch */
         int d0 = A->win[0]->d, nb = 0;
         size_t gsz = ipow(B200_NOS, d0), bsz = ipow(B200_NBAND, d0);
         while (nb < A->P && A->win[nb]->d == d0)
            nb++;
         A->nbatch = nb;
         A->d_specA = (double *)b200_malloc(sizeof(double) * gsz * (size_t)nb);
         A->d_specB = (double *)b200_malloc(sizeof(double) * gsz * (size_t)nb);
         A->d_band = (double *)b200_malloc(sizeof(double) * 2 * bsz * (size_t)nb);
         A->d_coef = (double *)b200_malloc(sizeof(double) * 2 * bsz * (size_t)nb);
         A->d_dcoef = (double *)b200_malloc(sizeof(double) * 2 * bsz * (size_t)nb);
         for (i = 0; i < nb; i++)
         {
            A->win[i]->ext_specA = A->d_specA + gsz * (size_t)i;
            A->win[i]->ext_specB = A->d_specB + gsz * (size_t)i;
            A->win[i]->ext_band = A->d_band + 2 * bsz * (size_t)i;
            A->win[i]->ext_coef = A->d_coef + 2 * bsz * (size_t)i;
            A->win[i]->ext_dcoef = A->d_dcoef + 2 * bsz * (size_t)i;
         }
      }
      A->blk_lo = 0;
      A->blk_hi = n;
      if (A->shard_mode == 2 && b200_comm_size() > 1)
      {
         A->blk_lo = (int)((long long)n * b200_comm_rank() / b200_comm_size());
         A->blk_hi = (int)((long long)n * (b200_comm_rank() + 1) / b200_comm_size());
      }
   }
   for (i = 0; i < A->P; i++)
   {
      b200_window *w = A->win[i];
      if (!b200_comm_owns_window(A, i))
      {
         /* window sharding: another rank tabulates and applies this window; the epilogue reads f and mu from window 0 */
         w->kernel_scale = k->_params[0];
         w->mu = k->_noise_level;
         cols += (size_t)n * A->dw;
         continue;
      }
      if (w->scale < 0.0)
      {
         if (window_setup_points(w, cols, n, A->blk_lo, A->blk_hi))
            return -1;
      }
      /* hyper-parameters are read from the handle at build time (:697-699) */
      if (window_build(w, kernel, k->_params[0], k->_params[1], k->_noise_level))
         return -1;
      cols += (size_t)n * A->dw; /* :703 -- the reference advances by dwindows columns per window */
   }
   A->kernel = kernel;
   A->built = 1;
   A->grad_tables_ready = 0;   /* the gradient factors depend on the length scale */
   additive_graphs_release(A); /* hyper-parameters are baked into the captured kernel arguments */
   if (additive_ensure_device(A) || b200_comm_build_local_windows(A))
      return -1;
   *Kp = (double *)str; /* :730-731 */
   *dKp = (double *)str;
   return 0;
}

/* nfft_interface.c:676-734 */
int Nfft4GPNFFTAdditiveKernelGaussianKernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc,
                                            int kc, double **Kp, double **dKp)
{
   (void)permr; (void)kr; (void)permc; (void)kc;
   return additive_kernel_build(str, 0, data, n, ldim, d, Kp, dKp);
}

/* nfft_interface.c:736-794 */
int Nfft4GPNFFTAdditiveKernelMatern12Kernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc,
                                            int kc, double **Kp, double **dKp)
{
   (void)permr; (void)kr; (void)permc; (void)kc;
   return additive_kernel_build(str, 1, data, n, ldim, d, Kp, dKp);
}

b200_additive *b200_additive_from_handle(void *data)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)data;
   b200_additive *A;
   if (!k)
      return NULL;
   A = (b200_additive *)k->_external;
   if (!A || A->magic != B200_MAGIC_ADDITIVE)
      return NULL;
   return A;
}

/* number of vector entries this rank owns: n, or its point block under point-block sharding (mode 2) */
static int additive_local_n(const b200_additive *A) { return A->blk_hi - A->blk_lo; }


/* ------------------------------------------------------------------ fused launches over the windows of an additive kernel
 * The streaming 3-D kernels run as ONE persistent wave.  Launching them once per window costs a fixed 20-30 us each time
 * (ramp-up, the final footprint flush / first footprint load of every warp, the tail of the wave); with all windows in one
 * launch that is paid once: the wave's warps are divided among the windows in proportion to their cost and every window's
 * chunk list is re-cut into that many ranges of equal cost. */
/* point-equivalents of the tensor pipe: 1 per point, 1 per chunk, `flush` per cell change (spread: the footprint is written
 * out with 1216 fp64 REDs; gather: it is re-read from the L2-resident grid) -- measured optima on a Gaussian cloud, where
 * the sparse periphery makes the difference (profiles/README.md); uniform clouds are insensitive */
#define FUSED_FLUSH_COST_SPREAD 60.0
#define FUSED_FLUSH_COST_GATHER 24.0
/* Work distribution of the fused launches: every warp of the persistent wave gets one large range up front and then
 * takes progressively smaller ones from a device counter (guided self-scheduling).  Equal static shares finish far
 * apart -- in-kernel clocks showed 326 k .. 641 k cycles per warp -- because the warp schedulers serve the resident warps
 * of a sub-partition unevenly and the cost model is only approximate.  NFFT4GP_B200_ROUNDS="w0,w1,..." overrides. */
#define FUSED_MAX_ROUNDS 8
#define FUSED_MIN_RANGE_CHUNKS 12 /* a range start costs about 1.5 chunks (ring reset, pipeline refill, final 10-plane flush) */
static int g_fused_rounds_env = -1; /* -1 not read, 0 no override */
static double g_fused_weight_env[FUSED_MAX_ROUNDS];
/* weights of the rounds for a wave whose warps get `chunks_per_warp` chunks each on average: 8:4:2:1:1 when the smallest
 * range still has FUSED_MIN_RANGE_CHUNKS chunks, static shares otherwise */
static int fused_rounds_choose(double chunks_per_warp, double *w)
{
   static const double full[5] = {8.0, 4.0, 2.0, 1.0, 1.0};
   int k, rounds;
   if (g_fused_rounds_env < 0)
   {
      const char *e = getenv("NFFT4GP_B200_ROUNDS");
      g_fused_rounds_env = 0;
      while (e && *e && g_fused_rounds_env < FUSED_MAX_ROUNDS)
      {
         char *end;
         double v = strtod(e, &end);
         if (end == e || !(v > 0.0))
            break;
         g_fused_weight_env[g_fused_rounds_env++] = v;
         e = (*end == ',') ? end + 1 : end;
      }
   }
   if (g_fused_rounds_env > 0)
   {
      memcpy(w, g_fused_weight_env, sizeof(double) * (size_t)g_fused_rounds_env);
      return g_fused_rounds_env;
   }
   /* measured (uniform cloud, one GPU): 217 chunks per warp (n = 10^6) dynamic = static, Gaussian cloud 18 % faster;
    * 108 chunks (n = 5 10^5) 3.5 % slower, 54 chunks 10 % slower with the shorter round lists: dynamic only for long lists */
   (void)k;
   rounds = (chunks_per_warp / 16.0 >= FUSED_MIN_RANGE_CHUNKS) ? 5 : 1;
   for (k = 0; k < rounds; k++)
      w[k] = full[5 - rounds + k];
   return rounds;
}

static double chunk_cost(const int *hc, int c, int prev_key, double flush)
{
   return (double)hc[4 * c + 1] + 1.0 + (hc[4 * c + 2] != prev_key ? flush : 0.0);
}

static void additive_fused_release(b200_additive *A)
{
   int i, k;
   for (k = 0; k < 3; k++)
   {
      b200_free(A->d_fused[k]);
      A->d_fused[k] = NULL;
   }
   for (k = 0; k < 2; k++)
   {
      if (A->d_fused_cfirst[k])
         for (i = 0; i < A->nbatch; i++)
            b200_free(A->d_fused_cfirst[k][i]);
      free(A->d_fused_cfirst[k]);
      A->d_fused_cfirst[k] = NULL;
   }
   b200_free((void *)A->d_prep_perm);
   b200_free(A->d_prep_a);
   b200_free(A->d_prep_n);
   A->d_prep_perm = NULL;
   A->d_prep_a = NULL;
   A->d_prep_n = NULL;
   A->fused_ready = 0;
}

/* Cut one window's chunk list (int4 descriptors {first point, count, cell key, 0}) into rounds * wcnt consecutive ranges:
 * round k holds wcnt ranges of the cost share weight[k] / sum(weight) / wcnt each, cost = points + 1 per chunk + `flush` per
 * cell change.  first[0] = 0, first[rounds * wcnt] = nch.  Pure host code (exported for the CPU tests as
 * Nfft4GPB200DebugPartitionChunks). */
void b200_partition_chunks(const int *hc, int nch, int wcnt, int rounds, const double *weight, double flush, int *first)
{
   int nr = rounds * wcnt, r = 0, prev = -1, c, k;
   double acc = 0.0, total = 0.0, wsum = 0.0, wacc = 0.0, *start = (double *)b200_xcalloc((size_t)nr + 1, sizeof(double));
   for (c = 0; c < nch; c++)
   {
      total += chunk_cost(hc, c, prev, flush);
      prev = hc[4 * c + 2];
   }
   for (k = 0; k < rounds; k++)
      wsum += weight[k];
   for (k = 0; k < rounds; k++)
   {
      for (c = 0; c < wcnt; c++)
         start[k * wcnt + c] = total * (wacc + weight[k] * (double)c / (double)wcnt) / wsum;
      wacc += weight[k];
   }
   first[0] = 0;
   prev = -1;
   for (c = 0; c < nch; c++)
   {
      while (r + 1 < nr && acc >= start[r + 1])
         first[++r] = c;
      acc += chunk_cost(hc, c, prev, flush);
      prev = hc[4 * c + 2];
   }
   while (r + 1 < nr)
      first[++r] = nch;
   first[nr] = nch;
   free(start);
}

/* cut every batch window's chunk list into ranges of equal cost; the warps of the wave are divided among the windows in
 * proportion to their cost.  dir 0 = spread, 1 = gather.  Fills d_fused[dir] (gather: output f) and d_fused_cfirst[dir]. */
static int additive_fused_partition(b200_additive *A, int dir, int total_warps)
{
   const double flush = dir ? FUSED_FLUSH_COST_GATHER : FUSED_FLUSH_COST_SPREAD;
   int nb = A->nbatch, i, c, assigned = 0, big = 0, rc = -1;
   double *cost = (double *)b200_xcalloc((size_t)nb, sizeof(double)), total = 0.0;
   int *wbeg = (int *)b200_xcalloc((size_t)nb + 1, sizeof(int)), *wcnt = (int *)b200_xcalloc((size_t)nb, sizeof(int));
   b200_mma_win *hw = (b200_mma_win *)b200_xcalloc((size_t)nb, sizeof(b200_mma_win));
   for (i = 0; i < nb; i++)
   {
      const int *hc = A->win[i]->h_chunks;
      int prev = -1;
      for (c = 0; c < A->win[i]->h_nchunk; c++)
      {
         cost[i] += chunk_cost(hc, c, prev, flush);
         prev = hc[4 * c + 2];
      }
      total += cost[i];
   }
   for (i = 0; i < nb; i++)
   {
      wcnt[i] = (int)(total_warps * cost[i] / total);
      if (wcnt[i] < 1)
         wcnt[i] = 1;
      assigned += wcnt[i];
      if (cost[i] > cost[big])
         big = i;
   }
   wcnt[big] += total_warps - assigned;
   if (wcnt[big] < 1)
      goto done;
   for (i = 0; i < nb; i++)
      wbeg[i + 1] = wbeg[i] + wcnt[i];
   A->d_fused_cfirst[dir] = (int **)b200_xcalloc((size_t)nb, sizeof(int *));
   for (i = 0; i < nb; i++)
   {
      /* rounds * wcnt consecutive ranges: round k holds wcnt ranges of the cost share weight[k] / wsum / wcnt each */
      const int *hc = A->win[i]->h_chunks;
      b200_win_dev *v = &A->win[i]->dev;
      int nr = A->fused_rounds * wcnt[i];
      int nch = A->win[i]->h_nchunk, *hr = (int *)b200_xcalloc((size_t)nr + 1, sizeof(int)), bad;
      b200_partition_chunks(hc, nch, wcnt[i], A->fused_rounds, A->fused_weight, flush, hr);
      A->d_fused_cfirst[dir][i] = (int *)b200_malloc(sizeof(int) * ((size_t)nr + 1));
      bad = !A->d_fused_cfirst[dir][i] || b200_h2d(A->d_fused_cfirst[dir][i], hr, sizeof(int) * ((size_t)nr + 1)) || b200_sync();
      free(hr);
      if (bad)
         goto done;
      hw[i].psi = v->psi;
      hw[i].a = v->a;
      hw[i].chunks = v->chunks;
      hw[i].cfirst = A->d_fused_cfirst[dir][i];
      hw[i].grid = v->grid;
      hw[i].f = v->f;
      hw[i].warp_begin = wbeg[i];
      hw[i].nwarps = wcnt[i];
   }
   A->d_fused[dir] = b200_malloc(sizeof(b200_mma_win) * (size_t)nb);
   rc = (!A->d_fused[dir] || b200_h2d(A->d_fused[dir], hw, sizeof(b200_mma_win) * (size_t)nb) || b200_sync()) ? -1 : 0;
done:
   free(cost);
   free(wbeg);
   free(wcnt);
   free(hw);
   return rc;
}

/* returns 1 when the fused path is available */
static int additive_fused_prepare(b200_additive *A)
{
   int nb = A->nbatch, i, c, total_warps = b200k_mma_warp_ranges();
   long long total_chunks = 0;
   int *hn;
   const int **hperm;
   double **ha;
   if (g_variant_spread3 >= 0 || g_variant_gather3 >= 0)
      return 0; /* an explicit kernel variant was requested (A/B runs): per-window launches */
   if (A->fused_ready)
      return A->fused_ready > 0;
   A->fused_ready = -1;
   if (nb < 2 || A->win[0]->d != 3 || (A->shard_mode == 1 && b200_comm_size() > 1))
      return 0;
   for (i = 0; i < nb; i++)
   {
      b200_win_dev *v = &A->win[i]->dev;
      if (!A->win[i]->h_chunks || v->nchunk <= 0 || v->nseg <= 0 || v->n / v->nseg < B200_MMA_MIN_POINTS_PER_CELL)
         return 0; /* a sparse or empty window: per-window launches (DFMA kernels with register-tile shifting) */
      total_chunks += v->nchunk;
   }
   /* per-window pointers of the batched prep kernel */
   hperm = (const int **)b200_xcalloc((size_t)nb, sizeof(int *));
   ha = (double **)b200_xcalloc((size_t)nb, sizeof(double *));
   hn = (int *)b200_xcalloc((size_t)nb, sizeof(int));
   A->prep_nmax = 0;
   for (i = 0; i < nb; i++)
   {
      hperm[i] = A->win[i]->dev.perm;
      ha[i] = A->win[i]->dev.a;
      hn[i] = A->win[i]->dev.n;
      if (hn[i] > A->prep_nmax)
         A->prep_nmax = hn[i];
   }
   A->d_prep_perm = (const int **)b200_malloc(sizeof(int *) * (size_t)nb);
   A->d_prep_a = (double **)b200_malloc(sizeof(double *) * (size_t)nb);
   A->d_prep_n = (int *)b200_malloc(sizeof(int) * (size_t)nb);
   c = !A->d_prep_perm || !A->d_prep_a || !A->d_prep_n || b200_h2d((void *)A->d_prep_perm, hperm, sizeof(int *) * (size_t)nb) ||
       b200_h2d(A->d_prep_a, ha, sizeof(double *) * (size_t)nb) || b200_h2d(A->d_prep_n, hn, sizeof(int) * (size_t)nb) ||
       (A->fused_rounds = fused_rounds_choose((double)total_chunks / (double)total_warps, A->fused_weight)) < 1 ||
       additive_fused_partition(A, 0, total_warps) || additive_fused_partition(A, 1, total_warps);
   free(hperm);
   free(ha);
   free(hn);
   if (c)
   {
      additive_fused_release(A);
      A->fused_ready = -1;
      return 0;
   }
   A->fused_warps = total_warps;
   A->fused_ready = 1;
   return 1;
}

/* derivative channel of the gather: the gather descriptors with f2 as the output */
static int additive_fused_channel1(b200_additive *A)
{
   int nb = A->nbatch, i, rc;
   b200_mma_win *hw;
   if (A->d_fused[2])
      return 0;
   hw = (b200_mma_win *)b200_xcalloc((size_t)nb, sizeof(b200_mma_win));
   if (b200_d2h(hw, A->d_fused[1], sizeof(b200_mma_win) * (size_t)nb))
   {
      free(hw);
      return -1;
   }
   for (i = 0; i < nb; i++)
      hw[i].f = A->win[i]->dev.f2;
   A->d_fused[2] = b200_malloc(sizeof(b200_mma_win) * (size_t)nb);
   rc = !A->d_fused[2] || b200_h2d(A->d_fused[2], hw, sizeof(b200_mma_win) * (size_t)nb) || b200_sync();
   free(hw);
   return rc ? -1 : 0;
}

/* spread of all windows, then (mode 2) the ONE collective of the matvec, then FFT / scale / inverse FFT / gather */
#define PHASE_SPREAD 1
#define PHASE_COLLECTIVE 2
#define PHASE_FINISH 4
#define PHASE_ALL 7
static int additive_apply_windows(b200_additive *A, const double *d_x, double s, int with_derivative, int phases)
{
   int i, fused = additive_fused_prepare(A);
   if (!(phases & PHASE_SPREAD))
      goto collective;
   if (fused)
   {
      /* batched prep, one memset of the contiguous grids, one persistent wave for the spread of all batch windows */
      size_t gsz = ipow(B200_NOS, 3) * (size_t)A->nbatch;
      if (b200k_prep_batch(A->nbatch, A->d_prep_perm, A->d_prep_a, A->d_prep_n, A->prep_nmax, d_x, s) ||
          b200k_spread_batch_begin(A->d_grids, gsz) || b200k_spread3_mma_batch((const b200_mma_win *)A->d_fused[0], A->nbatch, A->fused_warps, A->fused_rounds) ||
          b200k_spread_batch_end())
         return -1;
      for (i = A->nbatch; i < A->P; i++)
         if (window_spread(A->win[i], d_x, s))
            return -1;
   }
   else
      for (i = 0; i < A->P; i++)
         if (b200_comm_owns_window(A, i) && window_spread(A->win[i], d_x, s))
            return -1;
collective:
   if ((phases & PHASE_COLLECTIVE) && A->shard_mode == 2 && b200_comm_size() > 1 && b200_comm_allreduce_sum(A->d_grids, A->grids_doubles))
      return -1;
   if (!(phases & PHASE_FINISH))
      return 0;
   if (!(A->shard_mode == 1 && b200_comm_size() > 1) && A->nbatch > 1)
   {
      /* equally shaped leading windows: every FFT pass is ONE launch over the batch */
      b200_win_dev *v0 = &A->win[0]->dev;
      if (b200k_forward_fft_batch(v0, A->nbatch) || b200k_backward_fft_batch(v0, A->nbatch, 0))
         return -1;
      if (fused)
      {
         if (b200k_gather_batch_begin() || b200k_gather3_mma_batch((const b200_mma_win *)A->d_fused[1], A->nbatch, A->fused_warps, A->fused_rounds) ||
             b200k_gather_batch_end())
            return -1;
      }
      else
         for (i = 0; i < A->nbatch; i++)
            if (b200k_gather(&A->win[i]->dev, A->win[i]->dev.f, g_variant_gather3))
               return -1;
      if (with_derivative)
      {
         for (i = 0; i < A->nbatch; i++)
            if (!A->win[i]->dev.f2)
               A->win[i]->dev.f2 = (double *)b200_malloc(sizeof(double) * (size_t)(A->win[i]->dev.n ? A->win[i]->dev.n : 1));
         if (b200k_backward_fft_batch(v0, A->nbatch, 1))
            return -1;
         if (fused)
         {
            if (additive_fused_channel1(A) || b200k_gather_batch_begin() ||
                b200k_gather3_mma_batch((const b200_mma_win *)A->d_fused[2], A->nbatch, A->fused_warps, A->fused_rounds) || b200k_gather_batch_end())
               return -1;
         }
         else
            for (i = 0; i < A->nbatch; i++)
               if (b200k_gather(&A->win[i]->dev, A->win[i]->dev.f2, g_variant_gather3))
                  return -1;
      }
      for (i = A->nbatch; i < A->P; i++)
         if (window_finish(A->win[i], with_derivative))
            return -1;
      return 0;
   }
   for (i = 0; i < A->P; i++)
      if (b200_comm_owns_window(A, i) && window_finish(A->win[i], with_derivative))
         return -1;
   return 0;
}

/* device-resident body of Nfft4GPAdditiveNFFTMatSymv (nfft_interface.c:796-817).
 * n = length of d_x / d_y = the rank's block under point-block sharding, the full n otherwise. */
static void additive_graphs_release(b200_additive *A)
{
   int i;
   for (i = 0; i < 4; i++)
   {
      b200_graph_destroy(A->graph[i].exec);
      b200_graph_destroy(A->graph[i].exec2);
      A->graph[i].exec = A->graph[i].exec2 = NULL;
   }
   A->graph_next = 0;
}

/* all windows 1-D: spread (moments) -> batched FFTs -> coefficient tables -> gather with the fused epilogue */
static int additive_poly_matvec(b200_additive *A, int n, double alpha, const double *d_x, double beta, double *d_y, int grad)
{
   const double s = 1.0 / (double)A->P * alpha, f = A->win[0]->kernel_scale, ff = f * f, mu = A->win[0]->mu;
   const int sharded = (A->shard_mode == 1 && b200_comm_size() > 1);
   const b200_poly1d_dev *p = &A->poly;
   const size_t ny = grad ? 3 * (size_t)n : (size_t)n;
   double *out = d_y;
   if (sharded)
   {
      if (grad && !A->d_ypart3)
         A->d_ypart3 = (double *)b200_malloc(sizeof(double) * 3 * (size_t)A->n);
      out = grad ? A->d_ypart3 : A->d_ypart;
      if (!out)
         return -1;
   }
   if (p->nw > 0)
   {
      const b200_win_dev *v0 = &A->win[A->poly_win[0]]->dev;
      if (b200k_poly1d_spread(p, d_x, s, A->d_grids, B200_NOS) || b200k_forward_fft_batch(v0, p->nw) || b200k_backward_fft_batch(v0, p->nw, 0) ||
          b200k_poly1d_coef(p, 0, A->d_grids, B200_NOS, NULL))
         return -1;
      if (grad && (b200k_backward_fft_batch(v0, p->nw, 1) || b200k_poly1d_coef(p, 1, A->d_grids, B200_NOS, A->d_poly_wscale)))
         return -1;
      if (b200k_poly1d_gather(p, grad ? 2 : 1, d_x, s, grad ? 2.0 * f : ff, mu, ff, sharded ? 0.0 : beta, out, (size_t)n))
         return -1;
   }
   else if (b200_memset(out, 0, sizeof(double) * ny)) /* a rank without windows contributes zero */
      return -1;
   if (!sharded)
      return 0;
   /* window sharding: ONE allreduce of the partial result (n values, 3 n for the gradient blocks), then y = beta y + sum */
   if (b200_comm_allreduce_sum(out, ny))
      return -1;
   if (beta == 0.0)
      return b200_d2d(d_y, out, sizeof(double) * ny);
   if (beta != 1.0 && b200k_scale((int)ny, beta, d_y))
      return -1;
   return b200k_axpy((int)ny, 1.0, out, d_y);
}

static int additive_matvec_body(b200_additive *A, int n, double alpha, const double *d_x, double beta, double *d_y, int phases)
{
   double s = 1.0 / (double)A->P * alpha; /* :806 */
   double f = A->win[0]->kernel_scale;
   if (additive_apply_windows(A, d_x, s, 0, phases))
      return -1;
   if (!(phases & PHASE_FINISH))
      return 0;
   return b200_additive_epilogue(A, n, d_x, s, f * f, A->win[0]->mu, beta, d_y);
}

int b200_additive_matvec_dev(b200_additive *A, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   int i;
   if (!A->built || n != additive_local_n(A))
   {
      printf("Error: additive NFFT matvec on a kernel that was not built for n = %d.\n", n);
      return -1;
   }
   if (A->poly1d == 1)
      return additive_poly_matvec(A, n, alpha, d_x, beta, d_y, 0);
   /* CUDA graph of the launch sequence (fused tensor-core path only), one per argument signature.  Under point-block
    * sharding the collective is NOT captured: two graphs around a plainly enqueued allreduce (a captured NCCL kernel next
    * to the caller's own communicator deadlocked in a 2-GPU bench run). */
   if (!A->graph_off && !b200_outer_capturing() && !b200_prof_active() && !(A->shard_mode == 1 && b200_comm_size() > 1) &&
       additive_fused_prepare(A))
   {
      const int split = (A->shard_mode == 2 && b200_comm_size() > 1);
      long long before;
      void *exec, *exec2 = NULL;
      for (i = 0; i < 4; i++)
         if (A->graph[i].exec && A->graph[i].x == d_x && A->graph[i].y == d_y && A->graph[i].alpha == alpha && A->graph[i].beta == beta)
         {
            b200_launch_count_add(A->graph[i].launches);
            if (!split)
               return b200_graph_launch(A->graph[i].exec);
            return (b200_graph_launch(A->graph[i].exec) || additive_matvec_body(A, n, alpha, d_x, beta, d_y, PHASE_COLLECTIVE) ||
                    b200_graph_launch(A->graph[i].exec2)) ? -1 : 0;
         }
      if (getenv("NFFT4GP_B200_NO_GRAPH"))
         A->graph_off = 1;
      else if (b200_graph_begin() == 0)
      {
         int rc2 = 0;
         before = b200_launch_count();
         i = additive_matvec_body(A, n, alpha, d_x, beta, d_y, split ? PHASE_SPREAD : PHASE_ALL);
         exec = b200_graph_end();
         if (!i && exec && split)
         {
            if (b200_graph_begin() == 0)
            {
               rc2 = additive_matvec_body(A, n, alpha, d_x, beta, d_y, PHASE_FINISH);
               exec2 = b200_graph_end();
            }
            if (rc2 || !exec2)
            {
               b200_graph_destroy(exec);
               exec = NULL;
            }
         }
         if (i || rc2 || !exec)
         {
            /* capture refused: plain launches from now on */
            b200_graph_destroy(exec);
            b200_graph_destroy(exec2);
            A->graph_off = 1;
            if (i || rc2)
               return -1;
         }
         else
         {
            int slot = A->graph_next;
            A->graph_next = (A->graph_next + 1) % 4;
            b200_graph_destroy(A->graph[slot].exec);
            b200_graph_destroy(A->graph[slot].exec2);
            A->graph[slot].x = d_x;
            A->graph[slot].y = d_y;
            A->graph[slot].alpha = alpha;
            A->graph[slot].beta = beta;
            A->graph[slot].exec = exec;
            A->graph[slot].exec2 = exec2;
            A->graph[slot].launches = b200_launch_count() - before;
            if (!split)
               return b200_graph_launch(exec);
            return (b200_graph_launch(exec) || additive_matvec_body(A, n, alpha, d_x, beta, d_y, PHASE_COLLECTIVE) ||
                    b200_graph_launch(exec2)) ? -1 : 0;
         }
      }
      else
         A->graph_off = 1;
   }
   return additive_matvec_body(A, n, alpha, d_x, beta, d_y, PHASE_ALL);
}

/* device-resident body of Nfft4GPAdditiveNFFTGradMatSymv (nfft_interface.c:819-840) */
int b200_additive_grad_matvec_dev(b200_additive *A, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   double s = 1.0 / (double)A->P * alpha;
   double f = A->win[0]->kernel_scale, ff = f * f;
   double *c1;
   const double **fp;
   int i, rc;
   if (!A->built || n != additive_local_n(A))
   {
      printf("Error: additive NFFT matvec on a kernel that was not built for n = %d.\n", n);
      return -1;
   }
   if (A->poly1d == 1)
      return additive_poly_matvec(A, n, alpha, d_x, beta, d_y, 1);
   if (!A->grad_tables_ready)
   {
      /* once per kernel build: derivative-channel buffers, their pointer table and the factors c1_w = f^2 x gradient scale of the
       * window (nfft_interface.c:536-537) -- not per call (the allocation and the synchronous upload used to sit in every matvec) */
      const int sharded = (A->shard_mode == 1 && b200_comm_size() > 1);
      int nl = 0;
      c1 = (double *)b200_xmalloc(sizeof(double) * 2 * (size_t)A->P);
      fp = (const double **)b200_xmalloc(sizeof(double *) * 2 * (size_t)A->P);
      for (i = 0; i < A->P; i++)
      {
         b200_window *w = A->win[i];
         if (!b200_comm_owns_window(A, i))
         {
            c1[i] = 0.0;
            fp[i] = NULL;
            continue;
         }
         if (!w->dev.f2)
            w->dev.f2 = (double *)b200_malloc(sizeof(double) * (size_t)(w->dev.n ? w->dev.n : 1));
         c1[i] = ff * window_grad_scale(w);
         fp[i] = w->dev.f2;
         c1[A->P + nl] = c1[i];
         fp[A->P + nl] = fp[i];
         nl++;
      }
      rc = b200_h2d((void *)A->d_f2ptr, fp, sizeof(double *) * (size_t)A->P) || b200_h2d(A->d_c1w, c1, sizeof(double) * (size_t)A->P);
      if (!rc && sharded)
      {
         if (!A->d_f2ptr_local)
         {
            A->d_f2ptr_local = (const double **)b200_malloc(sizeof(double *) * (size_t)(nl ? nl : 1));
            A->d_c1w_local = (double *)b200_malloc(sizeof(double) * (size_t)(nl ? nl : 1));
            A->d_ypart3 = (double *)b200_malloc(sizeof(double) * 3 * (size_t)A->n);
         }
         rc = !A->d_f2ptr_local || !A->d_c1w_local || !A->d_ypart3 ||
              b200_h2d((void *)A->d_f2ptr_local, fp + A->P, sizeof(double *) * (size_t)nl) || b200_h2d(A->d_c1w_local, c1 + A->P, sizeof(double) * (size_t)nl);
      }
      rc = rc || b200_sync();
      free(c1);
      free(fp);
      if (rc)
         return -1;
      A->grad_tables_ready = 1;
   }
   if (additive_apply_windows(A, d_x, s, 1, PHASE_ALL))
      return -1;
   if (A->shard_mode == 1 && b200_comm_size() > 1)
   {
      /* window sharding: partial gradient blocks of the owned windows, ONE allreduce of 3 n doubles, y = beta y + sum */
      const size_t n3 = 3 * (size_t)n;
      if (b200k_combine_grad(n, A->nlocal, A->d_fptr_local, A->d_f2ptr_local, A->d_invptr_local, d_x, s, 2.0 * f, ff, A->d_c1w_local,
                             A->win[0]->mu, 0.0, A->d_ypart3) ||
          b200_comm_allreduce_sum(A->d_ypart3, n3))
         return -1;
      if (beta == 0.0)
         return b200_d2d(d_y, A->d_ypart3, sizeof(double) * n3);
      if (beta != 1.0 && b200k_scale((int)n3, beta, d_y))
         return -1;
      return b200k_axpy((int)n3, 1.0, A->d_ypart3, d_y);
   }
   return b200k_combine_grad(n, A->P, A->d_fptr, A->d_f2ptr, A->d_invptr, d_x, s, 2.0 * f, ff, A->d_c1w, A->win[0]->mu, beta, d_y);
}

/* reference-facing host-pointer call: x, y are FULL host vectors of length n (3n for the gradient); under point-block
 * sharding every rank moves and computes only its own block [blk_lo, blk_hi) of them */
static int additive_host_call(void *data, int n, double alpha, double *x, double beta, double *y, int grad)
{
   b200_additive *A = b200_additive_from_handle(data);
   int rc, nl, lo, b, nblk = grad ? 3 : 1;
   if (!A)
   {
      printf("Error: not an additive NFFT kernel handle.\n");
      return -1;
   }
   if (n != A->n || additive_ensure_device(A))
      return -1;
   nl = additive_local_n(A);
   lo = A->blk_lo;
   if (b200_h2d(A->d_x, x + lo, sizeof(double) * (size_t)nl))
      return -1;
   if (beta != 0.0)
      for (b = 0; b < nblk; b++)
         if (b200_h2d(A->d_y + (size_t)b * nl, y + (size_t)b * n + lo, sizeof(double) * (size_t)nl))
            return -1;
   rc = grad ? b200_additive_grad_matvec_dev(A, nl, alpha, A->d_x, beta, A->d_y) : b200_additive_matvec_dev(A, nl, alpha, A->d_x, beta, A->d_y);
   if (rc)
      return -1;
   for (b = 0; b < nblk; b++)
      if (b200_d2h(y + (size_t)b * n + lo, A->d_y + (size_t)b * nl, sizeof(double) * (size_t)nl))
         return -1;
   return 0;
}

/* nfft_interface.c:796-817 */
int Nfft4GPAdditiveNFFTMatSymv(void *data, int n, double alpha, double *x, double beta, double *y)
{
   return additive_host_call(data, n, alpha, x, beta, y, 0);
}

/* nfft_interface.c:819-840 */
int Nfft4GPAdditiveNFFTGradMatSymv(void *data, int n, double alpha, double *x, double beta, double *y)
{
   return additive_host_call(data, n, alpha, x, beta, y, 1);
}

int Nfft4GPB200AdditiveNFFTMatSymvDevice(void *data, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   b200_additive *A = b200_additive_from_handle(data);
   if (!A)
   {
      printf("Error: not an additive NFFT kernel handle.\n");
      return -1;
   }
   return b200_additive_matvec_dev(A, n, alpha, d_x, beta, d_y);
}

int Nfft4GPB200AdditiveNFFTGradMatSymvDevice(void *data, int n, double alpha, const double *d_x, double beta, double *d_y)
{
   b200_additive *A = b200_additive_from_handle(data);
   if (!A)
   {
      printf("Error: not an additive NFFT kernel handle.\n");
      return -1;
   }
   return b200_additive_grad_matvec_dev(A, n, alpha, d_x, beta, d_y);
}

/* nfft_interface.c:842-856 */
void Nfft4GPAdditiveNFFTKernelFree(void *str)
{
   nfft4gp_kernel *k = (nfft4gp_kernel *)str;
   b200_additive *A;
   int i;
   if (!k)
      return;
   A = (b200_additive *)k->_external;
   if (A)
   {
      for (i = 0; i < A->P; i++)
         window_free(A->win[i]);
      free(A->win);
      free(A->dims);
      b200_free(A->d_x);
      b200_free(A->d_y);
      b200_free((void *)A->d_fptr);
      b200_free((void *)A->d_f2ptr);
      b200_free((void *)A->d_invptr);
      b200_free(A->d_c1w);
      b200_free(A->d_ypart);
      b200_free(A->d_grids);
      b200_free(A->d_specA);
      b200_free(A->d_specB);
      b200_free(A->d_band);
      b200_free(A->d_coef);
      b200_free(A->d_dcoef);
      b200_free((void *)A->d_fptr_local);
      b200_free((void *)A->d_invptr_local);
      b200_free((void *)A->d_f2ptr_local);
      b200_free(A->d_c1w_local);
      b200_free(A->d_ypart3);
      additive_fused_release(A);
      additive_graphs_release(A);
      additive_poly_release(A);
      free(A);
   }
   if (k->_own_buffer)
      free(k->_buffer);
   free(k->_ibufferp);
   free(k);
}

/* ------------------------------------------------------------------ debug taps + knobs */
int Nfft4GPB200DebugWindowInfo(void *h, int w, int *dim, double *scale, int *nseg)
{
   b200_additive *A = b200_additive_from_handle(h);
   if (!A || w < 0 || w >= A->P)
      return -1;
   if (dim)
      *dim = A->win[w]->d;
   if (scale)
      *scale = A->win[w]->scale;
   if (nseg)
      *nseg = A->win[w]->dev.nseg;
   return 0;
}

static int debug_stage(void *h, int wi, double alpha, const double *x, double *out, int stage)
{
   b200_additive *A = b200_additive_from_handle(h);
   b200_window *w;
   b200_win_dev *v;
   size_t gsz;
   if (!A || wi < 0 || wi >= A->P || !A->built || additive_ensure_device(A))
      return -1;
   if (A->poly1d == 1)
   {
      printf("Error: the stage taps need the tap tables; this handle takes the table-free 1-D path (NFFT4GP_B200_POLY1D=0 disables it).\n");
      return -1;
   }
   w = A->win[wi];
   v = &w->dev;
   gsz = ipow(B200_NOS, w->d);
   if (b200_h2d(A->d_x, x, sizeof(double) * (size_t)A->n) || b200k_prep(v, A->d_x, alpha) || b200k_spread(v, g_variant_spread3))
      return -1;
   if (stage >= 1 && (b200k_forward_fft(v) || b200k_backward_fft(v, v->coef)))
      return -1;
   return b200_d2h(out, v->grid, sizeof(double) * gsz);
}

int Nfft4GPB200DebugSpread(void *h, int w, double alpha, const double *x, double *grid) { return debug_stage(h, w, alpha, x, grid, 0); }
int Nfft4GPB200DebugGridAfterFFT(void *h, int w, double alpha, const double *x, double *g) { return debug_stage(h, w, alpha, x, g, 1); }

int Nfft4GPB200DebugCoeff(void *h, int wi, int which, double *c)
{
   b200_additive *A = b200_additive_from_handle(h);
   b200_window *w;
   if (!A || wi < 0 || wi >= A->P || !A->built)
      return -1;
   w = A->win[wi];
   return b200_d2h(c, which ? w->dev.dcoef : w->dev.coef, sizeof(double) * 2 * ipow(B200_NBAND, w->d));
}

int Nfft4GPB200SetVariant(const char *name, int variant)
{
   if (!strcmp(name, "spread3d"))
      g_variant_spread3 = variant;
   else if (!strcmp(name, "gather3d"))
      g_variant_gather3 = variant;
   else
      return -1;
   return 0;
}
