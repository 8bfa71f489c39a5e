/* Thin C-ABI layer between the C host code (nfft_host.c, pcg_host.c, precond_host.c) and the
 * hand-written sm_100a kernels (*.cu).  Plain pointers and sizes only; every function returns 0 on
 * success and -1 after printing the CUDA error.  All launches go to ONE library stream. */
#ifndef NFFT4GP_B200_DEV_H
#define NFFT4GP_B200_DEV_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fixed NFFT parameters of the reference (SRC/external/nfft_interface.c:18-27) */
#define B200_NBAND 32 /* expansion degree N                  */
#define B200_NOS 64   /* oversampled FFT length (sigma = 2)   */
#define B200_MCUT 4   /* window cut-off m                     */
#define B200_TAPS 10  /* 2m + 2                               */
/* 3-D windows: average points per occupied cell from which the tensor-core spread / gather are used instead of the DFMA
 * kernels.  With the shared-memory footprint ring and the fused launches the tensor-core path wins down to ~ 2 points per
 * cell (n = 10^4 .. 10^5 measured, profiles/README.md); below that almost every chunk holds one point and 3 of the 4
 * k-slots of a DMMA are empty. */
#ifndef B200_MMA_MIN_POINTS_PER_CELL
#define B200_MMA_MIN_POINTS_PER_CELL 2
#endif

/* device-side state of one feature window */
typedef struct
{
   int d;        /* window dimension 1..3                                        */
   int n;        /* number of points                                             */
   int nseg;     /* number of (cell, warp-range) segments (information only)     */
   double *psi;  /* [n + pad][d*10] Kaiser-Bessel taps, cell-sorted point order; NULL on the table-free 2-D path */
   double *pv;   /* 2-D windows on the table-free path (nfft2d_poly.cu): [n][2] in-cell offsets v = 2 u - 1 of the sorted points */
   int *cells;   /* ... [ncell] packed keys of the occupied cells                 */
   int ncell;
   int f_orig;   /* ... the gather writes f / f2 in the ORIGINAL point order: the combine pass reads them without `inv` */
   double *xs;   /* [n][d] sorted scaled coordinates, kept only for 1-D / 2-D windows when the window is evaluated in the kernels
                    (NFFT4GP_B200_WINDOW_EVAL=1); NULL otherwise                  */
   int *key;     /* [n] packed cell index ((c1+32)<<12 | (c2+32)<<6 | (c3+32))    */
   int *perm;    /* [n] sorted position -> original point                        */
   int *inv;     /* [n] original point -> sorted position                        */
   int nchunk;   /* 3-D windows: number of chunk descriptors                      */
   int *chunks;  /* [nchunk + pad] int4 {first point, count <= 8, packed cell key, 0}: <= 8 consecutive points of ONE cell */
   int nwr;      /* number of warp ranges (= resident warps of the 3-D kernels, one persistent wave) */
   int *cfirst;  /* [nwr + 1] warp r owns chunks [cfirst[r], cfirst[r+1]): equal point counts per warp */
   double *a;    /* [n + pad] (alpha/P) * x[perm], sorted order                  */
   double *f;    /* [n] interpolation result, sorted order (kernel channel)      */
   double *f2;   /* [n] derivative-kernel channel (gradient matvec only)         */
   double *grid; /* [64^d] real oversampled grid (spread target, later Re h)     */
   double *specA, *specB, *band; /* complex work arrays of the pruned FFT passes */
   double *coef, *dcoef;         /* [32^d] complex  b_k / prod phi_hat(k_t)^2    */
} b200_win_dev;

/* ---- runtime ---- */
int b200_init(int device);
int b200_initialized(void);
int b200_device_count(void);
void *b200_malloc(size_t bytes);
void b200_free(void *p);
void *b200_host_alloc(size_t bytes);
void b200_host_free(void *p);
int b200_h2d(void *dst, const void *src, size_t bytes);
int b200_d2h(void *dst, const void *src, size_t bytes);
int b200_d2d(void *dst, const void *src, size_t bytes);
int b200_memset(void *dst, int v, size_t bytes);
int b200_sync(void);
void *b200_event_create(void);
void b200_event_destroy(void *ev);
int b200_event_record(void *ev);
float b200_event_elapsed_ms(void *a, void *b);
long long b200_launch_count(void);
void b200_launch_count_reset(void);
void b200_launch_count_add(long long k);
void *b200_stream_handle(void); /* cudaStream_t, for NCCL */
int b200_prof_active(void);
int b200_graph_begin(void);
void *b200_graph_end(void);
int b200_graph_launch(void *exec);
void b200_graph_destroy(void *exec);
void b200_outer_capture_set(int on);
int b200_outer_capturing(void);
int b200_d2h_async(void *dst, const void *src, size_t bytes);
double b200_fp64_tensor_peak_tflops(void);
void b200_prof_start(void);
int b200_prof_stop(double *sums_ms, long long *counts, int ncls);

/* ---- NFFT window kernels (nfft_kernels.cu) ---- */
int b200k_psi(const b200_win_dev *w, const double *d_xs /* [n][d] sorted scaled coords */);
int b200k_prep(const b200_win_dev *w, const double *d_x, double scale);
int b200k_spread(const b200_win_dev *w, int variant);
int b200k_forward_fft(const b200_win_dev *w);                                   /* grid -> band (unscaled) */
int b200k_backward_fft(const b200_win_dev *w, const double *d_coef);            /* band * coef -> grid (Re h) */
int b200k_gather(const b200_win_dev *w, double *d_f, int variant);
/* nb equally shaped windows whose work arrays are stored contiguously (slot = the size for dimension w->d) */
int b200k_forward_fft_batch(const b200_win_dev *w, int nb);
int b200k_backward_fft_batch(const b200_win_dev *w, int nb, int which);
/* nfft3d_warp.cu: warp-per-segment 3-D kernels */
int b200k_warp_ranges(void);
int b200k_spread3_warp(const b200_win_dev *w);
int b200k_gather3_warp(const b200_win_dev *w, double *d_f);
/* nfft12_stream.cu: 1-D / 2-D windows, warp-per-range with the footprint in registers */
int b200k_spread_stream(const b200_win_dev *w);
int b200k_gather_stream(const b200_win_dev *w, double *d_f);
/* nfft1d_poly.cu: ALL 1-D windows of an additive kernel without tap tables -- per-cell polynomial coefficients / moments.
 * Points are sorted by cell inside blocks of T consecutive points of the original order (once per handle). */
#define B200_POLY_RS 80 /* 65 run starts per block, padded to 160 bytes so that a row is 16-byte aligned (bulk staging) */
typedef struct
{
   int n, nblk, T, nw;          /* points of this rank, blocks of T points, windows handled by this rank */
   const double *xt;            /* [nw][nblk * T] scaled coordinates, sorted by cell inside every block */
   const unsigned short *lp;    /* [nw][nblk * T] position inside its block (original order) of the sorted point */
   const unsigned short *rs;    /* [nw][nblk][B200_POLY_RS] run starts: the points of cell key c of a block are [rs[c], rs[c + 1]) */
   double *mom;                 /* [nw][64][16] moments sum_i a_i v_i^k per cell (spread) */
   double *ctab;                /* [2][nw][64][16] polynomial coefficients per cell (gather): kernel / derivative channel */
} b200_poly1d_dev;
int b200k_poly1d_max_block(void);
int b200k_poly1d_wave_slots(void);
int b200k_poly1d_spread(const b200_poly1d_dev *p, const double *d_x, double s, double *d_grids, size_t grid_stride);
int b200k_poly1d_coef(const b200_poly1d_dev *p, int ch, const double *d_grids, size_t grid_stride, const double *d_wscale);
int b200k_poly1d_gather(const b200_poly1d_dev *p, int nch, const double *d_x, double s, double c0, double mu, double ff, double beta,
                        double *d_y, size_t ystride);
/* nfft2d_poly.cu: 2-D windows without tap tables (bivariate polynomial per cell, Vandermonde operands on the fp64 tensor cores) */
int b200k_poly2d_prepare(void);
int b200k_poly2d_offsets(const b200_win_dev *w, const double *d_xs);
int b200k_spread2_poly(const b200_win_dev *w, const double *d_x, double s); /* d_x != NULL: coefficient s * d_x[perm] fetched in the kernel */
int b200k_spread_fused_prep(const b200_win_dev *w, const double *d_x, double s);
int b200k_gather2_poly(const b200_win_dev *w, double *d_f);
/* nfft2d_mma.cu: 2-D windows on the fp64 tensor cores (same data as the stream kernels) */
int b200k_spread2_mma(const b200_win_dev *w);
int b200k_gather2_mma(const b200_win_dev *w, double *d_f);
/* nfft3d_mma.cu: the same streaming decomposition on the fp64 tensor cores (DMMA) */
int b200k_mma_warp_ranges(void);
int b200k_spread3_mma(const b200_win_dev *w);
int b200k_gather3_mma(const b200_win_dev *w, double *d_f);
/* all (equally shaped, dense) 3-D windows of an additive kernel in one launch: per-window pointers and the window's own
 * partition of its chunk list into `nwarps` warp ranges (cfirst has nwarps + 1 entries); warp_begin = prefix sum of nwarps */
typedef struct
{
   const double *psi;
   const double *a;
   const int *chunks;
   const int *cfirst;
   double *grid;
   double *f;
   int warp_begin;
   int nwarps;
} b200_mma_win;
/* `rounds` > 1: window w owns rounds * nwarps consecutive ranges (cfirst has rounds * nwarps + 1 entries); range
 * (round k, slot q of the wave) is cfirst[k * nwarps + q - warp_begin].  Every warp starts with the round-0 range of its
 * slot and then takes the remaining ranges in order from a device counter (dynamic: the warp schedulers do not serve the
 * resident warps evenly, so equal static shares do not end together). */
int b200k_spread3_mma_batch(const b200_mma_win *d_batch, int P, int total_warps, int rounds);
int b200k_gather3_mma_batch(const b200_mma_win *d_batch, int P, int total_warps, int rounds);
int b200k_spread_batch_begin(double *d_grids, size_t doubles);
int b200k_spread_batch_end(void);
int b200k_gather_batch_begin(void);
int b200k_gather_batch_end(void);
/* a_w = s * x[perm_w] for nb windows in one launch (device arrays of per-window pointers / sizes) */
int b200k_prep_batch(int nb, const int *const *d_perm, double *const *d_a, const int *d_n, int nmax, const double *d_x, double s);
/* y = beta*y + sum_w c0*(f_w[inv_w] + mu*s*x)   (matvec epilogue, nfft_interface.c:428-494, 803-814) */
int b200k_combine(int n, int P, const double *const *d_fptr, const int *const *d_invptr, const double *d_x,
                  double s, double c0, double mu, double beta, double *d_y);
/* gradient epilogue (nfft_interface.c:539-617, 826-837): y0 += f2*(f + mu a), y1 += c1*fd, y2 += ff*a */
int b200k_combine_grad(int n, int P, const double *const *d_fptr, const double *const *d_f2ptr, const int *const *d_invptr,
                       const double *d_x, double s, double f2, double ff, const double *d_c1w, double mu, double beta, double *d_y);

/* ---- vector kernels for PCG (vec_kernels.cu) ---- */
int b200k_dot(int n, const double *x, const double *y, double *d_out);          /* deterministic 2-stage */
int b200k_dot2(int n, const double *x, const double *y, const double *u, const double *v, double *d_out2);
int b200k_xpby(int n, const double *z, double beta, double *p);                 /* p = z + beta p */
/* x += alpha p ; r -= alpha q ; out = r.r   (fused PCG update, pcg.c:165-169) */
int b200k_pcg_update(int n, double alpha, const double *p, const double *q, double *x, double *r, double *d_out);
/* the same two updates with the step lengths formed on the device from device-resident dot products (see vec_kernels.cu) */
int b200k_xpby_dev(int n, const double *z, const double *d_rho, const double *d_rho_prev /* NULL: p = z */, double *p);
int b200k_pcg_update_dev(int n, const double *d_rho, const double *d_rho_prev, const double *d_pq, const double *p, const double *q,
                         double *x, double *r, double *d_out);
int b200k_copy(int n, const double *src, double *dst);
int b200k_fill(int n, double v, double *dst);
int b200k_scale(int n, double a, double *x);
int b200k_axpy(int n, double a, const double *x, double *y);
/* sharded applies: gather / scatter through an index array with holes (idx < 0), Nystrom coefficient, local factor rows */
/* device-resident modified Gram-Schmidt step, scaling by a device-side norm, x += Z coef (FGMRES) */
int b200k_mgs_step(int n, double *w, const double *v_prev, const double *d_t_prev, const double *v_next, double *d_out);
int b200k_scale_by_inv_norm(int n, double *w, const double *d_nrm2, double *d_t_out);
int b200k_multi_axpy(int n, int m, const double *Z, size_t ldz, const double *d_coef, double *x);
int b200k_gather_idx(int m, const int *idx, const double *src, double *dst);
int b200k_scatter_idx(int m, const int *idx, const double *src, double *dst);
int b200k_nys_coef(int k, const double *s, double inv_eta, const double *z, double *c);
int b200k_gather_rows_T(int nl, int k, const int *pos, const double *U, int ldu, double *UT);

/* ---- preconditioner kernels (precond_kernels.cu) ---- */
/* Nystrom: y = rhs(perm); z = U'y; out(perm) = y/eta + U((s - 1/eta) .* z) -- two passes over U (n x k col-major) */
int b200k_nys_apply(int n, int k, const int *perm, const double *U, const double *s, double eta,
                    const double *rhs, double *x, double *work /* n + k + partial sums */, size_t work_doubles);
size_t b200k_nys_work_doubles(int n, int k);
/* CSR y = A x (one warp per row) */
int b200k_csr_mv(int nrows, const int *ia, const int *ja, const double *a, const double *x, double *y);
/* gather / scatter by permutation */
int b200k_gather_perm(int n, const int *perm, const double *src, double *dst);  /* dst[i] = src[perm[i]] */
int b200k_scatter_perm(int n, const int *perm, const double *src, double *dst); /* dst[perm[i]] = src[i] */
/* dense column-major GEMV on device: y = alpha*op(A)x + beta*y, A m x n, lda = m */
int b200k_gemv(char trans, int m, int n, double alpha, const double *A, const double *x, double beta, double *y);
/* matrix stored in fp32 (m a multiple of 4), fp64 vectors and accumulation */
int b200k_gemv_f32(char trans, int m, int n, double alpha, const float *A, const double *x, double beta, double *y);
int b200k_convert_f64_f32(size_t count, const double *src, float *dst);
int b200k_convert_f32_f64(size_t count, const float *src, double *dst);
/* dense additive kernel block K(rows, cols) on device (kernels.c:3099-3494 call shapes):
 * Xw: window-major feature copy [P][dw][n] (missing features of a short last window flagged by dims[w] < dw) */
int b200k_dense_additive(const double *Xw, int n, int P, int dw, const int *d_dims, int kernel, double ff, double l,
                         double noise_diag, const int *rows, int nr, const int *cols, int nc, int lower_only,
                         double *K, int ldk);

/* ---- setup-side kernels (setup_kernels.cu) ---- */
/* column-major fp64 GEMM on the tensor cores (DMMA): C = alpha op(A) op(B) + beta C, op = 'N' | 'T' */
int b200k_dgemm(char transa, char transb, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                double beta, double *C, int ldc);
/* farthest point sampling of k points (first = nearest to d_center); X feature-major with leading dimension ld */
int b200k_fps(const double *d_X, int n, int ld, int d, const double *d_center, int k, int *d_perm_out);
/* KNN pattern of the FSAI: row i = its lfil-1 nearest earlier points + the diagonal (last); d_ia[n+1], d_ja[b200k_knn_nnz] */
long long b200k_knn_nnz(int n, int lfil);
int b200k_knn_pattern(const double *d_X, int n, int ld, int d, int lfil, int *d_ia, int *d_ja);
/* FSAI values for every row of a pattern (one small Cholesky solve per row); see setup_kernels.cu */
int b200k_fsai_rows(const double *d_Xw, const int *d_dims, int n_points, int P, int dw, int kernel, double ff, double l,
                    double noise_diag, int nrows, const int *d_ia, const int *d_ja, const int *d_rowmap, const double *d_Z, int kz,
                    int kmax, double *d_a, int *breakdown_out);
int b200k_fsai_rows_grad(const double *d_Xw, const int *d_dims, int n_points, int P, int dw, int kernel, double ff, double l,
                         double noise_diag, int nrows, const int *d_ia, const int *d_ja, int kmax, double *d_a, double *d_da, size_t da_stride,
                         int *breakdown_out);
/* sparse triangular solves with the FSAI factor: upper = 0: G x = rhs (rows of G, diagonal last); 1: G' x = rhs (rows of G', diagonal first) */
int b200k_sptrsv(int upper, int n, const int *d_ia, const int *d_ja, const double *d_a, const double *d_rhs, double *d_x);
int b200k_fsai_trace(int n, const int *d_ia, const double *d_a, const double *d_da, size_t stride, double *d_out3);
int b200k_csr_transpose(int nrows, int ncols, int nnz, const int *d_ia, const int *d_ja, const double *d_a, int *d_iat, int *d_jat,
                        double *d_at);
int b200k_gather_rows(int n, int ncol, const int *d_perm, const double *d_src, int lds, double *d_dst, int ldd);
int b200k_scale_columns(int n, int k, const double *d_scale, double *d_U);
/* blocked Cholesky of an SPD matrix (lower, ld n) on the device; d_Linv = explicit inverse of the factor (upper zero) */
int b200k_potrf_inverse(int n, double *d_A, double *d_Linv, int *info);
int b200k_add_diag(int n, double *d_A, int lda, double v);

#ifdef __cplusplus
}
#endif
#endif
