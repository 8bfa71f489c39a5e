/* Private host-side structures of libnfft4gp_b200 (C). */
#ifndef NFFT4GP_B200_HOST_H
#define NFFT4GP_B200_HOST_H

#include <stddef.h>
#include "b200_dev.h"

#define B200_MAGIC_WINDOW 0x42323057   /* "B20W" */
#define B200_MAGIC_SINGLE 0x42323053   /* "B20S" */
#define B200_MAGIC_ADDITIVE 0x42323041 /* "B20A" */
#define B200_MAGIC_NYS 0x4232304e      /* "B20N" */
#define B200_MAGIC_FSAI 0x42323046     /* "B20F" */
#define B200_MAGIC_AFN 0x42323058      /* "B20X" */
#define B200_MAGIC_DENSE 0x42323044    /* "B20D" */

/* One feature window: the analogue of the reference's str_adj (SRC/external/nfft_interface.h:35-57). */
typedef struct b200_window
{
   int magic;
   int kernel; /* 0 Gauss, 1 Matern-1/2 */
   int d;
   double sigma[2];
   double mu;
   int N, p, m;
   double eps;
   int n, max_n;     /* n = number of points this rank tabulates (local block under point-block sharding) */
   int n_global;     /* number of points of the window on all ranks */
   int NN;
   double scale;        /* < 0 until the points have been scaled (first kernel build) */
   double kernel_scale; /* f */
   int built;
   int *h_perm;         /* host copy of sorted -> original */
   int *h_chunks;       /* host copy of the 3-D chunk descriptors (int4 each), kept to re-partition them for fused launches */
   int h_nchunk;
   void *owner;         /* b200_single* for the single-window API, NULL inside an additive kernel */
   /* work arrays owned by the additive handle (all windows contiguous, so that FFT passes can be batched over windows and
    * the grids form ONE allreduce buffer); NULL = owned by the window */
   double *ext_grid, *ext_specA, *ext_specB, *ext_band, *ext_coef, *ext_dcoef;
   b200_win_dev dev;
} b200_window;

/* Private state behind nfft4gp_kernel._external of an additive NFFT kernel handle. */
typedef struct b200_additive
{
   int magic;
   int n, P, dw, skip_last;
   int kernel, built;
   b200_window **win;
   int *dims;
   double *d_x, *d_y; /* staging for the host-pointer API (n, 3n) */
   const double **d_fptr, **d_f2ptr;
   const int **d_invptr;
   double *d_c1w;
   /* multi-GPU */
   int shard_mode; /* 0 none, 1 by window, 2 by point block */
   int blk_lo, blk_hi;   /* mode 2: this rank's block of points [blk_lo, blk_hi); otherwise [0, n) */
   double *d_grids;      /* all windows' oversampled grids, contiguous: the ONE allreduce buffer of mode 2 */
   size_t grids_doubles;
   double *d_specA, *d_specB, *d_band, *d_coef, *d_dcoef; /* contiguous slots of the leading equally shaped windows */
   int nbatch;           /* number of leading windows with the dimension of window 0 (batched FFT passes) */
   /* fused spread / gather of the leading nbatch dense 3-D windows in one launch (nfft3d_mma.cu) */
   int fused_ready;      /* 0 = not examined, 1 = descriptors built, -1 = not applicable */
   int fused_warps;
   int fused_rounds;        /* rounds of the guided dynamic range scheduling (1 = static shares) and their cost weights */
   double fused_weight[8];
   void *d_fused[3];        /* b200_mma_win[nbatch]: spread, gather (f), gather of the derivative channel (f2) */
   int **d_fused_cfirst[2]; /* host arrays of nbatch device pointers (owned): range partitions for spread and gather */
   const int **d_prep_perm;
   double **d_prep_a;
   int *d_prep_n;
   int prep_nmax;
   /* CUDA graphs of the device-resident matvec, keyed by its arguments (the launch sequence is fixed per handle) */
   struct
   {
      const double *x;
      double *y;
      double alpha, beta;
      void *exec, *exec2; /* exec2: second half (after the collective) under point-block sharding */
      long long launches; /* kernels inside the graph (the launch counter is advanced by this on replay) */
   } graph[4];
   int graph_next, graph_off; /* graph_off: capture failed once or NFFT4GP_B200_NO_GRAPH is set */
   /* all windows 1-D (config C4): table-free polynomial path (nfft1d_poly.cu) over the windows this rank owns */
   int poly1d;              /* 0 = not examined, 1 = active, -1 = not applicable */
   b200_poly1d_dev poly;
   int *poly_win;           /* [poly.nw] window index of local slot li */
   double *d_poly_wscale;   /* [poly.nw] per-window factor of the derivative channel (f^2 x gradient scale) */
   int nlocal;     /* windows owned by this rank under mode 1 */
   const double **d_fptr_local;
   const int **d_invptr_local;
   double *d_ypart;
   /* gradient matvec: per-window pointer / factor tables are uploaded once per kernel build, not per call */
   int grad_tables_ready;
   const double **d_f2ptr_local; /* mode 1: derivative-channel results of the owned windows */
   double *d_c1w_local;
   double *d_ypart3;             /* mode 1: partial gradient blocks (3 n) */
} b200_additive;

void *b200_xmalloc(size_t bytes);
void *b200_xcalloc(size_t n, size_t sz);

b200_additive *b200_additive_from_handle(void *data);
void b200_partition_chunks(const int *hc, int nch, int wcnt, int rounds, const double *weight, double flush, int *first);
int b200_additive_matvec_dev(b200_additive *A, int n, double alpha, const double *d_x, double beta, double *d_y);
int b200_additive_grad_matvec_dev(b200_additive *A, int n, double alpha, const double *d_x, double beta, double *d_y);
int b200_window_matvec_dev(b200_window *w, int n, double alpha, const double *d_x, double beta, double *d_y);
int b200_window_grad_matvec_dev(b200_window *w, int n, double alpha, const double *d_x, double beta, double *d_y);

/* comm_host.c */
int b200_comm_rank(void);
int b200_comm_size(void);
int b200_comm_owns_window(const b200_additive *A, int w);
int b200_comm_allreduce_sum(double *d_buf, size_t count);
int b200_comm_build_local_windows(b200_additive *A);
int b200_additive_epilogue(b200_additive *A, int n, const double *d_x, double s, double ff, double mu, double beta, double *d_y);

/* precond_host.c: device-resident applies used by the device PCG */
typedef int (*b200_dev_solve)(void *prec, int n, double *d_x, const double *d_rhs);
typedef int (*b200_dev_matvec)(void *mat, int n, double alpha, const double *d_x, double beta, double *d_y);
b200_dev_solve b200_lookup_dev_solve(void *host_fn);
void b200_note_foreign_fn_on_native_handle(const void *prec_data);
int b200_dev_fn_is_native(void *fn);
b200_dev_matvec b200_lookup_dev_matvec(void *host_fn, void *mat_data, void **dev_handle);

/* pcg_host.c */
int b200_pcg_device(void *mat, int n, b200_dev_matvec mv, void *prec, b200_dev_solve ps, double *d_x, const double *d_rhs,
                    int maxits, int atol, double tol, double *prel_res, double **prel_res_v, int *piter, int print_level);

#endif
