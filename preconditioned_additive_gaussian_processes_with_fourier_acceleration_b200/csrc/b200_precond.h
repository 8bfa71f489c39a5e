/* Private preconditioner structures of libnfft4gp_b200 (C host side). */
#ifndef NFFT4GP_B200_PRECOND_H
#define NFFT4GP_B200_PRECOND_H

#include <stddef.h>
#include "nfft4gp_b200.h"

/* Nystrom: fields follow precond_nys (INC/_preconds.h:320-351) where they have a meaning here. */
typedef struct
{
   int magic;
   int k_setup;  /* requested rank */
   int own_perm;
   int *perm;    /* caller's permutation (host), landmarks first */
   int n, k;
   double eta;   /* mu * f^2 (nys.c:641) */
   int *d_perm;
   double *d_U;  /* n x k column-major */
   double *d_s;  /* 1 / (sigma_i^2 + eta) */
   double *d_work;
   size_t work_doubles;
   double *d_x, *d_rhs; /* staging for the host-pointer solve */
   int cap;
   double tset;
   /* point-block-sharded apply (one process per GPU): this rank's rows of U, transposed (k x nl), built on first use */
   int sh_lo, sh_hi;
   double *d_UT_loc, *d_sh_work; /* work: 2k */
   /* setup with require_grad (nys.c:555-610): what Dvp / Trace need, all in the permuted order of the setup */
   int has_grad;
   double f;          /* kernel scale of the setup: dK1/df = (2/f) K1 */
   double *d_gK1;     /* K1 = K(perm, perm[0:k]), n x k, no noise */
   double *d_gdK1l;   /* dK1/dl, n x k */
   double *d_gLinv;   /* inverse Cholesky factor of K11 + shift, k x k */
   double *d_gG[2];   /* L^{-1} (dK11/df) L^{-T}, L^{-1} (dK11/dl) L^{-T}  (the reference's _GdKG_data) */
   double *d_gvec;    /* 4 n + 8 k work */
   double traces[3];
   int traces_ready;
} b200_nys;

/* FSAI: G lower-triangular CSR (diagonal last in each row, INC/_preconds.h:271) and its explicit transpose. */
typedef struct
{
   int magic;
   int lfil;
   int n, nnz;
   int *d_ia, *d_ja;
   double *d_a;
   int *d_iat, *d_jat;
   double *d_at;
   double *d_work;
   double *d_x, *d_rhs;
   int cap;
   double tset;
   /* setup with require_grad (fsai.c:495-560): derivatives of the factor values in f, l, mu on the pattern of G and on the pattern
    * of G' (the reference's _dL_a, and its transpose so that dG' x needs no scatter), and 4 n doubles of work for Dvp */
   double *d_da;  /* [3][nnz] */
   double *d_dat; /* [3][nnz] */
   double *d_gwork;
} b200_fsai;

/* AFN (afn.h:27-58 design; MATLAB afn_setup.m / afn_solve.m). */
typedef struct
{
   int magic;
   int max_k, lfil;
   int est_m, est_num, noran; /* rank estimation: sample size (500), repetitions (5; 0 = use max_k), never switch to Nystrom */
   unsigned long long est_seed;
   int est_method;            /* first stage of the rank estimation: 0 = MATLAB afn_setup.m:163-228, 1 = reference C rankest.c:354-391 */
   int k_est;                 /* last estimated rank (before the cap), -1 when the estimation was skipped */
   int n, k;
   int *d_perm;
   double *d_Linv; /* k x k explicit inverse of chol(K11 + shift), lower, column-major */
   double *d_K12;  /* k x (n-k) column-major */
   float *d_K12f;  /* the same block in fp32 when k12_bits == 32 (then d_K12 is released after the setup) */
   int k12_bits;   /* 64 (default) or 32: storage of K12 for the two HBM-bound products of every apply */
   b200_fsai schur;
   /* optional: the two products with K12 of every apply through the NFFT operator instead of the dense block (precond_host.c) */
   void *k12_op;        /* b200_additive* of an NFFT kernel handle built with the hyper-parameters of the setup, or NULL */
   double *d_opz, *d_opt; /* two vectors of length n */
   void *nys;      /* set when the estimated rank is below max_k: Nystrom on the FPS ordering */
   double *d_work; /* 2n + 2k */
   double *d_x, *d_rhs;
   int cap;
   double tset;
   /* point-block-sharded apply, built on first use: owned landmarks / Schur points of the block [sh_lo, sh_hi) */
   int sh_ready, sh_lo, sh_hi, sh_plo, sh_phi; /* [sh_plo, sh_phi): this rank's range of Schur indices (0-based in n - k) */
   int *d_lm_idx;     /* [k]   local index of landmark p, -1 when another rank owns it */
   int *d_sc_idx;     /* [n2l] local index of Schur point sh_plo + j */
   double *d_sh_work; /* 4k + 2 (n - k) + 2 n2l */
} b200_afn;

int b200_nys_set_device_factors(b200_nys *s, int n, int k, const int *h_perm, double *d_U_owned, const double *h_s, double eta);
int b200_fsai_set_host_factors(b200_fsai *s, int n, const int *ia, const int *ja, const double *a);
/* takes ownership of the device CSR arrays and builds the explicit transpose on the device */
int b200_fsai_set_device_factors(b200_fsai *s, int n, int nnz, int *d_ia, int *d_ja, double *d_a);
/* takes ownership of d_da = [3][nnz] derivative values (require_grad) and builds their transposes */
int b200_fsai_set_device_grad(b200_fsai *s, double *d_da);
int b200_afn_set_device_factors(b200_afn *s, int n, int k, const int *h_perm, const double *h_L, double *d_K12_owned,
                                const int *ia, const int *ja, const double *a);
void b200_fsai_release(b200_fsai *s);
void b200_afn_release(b200_afn *s);
int b200_afn_finish_setup(b200_afn *s);
void b200_tril_inverse(int k, const double *L, double *Linv);
int b200_nys_solve_dev(void *vnys, int n, double *d_x, const double *d_rhs);
int b200_fsai_solve_dev(void *vf, int n, double *d_x, const double *d_rhs);
int b200_afn_solve_dev(void *va, int n, double *d_x, const double *d_rhs);
int b200_chol_solve_dev(void *vc, int n, double *d_x, const double *d_rhs);

/* dense_small.c: O(k^3) host factorizations of the k x k landmark block */
double b200_sym_fro_lower(int n, const double *A, int lda);
double b200_chol_shift(int n, double fro);
int b200_potrf_lower(int n, double *A, int lda);
int b200_syev(int n, double *A, double *w);
int b200_syev_values(int n, double *A, double *w);
int b200_stev(int n, double *d, double *e, double *Zm);
double b200_spd_norm2(int m, const double *S, int lds, double *work);

/* dense_host.c: device view of a dense kernel handle that this library can evaluate on the GPU */
typedef struct
{
   const double *d_Xw; /* window-major features [P * dw][n] on the device */
   const int *d_dims;  /* [P] actual dimension of every window */
   int n, P, dw;
   int kernel;         /* 0 Gauss, 1 Matern-1/2 */
   double f, l, mu;    /* hyper-parameters read from the handle at call time */
   int d_all;          /* number of feature columns of the caller's data (KNN / FPS use all of them) */
   int owns;           /* temporary upload (plain kernels): release with b200_dense_view_release */
} b200_dense_view;
/* returns 0 and fills *v when fkernel is one of this library's dense kernels; 1 when it is a foreign function */
int b200_dense_view_acquire(func_kernel fkernel, void *fkernel_params, const double *data, int n, int ldim, int d, b200_dense_view *v);
void b200_dense_view_release(b200_dense_view *v);
/* K(rows, cols) on the device: rows/cols device index arrays (NULL = 0..), column-major with leading dimension ldk */
int b200_dense_view_block(const b200_dense_view *v, double noise_diag, const int *d_rows, int nr, const int *d_cols, int nc,
                          int lower_only, double *d_K, int ldk);
int b200_dense_view_block_dl(const b200_dense_view *v, const int *d_rows, int nr, const int *d_cols, int nc, int lower_only, double *d_K,
                             int ldk);

#endif
