/* ORACLE (test infrastructure, not product code).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The NFFT3 fastsum boundary is a restatement (NFFT3 is absent), PINNED on the 400 numbers
 * the reference ships from its authors' real NFFT3 build (TESTS/TEST2/foo_0_nfft.txt; see nfft3_shim/fastsum.h and
 * tests/test_real_nfft3_pin.py); everything above that boundary is validated against the reference's own sources
 * compiled into oracle/_ref (see oracle/README.md).
 *
 * Self-contained CPU restatement of the hot path of the reference:
 *   - window setup and NFFT additive matvec / gradient matvec
 *       SRC/external/nfft_interface.c:129-263, 265-398, 400-497, 499-620, 622-840
 *   - PCG                         SRC/solvers/pcg.c:3-206
 *   - Nystrom / FSAI / AFN apply   SRC/preconds/nys.c:115-173, fsai.c:106-123 (+ matops.c:139-272),
 *                                  afn.c:109-136 (design only; MATLAB afn_solve.m:39-69)
 *   - vector ops                   SRC/linearalg/vecops.c:3-155
 */
#ifndef ORACLE_PATH_H
#define ORACLE_PATH_H
#ifdef __cplusplus
extern "C" {
#endif

typedef int (*oracle_symmatvec)(void *mat, int n, double alpha, double *x, double beta, double *y);
typedef int (*oracle_solve)(void *prec, int n, double *x, double *rhs);

/* ---- additive NFFT kernel ---- */
void *oracle_additive_create(const double *data, int n, int ldim, int d, const int *windows, int nwindows, int dwindows);
/* kernel: 0 Gaussian, 1 Matern-1/2.  Returns 0 / -1. */
int oracle_additive_build(void *h, int kernel, double f, double l, double mu);
int oracle_additive_matvec(void *h, int n, double alpha, double *x, double beta, double *y);
int oracle_additive_grad_matvec(void *h, int n, double alpha, double *x, double beta, double *y);
void oracle_additive_free(void *h);
/* introspection: per-window scale factor and dimension; scaled point-major coordinates */
int oracle_additive_nwindows(void *h);
int oracle_additive_window_dim(void *h, int w);
double oracle_additive_window_scale(void *h, int w);
const double *oracle_additive_window_points(void *h, int w);
/* kernel Fourier coefficients b_k (complex interleaved, N^d) of window w; which = 0 kernel, 1 derivative kernel */
const double *oracle_additive_window_bhat(void *h, int w, int which);

/* ---- PCG (pcg.c:3-206). rel_res_v must hold min(maxits,n)+1 doubles. ---- */
int oracle_pcg(void *mat, int n, oracle_symmatvec matvec, void *prec, oracle_solve precond,
               double *x, double *rhs, int maxits, int atol, double tol,
               double *prel_res, double *rel_res_v, int *piter);

/* ---- preconditioner applies on plain arrays ---- */
typedef struct { int n, k; const int *perm; const double *U; const double *s; double eta; double *work; } oracle_nys;
int oracle_nys_solve(void *nys, int n, double *x, double *rhs);
typedef struct { int n; const int *ia; const int *ja; const double *a; double *work; } oracle_fsai;
int oracle_fsai_solve(void *fsai, int n, double *x, double *rhs);
/* AFN 2x2 block apply: perm (n), L (k x k lower Cholesky factor of K11, column-major),
 * K12 (k x (n-k) column-major), FSAI factor G of the Schur complement on n-k points. */
typedef struct { int n, k; const int *perm; const double *L; const double *K12; oracle_fsai schur; double *work; } oracle_afn;
int oracle_afn_solve(void *afn, int n, double *x, double *rhs);

/* ---- dense additive kernel matrix (kernels.c:3099-3494 restated for the full-matrix call shape) ----
 * K (n x n column-major, full symmetric) = f^2 * ( mean_w k_w(x_i, x_j) + mu * delta_ij ) */
int oracle_dense_additive(const double *data, int n, int ldim, int d, const int *windows, int nwindows, int dwindows,
                          int kernel, double f, double l, double mu, double *K);
int oracle_dense_symv(const double *K, int n, double alpha, const double *x, double beta, double *y);

int oracle_num_threads(void);
void oracle_set_num_threads(int nthreads);

#ifdef __cplusplus
}
#endif
#endif
