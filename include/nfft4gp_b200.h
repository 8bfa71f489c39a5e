/*
 * nfft4gp_b200.h -- C ABI of libnfft4gp_b200.so, the B200-native (sm_100a) replacement for the
 * hot path of the nfft4gp C library: the NFFT-accelerated additive-kernel matrix-vector product
 * and the preconditioned CG solve built on it.
 *
 * Every entry point in part 1 has EXACTLY the name, argument meaning, return convention
 * (0 = ok, -1 = error + printf message) and ownership rules of the reference function it replaces;
 * the reference file:line is cited next to each declaration (paths relative to the reference
 * repository).  A maintainer of the reference swaps the objects listed in INTEGRATION.md for this
 * library and keeps every caller (solvers, GP optimizer, TESTS drivers) unchanged.
 *
 * Part 2 ("Nfft4GPB200*") are additive extensions: device selection, device-resident vectors,
 * factor upload for the preconditioners, multi-GPU sharding, and debug taps used by the parity tests.
 *
 * Plain C: pointers and sizes only.  All floating point is fp64, all indices are 32-bit int
 * (reference: NFFT4GP_DOUBLE = double, SRC/utils/utils.h:28-32).
 */
#ifndef NFFT4GP_B200_H
#define NFFT4GP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef NFFT4GP_DOUBLE
#define NFFT4GP_DOUBLE double
#endif

/* ------------------------------------------------------------------------------------------
 * Part 0: the plug-point types (INC/_solvers.h:23,51; INC/_linearalg.h:404; INC/_utils.h:430-438)
 * ------------------------------------------------------------------------------------------ */
typedef int (*func_solve)(void *precond, int n, double *x, double *rhs);                                  /* _solvers.h:23  */
typedef int (*func_symmatvec)(void *matrix, int n, double alpha, double *x, double beta, double *y);      /* _solvers.h:51  */
typedef int (*func_kernel)(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                           int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);                /* _linearalg.h:404 */
typedef int (*func_trace)(void *str, NFFT4GP_DOUBLE **tracesp);                                          /* _solvers.h:59  */
typedef NFFT4GP_DOUBLE (*func_logdet)(void *str);                                                        /* _solvers.h:67  */
typedef int (*func_dvp)(void *str, int n, int *mask, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE **yp);             /* _solvers.h:80  */
typedef void *(*func_create)(void);                                                                      /* _utils.h:430   */
typedef void (*func_free)(void *str);                                                                    /* _utils.h:438   */
typedef int (*precond_kernel_setup)(NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel,
                                    void *fkernel_params, int require_grad, void *precond);              /* _preconds.h:41-42 */

/* Generic kernel handle.  The callers of the reference write `_params[0]` (f), `_params[1]` (l) and
 * `_noise_level` (mu) DIRECTLY before every kernel build (SRC/optimizer/gp_loss.c:145-150,
 * TESTS/TEST1/foo.cpp:222-224), so the field order and types are part of the ABI
 * (INC/_linearalg.h:420-450).  `_external` points at this library's private state. */
#ifndef NFFT4GP_KERNEL_MAX_PARAMS
#define NFFT4GP_KERNEL_MAX_PARAMS 5
#endif
typedef struct NFFT4GP_KERNEL_STRUCT
{
   NFFT4GP_DOUBLE _params[NFFT4GP_KERNEL_MAX_PARAMS];
   int _iparams[NFFT4GP_KERNEL_MAX_PARAMS];
   int _max_n;
   int _omp;
   NFFT4GP_DOUBLE _noise_level;
   int _own_buffer;
   NFFT4GP_DOUBLE *_buffer;
   int _own_dbuffer;
   NFFT4GP_DOUBLE *_dbuffer;
   func_kernel _fkernel_buffer;
   int **_ibufferp;
   size_t *_libufferp;
   int _own_fkernel_buffer_params;
   void *_fkernel_buffer_params;
   size_t _ldwork;
   NFFT4GP_DOUBLE *_dwork;
   void *_external;
} nfft4gp_kernel, *pnfft4gp_kernel;

/* ------------------------------------------------------------------------------------------
 * Part 1a: NFFT interface (INC/_external.h, SRC/external/nfft_interface.c)
 * ------------------------------------------------------------------------------------------ */

/* nfft_interface.c:622-674 (_external.h:176).  `data` is feature-major (d columns of ldim values);
 * `windows` holds nwindows*dwindows feature ids read flat, -1 pads a short LAST window.  The handle
 * keeps the caller's `windows` pointer (not copied, not freed) and a private copy of the columns. */
void *Nfft4GPNFFTAdditiveKernelParamCreate(NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *windows, int nwindows, int dwindows);

/* nfft_interface.c:676-734 / 736-794 (_external.h:198,220): func_kernel-shaped plan (re)build.  Reads f, l, mu
 * from the handle; requires n == ldim and non-NULL Kp, dKp; ignores permr/permc; returns the handle itself in
 * *Kp and *dKp (freed by the no-op Nfft4GPNFFTKernelFree). */
int Nfft4GPNFFTAdditiveKernelGaussianKernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                            int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);
int Nfft4GPNFFTAdditiveKernelMatern12Kernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                            int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);

/* nfft_interface.c:796-817 (_external.h:233): y <- beta*y + alpha*f^2*((1/P) sum_w K_w + mu I) x.  x, y host, length n. */
int Nfft4GPAdditiveNFFTMatSymv(void *data, int n, NFFT4GP_DOUBLE alpha, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE beta, NFFT4GP_DOUBLE *y);
/* nfft_interface.c:819-840 (_external.h:246): y (3n) = [dK/df x ; dK/dl x ; dK/dmu x]. */
int Nfft4GPAdditiveNFFTGradMatSymv(void *data, int n, NFFT4GP_DOUBLE alpha, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE beta, NFFT4GP_DOUBLE *y);
/* nfft_interface.c:842-856 (_external.h:254) and :79-82 (_external.h:75, no-op). */
void Nfft4GPAdditiveNFFTKernelFree(void *str);
void Nfft4GPNFFTKernelFree(void *str);

/* single-window equivalents: nfft_interface.c:3-42, 44-77, 129-263, 265-398, 400-497, 499-620 (_external.h:60-161).
 * `data` for the kernel builders is the window's own feature-major n x d block; Mat(Grad)Symv take the
 * pointer returned in *Kp / *dKp. */
void *Nfft4GPNFFTKernelParamCreate(int max_n, int dim);
void Nfft4GPNFFTKernelParamFree(void *kernel);
int Nfft4GPNFFTKernelGaussianKernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                    int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);
int Nfft4GPNFFTKernelMatern12Kernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                    int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);
int Nfft4GPNFFTMatSymv(void *data, int n, NFFT4GP_DOUBLE alpha, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE beta, NFFT4GP_DOUBLE *y);
int Nfft4GPNFFTGradMatSymv(void *data, int n, NFFT4GP_DOUBLE alpha, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE beta, NFFT4GP_DOUBLE *y);

/* Preconditioned Lanczos solver (INC/_solvers.h:176, SRC/solvers/lanczos.c:3-419): solves A x = rhs and returns the Lanczos
 * tridiagonal matrix T (*TDp diagonal, *TEp sub-diagonal, *tsize rows; NULL on entry = callee allocates with malloc), which
 * keeps being extended up to wsize rows after convergence.  Bases, orthogonalisation (Nfft4GPModifiedGS2, matops.c:348-434) and
 * all vector operations on the GPU; single GPU. */
int Nfft4GPSolverLanczos(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc,
                         NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs, int wsize, int maxits, int atol, NFFT4GP_DOUBLE tol,
                         NFFT4GP_DOUBLE *prel_res, NFFT4GP_DOUBLE **prel_res_v, int *piter, int *tsize, NFFT4GP_DOUBLE **TDp,
                         NFFT4GP_DOUBLE **TEp, int print_level);
/* Stochastic Lanczos quadrature (INC/_solvers.h:216, lanczos.c:421-610): *logdet = log det(A) / n and (*dlogdetp)[3] = its
 * derivatives w.r.t. (f, l, mu) / n from nvecs Rademacher probes (radamacher: n x nvecs given probes, or NULL = rand() as
 * vecops.c:27-45).  dmatvec = the gradient matvec (y of length 3n).  tracefunc / logdetfunc / dvpfunc are only called when
 * prec_data != NULL (they belong to the preconditioner; this library exports none). */
int Nfft4GPLanczosQuadratureLogdet(void *mat_data, void *dmat_data, int n, func_symmatvec matvec, func_symmatvec dmatvec,
                                   void *prec_data, func_solve precondfunc, func_trace tracefunc, func_logdet logdetfunc,
                                   func_dvp dvpfunc, int maxits, int nvecs, NFFT4GP_DOUBLE *radamacher, int print_level,
                                   NFFT4GP_DOUBLE *logdet, NFFT4GP_DOUBLE **dlogdetp);

/* nfft_interface.c:84-128 (_external.h): single-window handle: forget the points / the points and the plans */
int Nfft4GPNFFTKernelParamRemovePoints(void *kernel);
int Nfft4GPNFFTKernelParamFreeNFFTKernel(void *kernel);

/* nfft_interface.c:858-871 (_external.h:260): [X1; X2] feature-major, malloc'ed */
NFFT4GP_DOUBLE *Nfft4GPNFFTAppendData(NFFT4GP_DOUBLE *X1, int n1, int ldim1, int d, NFFT4GP_DOUBLE *X2, int n2, int ldim2);

/* SRC/optimizer/transform.h:18-23 */
#ifndef NFFT4GP_TRANSFORM_TYPE_DEFINED
#define NFFT4GP_TRANSFORM_TYPE_DEFINED
typedef enum
{
   NFFT4GP_TRANSFORM_SOFTPLUS = 0,
   NFFT4GP_TRANSFORM_SIGMOID,
   NFFT4GP_TRANSFORM_EXP,
   NFFT4GP_TRANSFORM_IDENTITY
} nfft4gp_transform_type;
#endif

/* nfft_interface.c:873-1068 (_external.h:266): GP prediction with the NFFT additive kernel.  x[3] = untransformed (f, l, mu);
 * vfkernel_data = NFFT handle on the n training points, vfkernel_data_l = NFFT handle on data_all = [train; test];
 * mean through one FGMRES solve (kdim = maxits) and one matvec on the joint handle; *std_predictp (optional, may be NULL) through
 * one solve per test point.  *label_predictp / *std_predictp: NULL = callee allocates with malloc. */
int Nfft4GPAdditiveNFFTGpPredict(NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *data, NFFT4GP_DOUBLE *label, int n, int ldim, int d,
                                 NFFT4GP_DOUBLE *data_predict, int n_predict, int ldim_predict, NFFT4GP_DOUBLE *data_all,
                                 func_kernel fkernel, void *vfkernel_data, void *vfkernel_data_l, func_free kernel_data_free,
                                 func_symmatvec matvec, func_kernel precond_fkernel, void *precond_vfkernel_data,
                                 func_free precond_kernel_data_free, precond_kernel_setup precond_setup, func_solve precond_solve,
                                 void *precond_data, int atol, NFFT4GP_DOUBLE tol, int maxits, nfft4gp_transform_type transform,
                                 int print_level, NFFT4GP_DOUBLE *dwork, NFFT4GP_DOUBLE **label_predictp, NFFT4GP_DOUBLE **std_predictp);

/* ------------------------------------------------------------------------------------------
 * Part 1b: Krylov solver (INC/_solvers.h:102, SRC/solvers/pcg.c:3-206)
 * Same control flow and quirks as the reference (true-residual recheck with an extra matvec, *piter = 0
 * unless convergence is confirmed, rel_res_v[ii] overwritten with the ABSOLUTE true residual at the
 * recheck; *prel_res_v is callee-allocated with malloc, caller frees).  When `matvec` / `precondfunc` are
 * this library's own entry points the whole iteration runs device-resident with fused vector kernels;
 * any other function pointers are called on host vectors exactly as the reference does.
 * ------------------------------------------------------------------------------------------ */
int Nfft4GPSolverPcg(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc,
                     NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs, int maxits, int atol, NFFT4GP_DOUBLE tol,
                     NFFT4GP_DOUBLE *prel_res, NFFT4GP_DOUBLE **prel_res_v, int *piter, int print_level);

/* Flexible GMRES with restarts (INC/_solvers.h:137, SRC/solvers/fgmres.c:3-252) -- the solver the reference's GP loss,
 * prediction and TESTS drivers call.  Same control flow and quirks (restart scaled by the estimated residual norm, Givens
 * breakdown leaves without updating x, the end-of-cycle message is printed at print_level == 0, *prel_res_v has
 * maxits + 1 entries).  Krylov bases, Arnoldi step (modified Gram-Schmidt, no re-orthogonalisation) and all vector
 * operations run on the GPU; foreign function pointers are honoured as in Nfft4GPSolverPcg. */
int Nfft4GPSolverFgmres(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc,
                        NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs, int kdim, int maxits, int atol, NFFT4GP_DOUBLE tol,
                        NFFT4GP_DOUBLE *prel_res, NFFT4GP_DOUBLE **prel_res_v, int *piter, int print_level);

/* ------------------------------------------------------------------------------------------
 * Part 1c: preconditioners (INC/_preconds.h; SRC/preconds/nys.c, fsai.c, afn.c/afn.h)
 * Handles are opaque here (the reference's callers only use the functions).
 * ------------------------------------------------------------------------------------------ */
void *Nfft4GPPrecondNysCreate(void);                                     /* nys.c:3-33    (_preconds.h:358) */
void Nfft4GPPrecondNysFree(void *str);                                   /* nys.c:35-76   (:365) */
void Nfft4GPPrecondNysReset(void *str);                                  /* nys.c:78-98   (:372) */
void Nfft4GPPrecondNysSetRank(void *str, int k);                         /* nys.c:100-105 (:380) */
void Nfft4GPPrecondNysSetPerm(void *str, int *perm, int own_perm);       /* nys.c:107-113 (:389) */
int Nfft4GPPrecondNysSolve(void *vnys_mat, int n, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs);   /* nys.c:115-173 (:400) */
int Nfft4GPPrecondNysSetupWithKernel(NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel,
                                     void *fkernel_params, int require_grad, void *vnys_mat); /* nys.c:518-660 (:444) */
/* gradient side (what TESTS/TEST3/foo3.cpp hands to the GP problem); Dvp / Trace need a setup with require_grad != 0 and one
 * of this library's dense kernels.  The trace is formed from two k x k Gram matrices instead of k products of length n. */
int Nfft4GPPrecondNysDvp(void *vnys_mat, int n, int *mask, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE **yp); /* nys.c:175-330 (:413) */
int Nfft4GPPrecondNysTrace(void *vnys_mat, NFFT4GP_DOUBLE **tracesp);                          /* nys.c:332-484 (:421) */
NFFT4GP_DOUBLE Nfft4GPPrecondNysLogdet(void *vnys_mat);                                        /* nys.c:486-516 (:429) */

void *Nfft4GPPrecondFsaiCreate(void);                                    /* fsai.c:7-30    (_preconds.h:187) */
void Nfft4GPPrecondFsaiFree(void *str);                                  /* fsai.c:32-66   (:194) */
void Nfft4GPPrecondFsaiReset(void *str);                                 /* fsai.c:68-98   (:201) */
void Nfft4GPPrecondFsaiSetLfil(void *str, int lfil);                     /* fsai.c:100-104 (:209) */
int Nfft4GPPrecondFsaiSolve(void *vfsai_mat, int n, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs); /* fsai.c:106-123 (:220) */
int Nfft4GPPrecondFsaiSetupWithKernel(NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel,
                                      void *fkernel_params, int require_grad, void *vfsai_mat); /* fsai.c:302-673 (:264) */
NFFT4GP_DOUBLE Nfft4GPPrecondFsaiLogdet(void *vfsai_mat);                /* fsai.c:278-300 (:246) */
/* gradient side (set up with require_grad != 0): derivative factors, trace, derivative-vector products, and the two sparse
 * triangular solves with the factor -- synchronisation-free on the GPU (rows wait on the flags of the rows they depend on) */
int Nfft4GPPrecondFsaiDvp(void *vfsai_mat, int n, int *mask, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE **yp);  /* fsai.c:125-219 (:233) */
int Nfft4GPPrecondFsaiTrace(void *vfsai_mat, NFFT4GP_DOUBLE **tracesp);                             /* fsai.c:221-276 (:241) */
int Nfft4GPPrecondFsaiSetupWithKernelPattern(int *A_i, int *A_j, NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel,
                                             void *fkernel_params, int require_grad, void *vfsai_mat); /* fsai.c:314-673 (:282) */
int Nfft4GPPrecondFsaiInvL(void *vfsai_mat, int n, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs);          /* fsai.c:675-698 (:295) */
int Nfft4GPPrecondFsaiInvLT(void *vfsai_mat, int n, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs);         /* fsai.c:700-721 (:307) */
/* read the derivative factor values back (parity tests): dG/df, dG/dl, dG/dmu on the pattern of G, 3 * nnz values */
int Nfft4GPB200PrecondFsaiGetGradFactors(void *fsai, double *da);

/* Dense Cholesky (the exact solver of the small-n experiments), SRC/preconds/chol.c; with require_grad the setup keeps
 * K and dK/dl in HBM and Trace / Dvp / Logdet are available (the reference's Dvp dereferences a NULL mask copy when the
 * caller passes a mask, chol.c:145-150; here a given mask is honoured). */
void *Nfft4GPPrecondCholCreate(void);                                    /* chol.c:8-27    (_preconds.h:79) */
void Nfft4GPPrecondCholFree(void *str);                                  /* chol.c:29-64   (:86) */
void Nfft4GPPrecondCholReset(void *str);                                 /* chol.c:66-99   (:93) */
void Nfft4GPPrecondCholSetStable(void *str, int stable);                 /* chol.c:101-109 (:100) */
int Nfft4GPPrecondCholSolve(void *vchol_mat, int n, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs); /* chol.c:111-136 (:111) */
int Nfft4GPPrecondCholDvp(void *vchol_mat, int n, int *mask, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE **yp); /* chol.c:138-229 (:124) */
int Nfft4GPPrecondCholTrace(void *vchol_mat, NFFT4GP_DOUBLE **tracesp);  /* chol.c:231-291 (:132) */
NFFT4GP_DOUBLE Nfft4GPPrecondCholLogdet(void *vchol_mat);                /* chol.c:293-311 (:140) */
int Nfft4GPPrecondCholSetupWithKernel(NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel,
                                      void *fkernel_params, int require_grad, void *vchol_mat); /* chol.c:428-640 (:156) */

/* AFN: SRC/preconds/afn.c is not compiled by the reference (SURVEY.md finding 2); names follow afn.h:65-132,
 * the algorithm follows MATLAB afn_setup.m:30-109 / afn_solve.m:39-69. */
void *Nfft4GPPrecondAFNCreate(void);
void Nfft4GPPrecondAFNFree(void *str);
void Nfft4GPPrecondAFNSetRank(void *str, int max_k);
void Nfft4GPPrecondAFNSetLfil(void *str, int lfil);
int Nfft4GPPrecondAFNSolve(void *vafn_mat, int n, NFFT4GP_DOUBLE *x, NFFT4GP_DOUBLE *rhs);
/* afn.h:99 / MATLAB afn_logdet.m: log det M = 2 (sum log L_ii + sum log 1 / G_ii); the Nystrom form delegates to Nfft4GPPrecondNysLogdet */
NFFT4GP_DOUBLE Nfft4GPPrecondAFNLogdet(void *vafn_mat);
/* afn.h:91 / MATLAB afn_trace.m: tr(M^{-1} dM / dtheta).  Available when the setup chose the Nystrom form (rank below the maximum,
 * set up with require_grad; delegates to Nfft4GPPrecondNysTrace); in the block form only the f component has a closed form (2 n / f)
 * and the call returns -1 for the others (the derivative of the Schur-complement FSAI rows is not built). */
int Nfft4GPPrecondAFNTrace(void *vafn_mat, NFFT4GP_DOUBLE **tracesp);
int Nfft4GPPrecondAFNSetupWithKernel(NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel,
                                     void *fkernel_params, int require_grad, void *vafn_mat);

/* Dense additive kernel used by the preconditioner setups (SRC/linearalg/kernels.c:3046-3494; _linearalg.h:717,740).
 * Same func_kernel contract: full lower (permr NULL), K(permr,permr) lower, K(permr,permc) full, column-major;
 * dKp (optional) receives the three gradient blocks stacked [d/df; d/dl; d/dmu] in the same shape. */
void *Nfft4GPKernelAdditiveKernelParamCreate(NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *windows, int nwindows,
                                             int dwindows, func_kernel fkernel);
int Nfft4GPKernelAdditiveKernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);
int Nfft4GPKernelGaussianKernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);   /* kernels.c:680  */
int Nfft4GPKernelMatern12Kernel(void *str, NFFT4GP_DOUBLE *data, int n, int ldim, int d, int *permr, int kr,
                                int *permc, int kc, NFFT4GP_DOUBLE **Kp, NFFT4GP_DOUBLE **dKp);   /* kernels.c:2390 */
void *Nfft4GPKernelParamCreate(int max_n, int omp);                                               /* kernels.c:404 (_linearalg.h:459) */
void Nfft4GPKernelParamFree(void *str);                                                           /* kernels.c:443 (_linearalg.h:467) */
void Nfft4GPKernelFree(void *str);                                                                /* func_free of *Kp (_linearalg.h:475) */

/* ------------------------------------------------------------------------------------------
 * Part 2: B200 extensions
 * ------------------------------------------------------------------------------------------ */

/* Select the CUDA device (default: device 0, or LOCAL_RANK when set).  Fails loudly (-1 + message) when no
 * sm_100 device is visible: there is NO CPU fallback anywhere in this library. */
int Nfft4GPB200Init(int device);
int Nfft4GPB200DeviceCount(void);
const char *Nfft4GPB200Version(void);
/* number of kernels this library has launched since load / since the last reset (bench.py: gpu_launches) */
long long Nfft4GPB200LaunchCount(void);
void Nfft4GPB200LaunchCountReset(void);

/* pinned host and device memory helpers for callers that want device-resident vectors */
void *Nfft4GPB200HostAlloc(size_t bytes);
void Nfft4GPB200HostFree(void *p);
void *Nfft4GPB200DeviceAlloc(size_t bytes);
void Nfft4GPB200DeviceFree(void *p);
int Nfft4GPB200MemcpyH2D(void *dst_dev, const void *src_host, size_t bytes);
int Nfft4GPB200MemcpyD2H(void *dst_host, const void *src_dev, size_t bytes);
int Nfft4GPB200DeviceSynchronize(void);
/* CUDA-event timing on the library's stream (the stream every kernel of this library is launched on) */
void *Nfft4GPB200EventCreate(void);
void Nfft4GPB200EventDestroy(void *ev);
int Nfft4GPB200EventRecord(void *ev);
float Nfft4GPB200EventElapsedMs(void *start, void *stop); /* synchronises on `stop` */

/* Per-kernel-class timing with CUDA events recorded on the library stream around every launch (bench.py roofline).
 * classes: 0 prep, 1 spread, 2 fft forward, 3 fft inverse(+scale), 4 gather, 5 combine, 6 vector ops, 7 Nystrom,
 * 8 CSR SpMV, 9 GEMV, 10 other.  Stop synchronises, fills sums_ms[ncls] / counts[ncls], returns the region count. */
void Nfft4GPB200ProfileStart(void);
int Nfft4GPB200ProfileStop(double *sums_ms, long long *counts, int ncls);

/* Device-resident variants of the plug points: same semantics, x / y / rhs are DEVICE pointers. */
int Nfft4GPB200AdditiveNFFTMatSymvDevice(void *data, int n, double alpha, const double *d_x, double beta, double *d_y);
int Nfft4GPB200AdditiveNFFTGradMatSymvDevice(void *data, int n, double alpha, const double *d_x, double beta, double *d_y);
int Nfft4GPB200PrecondSolveDevice(void *prec, func_solve which, int n, double *d_x, const double *d_rhs);
/* PCG with device-resident x (in: initial guess, out: solution) and rhs. Other arguments as Nfft4GPSolverPcg. */
int Nfft4GPB200SolverPcgDevice(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc,
                               double *d_x, const double *d_rhs, int maxits, int atol, double tol,
                               double *prel_res, double **prel_res_v, int *piter, int print_level);

int Nfft4GPB200SolverFgmresDevice(void *mat_data, int n, func_symmatvec matvec, void *prec_data, func_solve precondfunc,
                                  double *d_x, const double *d_rhs, int kdim, int maxits, int atol, double tol,
                                  double *prel_res, double **prel_res_v, int *piter, int print_level);

/* Upload precomputed factors (e.g. produced by the reference's CPU setup) instead of running the setup here. */
int Nfft4GPB200PrecondNysSetFactors(void *nys, int n, int k, const int *perm, const double *U /* n x k col-major */,
                                    const double *s /* k */, double eta);
int Nfft4GPB200PrecondFsaiSetFactors(void *fsai, int n, const int *ia, const int *ja, const double *a);
int Nfft4GPB200PrecondAFNSetFactors(void *afn, int n, int k, const int *perm, const double *L /* k x k lower, col-major */,
                                    const double *K12 /* k x (n-k) col-major */, const int *ia, const int *ja, const double *a);
/* read factors back (parity tests of the setups); any output pointer may be NULL */
int Nfft4GPB200PrecondNysGetFactors(void *nys, int *n, int *k, int *perm, double *U, double *s, double *eta);
int Nfft4GPB200PrecondFsaiGetFactors(void *fsai, int *n, int *nnz, int *ia, int *ja, double *a);
int Nfft4GPB200PrecondAFNGetFactors(void *afn, int *n, int *k, int *k_est, int *is_nys, int *perm, double *Linv /* k x k */,
                                    double *K12 /* k x (n-k) */, int *nnz, int *ia, int *ja, double *a);
/* AFN rank estimation (afn_setup.m:111-293): m sampled points (default 500), numests repetitions (default 5; 0 = skip the
 * estimation and use the maximum rank), noran != 0 = never switch to the Nystrom form, seed of the sub-sampling generator */
int Nfft4GPB200PrecondAFNSetRankEstimation(void *afn, int m, int numests, int noran, unsigned long long seed);
/* first stage of the rank estimation: 0 (default) = MATLAB afn_setup.m:163-228 (bisection on the 2-norm Nystrom error),
 * 1 = the reference's compiled Nfft4GPRankestNysScaled (SRC/linearalg/rankest.c:354-391: 50-point scan, Frobenius norm) */
int Nfft4GPB200PrecondAFNSetRankMethod(void *afn, int method);
/* One sample of Nfft4GPRankestNysScaledEstimateRank (rankest.c:248-352) on given sample rows (the reference draws them with
 * rand()); fkernel must be one of this library's dense kernels.  Returns the rank scaled to n, or -1. */
int Nfft4GPB200RankestNysScaledSample(NFFT4GP_DOUBLE *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params,
                                      const int *rows, int m, int max_rank);
/* Storage of the AFN K12 block (k x (n-k), read twice by every apply and the dominant HBM traffic of a PCG iteration at
 * large rank): bits = 64 (default) or 32.  With 32 the block is kept in fp32 (k a multiple of 4), vectors and sums stay fp64
 * and the apply stays symmetric; the preconditioner changes by O(1e-7) relative, the solution PCG converges to does not. */
int Nfft4GPB200PrecondAFNSetK12Storage(void *afn, int bits);
/* The two K12 products of every AFN apply through the NFFT operator instead of the dense k x (n-k) block: K21 y and K12 y2 are
 * rows of K applied to vectors that vanish on the other block, i.e. two fast O(n) matvecs (afn.h:27-58 keeps A12 behind a function
 * pointer, _fA12_matvec).  nfft_kernel: the *Kp of an additive NFFT kernel built on the same data and hyper-parameters (NULL: back to
 * the dense block); release_block != 0 frees the dense block after the setup.  fp64 throughout; K12 is then the NFFT approximation
 * of the kernel block (3e-8 relative at a resolved length scale) -- the preconditioner changes at that level, the solution PCG
 * converges to does not.  Default: dense block (the reference's form). */
int Nfft4GPB200PrecondAFNSetK12Operator(void *afn, void *nfft_kernel, int release_block);
/* setup building blocks, exposed for the parity tests (host in / host out; the device ones need a GPU) */
int Nfft4GPB200DebugPartitionChunks(const int *chunks /* int4 per chunk */, int nchunk, int wcnt, int rounds, const double *weight,
                                    double flush, int *first /* rounds * wcnt + 1 */);
/* plan build of the table-free 1-D path (host only): cell sort inside blocks of T consecutive points of the original order */
int Nfft4GPB200DebugPoly1dBlockSort(const double *xt, int n, int T, double *out_xt /* nblk*T */, unsigned short *out_lp /* nblk*T */,
                                    unsigned short *out_rs /* nblk*80: 65 run starts per block, rows padded to 160 bytes */);
int Nfft4GPB200HostPotrf(int n, double *A /* lower, column-major, in place */);
int Nfft4GPB200HostSyev(int n, double *A /* in: symmetric; out: eigenvectors */, double *w /* ascending */, int want_vectors);
int Nfft4GPB200HostStev(int n, double *d /* diagonal -> eigenvalues */, double *e /* sub-diagonal, destroyed */, double *Z /* n x n */);
int Nfft4GPB200Dgemm(char transa, char transb, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                     double beta, double *C, int ldc);
int Nfft4GPB200Fps(const double *data, int n, int ldim, int d, int k, int *perm_out /* k */);
int Nfft4GPB200KnnPattern(const double *data, int n, int ldim, int d, int lfil, int *ia /* n+1 */, int *ja /* nnz */);

/* Multi-GPU (one process per GPU).  The NCCL communicator is created from a unique id that the caller
 * broadcasts (128 bytes).  mode 0: replicate; 1: shard by feature window, allreduce(sum) of y per matvec;
 * 2: shard by point block, allreduce(sum) of the oversampled grids per matvec; x and y stay sharded: the device entry points
 *    take the rank's block, and the host-pointer calls read / fill the entries [n r / R, n (r + 1) / R) of the full-length
 *    host vectors on rank r (the caller assembles y across ranks if it needs all of it on one rank). */
int Nfft4GPB200CommGetUniqueId(void *id128);
int Nfft4GPB200CommInit(int rank, int nranks, const void *id128);
int Nfft4GPB200CommFinalize(void);
int Nfft4GPB200SetSharding(void *additive_handle, int mode);
/* Stochastic Lanczos quadrature on several GPUs (BASELINE configs[4]): with on != 0, Nfft4GPLanczosQuadratureLogdet distributes its
 * probe vectors over the ranks (probe i on rank i mod R; every rank holds the whole, unsharded operator) and completes the four
 * sums with one allreduce of four doubles.  Same result as a single-process run, including the rand() stream of the probes. */
int Nfft4GPB200SetProbeSharding(int on);

/* fp64 tensor-core (mma.sync.m8n8k4.f64 = DMMA.8x8x4) throughput of the current GPU measured now, in TFLOP/s (a 2 ms kernel with
 * register operands): the roof bench.py reports the 3-D spread / gather kernels against.  Negative on error. */
double Nfft4GPB200MeasureFp64TensorPeak(void);

/* Debug taps for the stage-by-stage parity tests (host output buffers). */
int Nfft4GPB200DebugWindowInfo(void *additive_handle, int w, int *dim, double *scale, int *nseg);
int Nfft4GPB200DebugSpread(void *additive_handle, int w, double alpha, const double *x, double *grid /* 64^dim */);
int Nfft4GPB200DebugCoeff(void *additive_handle, int w, int which, double *c /* 2 * 32^dim: b_k/phi_hat^2 interleaved */);
int Nfft4GPB200DebugGridAfterFFT(void *additive_handle, int w, double alpha, const double *x, double *h /* 64^dim */);
/* tuning knobs (bench / A-B tests): name in {"spread3d","gather3d"}; variant -1 = automatic (default), 0 = generic
 * thread-per-point, 1 = DFMA warp-per-range with register-tile shifting, 5 = fp64 tensor-core (DMMA) warp-per-range */
int Nfft4GPB200SetVariant(const char *name, int variant);

#ifdef __cplusplus
}
#endif
#endif /* NFFT4GP_B200_H */
