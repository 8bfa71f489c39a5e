/* ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the part of NFFT3 applications/fastsum/fastsum.h that the
 * reference binds (SRC/external/nfft_interface.c:223-254, 359-390 plan build;
 * :426, :533-534 execute; :56-59, 93-98, 110-113 teardown; public fields touched:
 * x, y, alpha, f, kernel_param at :248-251, 412, 436, 536).
 *
 * NFFT3 is a third-party dependency absent from /root/reference and from this
 * image, un-pinned by the reference (API evidence: >= 3.5.2).  This file restates
 * its published algorithm (Potts/Steidl fastsum on top of the Keiner/Kunis/Potts
 * NFFT with the default Kaiser-Bessel window).  The reference holds no golden vector AT this boundary; parity is pinned
 * one level up, on the 400 numbers it ships from its authors' real NFFT3 build (TESTS/TEST2/foo_0_nfft.txt, reproduced
 * to the printed precision: tests/test_real_nfft3_pin.py), and arbitrated by the reference's dense additive kernel
 * (see oracle/README.md).
 */
#ifndef ORACLE_NFFT3_SHIM_FASTSUM_H
#define ORACLE_NFFT3_SHIM_FASTSUM_H

#include "config.h"
#include <complex.h>
#include "nfft3.h"
#include "infft.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef C (*kernel)(R, int, const R *);

#define EXACT_NEARFIELD (1U << 0)
#define NEARFIELD_BOXES (1U << 1)
#define STORE_PERMUTATION_X_ALPHA (1U << 2)

/* One NFFT plan (node set + window tables) of the shim. */
typedef struct
{
   int M;        /* number of nodes                                   */
   R *psi;       /* PRE_PSI table, [M][d][2m+2]                        */
   int *cell;    /* floor(n_os * x) per node and dimension, [M][d]     */
   int ready;
} shim_nfft_nodes;

typedef struct fastsum_plan_
{
   /* --- public fields (same names as NFFT3) --- */
   int d;            /* spatial dimension                              */
   int N_total;      /* number of source nodes                         */
   int M_total;      /* number of target nodes                         */
   C *alpha;         /* source coefficients                            */
   C *f;             /* target evaluations                             */
   R *x;             /* source nodes, point-major N_total x d          */
   R *y;             /* target nodes, point-major M_total x d          */
   kernel k;         /* radial kernel                                  */
   R *kernel_param;  /* its parameters                                 */
   unsigned flags;
   int n;            /* expansion degree (N of the reference, 32)      */
   C *b;             /* Fourier coefficients of the kernel, n^d        */
   C *f_hat;         /* n^d                                            */
   int p;
   R eps_I;
   R eps_B;
   /* --- shim-private --- */
   int n_os_src, m_src, n_os_tgt, m_tgt; /* oversampled FFT length, cut-off */
   shim_nfft_nodes src, tgt;
   C *g;             /* oversampled grid work array, n_os^d            */
   int b_ready;
} fastsum_plan;

void fastsum_init_guru_kernel(fastsum_plan *ths, int d, kernel k, R *param,
                              unsigned flags, int nn, int p, R eps_I, R eps_B);
void fastsum_init_guru_source_nodes(fastsum_plan *ths, int N_total, int nn_oversampled, int m);
void fastsum_init_guru_target_nodes(fastsum_plan *ths, int M_total, int nn_oversampled, int m);
void fastsum_precompute_kernel(fastsum_plan *ths);
void fastsum_precompute_source_nodes(fastsum_plan *ths);
void fastsum_precompute_target_nodes(fastsum_plan *ths);
void fastsum_precompute(fastsum_plan *ths);
void fastsum_trafo(fastsum_plan *ths);
void fastsum_finalize_source_nodes(fastsum_plan *ths);
void fastsum_finalize_target_nodes(fastsum_plan *ths);
void fastsum_finalize_kernel(fastsum_plan *ths);

/* shim introspection used only by the parity tests: raw access to the
 * Kaiser-Bessel window and to the individual pipeline stages. */
R shim_kb_phi(R x, int n_os, int m);            /* phi(x), window in space      */
R shim_kb_phi_hut(int k, int n_os, int m);      /* phi_hat(k)                   */
void shim_fft_nd(C *data, int d, int n, int sign); /* unnormalised, sign = -1 forward */

#ifdef __cplusplus
}
#endif
#endif
