/* ORACLE (test infrastructure).  Stand-in for NFFT3 applications/fastsum/kernels.h.
 * Only the four radial kernels named at SRC/external/nfft_interface.c:227,230,363,366
 * are restated.  The formulas are fixed by the reference's own gradient scale
 * factors (nfft_interface.c:536) and dense twins (SRC/linearalg/kernels.c:495,522,
 * 2218-2222); see SURVEY.md section 8(c) item 2. */
#ifndef ORACLE_NFFT3_SHIM_KERNELS_H
#define ORACLE_NFFT3_SHIM_KERNELS_H
#include "infft.h"
#ifdef __cplusplus
extern "C" {
#endif
C gaussian(R x, int der, const R *param);          /* exp(-x^2/c^2)            */
C xx_gaussian(R x, int der, const R *param);       /* x^2/c^2 * exp(-x^2/c^2)  */
C laplacian_rbf(R x, int der, const R *param);     /* exp(-|x|/c)              */
C der_laplacian_rbf(R x, int der, const R *param); /* |x|/c * exp(-|x|/c)      */
#ifdef __cplusplus
}
#endif
#endif
