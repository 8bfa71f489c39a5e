/* ORACLE (test infrastructure).  Stand-in for the two allocation entry points
 * of nfft3.h used at SRC/external/nfft_interface.c:223-224,94,97. */
#ifndef ORACLE_NFFT3_SHIM_NFFT3_H
#define ORACLE_NFFT3_SHIM_NFFT3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void *nfft_malloc(size_t n);
void nfft_free(void *p);
#ifdef __cplusplus
}
#endif
#endif
