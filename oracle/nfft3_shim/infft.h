/* ORACLE (test infrastructure).  Stand-in for NFFT3's infft.h: the reference
 * uses only CREAL (SRC/external/nfft_interface.c:436-489,547-612) and the
 * R / C type names. */
#ifndef ORACLE_NFFT3_SHIM_INFFT_H
#define ORACLE_NFFT3_SHIM_INFFT_H
#include <complex.h>
#include <math.h>
typedef double R;
typedef double _Complex C;
#define K(x) ((R) x)
#define CREAL(z) creal(z)
#define CIMAG(z) cimag(z)
#endif
