/* ORACLE (test infrastructure, not product code).
 * Stand-in for NFFT3's private config.h, which the reference includes at
 * SRC/external/nfft_interface.h:10.  The only macro the reference consumes is
 * HAVE_COMPLEX_H (nfft_interface.h:16-18).  NFFT3 itself (github.com/NFFT/nfft,
 * >= 3.5.2 by API evidence, version un-pinned by the reference) is absent from
 * this environment; the restatement is pinned on the reference's shipped TEST2 output (see fastsum.h). */
#ifndef ORACLE_NFFT3_SHIM_CONFIG_H
#define ORACLE_NFFT3_SHIM_CONFIG_H
#define HAVE_COMPLEX_H 1
#endif
