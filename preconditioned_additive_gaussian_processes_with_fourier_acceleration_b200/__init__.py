"""nfft4gp_b200: B200-native (sm_100a) NFFT additive-kernel matvec + preconditioned CG behind the nfft4gp C interface.

The product is the C-ABI shared library ``libnfft4gp_b200.so`` (built in-tree from ``csrc/``; see
``include/nfft4gp_b200.h``).  This package is only the Python-side loader and a thin ctypes mirror of the C interface,
used by tests and bench.py.  There is no CPU fallback: importing works without a GPU (so that symbols can be checked),
but every compute entry point fails loudly when the library or a CUDA device is missing.
"""
from .capi import (  # noqa: F401
    LIB_PATH,
    AdditiveNFFTKernel,
    AfnPrecond,
    CholPrecond,
    DenseAdditiveKernel,
    DensePlainKernel,
    FsaiPrecond,
    NysPrecond,
    build_library,
    gp_predict,
    lib,
    lanczos_quadrature_logdet,
    solver_fgmres,
    solver_lanczos,
    solver_pcg,
    rankest_nys_scaled_sample,
)

__all__ = ["LIB_PATH", "AdditiveNFFTKernel", "DenseAdditiveKernel", "DensePlainKernel", "NysPrecond", "FsaiPrecond", "AfnPrecond", "CholPrecond", "build_library", "lib", "solver_pcg", "rankest_nys_scaled_sample", "solver_fgmres", "gp_predict", "solver_lanczos", "lanczos_quadrature_logdet"]
