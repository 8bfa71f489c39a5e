"""Link-time drop-in check (this container only: needs /root/reference and the reference objects compiled by `make -C oracle
ref`): each of the reference's seven TESTS drivers is linked, unmodified, against the UNCHANGED reference objects that
INTEGRATION.md keeps (utils, vector / vecops / matops, optimizer, precond.o, solvers.o) plus libnfft4gp_b200.so -- with
nfft_interface.o, pcg.o, fgmres.o, lanczos.o, chol.o, nys.o, fsai.o, kernels.o, ordering.o, rankest.o and all of NFFT3 / FFTW3
left out.  Every symbol those drivers and kept objects need from the dropped ones must come out of the library."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
PKG = os.path.join(ROOT, "preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200")
DROPPED = ("external__nfft_interface", "solvers__pcg", "solvers__fgmres", "solvers__lanczos", "preconds__chol", "preconds__nys",
           "preconds__fsai", "linearalg__kernels", "linearalg__ordering", "linearalg__rankest", "shim")
DRIVERS = ["TEST1/foo.cpp", "TEST2/foo1.cpp", "TEST2/foo2.cpp", "TEST3/foo1.cpp", "TEST3/foo2.cpp", "TEST3/foo3.cpp", "TEST4/foo.cpp"]


def _blas():
    import sysconfig
    libs = glob.glob(os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs", "libopenblasp-*.so"))
    return libs[0] if libs else None


@pytest.mark.parametrize("driver", DRIVERS)
def test_reference_driver_links_against_the_library(driver, tmp_path):
    if not os.path.isdir(os.path.join(REF, "TESTS")) or not os.path.isdir(OBJ) or _blas() is None:
        pytest.skip("needs /root/reference, oracle/_ref/obj (make -C oracle ref) and the bundled OpenBLAS")
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    pkg.build_library()
    keep = [o for o in sorted(glob.glob(os.path.join(OBJ, "*.o"))) if not os.path.basename(o).startswith(DROPPED)]
    assert len(keep) >= 12
    exe = os.path.join(str(tmp_path), "driver.ex")
    blas = _blas()
    cmd = ["g++", "-O2", "-fopenmp", "-DNFFT4GP_USING_OPENMP", "-w", f"-I{REF}/INC", "-I" + os.path.join(ROOT, "oracle", "nfft3_shim"),
           os.path.join(REF, "TESTS", driver)] + keep + ["-o", exe, f"-L{PKG}", "-lnfft4gp_b200", f"-Wl,-rpath,{PKG}", blas,
           f"-Wl,-rpath,{os.path.dirname(blas)}", f"-Wl,-rpath-link,{os.path.dirname(blas)}", "-Wl,--allow-shlib-undefined", "-lm"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    undefined = sorted({l.split("undefined reference to")[1].strip() for l in p.stderr.splitlines() if "undefined reference to" in l})
    assert p.returncode == 0 and os.path.exists(exe), f"unresolved: {undefined[:20]}"
    imported = {l.split()[-1] for l in subprocess.run(["nm", "-D", exe], capture_output=True, text=True).stdout.splitlines()
                if " U Nfft4GP" in l}
    exported = {l.split()[-1] for l in subprocess.run(["nm", "-D", os.path.join(PKG, "libnfft4gp_b200.so")], capture_output=True,
                                                     text=True).stdout.splitlines() if " T Nfft4GP" in l}
    assert imported and imported <= exported, sorted(imported - exported)
    if driver == "TEST3/foo3.cpp":   # the preconditioned NFFT GP: everything foo3.cpp:262-330 hands to the GP problem
        for sym in ("Nfft4GPNFFTAdditiveKernelParamCreate", "Nfft4GPAdditiveNFFTMatSymv", "Nfft4GPAdditiveNFFTGradMatSymv",
                    "Nfft4GPKernelAdditiveKernel", "Nfft4GPPrecondNysSetupWithKernel", "Nfft4GPPrecondNysSolve", "Nfft4GPPrecondNysTrace",
                    "Nfft4GPPrecondNysLogdet", "Nfft4GPPrecondNysDvp", "Nfft4GPAdditiveNFFTGpPredict"):
            assert sym in imported, sym
