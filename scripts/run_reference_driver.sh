#!/bin/bash
# Build one of the reference's OWN TESTS drivers, unmodified, against libnfft4gp_b200.so plus the reference objects that
# INTEGRATION.md keeps (no NFFT3, no FFTW3), into examples/ref_link/ (git-ignored; travels to the GPU box with gpurun).
# Needs /root/reference and `make -C oracle ref` (the reference objects).  usage: scripts/run_reference_driver.sh TEST2/foo2.cpp
# then e.g.:  gpurun -- 'cd examples/ref_link && ./foo2.ex foo ./data/foo ./data/foo.window -1 -1 0 0 0 0 10 10 10 10 1 1'
set -e
cd "$(dirname "$0")/.."
drv=${1:-TEST2/foo2.cpp}
PKG=preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200
B=$(python -c "import sysconfig;print(sysconfig.get_paths()['purelib'])")/opencv_python_headless.libs
KEEP=$(ls oracle/_ref/obj/*.o | grep -v "external__nfft_interface\|solvers__pcg\|solvers__fgmres\|solvers__lanczos\|preconds__chol\|preconds__nys\|preconds__fsai\|linearalg__kernels\|linearalg__ordering\|linearalg__rankest\|shim")
# TEST2 keeps its place (tests/test_reference_driver_gpu.py); the other drivers get a directory of their own (each has a data/ folder)
dst=examples/ref_link
if [ "$(dirname $drv)" != "TEST2" ]; then dst=examples/ref_link/$(dirname $drv); fi
mkdir -p $dst
cp -r /root/reference/TESTS/$(dirname $drv)/data $dst/ && chmod -R u+w $dst
out=$dst/$(basename ${drv%.cpp}).ex
g++ -O2 -fopenmp -DNFFT4GP_USING_OPENMP -w -I/root/reference/INC -Ioracle/nfft3_shim /root/reference/TESTS/$drv $KEEP -o $out \
   -L$PKG -lnfft4gp_b200 -Wl,--disable-new-dtags -Wl,-rpath,/root/repo/$PKG $(ls $B/libopenblasp-*.so) -Wl,-rpath,$B -Wl,-rpath-link,$B \
   -Wl,--allow-shlib-undefined -lm
echo built $out
