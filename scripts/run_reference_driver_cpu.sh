#!/bin/bash
# Build one of the reference's OWN TESTS drivers, unmodified, against the reference's own CPU code (oracle/_ref/libnfft4gp_ref.so =
# /root/reference/SRC compiled unmodified on the NFFT3 fastsum restatement) into /tmp/ref_cpu/<TESTn>/ -- the generator of the
# expected outputs under tests/golden/*_driver_reference_cpu.npz (scripts/make_driver_golden.py).  usage: ... TEST4/foo.cpp
set -e
cd "$(dirname "$0")/.."
drv=${1:-TEST4/foo.cpp}
B=$(python -c "import sysconfig;print(sysconfig.get_paths()['purelib'])")/opencv_python_headless.libs
dst=/tmp/ref_cpu/$(dirname $drv)
mkdir -p $dst
cp -r /root/reference/TESTS/$(dirname $drv)/data $dst/ && chmod -R u+w $dst
out=$dst/$(basename ${drv%.cpp}).ex
g++ -O2 -fopenmp -DNFFT4GP_USING_OPENMP -w -I/root/reference/INC -Ioracle/nfft3_shim /root/reference/TESTS/$drv -o $out \
   -L$PWD/oracle/_ref -lnfft4gp_ref -Wl,-rpath,$PWD/oracle/_ref $(ls $B/libopenblasp-*.so) -Wl,-rpath,$B -Wl,-rpath-link,$B -Wl,--allow-shlib-undefined -lm
echo built $out
