#!/bin/bash
# A/B harness (development aid): build a variant of ONE kernel file locally with extra nvcc defines and link it with the
# other, unchanged objects into ab/libnfft4gp_b200_<tag>.so (travels to the GPU box; select it with NFFT4GP_B200_LIB).
# usage: scripts/ab_local.sh <tag> <file.cu> "<extra nvcc flags>"
# nfft3d_mma.cu is built as its two halves (spread: -DMMA_PART=1 -Xptxas -O1, gather: -DMMA_PART=2), as in csrc/Makefile.
set -e
cd "$(dirname "$0")/.."
CS=preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200/csrc
tag=$1; src=$2; extra=$3
mkdir -p ab/obj_$tag
make -C $CS -j8 > /dev/null
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -Iinclude -I$CS"
if [ "$src" == "nfft3d_mma.cu" ]; then
   $NV -DMMA_PART=1 -Xptxas -O1 $extra -c $CS/$src -o ab/obj_$tag/nfft3d_mma_spread.o
   $NV -DMMA_PART=2 $extra -c $CS/$src -o ab/obj_$tag/nfft3d_mma_gather.o
else
   $NV $extra -c $CS/$src -o ab/obj_$tag/${src%.cu}.o
fi
objs=""
for o in $CS/build/*.o; do
   b=$(basename $o)
   if [ -f ab/obj_$tag/$b ]; then objs="$objs ab/obj_$tag/$b"; else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ab/libnfft4gp_b200_$tag.so $objs \
   -cudart shared -Xlinker -Bsymbolic -Xlinker --no-undefined -ldl -lm
echo built ab/libnfft4gp_b200_$tag.so
