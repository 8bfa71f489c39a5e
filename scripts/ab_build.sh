#!/bin/bash
# A/B harness (development aid): rebuild the library on the GPU box with extra nvcc defines and time the matvec.
# usage: scripts/ab_build.sh "<NVEXTRA flags>" [quick_time args...]
set -e
cd "$(dirname "$0")/.."
CS=preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200/csrc
touch $CS/*.cu $CS/*.c
make -C $CS -j8 NVEXTRA="$1" CFLAGS_EXTRA="$1" > /dev/null 2>&1
shift
python scripts/quick_time.py "$@" 2>&1 | grep "variant"
