#!/bin/bash
# run scripts/quick_time.py once per A/B library in ab/ (see ab_local.sh); usage: scripts/ab_run.sh "<tags>" [quick_time args]
cd "$(dirname "$0")/.."
tags=$1; shift
for t in $tags; do
   echo "== $t"
   NFFT4GP_B200_LIB=$PWD/ab/libnfft4gp_b200_$t.so timeout 300 python scripts/quick_time.py "$@" 2>&1 | grep "variant"
done
