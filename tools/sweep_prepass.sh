#!/bin/bash
# Times the pre-pass for several (threads per CTA, channels per thread item) variants: recompiles only corr_prepass.cu and relinks.
set -u
P=temporally-consistent-stereo-matching_b200
for cfg in "256 16" "128 64" "128 32" "256 32" "512 16"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_PRE_THREADS=$1 -DTCS_PRE_GROUP=$2 \
       -I include -I $P/csrc -c $P/csrc/corr_prepass.cu -o $P/build/corr_prepass.o || exit 1
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  echo "== threads $1 group $2"
  python tools/time_prepass.py
done
