#!/bin/bash
# Tensor-core fused lookup + 1x1: resident CTAs per SM (register cap) sweep; recompiles only corr_lookup.cu and relinks.
set -u
P=temporally-consistent-stereo-matching_b200
for mb in 4 3 5; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_ENCODE_TC_MINBLOCKS=$mb \
       -I include -I $P/csrc -c $P/csrc/corr_lookup.cu -o $P/build/corr_lookup.o || exit 1
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  echo "== min blocks $mb"; python tools/time_encode.py 2>&1 | tail -2
done
