#!/bin/bash
# Lookup kernel: resident CTAs per SM (register cap) sweep; recompiles only corr_lookup.cu and relinks.
set -u
P=temporally-consistent-stereo-matching_b200
for mb in 1 8 9 10; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_LOOKUP_MINBLOCKS=$mb \
       -I include -I $P/csrc -c $P/csrc/corr_lookup.cu -o $P/build/corr_lookup.o || exit 1
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  echo -n "== min blocks $mb: "
  python bench.py --skip-cpu --skip-gpu-reference --skip-e2e --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lookups_ms %.4f step %.4f' % (d['phases']['lookups_ms'], d['ms_per_step']))"
done
