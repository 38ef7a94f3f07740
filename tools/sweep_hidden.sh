#!/bin/bash
# Times the warp phase for several (channels per thread, loads batched) variants of warp_hidden3_kernel: recompiles only warp.cu and relinks.
set -u
P=temporally-consistent-stereo-matching_b200
for cfg in "32 8" "16 8" "64 8" "32 16" "64 16" "32 4" "128 8"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTCS_HIDDEN_CH=$1 -DTCS_HIDDEN_BATCH=$2 \
       -I include -I $P/csrc -c $P/csrc/warp.cu -o $P/build/warp.o || exit 1
  nvcc -shared -o $P/libtcs_b200.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a -cudart static || exit 1
  echo -n "== channels/thread $1 batch $2: "
  python bench.py --skip-cpu --skip-gpu-reference --skip-e2e --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('warp_ms %.4f step %.4f' % (d['phases']['warp_ms'], d['ms_per_step']))"
done
